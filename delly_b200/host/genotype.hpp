// genotype.hpp — short-read genotyping realignment, the batch seam of the reference
// (AlignJob / AlignResult / process_batch, src/coverage.h:87-115, :412-454), with the two
// _editDistanceHW calls per job replaced by ONE dgpu_edit_distance call for the whole batch.
// Also _editDistanceNW for the long-read path (src/genotype.h:22-29).
#pragma once
#include <string>
#include <vector>

#include "../../include/dgpu.h"
#include "types.hpp"

namespace dellyb200 {

struct AlignJob {  // src/coverage.h:87-96
  std::string consProbe, refProbe, sequence;
  uint32_t fileIndex, svId;
  uint8_t qual;
};

struct AlignResult {  // src/coverage.h:98-105
  uint32_t fileIndex = 0, svId = 0;
  char type = 'N';  // 'R' ref, 'A' alt, 'N' none
  uint8_t qual = 0;
};

// k exactly as edlibNewAlignConfig receives it at src/coverage.h:111: int(2 * 0.95f * |q|) evaluated in float
inline int32_t _hwBound(Config const& c, std::size_t qsize) { return (int32_t) (2 * c.flankQuality * qsize); }

// score of _editDistanceHW (src/coverage.h:109-113) from the device's edit distance
inline double _hwScore(Config const& c, std::size_t qsize, int32_t editDistance) {
  if (editDistance == -1) return 0;
  return ((1.0 - c.flankQuality) * (double) qsize) / (double) (editDistance + 1);
}

// process_batch (src/coverage.h:412-441): results[i] for jobs[i]; the merge into countMap (:442-453) stays with the caller.
inline int processBatch(dgpu_ctx* ctx, Config const& c, std::vector<AlignJob> const& jobs, std::vector<AlignResult>& results) {
  const std::size_t J = jobs.size();
  results.assign(J, AlignResult());
  if (!J) return DGPU_OK;
  std::string arena;
  std::vector<uint32_t> qo(2 * J), ql(2 * J), to(2 * J), tl(2 * J);
  std::vector<int32_t> k(2 * J), dist(2 * J);
  for (std::size_t i = 0; i < J; ++i) {
    const AlignJob& j = jobs[i];
    const uint32_t so = (uint32_t) arena.size(); arena += j.sequence;
    qo[2 * i] = (uint32_t) arena.size(); ql[2 * i] = (uint32_t) j.consProbe.size(); arena += j.consProbe;
    qo[2 * i + 1] = (uint32_t) arena.size(); ql[2 * i + 1] = (uint32_t) j.refProbe.size(); arena += j.refProbe;
    to[2 * i] = to[2 * i + 1] = so; tl[2 * i] = tl[2 * i + 1] = (uint32_t) j.sequence.size();
    k[2 * i] = _hwBound(c, j.consProbe.size()); k[2 * i + 1] = _hwBound(c, j.refProbe.size());
  }
  int rc = dgpu_edit_distance(ctx, (const uint8_t*) arena.data(), arena.size(), qo.data(), ql.data(), to.data(), tl.data(), k.data(),
                              DGPU_MODE_HW, 2 * J, dist.data(), nullptr);
  if (rc) return rc;
  for (std::size_t i = 0; i < J; ++i) {
    const double scoreAlt = _hwScore(c, jobs[i].consProbe.size(), dist[2 * i]);
    const double scoreRef = _hwScore(c, jobs[i].refProbe.size(), dist[2 * i + 1]);
    if ((scoreRef > 0.7) || (scoreAlt > 0.7)) {
      results[i].svId = jobs[i].svId;
      results[i].fileIndex = jobs[i].fileIndex;
      const bool isRef = scoreRef > scoreAlt;
      results[i].type = isRef ? 'R' : 'A';
      results[i].qual = (uint8_t) std::min(255, std::min((int) ((isRef ? scoreRef : scoreAlt) * 35), (int) jobs[i].qual));
    }
  }
  return DGPU_OK;
}

// Batched _editDistanceNW (src/genotype.h:22-29): global edit distance of (query, target) pairs.
inline int editDistanceNWBatch(dgpu_ctx* ctx, std::vector<std::pair<std::string, std::string> > const& pairs, std::vector<int32_t>& dist) {
  dist.assign(pairs.size(), 0);
  if (pairs.empty()) return DGPU_OK;
  std::string arena;
  std::vector<uint32_t> qo, ql, to, tl;
  for (auto const& p : pairs) {
    qo.push_back((uint32_t) arena.size()); ql.push_back((uint32_t) p.first.size()); arena += p.first;
    to.push_back((uint32_t) arena.size()); tl.push_back((uint32_t) p.second.size()); arena += p.second;
  }
  return dgpu_edit_distance(ctx, (const uint8_t*) arena.data(), arena.size(), qo.data(), ql.data(), to.data(), tl.data(), nullptr, DGPU_MODE_NW,
                            pairs.size(), dist.data(), nullptr);
}

}  // namespace dellyb200
