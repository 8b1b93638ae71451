// genotype.hpp — short-read genotyping realignment, the batch seam of the reference
// (AlignJob / AlignResult / process_batch, src/coverage.h:87-115, :412-454), with the two
// _editDistanceHW calls per job replaced by ONE dgpu_edit_distance call for the whole batch.
// Also _editDistanceNW for the long-read path (src/genotype.h:22-29).
#pragma once
#include <string>
#include <vector>

#include "../../include/dgpu.h"
#include "split.hpp"
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

// ---- REF / ALT probes of the genotyping pass (src/coverage.h:117-263) ------------------------------------------

struct BpRegion {  // src/coverage.h:49-67; sorted by bppos only (std::sort, unstable for ties like the reference)
  int32_t regionStart = 0, regionEnd = 0, bppos = 0, homLeft = 0, homRight = 0, svt = 0;
  uint32_t id = 0;
  uint8_t bpPoint = 0;
  BpRegion() {}
  BpRegion(int32_t rs, int32_t re, int32_t bpos, int32_t hl, int32_t hr, int32_t s, uint32_t identifier, uint8_t bpp)
      : regionStart(rs), regionEnd(re), bppos(bpos), homLeft(hl), homRight(hr), svt(s), id(identifier), bpPoint(bpp) {}
  bool operator<(BpRegion const& o) const { return bppos < o.bppos; }
};

// src/coverage.h:117-160: which end of the reference gap a breakpoint's REF probe is centred on
inline bool _cutAtREnd(unsigned int bpPoint, int32_t svt) {
  const bool flipped = _translocation(svt) ? (_getSpanOrientation(svt) == 3) : (svt == 3);
  return flipped ? !bpPoint : (bpPoint != 0);
}
inline int32_t _cutRefStart(int32_t rStart, int32_t rEnd, int32_t offset, unsigned int bpPoint, int32_t svt) {
  return (_cutAtREnd(bpPoint, svt) ? rEnd : rStart) - offset;
}
inline int32_t _cutRefEnd(int32_t rStart, int32_t rEnd, int32_t offset, unsigned int bpPoint, int32_t svt) {
  return (_cutAtREnd(bpPoint, svt) ? rEnd : rStart) + offset;
}

inline std::string _addID(int32_t svt) {  // src/util.h:237-246
  if (svt == 0 || svt == 1) return "INV";
  if (svt == 2) return "DEL";
  if (svt == 3) return "DUP";
  if (svt == 4) return "INS";
  if (svt == 9) return "CNV";
  return "BND";
}

// symbolic / breakend ALT allele (src/util.h:253-272)
inline std::string _addAlleles(std::string const& ref, std::string const& chr2, StructuralVariantRecord const& sv, int32_t svt) {
  if (_translocation(svt)) {
    const uint8_t ct = _getSpanOrientation(svt);
    const std::string pos = std::to_string(sv.svEnd);
    if (ct == 0) return ref + "," + ref + "]" + chr2 + ":" + pos + "]";
    if (ct == 1) return ref + "," + "[" + chr2 + ":" + pos + "[" + ref;
    if (ct == 2) return ref + "," + ref + "[" + chr2 + ":" + pos + "[";
    if (ct == 3) return ref + "," + "]" + chr2 + ":" + pos + "]" + ref;
  }
  return ref + ",<" + _addID(svt) + ">";
}

// std::string::substr with the reference's argument conversion (int32 -> size_t): pos > size throws there
// (std::out_of_range aborts delly); here it is reported so a batch never dies half-way.
inline bool _substrChecked(std::string const& s, int32_t start, int32_t len, std::string& out) {
  const std::size_t pos = (std::size_t) start;   // negative start wraps like the implicit conversion does
  if (pos > s.size()) return false;
  out = s.substr(pos, (std::size_t) len);
  return true;
}

// _generateProbes (src/coverage.h:164-263) for a whole SV list: the per-precise-SV _consRefAlignment + _findSplit is ONE
// consRefAlignmentBatch call (one dgpu_long_needle + the splitAlign rounds) instead of |svs| sequential DPs.
//   svs            in/out: alleles filled when empty (:185-187)
//   refProbeArr / consProbeArr   [2][svs.size()], indexed by sv.id like the reference
//   bpRegion       [n_targets], each sorted by bppos; svOnChr[n_targets]
//   bad            ids whose probe cut fell outside the string (the reference would have thrown)
inline int generateProbesBatch(dgpu_ctx* ctx, Config const& c, std::vector<uint32_t> const& target_len, std::vector<std::string> const& target_name,
                               std::vector<const char*> const& chrseq, std::vector<StructuralVariantRecord>& svs,
                               std::vector<std::vector<std::string> >& refProbeArr, std::vector<std::vector<std::string> >& consProbeArr,
                               std::vector<std::vector<BpRegion> >& bpRegion, std::vector<bool>& svOnChr, std::vector<uint32_t>& bad) {
  const std::size_t N = svs.size();
  refProbeArr.assign(2, std::vector<std::string>(N));
  consProbeArr.assign(2, std::vector<std::string>(N));
  bpRegion.assign(target_len.size(), std::vector<BpRegion>());
  svOnChr.assign(target_len.size(), false);
  bad.clear();
  std::vector<uint32_t> idx;
  std::vector<std::string> refs;
  for (std::size_t i = 0; i < N; ++i) {
    StructuralVariantRecord& sv = svs[i];
    svOnChr[sv.chr] = true;
    svOnChr[sv.chr2] = true;
    if (sv.alleles.empty()) sv.alleles = _addAlleles(detail::upperSlice(chrseq[sv.chr], sv.svStart - 1, sv.svStart), target_name[sv.chr2], sv, sv.svt);
    if (!sv.precise) continue;
    Breakpoint bp(sv);
    if (sv.svt == 4) _initBreakpoint(target_len, bp, std::max((int32_t) ((sv.consensus.size() - sv.insLen) / 3), c.minimumFlankSize), sv.svt);
    else _initBreakpoint(target_len, bp, (int32_t) sv.consensus.size(), sv.svt);
    // translocations: the chr2 part is cut first, with the breakpoint initialised from the full consensus length (:193-197)
    // (the reference visits contigs in index order and fills the chr2 part while on contig chr2, so it exists only if chr2 < chr)
    if (sv.chr != sv.chr2 && sv.chr2 < sv.chr) {
      Breakpoint bp2(sv);
      _initBreakpoint(target_len, bp2, (int32_t) sv.consensus.size(), sv.svt);
      bp.part1 = _getSVRef(c, chrseq[sv.chr2], bp2, sv.chr2, sv.svt);
    }
    refs.push_back(_getSVRef(c, chrseq[sv.chr], bp, sv.chr, sv.svt));
    idx.push_back((uint32_t) i);
  }
  if (!idx.empty()) {
    std::vector<int32_t> svts;
    std::vector<const std::string*> cp, rp;
    for (std::size_t k = 0; k < idx.size(); ++k) { svts.push_back(svs[idx[k]].svt); cp.push_back(&svs[idx[k]].consensus); rp.push_back(&refs[k]); }
    std::vector<uint8_t> aok;
    std::vector<TAlign> aligns;
    int rc = consRefAlignmentBatch(ctx, svts, cp, rp, aok, aligns);
    if (rc) return rc;
    // The reference walks contigs in index order and, per contig, the SVs whose chr is that contig (:176-254): breakpoint
    // regions are appended in that order before the (unstable) sort by bppos, so the same order is kept here.
    std::vector<uint32_t> order(idx.size());
    for (std::size_t k = 0; k < idx.size(); ++k) order[k] = (uint32_t) k;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return svs[idx[a]].chr < svs[idx[b]].chr; });
    for (uint32_t k : order) {
      if (!aok[k]) continue;
      StructuralVariantRecord const& sv = svs[idx[k]];
      std::string const& svRefStr = refs[k];
      AlignDescriptor ad;
      if (!_findSplit(c, sv.consensus, svRefStr, aligns[k], ad, sv.svt)) continue;
      for (unsigned int bpPoint = 0; bpPoint < 2; ++bpPoint) {
        int32_t regionChr, regionStart, regionEnd, cutConsStart, cutConsEnd, bppos;
        if (bpPoint) {
          regionChr = sv.chr2;
          regionStart = std::max(0, sv.svEnd - c.minimumFlankSize);
          regionEnd = (int32_t) std::min((uint32_t) (sv.svEnd + c.minimumFlankSize), target_len[sv.chr2]);
          cutConsStart = ad.cEnd - ad.homLeft - c.minimumFlankSize;
          cutConsEnd = ad.cEnd + ad.homRight + c.minimumFlankSize;
          bppos = sv.svEnd;
        } else {
          regionChr = sv.chr;
          regionStart = std::max(0, sv.svStart - c.minimumFlankSize);
          regionEnd = (int32_t) std::min((uint32_t) (sv.svStart + c.minimumFlankSize), target_len[sv.chr]);
          cutConsStart = ad.cStart - ad.homLeft - c.minimumFlankSize;
          cutConsEnd = ad.cStart + ad.homRight + c.minimumFlankSize;
          bppos = sv.svStart;
        }
        const int32_t cutRefStart = _cutRefStart(ad.rStart, ad.rEnd, ad.homLeft + c.minimumFlankSize, bpPoint, sv.svt);
        const int32_t cutRefEnd = _cutRefEnd(ad.rStart, ad.rEnd, ad.homRight + c.minimumFlankSize, bpPoint, sv.svt);
        if (!_substrChecked(sv.consensus, cutConsStart, cutConsEnd - cutConsStart, consProbeArr[bpPoint][sv.id]) ||
            !_substrChecked(svRefStr, cutRefStart, cutRefEnd - cutRefStart, refProbeArr[bpPoint][sv.id])) {
          bad.push_back(sv.id);
          break;
        }
        bpRegion[regionChr].push_back(BpRegion(regionStart, regionEnd, bppos, ad.homLeft, ad.homRight, sv.svt, sv.id, (uint8_t) bpPoint));
      }
    }
  }
  for (auto& r : bpRegion) std::sort(r.begin(), r.end());
  return DGPU_OK;
}

}  // namespace dellyb200
