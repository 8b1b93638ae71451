// svanno.hpp — reference-based SV annotation of the long-read path, the host mirror of
//   src/svanno.h:38-53   detectTandemRepeat (first period whose lag-p self-match fraction reaches the threshold)
//   src/svanno.h:55-238  annotateSV: breakpoint homology / target-site duplication, mobile-element / NUMT / LTR / HERVK
//                        classification of inserted sequence, tandem-repeat classification of insertions and deletions
// called once per SV from genotypeLR (src/genotype.h:155-163) for every non-translocation SV of the chromosome in memory.
//
// The reference runs up to fourteen edlibAlign calls per insertion, one SV at a time (HW mode, distance only, k = -1:
// src/svanno.h:154,164,209). Here the SVs of a chromosome are annotated together in two device rounds:
//   round 1: every (inserted sequence, template) pair, forward and reverse-complemented template — ONE dgpu_edit_distance
//            call; the templates sit once in the arena and every job points at them;
//   round 2: the flank-repeat templates of the insertions round 1 left unclassified (both flanks; the right flank's
//            distance is only consulted if the left one did not classify, which is the reference's loop order).
// Homology scans and the autocorrelation are byte loops over the chromosome and stay on the host.
//
// The template sequences (class MEI, src/svanno.h:30-36) are DATA of the reference and are passed in by the caller
// (MeiTemplates); they are not part of this repository.
#pragma once
#include <algorithm>
#include <cctype>
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

#include "../../include/dgpu.h"
#include "split.hpp"
#include "types.hpp"

namespace dellyb200 {

struct MeiTemplates {
  std::string seq[7];  // [1..6] = Alu, LINE1, SVA, NUMT, solo LTR, HERVK (SVAnno::seqType numbering, src/tags.h:84); [0] unused
  std::string polyA;   // appended to the three retrotransposon templates (seqtype < 4, src/svanno.h:142)
};

struct AnnoConfig {  // src/tegua.h:63-64, defaults :242-243
  float meiMinFrac = 0.8f;
  float trMinFrac = 0.85f;
};

namespace detail {
inline int upc(char c) { return std::toupper((unsigned char) c); }
inline std::string upperCopy(const char* b, const char* e) {
  std::string s(b, e);
  for (char& ch : s) ch = (char) std::toupper((unsigned char) ch);
  return s;
}
}  // namespace detail

// src/svanno.h:38-53
inline std::pair<int32_t, float> detectTandemRepeat(std::string const& s, int32_t maxPeriod = 100, float minFraction = 0.85f) {
  const int32_t n = (int32_t) s.size();
  if (n < 10) return std::make_pair(0, 0.0f);
  const int32_t last = std::min(maxPeriod, n / 2);
  for (int32_t period = 1; period <= last; ++period) {
    int32_t same = 0;
    for (int32_t i = period; i < n; ++i) same += (s[i] == s[i - period]);
    if ((float) same / (float) (n - period) >= minFraction) return std::make_pair(period, (float) n / (float) period);
  }
  return std::make_pair(0, 0.0f);
}

// annotateSV for the SVs `which` (indices into svs), each against its own chromosome chrseq[sv.chr] of length
// target_len[sv.chr] (the reference walks the chromosomes one at a time, src/genotype.h:139-163; the annotation of one SV
// reads nothing but its own chromosome, so the SVs of all chromosomes can share the two device rounds). Fills svs[i].anno.
// Returns DGPU_OK or the device error; with no inserted sequence to classify no device call is made (ctx may be null then).
inline int annotateSVBatch(dgpu_ctx* ctx, AnnoConfig const& c, MeiTemplates const& mei, std::vector<const char*> const& chrseq,
                           std::vector<uint32_t> const& target_len, std::vector<StructuralVariantRecord>& svs, std::vector<int32_t> const& which) {
  static const int32_t minInsLen[7] = {0, 100, 150, 1000, 100, 300, 300};  // src/svanno.h:130
  const int32_t maxEdits = 2;

  // ---- pass 1 (host): inserted sequence, homology; collect the insertions that need template distances
  std::vector<std::string> insSeq(which.size());
  std::vector<std::size_t> withIns;
  for (std::size_t w = 0; w < which.size(); ++w) {
    StructuralVariantRecord& sv = svs[which[w]];
    const char* seq = chrseq[sv.chr];
    const int32_t chrLen = (int32_t) target_len[sv.chr];
    std::string altStr;
    if (sv.svt == 4 && !sv.alleles.empty()) {  // :62-73, the sequence-resolved ALT of "REF,ALT"
      const std::size_t comma = sv.alleles.find(',');
      if (comma != std::string::npos) {
        const std::string alt = sv.alleles.substr(comma + 1);
        if (alt.size() > 1 && alt[0] != '<') {
          altStr = detail::upperCopy(alt.data(), alt.data() + alt.size());
          insSeq[w] = altStr.substr(1);
          if (insSeq[w].size() > 15000) insSeq[w].resize(15000);
        }
      }
    }
    if (sv.svt == 4) {  // :76-88 target-site duplication: ALT prefix against the reference right of the insertion point
      if (!altStr.empty() && sv.svStart >= 1) {  // svStart 0 would read the base before the chromosome (the reference does; never a real call)
        const int32_t limit = std::min(std::min((int32_t) 100, (int32_t) altStr.size()), chrLen - (sv.svStart - 1));
        int32_t edits = 0;
        for (int32_t h = 1; h <= limit; ++h) {
          if (detail::upc(altStr[h - 1]) != detail::upc(seq[sv.svStart - 1 + h - 1])) {
            if (++edits > maxEdits) break;
          }
          sv.anno.homLen = h;
        }
      }
    } else if (sv.svEnd > sv.svStart) {  // :89-121 both breakpoints against each other, leftwards then rightwards
      const int32_t half = (sv.svEnd - sv.svStart) / 2;
      for (int dir = 0; dir < 2; ++dir) {
        int32_t limit = dir ? std::min(std::min((int32_t) 10000, half), chrLen - sv.svEnd - 1) : std::min(std::min((int32_t) 10000, half), sv.svStart);
        if (dir && limit < 0) limit = 0;
        int32_t edits = 0, lastGood = 0;
        for (int32_t h = 1; h <= limit; ++h) {
          const int32_t a = dir ? sv.svStart + h : sv.svStart - h;
          const int32_t b = dir ? sv.svEnd + h : sv.svEnd - h;
          if (detail::upc(seq[a]) != detail::upc(seq[b])) ++edits;
          const double identity = 1.0 - (double) edits / h;
          if (identity >= 0.90) {
            if (!dir || h > sv.anno.homLen) sv.anno.homLen = h;
            lastGood = h;
          } else if (h - lastGood > 100 && identity < 0.75) break;
        }
      }
    }
    if (!insSeq[w].empty()) withIns.push_back(w);
  }

  // ---- round 1 (device): inserted sequence x {template, reverse-complemented template}, HW distance, unbounded
  struct Tpl { std::string fwd, rev; uint32_t offF = 0, offR = 0; };
  Tpl tpl[7];
  std::string arena;
  std::vector<uint32_t> qo, ql, to, tl;
  struct MeiJob { std::size_t w; int32_t type; int32_t qlen, tlen; bool longIns; };
  std::vector<MeiJob> meiJobs;
  if (!withIns.empty()) {
    for (int t = 1; t <= 6; ++t) {
      tpl[t].fwd = mei.seq[t];
      if (t < 4) tpl[t].fwd += mei.polyA;
      tpl[t].rev = tpl[t].fwd;
      reverseComplement(tpl[t].rev);
      tpl[t].offF = (uint32_t) arena.size(); arena += tpl[t].fwd;
      tpl[t].offR = (uint32_t) arena.size(); arena += tpl[t].rev;
    }
    for (std::size_t w : withIns) {
      const uint32_t insOff = (uint32_t) arena.size();
      const uint32_t insLen = (uint32_t) insSeq[w].size();
      arena += insSeq[w];
      for (int t = 1; t <= 6; ++t) {
        if ((int32_t) insLen < minInsLen[t]) continue;
        const uint32_t tplLen = (uint32_t) tpl[t].fwd.size();
        const bool longIns = insLen > tplLen;  // the shorter of the two is the query (:147-151)
        for (int strand = 0; strand < 2; ++strand) {
          const uint32_t tOff = strand ? tpl[t].offR : tpl[t].offF;
          qo.push_back(longIns ? tOff : insOff); ql.push_back(longIns ? tplLen : insLen);
          to.push_back(longIns ? insOff : tOff); tl.push_back(longIns ? insLen : tplLen);
        }
        meiJobs.push_back(MeiJob{w, t, (int32_t) (longIns ? tplLen : insLen), (int32_t) (longIns ? insLen : tplLen), longIns});
      }
    }
  }
  std::vector<int32_t> dist(qo.size());
  if (!qo.empty()) {
    if (!ctx) return DGPU_ERR_NODEVICE;
    int rc = dgpu_edit_distance(ctx, (const uint8_t*) arena.data(), arena.size(), qo.data(), ql.data(), to.data(), tl.data(), nullptr, DGPU_MODE_HW,
                                qo.size(), dist.data(), nullptr);
    if (rc) return rc;
  }
  // best class per insertion, in template order (:124-180)
  std::vector<double> bestId(which.size(), (double) c.meiMinFrac), bestFwd(which.size(), 0.0), bestRev(which.size(), 0.0);
  std::vector<int32_t> bestType(which.size(), 0);
  for (std::size_t j = 0; j < meiJobs.size(); ++j) {
    MeiJob const& m = meiJobs[j];
    const double fwdId = (dist[2 * j] >= 0) ? 1.0 - (double) dist[2 * j] / m.qlen : 0.0;
    const double revId = (dist[2 * j + 1] >= 0) ? 1.0 - (double) dist[2 * j + 1] / m.qlen : 0.0;
    const double coverage = m.longIns ? ((double) m.qlen / m.tlen) : 1.0;
    const double effective = std::min(std::max(fwdId, revId), coverage);
    if (effective > bestId[m.w]) { bestId[m.w] = effective; bestType[m.w] = m.type; bestFwd[m.w] = fwdId; bestRev[m.w] = revId; }
  }

  // ---- pass 2 (host): classified insertions; tandem repeats by autocorrelation; flank-repeat templates for the rest
  struct FlankJob { std::size_t w; int32_t side, period; };
  std::vector<FlankJob> flankJobs;
  arena.clear(); qo.clear(); ql.clear(); to.clear(); tl.clear();
  for (std::size_t w = 0; w < which.size(); ++w) {
    StructuralVariantRecord& sv = svs[which[w]];
    const char* seq = chrseq[sv.chr];
    const int32_t chrLen = (int32_t) target_len[sv.chr];
    std::string const& ins = insSeq[w];
    if (!ins.empty()) {
      if (bestType[w] > 0) {
        sv.anno.seqType = bestType[w];
        sv.anno.isRC = bestRev[w] > bestFwd[w];
      } else if (ins.size() >= 10) {
        const std::pair<int32_t, float> tr = detectTandemRepeat(ins, 100, c.trMinFrac);  // :186
        if (tr.first > 0) {
          sv.anno.seqType = 7; sv.anno.trPeriod = tr.first; sv.anno.trCopies = tr.second;
        } else {
          const int32_t flank = 200;
          uint32_t insOff = 0;
          bool insPlaced = false;
          for (int32_t side = 0; side < 2; ++side) {  // :192-217 repeat unit of the left / right reference flank
            std::string win;
            if (side == 0) {
              const int32_t beg = std::max(0, sv.svStart - flank);
              if (sv.svStart - beg < 40) continue;
              win = detail::upperCopy(seq + beg, seq + sv.svStart);
            } else {
              const int32_t end = std::min(chrLen, sv.svStart + flank);
              if (end - sv.svStart < 40) continue;
              win = detail::upperCopy(seq + sv.svStart, seq + end);
            }
            const int32_t pr = detectTandemRepeat(win, 100, c.trMinFrac).first;
            if (pr <= 0 || (int32_t) win.size() < pr) continue;
            const std::string unit = side ? win.substr(0, pr) : win.substr(win.size() - pr);
            if (!insPlaced) { insOff = (uint32_t) arena.size(); arena += ins; insPlaced = true; }
            const uint32_t tOff = (uint32_t) arena.size();
            uint32_t tLen = 0;
            while ((int32_t) tLen < (int32_t) ins.size() + pr) { arena += unit; tLen += (uint32_t) unit.size(); }
            qo.push_back(insOff); ql.push_back((uint32_t) ins.size()); to.push_back(tOff); tl.push_back(tLen);
            flankJobs.push_back(FlankJob{w, side, pr});
          }
        }
      }
    } else if (sv.svt == 2) {  // :225-236 deleted sequence itself
      const int32_t delLen = sv.svEnd - sv.svStart;
      if (delLen >= 10 && delLen <= 50000) {
        const std::pair<int32_t, float> tr = detectTandemRepeat(detail::upperCopy(seq + sv.svStart, seq + sv.svEnd), 100, c.trMinFrac);
        if (tr.first > 0) { sv.anno.seqType = 7; sv.anno.trPeriod = tr.first; sv.anno.trCopies = tr.second; }
      }
    }
  }

  // ---- round 2 (device): inserted sequence inside the tiled flank unit
  if (!flankJobs.empty()) {
    if (!ctx) return DGPU_ERR_NODEVICE;
    dist.assign(flankJobs.size(), 0);
    int rc = dgpu_edit_distance(ctx, (const uint8_t*) arena.data(), arena.size(), qo.data(), ql.data(), to.data(), tl.data(), nullptr, DGPU_MODE_HW,
                                flankJobs.size(), dist.data(), nullptr);
    if (rc) return rc;
    for (std::size_t j = 0; j < flankJobs.size(); ++j) {  // jobs are in (SV, side) order: the left flank wins if both qualify
      FlankJob const& f = flankJobs[j];
      StructuralVariantRecord& sv = svs[which[f.w]];
      if (sv.anno.seqType == 7) continue;
      const double insLen = (double) insSeq[f.w].size();
      const double identity = (dist[j] >= 0) ? (1.0 - (double) dist[j] / insLen) : 0.0;
      if (identity >= 0.70) {
        sv.anno.seqType = 7;
        sv.anno.trPeriod = f.period;
        sv.anno.trCopies = (float) insSeq[f.w].size() / (float) f.period;
      }
    }
  }
  return DGPU_OK;
}

// The annotation step of genotypeLR (src/genotype.h:155-163) for a whole call set: every SV that is not a translocation,
// against the chromosome of its first breakpoint.
inline int annotateSVs(dgpu_ctx* ctx, AnnoConfig const& c, MeiTemplates const& mei, std::vector<const char*> const& chrseq,
                       std::vector<uint32_t> const& target_len, std::vector<StructuralVariantRecord>& svs) {
  std::vector<int32_t> which;
  for (std::size_t i = 0; i < svs.size(); ++i)
    if (!_translocation(svs[i].svt)) which.push_back((int32_t) i);
  return annotateSVBatch(ctx, c, mei, chrseq, target_len, svs, which);
}

}  // namespace dellyb200
