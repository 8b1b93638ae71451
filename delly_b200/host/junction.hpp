// junction.hpp — from per-read alignment junctions to split-read SV candidates.
// Same entry points and results as the reference's junction selection:
//   _insertJunction        src/junction.h:22-41   (here from explicit flag / position / length fields)
//   cigarJunctions         the CIGAR walk of src/shortpe.h:360-389 (soft/hard clips -> junctions)
//   selectDeletions / Duplications / Inversions / Insertions / Translocations   src/junction.h:60-278
//   bridgeInsertions       src/junction.h:282-316
//   fetchSVs               src/junction.h:463-475
// readBp maps a read id to its junctions sorted by (seqpos, refidx, refpos, scleft) — Junction::operator<.
// Every selector looks at ordered junction pairs (i<j) of one read; the five reference loops differ only in
// the pair predicate and in the record they emit, which is how they are written here.
#pragma once
#include <map>
#include <unordered_set>
#include <vector>

#include "types.hpp"

namespace dellyb200 {

typedef std::vector<Junction> TJunctionVector;
typedef std::vector<std::vector<SRBamRecord> > TSvtSRBamRecord;

// BAM flag bits used below (SAM spec)
constexpr uint32_t BAM_FUNMAP_ = 0x4, BAM_FREVERSE_ = 0x10, BAM_FSECONDARY_ = 0x100, BAM_FQCFAIL_ = 0x200, BAM_FDUP_ = 0x400,
                   BAM_FSUPPLEMENTARY_ = 0x800;

// src/junction.h:22-41. seqlen = query length incl. clips (readLength, src/util.h:430-438).
template <typename TReadBp>
inline void _insertJunction(TReadBp& readBp, std::size_t seed, uint32_t flag, int32_t tid, int32_t pos, uint8_t mapq, int32_t seqlen, int32_t rp,
                            int32_t sp, bool scleft) {
  const bool rev = (flag & BAM_FREVERSE_) != 0;
  int32_t readStart = pos;
  if (flag & (BAM_FQCFAIL_ | BAM_FDUP_ | BAM_FUNMAP_ | BAM_FSECONDARY_ | BAM_FSUPPLEMENTARY_)) readStart = -1;
  if (sp > seqlen) return;
  readBp[seed].push_back(Junction(!rev, scleft, tid, readStart, rp, rev ? seqlen - sp : sp, mapq));
}

// The CIGAR scan of src/shortpe.h:360-389: deletions / insertions longer than minRefSep yield a junction pair
// (right-clip-like before the event, left-clip-like after it); soft/hard clips longer than minClip yield one
// junction (left clip if nothing of the read has been consumed yet). cigar = (op,len) pairs with BAM op codes
// (0 M, 1 I, 2 D, 3 N, 4 S, 5 H, 7 =, 8 X).
template <typename TReadBp>
inline void cigarJunctions(TReadBp& readBp, std::size_t seed, uint32_t flag, int32_t tid, int32_t pos, uint8_t mapq,
                           std::vector<std::pair<uint8_t, uint32_t> > const& cigar, uint32_t minClip, uint32_t minRefSep) {
  int32_t seqlen = 0;
  for (auto const& c : cigar)
    if (c.first == 0 || c.first == 7 || c.first == 8 || c.first == 1 || c.first == 4 || c.first == 5) seqlen += (int32_t) c.second;
  auto put = [&](uint32_t rp, uint32_t sp, bool scleft) { _insertJunction(readBp, seed, flag, tid, pos, mapq, seqlen, (int32_t) rp, (int32_t) sp, scleft); };
  uint32_t rp = (uint32_t) pos, sp = 0;
  for (auto const& c : cigar) {
    const uint8_t op = c.first;
    const uint32_t len = c.second;
    if (op == 0 || op == 7 || op == 8) { sp += len; rp += len; }
    else if (op == 2) { if (len > minRefSep) put(rp, sp, false); rp += len; if (len > minRefSep) put(rp, sp, true); }
    else if (op == 1) { if (len > minRefSep) put(rp, sp, false); sp += len; if (len > minRefSep) put(rp, sp, true); }
    else if (op == 4 || op == 5) {
      const bool scleft = (sp == 0);
      const uint32_t finalsp = scleft ? sp + len : sp;
      sp += len;
      if (len > minClip) put(rp, finalsp, scleft);
    } else if (op == 3) rp += len;
  }
}

// The CIGAR scan of the long-read path (findJunctions, src/junction.h:363-437): like cigarJunctions, but a deletion /
// insertion longer than minRefSep is EXTENDED over following short aligned stretches and further indels of the same kind
// as long as (aligned bases since the event) / (event length + reference (resp. read) bases since) stays <= indelExtension
// (a float in the reference's config, compared in double); the closing junction carries the extended coordinates.
template <typename TReadBp>
inline void cigarJunctionsLR(TReadBp& readBp, std::size_t seed, uint32_t flag, int32_t tid, int32_t pos, uint8_t mapq,
                             std::vector<std::pair<uint8_t, uint32_t> > const& cigar, uint32_t minClip, uint32_t minRefSep, float indelExtension) {
  int32_t seqlen = 0;
  for (auto const& c : cigar)
    if (c.first == 0 || c.first == 7 || c.first == 8 || c.first == 1 || c.first == 4 || c.first == 5) seqlen += (int32_t) c.second;
  auto put = [&](uint32_t rp, uint32_t sp, bool scleft) { _insertJunction(readBp, seed, flag, tid, pos, mapq, seqlen, (int32_t) rp, (int32_t) sp, scleft); };
  auto isAligned = [](uint8_t op) { return op == 0 || op == 7 || op == 8; };
  uint32_t rp = (uint32_t) pos, sp = 0;
  for (std::size_t i = 0; i < cigar.size(); ++i) {
    const uint8_t op = cigar[i].first;
    const uint32_t len = cigar[i].second;
    if (isAligned(op)) { sp += len; rp += len; }
    else if (op == 2) {  // deletion (:367-394)
      if (len > minRefSep) put(rp, sp, false);
      rp += len;
      if (len > minRefSep) {
        const uint32_t spOrig = sp;
        uint32_t rpTmp = rp, spTmp = sp, dlen = len;
        for (std::size_t j = i + 1; j < cigar.size(); ++j) {
          const uint8_t oj = cigar[j].first;
          const uint32_t lj = cigar[j].second;
          if (isAligned(oj)) {
            spTmp += lj; rpTmp += lj;
            if ((double) (spTmp - sp) / (double) (dlen + (rpTmp - rp)) > indelExtension) break;
          } else if (oj == 2) {
            rpTmp += lj;
            if (lj > minRefSep) { dlen += (rpTmp - rp); rp = rpTmp; sp = spTmp; i = j; }
          } else if (oj == 1) {
            if (lj > minRefSep) break;
            spTmp += lj;
          } else break;
        }
        put(rp, spOrig, true);
      }
    } else if (op == 1) {  // insertion (:395-424)
      if (len > minRefSep) put(rp, sp, false);
      sp += len;
      if (len > minRefSep) {
        const uint32_t rpOrig = rp;
        uint32_t rpTmp = rp, spTmp = sp, ilen = len;
        for (std::size_t j = i + 1; j < cigar.size(); ++j) {
          const uint8_t oj = cigar[j].first;
          const uint32_t lj = cigar[j].second;
          if (isAligned(oj)) {
            spTmp += lj; rpTmp += lj;
            if ((double) (rpTmp - rp) / (double) (ilen + (spTmp - sp)) > indelExtension) break;
          } else if (oj == 2) {
            if (lj > minRefSep) break;
            rpTmp += lj;
          } else if (oj == 1) {
            spTmp += lj;
            if (lj > minRefSep) { ilen += (spTmp - sp); rp = rpTmp; sp = spTmp; i = j; }
          } else break;
        }
        put(rpOrig, sp, true);
      }
    } else if (op == 3) rp += len;
    else if (op == 4 || op == 5) {
      const bool scleft = (sp == 0);
      const uint32_t finalsp = scleft ? sp + len : sp;
      sp += len;
      if (len > minClip) put(rp, finalsp, scleft);
    }
  }
}

inline int32_t _selectReadStart(TJunctionVector const& jc) {
  for (auto const& j : jc) if (j.rstart != -1) return j.rstart;
  return -1;
}
inline int32_t _selectPrimaryChr(TJunctionVector const& jc) {
  for (auto const& j : jc) if (j.rstart != -1) return j.refidx;
  return -1;
}

namespace detail {
enum PairVerdict { PAIR_NEXT_J = 0, PAIR_STOP_I = 1 };

// Visit junction pairs (i<j) of every multi-junction read in seqpos order; stop the inner scan once the
// junctions are more than maxReadSep apart in the read (when `windowed`) or when the visitor asks to.
template <typename TReadBp, typename TVisit>
inline void forJunctionPairs(Config const& c, TReadBp const& readBp, bool windowed, TVisit visit) {
  for (typename TReadBp::const_iterator it = readBp.begin(); it != readBp.end(); ++it) {
    TJunctionVector const& jv = it->second;
    if (jv.size() < 2) continue;
    for (uint32_t i = 0; i < jv.size(); ++i)
      for (uint32_t j = i + 1; j < jv.size(); ++j) {
        if (windowed && (uint32_t) (jv[j].seqpos - jv[i].seqpos) > c.maxReadSep) break;
        if (visit(it->first, jv, jv[i], jv[j]) == PAIR_STOP_I) break;
      }
  }
}

// SRBamRecord of a junction pair, left breakpoint first
inline SRBamRecord pairRecord(Junction const& a, Junction const& b, int32_t rst, Junction const& i, Junction const& j, int32_t inslen, std::size_t id) {
  const int32_t qval = (int32_t) (((int32_t) i.qual + (int32_t) j.qual) / 2);
  return SRBamRecord(a.refidx, a.refpos, b.refidx, b.refpos, rst, std::min(j.seqpos, i.seqpos), qval, inslen, id);
}
}  // namespace detail

// src/junction.h:60-110 — same contig and strand, opposing clips, reference gap longer than the read gap
template <typename TReadBp>
inline void selectDeletions(Config const& c, TReadBp const& readBp, TSvtSRBamRecord& br) {
  detail::forJunctionPairs(c, readBp, true, [&](std::size_t id, TJunctionVector const& jv, Junction const& i, Junction const& j) {
    if (!((j.refidx == i.refidx) && (j.forward == i.forward) && (i.scleft != j.scleft))) return detail::PAIR_NEXT_J;
    const bool iLeft = i.refpos <= j.refpos;           // i is the left breakpoint on the reference
    const int32_t dRef = std::abs(j.refpos - i.refpos), dSeq = j.seqpos - i.seqpos;
    // the clip orientation a true deletion shows, per strand, decides the sign of the read-gap correction
    int32_t dellen = 0;
    if (i.forward) { if (!i.scleft) dellen = iLeft ? dRef - dSeq : 0; else dellen = iLeft ? 0 : dRef + dSeq; }
    else { if (i.scleft) dellen = iLeft ? 0 : dRef - dSeq; else dellen = iLeft ? dRef + dSeq : 0; }
    if (dellen <= (int32_t) c.minRefSep) return detail::PAIR_NEXT_J;
    const int32_t rst = _selectReadStart(jv);
    if (iLeft) { if (!i.scleft && j.scleft) br[2].push_back(detail::pairRecord(i, j, rst, i, j, std::abs(dSeq), id)); }
    else { if (i.scleft && !j.scleft) br[2].push_back(detail::pairRecord(j, i, rst, i, j, std::abs(dSeq), id)); }
    return detail::PAIR_STOP_I;  // only the first SV split per junction (src/junction.h:103)
  });
}

// src/junction.h:114-146
template <typename TReadBp>
inline void selectDuplications(Config const& c, TReadBp const& readBp, TSvtSRBamRecord& br) {
  detail::forJunctionPairs(c, readBp, true, [&](std::size_t id, TJunctionVector const& jv, Junction const& i, Junction const& j) {
    if (!((j.refidx == i.refidx) && (j.forward == i.forward) && (i.scleft != j.scleft))) return detail::PAIR_NEXT_J;
    if (!((uint32_t) std::abs(j.refpos - i.refpos) > c.minRefSep)) return detail::PAIR_NEXT_J;
    const int32_t rst = _selectReadStart(jv), il = std::abs(j.seqpos - i.seqpos);
    if (i.refpos <= j.refpos) { if (i.scleft && !j.scleft) br[3].push_back(detail::pairRecord(i, j, rst, i, j, il, id)); }
    else { if (!i.scleft && j.scleft) br[3].push_back(detail::pairRecord(j, i, rst, i, j, il, id)); }
    return detail::PAIR_NEXT_J;
  });
}

// src/junction.h:149-180 — same contig, opposite strands, agreeing clips; left clip => 5to5 (svt 1), right => 3to3 (svt 0)
template <typename TReadBp>
inline void selectInversions(Config const& c, TReadBp const& readBp, TSvtSRBamRecord& br) {
  detail::forJunctionPairs(c, readBp, true, [&](std::size_t id, TJunctionVector const& jv, Junction const& i, Junction const& j) {
    if (!((j.refidx == i.refidx) && (j.forward != i.forward) && (i.scleft == j.scleft))) return detail::PAIR_NEXT_J;
    if (!((uint32_t) std::abs(j.refpos - i.refpos) > c.minRefSep)) return detail::PAIR_NEXT_J;
    const int32_t rst = _selectReadStart(jv), il = std::abs(j.seqpos - i.seqpos);
    const bool iLeft = i.refpos <= j.refpos;
    br[i.scleft ? 1 : 0].push_back(iLeft ? detail::pairRecord(i, j, rst, i, j, il, id) : detail::pairRecord(j, i, rst, i, j, il, id));
    return detail::PAIR_NEXT_J;
  });
}

// src/junction.h:183-224 — small reference footprint, large separation in the read
template <typename TReadBp>
inline void selectInsertions(Config const& c, TReadBp const& readBp, TSvtSRBamRecord& br) {
  detail::forJunctionPairs(c, readBp, false, [&](std::size_t id, TJunctionVector const& jv, Junction const& i, Junction const& j) {
    if (!((j.refidx == i.refidx) && (j.forward == i.forward) && (i.scleft != j.scleft))) return detail::PAIR_NEXT_J;
    if (!((uint32_t) std::abs(j.refpos - i.refpos) < c.maxReadSep)) return detail::PAIR_NEXT_J;
    const int32_t dSeq = j.seqpos - i.seqpos, dRef = j.refpos - i.refpos;  // dRef signed: both branches of the reference reduce to this
    int32_t isizelen = 0;
    if (i.forward) { if (!i.scleft) isizelen = dSeq - dRef; }
    else { if (i.scleft) isizelen = dSeq + dRef; }
    if (!((isizelen > (int32_t) c.minRefSep) && (isizelen <= std::max(i.seqpos, j.seqpos)))) return detail::PAIR_NEXT_J;
    const int32_t rst = _selectReadStart(jv);
    br[4].push_back((i.refpos <= j.refpos) ? detail::pairRecord(i, j, rst, i, j, isizelen, id) : detail::pairRecord(j, i, rst, i, j, isizelen, id));
    return detail::PAIR_STOP_I;
  });
}

// src/junction.h:228-278 — junctions on different contigs; the record's chr is the HIGHER contig index
template <typename TReadBp>
inline void selectTranslocations(Config const& c, TReadBp const& readBp, TSvtSRBamRecord& br) {
  detail::forJunctionPairs(c, readBp, true, [&](std::size_t id, TJunctionVector const& jv, Junction const& i, Junction const& j) {
    if (j.refidx == i.refidx) return detail::PAIR_NEXT_J;
    Junction const& lowChr = (i.refidx < j.refidx) ? i : j;   // "chr1ev"
    Junction const& highChr = (i.refidx < j.refidx) ? j : i;  // "chr2ev"
    int32_t ct;
    if (lowChr.forward == highChr.forward) {
      if (lowChr.scleft == highChr.scleft) return detail::PAIR_NEXT_J;
      ct = lowChr.scleft ? 2 : 3;  // 3to5 : 5to3
    } else {
      if (lowChr.scleft != highChr.scleft) return detail::PAIR_NEXT_J;
      ct = lowChr.scleft ? 1 : 0;  // 5to5 : 3to3
    }
    SRBamRecord rec = detail::pairRecord(highChr, lowChr, _selectReadStart(jv), i, j, std::abs(j.seqpos - i.seqpos), id);
    rec.primaryChr = _selectPrimaryChr(jv);
    br[DELLY_SVT_TRANS + ct].push_back(rec);
    return detail::PAIR_NEXT_J;
  });
}

// src/junction.h:282-316 — single-junction reads that end inside a known insertion footprint support it too
template <typename TReadBp>
inline void bridgeInsertions(TReadBp const& readBp, TSvtSRBamRecord& br) {
  std::unordered_set<std::size_t> readIds;
  std::map<std::pair<uint32_t, uint32_t>, int32_t> pins;  // (contig, position) -> running mean insertion length
  const std::size_t n0 = br[4].size();
  for (std::size_t i = 0; i < n0; ++i) {
    readIds.insert(br[4][i].id);
    for (int32_t k = br[4][i].pos; k <= br[4][i].pos2; ++k) {
      auto key = std::make_pair((uint32_t) br[4][i].chr, (uint32_t) k);
      auto it = pins.find(key);
      if (it == pins.end()) pins.insert(std::make_pair(key, br[4][i].inslen));
      else it->second = (it->second + br[4][i].inslen) / 2;
    }
  }
  if (pins.empty()) return;
  for (typename TReadBp::const_iterator it = readBp.begin(); it != readBp.end(); ++it) {
    if (readIds.count(it->first)) continue;
    for (auto const& j : it->second) {
      auto hit = pins.find(std::make_pair((uint32_t) j.refidx, (uint32_t) j.refpos));
      if (hit == pins.end()) continue;
      br[4].push_back(SRBamRecord(j.refidx, j.refpos, j.refidx, j.refpos + 1, _selectReadStart(it->second), j.seqpos, (int32_t) j.qual, hit->second, it->first));
    }
  }
}

// src/junction.h:463-475 with an empty svtset (all SV types)
template <typename TReadBp>
inline void fetchSVs(Config const& c, TReadBp& readBp, TSvtSRBamRecord& br) {
  if (c.wantSvt(2)) selectDeletions(c, readBp, br);
  if (c.wantSvt(3)) selectDuplications(c, readBp, br);
  if (c.svtset.empty() || c.svtset.count(0) || c.svtset.count(1)) selectInversions(c, readBp, br);
  if (c.wantSvt(4)) { selectInsertions(c, readBp, br); bridgeInsertions(readBp, br); }
  if (c.svtset.empty() || c.svtset.count(5) || c.svtset.count(6) || c.svtset.count(7) || c.svtset.count(8)) selectTranslocations(c, readBp, br);
}

}  // namespace dellyb200
