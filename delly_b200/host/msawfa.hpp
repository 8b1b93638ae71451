// msawfa.hpp — long-read insertion consensus: msaWfa (src/assemble.h:549-725) with fillKmerTable /
// bestDiagonal (:502-545), buildSuperstring (:90-134), consensusWfa (:262-336), convertAlignment in HW mode (:24-88) and
// _trimConsensus (:339-365), batched over all insertion clusters of a chromosome. Every edlibAlign call of the reference
// becomes one slot of a batched device call; the cheap string bookkeeping between them stays on the host:
//   1 dgpu_edit_distance call   : NW distance of the k-mer-diagonal overlap of every read pair of every cluster (:556-579)
//   host                        : median-distance centroid, order by distance to it, drop the poorest 20 % (:582-607)
//   <= maxReads-1 rounds        : superstring extension; round i aligns the overlap of every cluster's superstring and
//                                 its i-th selected read with ONE dgpu_edit_path call (NW, :609-668)
//   <= maxReads rounds          : progressive alignment; round i places the i-th selected read inside the IUPAC
//                                 consensus of the growing alignment with ONE dgpu_edit_path_ex call (HW, 20 extra
//                                 equality pairs, :671-694)
//   host                        : coverage-thresholded consensus (src/msa.h:111-173)
//   2 calls                     : _trimConsensus — orientation by two HW distances, then prefix / suffix anchors (HW paths)
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>
#include <string>
#include <vector>

#include "../../include/dgpu.h"
#include "msaedlib.hpp"
#include "split.hpp"
#include "splitalign.hpp"

namespace dellyb200 {

constexpr uint32_t DELLY_KMER_ = 7;                                             // src/tags.h:19
constexpr uint32_t DELLY_DUPLICATE_ = std::numeric_limits<uint32_t>::max();     // src/tags.h:15
constexpr uint32_t KMER_TABLE_ = 65536;                                         // pow(4, DELLY_KMER + 1), src/assemble.h:504

// src/assemble.h:475-497
inline uint32_t charToInt(char c) {
  switch (c) {
    case 'A': case 'B': return 0;
    case 'C': case 'D': return 1;
    case 'G': case 'E': return 2;
    case 'T': case 'F': return 3;
  }
  return 0;
}

// src/assemble.h:501-520 — the (hash, 1-based position) sequence fillKmerTable walks through: one entry per 7-mer
struct KmerHit { uint32_t hash, pos; };
inline void kmerHits(std::string const& s, std::vector<KmerHit>& hits) {
  const uint32_t len = (uint32_t) s.size();
  hits.clear();
  uint32_t hash = 0;
  for (uint32_t ki = 0; ki < len && ki < DELLY_KMER_; ++ki) { hash *= 4; hash += charToInt(s[ki]); }
  for (uint32_t ki = DELLY_KMER_; ki < len; ++ki) {
    hits.push_back(KmerHit{hash, ki - DELLY_KMER_ + 1});
    hash -= charToInt(s[ki - DELLY_KMER_]) * 4 * 4 * 4 * 4 * 4 * 4;
    hash *= 4;
    hash += charToInt(s[ki]);
  }
  hits.push_back(KmerHit{hash, len - DELLY_KMER_ + 1});
}

// fillKmerTable proper: position of every unique 7-mer, DELLY_DUPLICATE for repeated ones, 0 for absent ones.
// `uniq` receives the unique ones (table value == their position), which is all bestDiagonal ever looks at.
inline void fillKmerTable(std::string const& s, std::vector<uint32_t>& kmerpos, std::vector<KmerHit>& uniq) {
  std::vector<KmerHit> hits;
  kmerHits(s, hits);
  kmerpos.assign(KMER_TABLE_, 0);
  for (KmerHit const& h : hits) {
    if (kmerpos[h.hash]) kmerpos[h.hash] = DELLY_DUPLICATE_;
    else kmerpos[h.hash] = h.pos;
  }
  uniq.clear();
  for (KmerHit const& h : hits) if (kmerpos[h.hash] == h.pos && h.pos != 0 && h.pos != DELLY_DUPLICATE_) uniq.push_back(h);
}

// src/assemble.h:522-545 — diagonal (offset of J inside I) with the most shared unique k-mers in a 20-wide window.
// The reference scans all 4^8 table slots; a slot contributes iff it is a unique k-mer of both, so walking J's unique
// k-mers and probing I's table counts exactly the same diagonals.
inline int32_t bestDiagonal(std::vector<uint32_t> const& kmerHitI, std::vector<KmerHit> const& uniqJ, uint32_t lenI, uint32_t lenJ) {
  std::vector<uint32_t> diag(lenI + lenJ, 0);
  for (KmerHit const& h : uniqJ) {
    const uint32_t hi = kmerHitI[h.hash];
    if (hi && hi != DELLY_DUPLICATE_) ++diag[lenJ + hi - h.pos];
  }
  const uint32_t window = 20;
  uint32_t windowVal = 0;
  for (uint32_t d = 0; d < diag.size() && d < window; ++d) windowVal += diag[d];
  uint32_t bestDiag = window / 2;
  uint32_t bestWindowVal = windowVal;
  for (uint32_t d = window; d < diag.size(); ++d) {
    windowVal -= diag[d - window];
    windowVal += diag[d];
    if (windowVal > bestWindowVal) { bestWindowVal = windowVal; bestDiag = d - window / 2; }
  }
  return (int32_t) bestDiag - (int32_t) lenJ;
}

// The clipping geometry both loops of msaWfa derive from the diagonal (src/assemble.h:563-573, :618-639)
struct DiagOverlap {
  uint32_t seqlen = 0, preI = 0, postI = 0, preJ = 0, postJ = 0, startI = 0, startJ = 0;
};
inline DiagOverlap diagOverlap(int32_t bestDiag, uint32_t lenI, uint32_t lenJ) {
  DiagOverlap o;
  if (bestDiag >= 0) {
    o.seqlen = std::min(lenI - (uint32_t) bestDiag, lenJ);
    o.preI = (uint32_t) bestDiag; o.postI = lenI - ((uint32_t) bestDiag + o.seqlen);
    o.preJ = 0; o.postJ = lenJ - o.seqlen;
    o.startI = (uint32_t) bestDiag; o.startJ = 0;
  } else {
    o.seqlen = std::min(lenJ + bestDiag, lenI);
    o.preI = 0; o.postI = lenI - o.seqlen;
    o.preJ = (uint32_t) (-1 * bestDiag); o.postJ = lenJ - ((uint32_t) (-1 * bestDiag) + o.seqlen);
    o.startI = 0; o.startJ = (uint32_t) (-1 * bestDiag);
  }
  return o;
}

// src/assemble.h:90-134 — walk the overlap alignment, switch source sequence at the middle of the path
inline void buildSuperstring(std::string const& seqI, std::string const& seqJ, std::string& outStr, std::string const& ops, uint32_t preI, uint32_t postI,
                             uint32_t preJ, uint32_t postJ) {
  int32_t iIdx = 0, jIdx = 0;
  bool firstSeq = false;
  if (preI > preJ) {
    firstSeq = true;
    for (uint32_t j = 0; j < preI; ++j) outStr += seqI[iIdx++];
    jIdx += (int32_t) preJ;
  } else {
    iIdx += (int32_t) preI;
    for (uint32_t j = 0; j < preJ; ++j) outStr += seqJ[jIdx++];
  }
  const int32_t alen = (int32_t) ops.size();
  const int32_t bp = alen / 2;
  for (int32_t j = 0; j < alen; ++j) {
    if (bp == j) firstSeq = !firstSeq;
    if (ops[j] == 2) {  // EDLIB_EDOP_DELETE
      if (!firstSeq) outStr += seqJ[jIdx];
      ++jIdx;
    } else if (ops[j] == 1) {  // EDLIB_EDOP_INSERT
      if (firstSeq) outStr += seqI[iIdx];
      ++iIdx;
    } else {
      if (firstSeq) outStr += seqI[iIdx];
      else outStr += seqJ[jIdx];
      ++iIdx; ++jIdx;
    }
  }
  if (postI > postJ) { for (uint32_t j = 0; j < postI; ++j) outStr += seqI[iIdx++]; }
  else { for (uint32_t j = 0; j < postJ; ++j) outStr += seqJ[jIdx++]; }
}

// src/assemble.h:262-336 — like consensusEdlib but a row only votes between its first and last aligned column
inline void consensusWfa(TAlign const& align, std::string& cons) {
  const std::size_t R = align.size(), L = R ? align[0].size() : 0;
  std::vector<uint32_t> readStart(R, (uint32_t) L), readEnd(R, 0);
  for (std::size_t i = 0; i < R; ++i)
    for (std::size_t j = 0; j < L; ++j)
      if (align[i][j] != '-') {
        if (j < readStart[i]) readStart[i] = (uint32_t) j;
        if (j > readEnd[i]) readEnd[i] = (uint32_t) j;
      }
  cons.assign(L, '-');
  for (std::size_t j = 0; j < L; ++j) {
    int32_t count[5] = {0, 0, 0, 0, 0};
    for (std::size_t i = 0; i < R; ++i) {
      if (j >= readStart[i] && j <= readEnd[i]) {
        switch (align[i][j]) {
          case 'A': case 'a': ++count[0]; break;
          case 'C': case 'c': ++count[1]; break;
          case 'G': case 'g': ++count[2]; break;
          case 'T': case 't': ++count[3]; break;
          default: ++count[4]; break;
        }
      }
    }
    uint32_t maxIdx = 0, sndIdx = 1;
    if (count[maxIdx] < count[sndIdx]) { maxIdx = 1; sndIdx = 0; }
    for (uint32_t i = 2; i < 5; ++i) {
      if (count[i] > count[maxIdx]) { sndIdx = maxIdx; maxIdx = i; }
      else if (count[i] > count[sndIdx]) sndIdx = i;
    }
    if (2 * count[sndIdx] < count[maxIdx]) cons[j] = (maxIdx < 4) ? "ACGT"[maxIdx] : '-';
    else {
      static const char amb[5][5] = {{0, 'M', 'R', 'W', 'B'}, {0, 0, 'S', 'Y', 'D'}, {0, 0, 0, 'K', 'E'}, {0, 0, 0, 0, 'F'}, {0, 0, 0, 0, 0}};
      const uint32_t k1 = std::min(maxIdx, sndIdx), k2 = std::max(maxIdx, sndIdx);
      cons[j] = amb[k1][k2] ? amb[k1][k2] : '-';
    }
  }
}

// src/assemble.h:24-88 for EDLIB_MODE_HW: add `query` as a new last row; the part of the alignment left and right of the
// infix the query was placed in is carried over with gaps in the new row
inline void convertAlignmentHW(std::string const& query, TAlign& align, EdPath const& cigar) {
  const std::size_t R = align.size(), W = R ? align[0].size() : 0;
  const int32_t alen = (int32_t) cigar.ops.size();
  int32_t tIdx = cigar.end, qIdx = -1;
  uint32_t missingEnd = 0, missingStart = 0;
  if (tIdx < (int32_t) W) missingEnd = (uint32_t) ((int32_t) W - tIdx - 1);
  for (int32_t i = 0; i < alen; ++i) if (cigar.ops[i] != 1) --tIdx;
  if (tIdx >= 0) missingStart = (uint32_t) (tIdx + 1);
  const std::size_t L = (std::size_t) missingStart + alen + missingEnd;
  TAlign out(R + 1, std::string(L, '-'));
  for (uint32_t j = 0; j < missingStart; ++j)
    for (std::size_t r = 0; r < R; ++r) out[r][j] = align[r][j];
  for (int32_t j = 0; j < alen; ++j) {
    if (cigar.ops[j] != 1) { ++tIdx; for (std::size_t r = 0; r < R; ++r) out[r][j + missingStart] = align[r][tIdx]; }
    if (cigar.ops[j] != 2) out[R][j + missingStart] = query[++qIdx];
  }
  for (std::size_t j = (std::size_t) alen + missingStart; j < L; ++j) {
    ++tIdx;
    for (std::size_t r = 0; r < R; ++r) out[r][j] = align[r][tIdx];
  }
  align.swap(out);
}

// Batched msaWfa: clusters[i] = reads of insertion i (caller's order); prefix[i] / suffix[i] = the reference flanks
// (both empty: the 5 % end trim is applied instead of _trimConsensus). consensus[i], rows[i] = msaWfa's cs and return value.
inline int msaWfaBatch(dgpu_ctx* ctx, Config const& c, std::vector<std::vector<std::string> > const& clusters, std::vector<std::string> const& prefix,
                       std::vector<std::string> const& suffix, std::vector<std::string>& consensus, std::vector<int>& rows) {
  static const uint8_t iupac[40] = {'M', 'A', 'M', 'C', 'R', 'A', 'R', 'G', 'W', 'A', 'W', 'T', 'B', 'A', 'B', '-', 'S', 'C', 'S', 'G',
                                    'Y', 'C', 'Y', 'T', 'D', 'C', 'D', '-', 'K', 'G', 'K', 'T', 'E', 'G', 'E', '-', 'F', 'T', 'F', '-'};
  const std::size_t N = clusters.size();
  consensus.assign(N, std::string());
  rows.assign(N, 0);
  if (!N) return DGPU_OK;
  if (prefix.size() != N || suffix.size() != N) return DGPU_ERR_ARG;
  int rc;
  // ---- pairwise scores on the best k-mer diagonal (src/assemble.h:551-580) ---------------------------------------------
  std::vector<std::vector<int32_t> > edit(N);
  {
    std::string arena;
    std::vector<std::vector<uint32_t> > off(N);
    for (std::size_t i = 0; i < N; ++i)
      for (std::string const& r : clusters[i]) { off[i].push_back((uint32_t) arena.size()); arena += r; }
    std::vector<uint32_t> qo, ql, to, tl;
    std::vector<std::vector<uint32_t> > tabs;
    std::vector<std::vector<KmerHit> > uniq;
    for (std::size_t i = 0; i < N; ++i) {
      const std::size_t S = clusters[i].size();
      tabs.assign(S, std::vector<uint32_t>());
      uniq.assign(S, std::vector<KmerHit>());
      for (std::size_t a = 0; a < S; ++a) fillKmerTable(clusters[i][a], tabs[a], uniq[a]);
      for (std::size_t a = 0; a < S; ++a)
        for (std::size_t b = a + 1; b < S; ++b) {
          const uint32_t lenI = (uint32_t) clusters[i][a].size(), lenJ = (uint32_t) clusters[i][b].size();
          const DiagOverlap o = diagOverlap(bestDiagonal(tabs[a], uniq[b], lenI, lenJ), lenI, lenJ);
          qo.push_back(off[i][a] + o.startI); ql.push_back(o.seqlen);
          to.push_back(off[i][b] + o.startJ); tl.push_back(o.seqlen);
        }
    }
    std::vector<int32_t> dist(qo.size());
    if (!qo.empty()) {
      rc = dgpu_edit_distance(ctx, (const uint8_t*) arena.data(), arena.size(), qo.data(), ql.data(), to.data(), tl.data(), nullptr, DGPU_MODE_NW, qo.size(),
                              dist.data(), nullptr);
      if (rc) return rc;
    }
    std::size_t k = 0;
    for (std::size_t i = 0; i < N; ++i) {
      const std::size_t S = clusters[i].size();
      edit[i].assign(S * S, 0);
      for (std::size_t a = 0; a < S; ++a)
        for (std::size_t b = a + 1; b < S; ++b, ++k) {
          if (!ql[k]) return DGPU_ERR_ARG;  // the reference divides by the overlap length here (src/assemble.h:575)
          const int32_t score = (int32_t) ((dist[k] * 1000) / (int32_t) ql[k]);
          edit[i][a * S + b] = edit[i][b * S + a] = score;
        }
    }
  }
  // ---- centroid, ordering, selection (src/assemble.h:582-607) ----------------------------------------------------------
  std::vector<std::vector<uint32_t> > sel(N);
  std::vector<std::string> superStr(N);
  std::size_t maxSel = 0;
  for (std::size_t i = 0; i < N; ++i) {
    std::vector<std::string> const& sps = clusters[i];
    const std::size_t S = sps.size();
    if (!S) continue;
    uint32_t bestIdx = 0;
    int32_t bestVal = (int32_t) sps[0].size();
    for (uint32_t a = 0; a < S; ++a) {
      std::vector<int32_t> d(edit[i].begin() + a * S, edit[i].begin() + (a + 1) * S);
      std::sort(d.begin(), d.end());
      if (d[S / 2] < bestVal) { bestVal = d[S / 2]; bestIdx = a; }
    }
    std::vector<std::pair<int32_t, int32_t> > qs;
    qs.push_back(std::make_pair(0, (int32_t) bestIdx));
    for (uint32_t j = 0; j < S; ++j) if (j != bestIdx) qs.push_back(std::make_pair(edit[i][bestIdx * S + j], (int32_t) j));
    std::sort(qs.begin(), qs.end());
    uint32_t lastIdx = (uint32_t) (0.8 * qs.size());
    if (lastIdx < 3) lastIdx = 3;
    for (uint32_t k = 0; k < qs.size() && k < lastIdx; ++k) sel[i].push_back((uint32_t) qs[k].second);
    superStr[i] = sps[sel[i][0]];
    maxSel = std::max(maxSel, sel[i].size());
  }
  // ---- superstring (src/assemble.h:609-668): one NW path round per selected read ----------------------------------------------
  {
    std::vector<uint32_t> kmerI, kmerJ;
    std::vector<KmerHit> uniqI, uniqJ;
    for (std::size_t round = 1; round < maxSel; ++round) {
      std::vector<uint32_t> who;
      std::vector<DiagOverlap> geo;
      std::vector<std::pair<std::string, std::string> > qt;
      for (std::size_t i = 0; i < N; ++i) {
        if (sel[i].size() <= round) continue;
        std::string const& rd = clusters[i][sel[i][round]];
        const uint32_t lenI = (uint32_t) superStr[i].size(), lenJ = (uint32_t) rd.size();
        fillKmerTable(superStr[i], kmerI, uniqI);
        fillKmerTable(rd, kmerJ, uniqJ);
        const DiagOverlap o = diagOverlap(bestDiagonal(kmerI, uniqJ, lenI, lenJ), lenI, lenJ);
        if (o.preI > o.preJ && o.postI > o.postJ) continue;                           // nested alignment
        if (o.preJ > o.preI && o.postJ > o.postI) { superStr[i] = rd; continue; }     // nested, new sequence longer
        who.push_back((uint32_t) i);
        geo.push_back(o);
        qt.push_back(std::make_pair(superStr[i].substr(o.startI, o.seqlen), rd.substr(o.startJ, o.seqlen)));
      }
      std::vector<EdPath> paths;
      if ((rc = editPathBatch(ctx, qt, DGPU_MODE_NW, paths))) return rc;
      for (std::size_t k = 0; k < who.size(); ++k) {
        const std::size_t i = who[k];
        if (paths[k].failed) { sel[i].clear(); superStr[i].clear(); continue; }   // beyond a device limit: this cluster yields no consensus (counted)
        std::string outStr;
        buildSuperstring(superStr[i], clusters[i][sel[i][round]], outStr, paths[k].ops, geo[k].preI, geo[k].postI, geo[k].preJ, geo[k].postJ);
        superStr[i].swap(outStr);
      }
    }
  }
  // ---- progressive alignment against the superstring (src/assemble.h:671-694) ---------------------------------------------------
  std::vector<TAlign> aligns(N);
  for (std::size_t i = 0; i < N; ++i) if (!sel[i].empty()) aligns[i].assign(1, superStr[i]);
  for (std::size_t round = 0; round < maxSel; ++round) {
    std::vector<uint32_t> who;
    std::vector<std::pair<std::string, std::string> > qt;
    for (std::size_t i = 0; i < N; ++i) {
      if ((sel[i].size() <= round) || aligns[i].empty()) continue;
      std::string alignStr;
      consensusWfa(aligns[i], alignStr);
      who.push_back((uint32_t) i);
      qt.push_back(std::make_pair(clusters[i][sel[i][round]], alignStr));
    }
    std::vector<EdPath> paths;
    if ((rc = editPathBatch(ctx, qt, DGPU_MODE_HW, paths, iupac, 20))) return rc;
    for (std::size_t k = 0; k < who.size(); ++k) {
      if (paths[k].failed) { aligns[who[k]].clear(); sel[who[k]].clear(); continue; }
      if (!aligns[who[k]].empty()) convertAlignmentHW(qt[k].first, aligns[who[k]], paths[k]);
    }
  }
  // ---- consensus (src/assemble.h:704-706) -------------------------------------------------------------------------------------
  std::vector<std::string> cs(N);
  for (std::size_t i = 0; i < N; ++i) {
    if (aligns[i].empty()) continue;
    std::string gapped;
    consensusRows(c, aligns[i], gapped, cs[i]);
    rows[i] = (int) sel[i].size();
  }
  // ---- orientation + trimming (src/assemble.h:709-720, _trimConsensus :339-365) -------------------------------------------------
  std::vector<uint32_t> trimmed;
  for (std::size_t i = 0; i < N; ++i) {
    if (aligns[i].empty()) continue;
    if (!prefix[i].empty() && !suffix[i].empty()) trimmed.push_back((uint32_t) i);
    else {
      int32_t trim = (int32_t) (0.05 * cs[i].size());
      if (trim > 50) trim = 50;
      const int32_t len = (int32_t) cs[i].size() - 2 * trim;
      if (len > 100) cs[i] = cs[i].substr(trim, len);
    }
  }
  if (!trimmed.empty()) {
    // forward vs reverse-complemented prefix inside the consensus
    std::string arena;
    std::vector<uint32_t> qo, ql, to, tl;
    for (uint32_t i : trimmed) {
      std::string prefixRev = prefix[i];
      reverseComplement(prefixRev);
      const uint32_t t0 = (uint32_t) arena.size();
      arena += cs[i];
      qo.push_back((uint32_t) arena.size()); ql.push_back((uint32_t) prefix[i].size()); arena += prefix[i];
      to.push_back(t0); tl.push_back((uint32_t) cs[i].size());
      qo.push_back((uint32_t) arena.size()); ql.push_back((uint32_t) prefixRev.size()); arena += prefixRev;
      to.push_back(t0); tl.push_back((uint32_t) cs[i].size());
    }
    std::vector<int32_t> dist(qo.size());
    rc = dgpu_edit_distance(ctx, (const uint8_t*) arena.data(), arena.size(), qo.data(), ql.data(), to.data(), tl.data(), nullptr, DGPU_MODE_HW, qo.size(),
                            dist.data(), nullptr);
    if (rc) return rc;
    std::vector<std::pair<std::string, std::string> > qt;
    for (std::size_t k = 0; k < trimmed.size(); ++k) {
      const uint32_t i = trimmed[k];
      if (dist[2 * k] > dist[2 * k + 1]) reverseComplement(cs[i]);
      qt.push_back(std::make_pair(prefix[i], cs[i]));
      qt.push_back(std::make_pair(suffix[i], cs[i]));
    }
    std::vector<EdPath> paths;
    if ((rc = editPathBatch(ctx, qt, DGPU_MODE_HW, paths))) return rc;
    for (std::size_t k = 0; k < trimmed.size(); ++k) {
      const uint32_t i = trimmed[k];
      if (paths[2 * k].failed || paths[2 * k + 1].failed) { cs[i].clear(); continue; }
      const uint32_t csStart = infixStart(paths[2 * k]);
      const uint32_t csEnd = infixEnd(paths[2 * k + 1]);
      if (csStart < csEnd && csEnd < cs[i].size()) cs[i] = cs[i].substr(csStart, csEnd - csStart);
    }
  }
  for (std::size_t i = 0; i < N; ++i) consensus[i].swap(cs[i]);
  return DGPU_OK;
}

}  // namespace dellyb200
