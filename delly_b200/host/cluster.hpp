// cluster.hpp — split-read and paired-end clustering that seeds the realignment candidates
// (SURVEY.md §8 a12, a13). Same entry points and results as the reference:
//   cluster(c, br, sv, svt)                     src/cluster.h:324-442  (+ _searchCliques :207-321)
//   cluster(c, bamRecord, svs, varisize, svt)   src/cluster.h:528-629  (+ _searchCliques :445-524,
//                                               _initClique/_updateClique :68-204)
// Structure here: both variants share one ComponentGraph (vertex labels + per-component edge lists with
// the reference's merge order and graphPruning cap), a windowed sweep that feeds it, and a greedy clique
// search over the weight-sorted edges. Clique growth is inherently sequential per component, so this stays
// on the host; what moves to the GPU is everything the clusters then trigger.
#pragma once
#include <cmath>
#include <iostream>
#include <map>
#include <set>
#include <vector>

#include "../../include/dgpu.h"
#include "types.hpp"

namespace dellyb200 {

namespace detail {

// Connected components over a sliding window of sorted records, labelled like the reference: a new edge
// either creates a component, extends one, or merges two (the smaller label survives, the other list is
// appended to it); at most `pruning` edges are kept per component.
class ComponentGraph {
 public:
  typedef std::map<uint32_t, std::vector<EdgeRecord> > TCompEdgeList;
  ComponentGraph(std::size_t nvertices, uint32_t pruning) : comp_(nvertices, 0), pruning_(pruning) {}
  bool empty() const { return edges_.empty(); }
  TCompEdgeList& components() { return edges_; }
  void clearEdges() { edges_.clear(); }
  void resetLabels(std::size_t lo, std::size_t hi) { for (std::size_t k = lo; k < hi; ++k) comp_[k] = 0; }
  void resetCounter() { numComp_ = 0; }
  // relabelLo..relabelHi = the window that can still hold vertices of a live component
  void connect(uint32_t i, uint32_t j, uint32_t weight, std::size_t relabelLo, std::size_t relabelHi) {
    uint32_t label;
    if (!comp_[i] && !comp_[j]) {
      label = ++numComp_;
      comp_[i] = comp_[j] = label;
      edges_.insert(std::make_pair(label, std::vector<EdgeRecord>()));
    } else if (!comp_[i]) { label = comp_[i] = comp_[j]; }
    else if (!comp_[j]) { label = comp_[j] = comp_[i]; }
    else if (comp_[i] == comp_[j]) { label = comp_[i]; }
    else {
      label = std::min(comp_[i], comp_[j]);
      const uint32_t other = std::max(comp_[i], comp_[j]);
      for (std::size_t k = relabelLo; k <= relabelHi; ++k)
        if (comp_[k] == other) comp_[k] = label;
      TCompEdgeList::iterator keep = edges_.find(label), gone = edges_.find(other);
      keep->second.insert(keep->second.end(), gone->second.begin(), gone->second.end());
      edges_.erase(gone);
    }
    std::vector<EdgeRecord>& el = edges_.find(label)->second;
    if (el.size() < pruning_) el.push_back(EdgeRecord{i, j, weight});
  }
 private:
  std::vector<uint32_t> comp_;
  uint32_t pruning_;
  uint32_t numComp_ = 0;
  TCompEdgeList edges_;
};

// wiggle / window size of one SR record (src/cluster.h:362-370 and :239-248)
inline uint32_t srVariability(Config const& c, int32_t svt, double span, int32_t inslen) {
  uint32_t v = c.maxReadSep;
  if (_translocation(svt)) return 2 * c.maxReadSep;
  uint32_t svvar = (uint32_t) std::abs(0.1 * span);
  if (svt == 4) svvar = (uint32_t) std::abs(0.1 * inslen);
  if (v < svvar) v = svvar;
  if (v > 1000) v = 1000;
  return v;
}

// Greedy clique over one component's edges (already sorted by weight, source, target): repeatedly take the
// first edge with exactly one endpoint inside that `accept` admits; rejected vertices are never retried.
template <typename TAccept>
inline void growClique(std::vector<EdgeRecord> const& edges, std::set<std::size_t>& clique, TAccept accept) {
  std::set<std::size_t> incompatible;
  bool grew = true;
  while (grew) {
    grew = false;
    for (std::vector<EdgeRecord>::const_iterator e = edges.begin(); !grew && e != edges.end(); ++e) {
      const bool sIn = clique.count(e->source) != 0, tIn = clique.count(e->target) != 0;
      if (sIn == tIn) continue;
      const std::size_t v = sIn ? e->target : e->source;
      if (incompatible.count(v)) continue;
      int verdict = accept(v);  // 1 accept, 0 reject for good, -1 skip without marking
      if (verdict > 0) { clique.insert(v); grew = true; }
      else if (verdict == 0) incompatible.insert(v);
    }
  }
}

inline void searchCliquesSR(Config const& c, ComponentGraph::TCompEdgeList& compEdge, std::vector<SRBamRecord>& br,
                            std::vector<StructuralVariantRecord>& sv, int32_t svt) {
  for (ComponentGraph::TCompEdgeList::iterator comp = compEdge.begin(); comp != compEdge.end(); ++comp) {
    std::vector<EdgeRecord>& edges = comp->second;
    std::sort(edges.begin(), edges.end());
    const SRBamRecord& seed = br[edges.front().source];
    std::set<std::size_t> clique, seeds;
    clique.insert(edges.front().source);
    seeds.insert(seed.id);
    const int32_t chr = seed.chr, chr2 = seed.chr2;
    int32_t ciposlow = seed.pos, ciposhigh = seed.pos, ciendlow = seed.pos2, ciendhigh = seed.pos2;
    uint64_t pos = seed.pos, pos2 = seed.pos2;
    int32_t mapq = seed.qual, inslen = seed.inslen;
    const uint32_t wiggle = srVariability(c, svt, (double) (pos2 - pos), inslen);
    growClique(edges, clique, [&](std::size_t v) -> int {
      if (seeds.count(br[v].id)) return -1;  // same read already in: skipped, not marked incompatible
      const int32_t lo1 = std::min(br[v].pos, ciposlow), hi1 = std::max(br[v].pos, ciposhigh);
      const int32_t lo2 = std::min(br[v].pos2, ciendlow), hi2 = std::max(br[v].pos2, ciendhigh);
      if (((hi1 - lo1) < (int32_t) wiggle) && ((hi2 - lo2) < (int32_t) wiggle) && (!_translocation(svt) || br[v].chr2 == chr2)) {
        seeds.insert(br[v].id);
        ciposlow = lo1; ciposhigh = hi1; ciendlow = lo2; ciendhigh = hi2;
        pos += br[v].pos; pos2 += br[v].pos2; mapq += br[v].qual; inslen += br[v].inslen;
        return 1;
      }
      return 0;
    });
    // both mates of a pair carry consecutive ids: count the fragment once (src/cluster.h:289-298)
    std::size_t prevSeed = 0;
    uint32_t cliqSize = 0;
    std::set<std::size_t> selected;
    for (std::size_t s : seeds) {
      if (prevSeed + 1 != s) { selected.insert(s); ++cliqSize; }
      prevSeed = s;
    }
    if (cliqSize < c.minCliqueSize) continue;
    const int32_t svStart = (int32_t) (pos / (uint64_t) clique.size());
    const int32_t svEnd = (int32_t) (pos2 / (uint64_t) clique.size());
    const int32_t svInsLen = (int32_t) (inslen / (int32_t) clique.size());
    if (!_svSizeCheck(svStart, svEnd, svt, svInsLen)) continue;
    if ((ciposlow > svStart) || (ciposhigh < svStart) || (ciendlow > svEnd) || (ciendhigh < svEnd))
      std::cerr << "Warning: Confidence intervals out of bounds: " << ciposlow << ',' << svStart << ',' << ciposhigh << ':' << ciendlow << ','
                << svEnd << ',' << ciendhigh << std::endl;
    const int32_t svid = (int32_t) sv.size();
    sv.push_back(StructuralVariantRecord(chr, svStart, chr2, svEnd, ciposlow - svStart, ciposhigh - svStart, ciendlow - svEnd, ciendhigh - svEnd,
                                         (int32_t) cliqSize, mapq / (int32_t) clique.size(), mapq, svInsLen, svt, svid));
    for (std::size_t v : clique)
      if (selected.count(br[v].id)) br[v].svid = svid;
  }
}

// src/cluster.h:68-111
inline void initClique(BamAlignRecord const& el, int32_t& svStart, int32_t& svEnd, int32_t& wiggle, int32_t svt) {
  if (_translocation(svt)) {
    const uint8_t ct = _getSpanOrientation(svt);
    svStart = (ct % 2 == 0) ? el.pos + el.alen : el.pos;
    const bool mateEnd = ((ct % 2 == 0) != (ct >= 2));  // ct 0 and 3 take the mate's end, 1 and 2 its start
    svEnd = mateEnd ? el.mpos + el.malen : el.mpos;
    wiggle = el.maxNormalISize;
    return;
  }
  switch (svt) {
    case 0: svStart = el.mpos + el.malen; svEnd = el.pos + el.alen; wiggle = el.maxNormalISize - std::max(el.alen, el.malen); break;
    case 1: svStart = el.mpos; svEnd = el.pos; wiggle = el.maxNormalISize - std::max(el.alen, el.malen); break;
    case 2: svStart = el.mpos + el.malen; svEnd = el.pos; wiggle = -el.maxNormalISize; break;
    case 3: svStart = el.mpos; svEnd = el.pos + el.alen; wiggle = el.maxNormalISize; break;
    default: break;
  }
}

// src/cluster.h:113-204
inline bool updateClique(BamAlignRecord const& el, int32_t& svStart, int32_t& svEnd, int32_t& wiggle, int32_t svt) {
  int32_t ns, ne, nw;
  bool valid = false;
  if (_translocation(svt)) {
    const int ct = _getSpanOrientation(svt);
    nw = wiggle;
    if (ct % 2 == 0) { ns = std::max(svStart, el.pos + el.alen); nw -= (ns - svStart); }
    else { ns = std::min(svStart, el.pos); nw -= (svStart - ns); }
    const bool mateEnd = ((ct % 2 == 0) != (ct >= 2));
    if (mateEnd) { ne = std::max(svEnd, el.mpos + el.malen); nw -= (ne - svEnd); }
    else { ne = std::min(svEnd, el.mpos); nw -= (svEnd - ne); }
    valid = nw > 0;
  } else if (svt == 0 || svt == 1) {
    int32_t change;
    if (svt == 0) {
      ns = std::max(svStart, el.mpos + el.malen);
      ne = std::max(svEnd, el.pos + el.alen);
      nw = std::min(el.maxNormalISize - (ns - el.mpos), el.maxNormalISize - (ne - el.pos));
      change = wiggle - std::max(ns - svStart, ne - svEnd);
    } else {
      ns = std::min(svStart, el.mpos);
      ne = std::min(svEnd, el.pos);
      nw = std::min(el.maxNormalISize - (el.mpos + el.malen - ns), el.maxNormalISize - (el.pos + el.alen - ne));
      change = wiggle - std::max(svStart - ns, svEnd - ne);
    }
    if (change < nw) nw = change;
    valid = (ns < ne) && (nw >= 0);
  } else if (svt == 2) {
    ns = std::max(svStart, el.mpos + el.malen);
    ne = std::min(svEnd, el.pos);
    nw = el.pos + el.alen - el.mpos - el.maxNormalISize - (ne - ns);
    const int32_t change = wiggle + (svEnd - svStart) - (ne - ns);
    if (change > nw) nw = change;
    valid = (ns < ne) && (nw <= 0);
  } else if (svt == 3) {
    ns = std::min(svStart, el.mpos);
    ne = std::max(svEnd, el.pos + el.alen);
    nw = el.pos - (el.mpos + el.malen) + el.maxNormalISize - (ne - ns);
    const int32_t change = wiggle - ((ne - ns) - (svEnd - svStart));
    if (change < nw) nw = change;
    valid = (ns < ne) && (nw >= 0);
  } else return false;
  if (valid) { svStart = ns; svEnd = ne; wiggle = nw; }
  return valid;
}

inline void searchCliquesPE(Config const& c, ComponentGraph::TCompEdgeList& compEdge, std::vector<BamAlignRecord> const& rec,
                            std::vector<StructuralVariantRecord>& svs, int32_t svt) {
  for (ComponentGraph::TCompEdgeList::iterator comp = compEdge.begin(); comp != compEdge.end(); ++comp) {
    std::vector<EdgeRecord>& edges = comp->second;
    std::sort(edges.begin(), edges.end());
    const BamAlignRecord& seed = rec[edges.front().source];
    int32_t svStart = -1, svEnd = -1, wiggle = 0;
    initClique(seed, svStart, svEnd, wiggle, svt);
    if ((seed.tid == seed.mtid) && (svStart >= svEnd)) continue;
    std::set<std::size_t> clique;
    clique.insert(edges.front().source);
    growClique(edges, clique, [&](std::size_t v) -> int { return updateClique(rec[v], svStart, svEnd, wiggle, svt) ? 1 : 0; });
    if (!((clique.size() >= c.minCliqueSize) && _svSizeCheck(svStart, svEnd, svt))) continue;
    StructuralVariantRecord r;
    r.chr = seed.tid; r.chr2 = seed.mtid;
    const int32_t bnd = (svt < DELLY_SVT_TRANS) ? 0 : 1;  // BND positions are 1-based (src/cluster.h:496-500)
    r.svStart = (int32_t) ((uint32_t) svStart + bnd);
    r.svEnd = (int32_t) ((uint32_t) svEnd + bnd);
    r.peSupport = (int32_t) clique.size();
    const int32_t ci = std::max(std::abs(wiggle), 50);
    r.ciposlow = -ci; r.ciposhigh = ci; r.ciendlow = -ci; r.ciendhigh = ci;
    std::vector<uint8_t> mq;
    for (std::size_t v : clique) { mq.push_back(rec[v].MapQuality); r.mapq += rec[v].MapQuality; }
    std::sort(mq.begin(), mq.end());
    r.peMapQuality = mq[mq.size() / 2];
    r.svt = svt;
    svs.push_back(r);
  }
}

}  // namespace detail

// Split-read clustering (src/cluster.h:324-442). br must be sorted (SRBamRecord::operator<).
inline void cluster(Config const& c, std::vector<SRBamRecord>& br, std::vector<StructuralVariantRecord>& sv, int32_t svt) {
  detail::ComponentGraph g(br.size(), c.graphPruning);
  for (int32_t refIdx = 0; refIdx < c.nchr; ++refIdx) {
    const uint32_t lo = (uint32_t) (std::lower_bound(br.begin(), br.end(), refIdx, [](SRBamRecord const& r, int32_t v) { return r.chr < v; }) - br.begin());
    const uint32_t hi = (uint32_t) (std::upper_bound(br.begin(), br.end(), refIdx, [](int32_t v, SRBamRecord const& r) { return v < r.chr; }) - br.begin());
    if (lo >= hi) continue;
    g.resetCounter();
    std::size_t lastConnected = lo, lastConnectedStart = lo;
    for (uint32_t i = lo; i < hi; ++i) {
      if (i > lastConnected && !g.empty()) {  // nothing later can join the live components: flush them
        detail::searchCliquesSR(c, g.components(), br, sv, svt);
        lastConnectedStart = lastConnected;
        g.clearEdges();
      }
      const uint32_t varisize = detail::srVariability(c, svt, (double) (br[i].pos2 - br[i].pos), br[i].inslen);
      for (uint32_t j = i + 1; j < hi; ++j) {
        if ((uint32_t) (br[j].pos - br[i].pos) > varisize) break;
        if ((svt == 4) && ((uint32_t) std::abs(br[j].inslen - br[i].inslen) > varisize)) continue;
        if (_translocation(svt) && (br[j].chr2 != br[i].chr2)) continue;
        if ((uint32_t) std::abs(br[j].pos2 - br[i].pos2) < varisize) {
          if (j > lastConnected) lastConnected = j;
          const uint32_t w = (uint32_t) (std::abs(br[j].pos2 - br[i].pos2) + std::abs(br[j].pos - br[i].pos));
          g.connect(i, j, w, lastConnectedStart, lastConnected);
        }
      }
    }
    if (!g.empty()) {
      detail::searchCliquesSR(c, g.components(), br, sv, svt);
      g.clearEdges();
    }
    g.resetLabels(lo, hi);
  }
}

// Paired-end clustering (src/cluster.h:528-629). bamRecord must be sorted (BamAlignRecord::operator<).
inline void cluster(Config const& c, std::vector<BamAlignRecord>& bamRecord, std::vector<StructuralVariantRecord>& svs, uint32_t varisize, int32_t svt) {
  detail::ComponentGraph g(bamRecord.size(), c.graphPruning);
  std::size_t lastConnected = 0, lastConnectedStart = 0;
  const std::size_t n = bamRecord.size();
  for (std::size_t i = 0; i < n; ++i) {
    if (i > lastConnected && !g.empty()) {
      detail::searchCliquesPE(c, g.components(), bamRecord, svs, svt);
      lastConnectedStart = lastConnected;
      g.clearEdges();
    }
    const BamAlignRecord& a = bamRecord[i];
    const int32_t aMin = _minCoord(a.pos, a.mpos, svt), aMax = _maxCoord(a.pos, a.mpos, svt);
    for (std::size_t j = i + 1; j < n; ++j) {
      const BamAlignRecord& b = bamRecord[j];
      const int32_t bMin = _minCoord(b.pos, b.mpos, svt), bMax = _maxCoord(b.pos, b.mpos, svt);
      if (!((uint32_t) std::abs(bMin + b.alen - aMin) <= varisize)) break;
      if (a.mtid != b.mtid) continue;
      if (_pairsDisagree(aMin, aMax, a.alen, a.maxNormalISize, bMin, bMax, b.alen, b.maxNormalISize, svt)) continue;
      if (j > lastConnected) lastConnected = j;
      const uint32_t w = (uint32_t) (std::log2((double) std::abs(std::abs((bMin - aMin) - (bMax - aMax)) - std::abs(a.Median - b.Median)) + 1));
      g.connect((uint32_t) i, (uint32_t) j, w, lastConnectedStart, lastConnected);
    }
  }
  if (!g.empty()) {
    detail::searchCliquesPE(c, g.components(), bamRecord, svs, svt);
    g.clearEdges();
  }
}


// ---- the same two functions with the pair scan on the device -------------------------------------------------------------
// dgpu_cluster_edges_* returns, per record, the (target, weight) pairs the scans above would connect, in the same order;
// the sequential part (component labels, pruning, clique search, flush points) is the code above, unchanged.

namespace detail {
struct EdgeCsr {
  std::vector<uint32_t> off, j, w;
};
template <typename TCall>
inline int fetchEdges(std::size_t n, EdgeCsr& e, TCall call) {
  e.off.assign(n + 1, 0);
  uint64_t cap = std::max<std::size_t>(4 * n, 1024), total = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    e.j.resize(cap); e.w.resize(cap);
    const int rc = call(e.off.data(), e.j.data(), e.w.data(), cap, &total);
    if (rc == DGPU_OK) { e.j.resize(total); e.w.resize(total); return DGPU_OK; }
    if (rc != DGPU_ERR_CAPACITY || total <= cap) return rc;
    cap = total;
  }
  return DGPU_ERR_CAPACITY;
}
}  // namespace detail

inline int clusterGpu(dgpu_ctx* ctx, Config const& c, std::vector<SRBamRecord>& br, std::vector<StructuralVariantRecord>& sv, int32_t svt) {
  const std::size_t n = br.size();
  if (!n) return DGPU_OK;
  std::vector<int32_t> col(5 * n);
  for (std::size_t i = 0; i < n; ++i) { col[i] = br[i].chr; col[n + i] = br[i].pos; col[2 * n + i] = br[i].chr2; col[3 * n + i] = br[i].pos2; col[4 * n + i] = br[i].inslen; }
  detail::EdgeCsr e;
  int rc = detail::fetchEdges(n, e, [&](uint32_t* off, uint32_t* ej, uint32_t* ew, uint64_t cap, uint64_t* total) {
    return dgpu_cluster_edges_sr(ctx, col.data(), col.data() + n, col.data() + 2 * n, col.data() + 3 * n, col.data() + 4 * n, n, svt, c.maxReadSep, off, ej, ew, cap, total);
  });
  if (rc) return rc;
  detail::ComponentGraph g(n, c.graphPruning);
  for (int32_t refIdx = 0; refIdx < c.nchr; ++refIdx) {
    const uint32_t lo = (uint32_t) (std::lower_bound(br.begin(), br.end(), refIdx, [](SRBamRecord const& r, int32_t v) { return r.chr < v; }) - br.begin());
    const uint32_t hi = (uint32_t) (std::upper_bound(br.begin(), br.end(), refIdx, [](int32_t v, SRBamRecord const& r) { return v < r.chr; }) - br.begin());
    if (lo >= hi) continue;
    g.resetCounter();
    std::size_t lastConnected = lo, lastConnectedStart = lo;
    for (uint32_t i = lo; i < hi; ++i) {
      if (i > lastConnected && !g.empty()) {
        detail::searchCliquesSR(c, g.components(), br, sv, svt);
        lastConnectedStart = lastConnected;
        g.clearEdges();
      }
      for (uint32_t k = e.off[i]; k < e.off[i + 1]; ++k) {
        const uint32_t j = e.j[k];
        if (j > lastConnected) lastConnected = j;
        g.connect(i, j, e.w[k], lastConnectedStart, lastConnected);
      }
    }
    if (!g.empty()) {
      detail::searchCliquesSR(c, g.components(), br, sv, svt);
      g.clearEdges();
    }
    g.resetLabels(lo, hi);
  }
  return DGPU_OK;
}

inline int clusterGpu(dgpu_ctx* ctx, Config const& c, std::vector<BamAlignRecord>& bamRecord, std::vector<StructuralVariantRecord>& svs, uint32_t varisize,
                      int32_t svt) {
  const std::size_t n = bamRecord.size();
  if (!n) return DGPU_OK;
  std::vector<int32_t> col(6 * n);
  for (std::size_t i = 0; i < n; ++i) {
    BamAlignRecord const& r = bamRecord[i];
    col[i] = r.pos; col[n + i] = r.mpos; col[2 * n + i] = r.mtid; col[3 * n + i] = r.alen; col[4 * n + i] = r.Median; col[5 * n + i] = r.maxNormalISize;
  }
  detail::EdgeCsr e;
  int rc = detail::fetchEdges(n, e, [&](uint32_t* off, uint32_t* ej, uint32_t* ew, uint64_t cap, uint64_t* total) {
    return dgpu_cluster_edges_pe(ctx, col.data(), col.data() + n, col.data() + 2 * n, col.data() + 3 * n, col.data() + 4 * n, col.data() + 5 * n, n, svt, varisize, off, ej,
                                 ew, cap, total);
  });
  if (rc) return rc;
  detail::ComponentGraph g(n, c.graphPruning);
  std::size_t lastConnected = 0, lastConnectedStart = 0;
  for (std::size_t i = 0; i < n; ++i) {
    if (i > lastConnected && !g.empty()) {
      detail::searchCliquesPE(c, g.components(), bamRecord, svs, svt);
      lastConnectedStart = lastConnected;
      g.clearEdges();
    }
    for (uint32_t k = e.off[i]; k < e.off[i + 1]; ++k) {
      const uint32_t j = e.j[k];
      if (j > lastConnected) lastConnected = j;
      g.connect((uint32_t) i, j, e.w[k], lastConnectedStart, lastConnected);
    }
  }
  if (!g.empty()) {
    detail::searchCliquesPE(c, g.components(), bamRecord, svs, svt);
    g.clearEdges();
  }
  return DGPU_OK;
}

}  // namespace dellyb200
