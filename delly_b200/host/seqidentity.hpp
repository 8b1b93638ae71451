// seqidentity.hpp — sequence identity between the inserted / junction sequences of two SV calls, the edit-distance call site
// of `delly merge`. Host mirror of
//   src/merge.h:187-208  _minRotation      (lexicographically smallest rotation of a string)
//   src/merge.h:210-223  _seqIdentity      (1 - NW edit distance / longer length, bounded by k = (1 - minId) * longer length)
//   src/merge.h:226-243  _bestSeqIdentity  (as is; else b rotated by the position offset; else both in canonical rotation)
// The reference aligns one pair at a time inside _svMatch; here all pairs of a comparison round go to the device in ONE
// dgpu_edit_distance call (NW, per-job bound), and the up to three attempts of _bestSeqIdentity become three rounds over
// the pairs that are still undecided.
#pragma once
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/dgpu.h"

namespace dellyb200 {

// The smallest rotation is unique as a string, so any exact method gives the reference's result; this one compares
// candidate start positions pairwise (Duval-style), linear time.
inline std::string _minRotation(std::string const& s) {
  const std::size_t n = s.size();
  if (n < 2) return s;
  std::size_t i = 0, j = 1, k = 0;
  while (i < n && j < n && k < n) {
    const char a = s[(i + k) % n], b = s[(j + k) % n];
    if (a == b) { ++k; continue; }
    if (a > b) i += k + 1; else j += k + 1;
    if (i == j) ++j;
    k = 0;
  }
  const std::size_t start = std::min(i, j);
  return s.substr(start) + s.substr(0, start);
}

struct IdentityPair {
  std::string a, b;
  int32_t posOff = 0;   // only _bestSeqIdentity reads it
};

// _seqIdentity for every pair: -1 for an empty sequence; 0 when the bound was exceeded
inline int seqIdentityBatch(dgpu_ctx* ctx, std::vector<IdentityPair> const& pairs, double minId, std::vector<double>& out) {
  out.assign(pairs.size(), -1.0);
  std::string arena;
  std::vector<uint32_t> qo, ql, to, tl, idx;
  std::vector<int32_t> kk;
  for (std::size_t i = 0; i < pairs.size(); ++i) {
    IdentityPair const& p = pairs[i];
    if (p.a.empty() || p.b.empty()) continue;
    const int32_t maxlen = (int32_t) std::max(p.a.size(), p.b.size());
    int32_t k = -1;
    if (minId > 0.0 && minId < 1.0) k = (int32_t) ((1.0 - minId) * maxlen);
    qo.push_back((uint32_t) arena.size()); ql.push_back((uint32_t) p.a.size()); arena += p.a;
    to.push_back((uint32_t) arena.size()); tl.push_back((uint32_t) p.b.size()); arena += p.b;
    kk.push_back(k); idx.push_back((uint32_t) i);
  }
  if (idx.empty()) return DGPU_OK;
  if (!ctx) return DGPU_ERR_NODEVICE;
  std::vector<int32_t> dist(idx.size());
  const int rc = dgpu_edit_distance(ctx, (const uint8_t*) arena.data(), arena.size(), qo.data(), ql.data(), to.data(), tl.data(), kk.data(), DGPU_MODE_NW, idx.size(),
                                    dist.data(), nullptr);
  if (rc) return rc;
  for (std::size_t j = 0; j < idx.size(); ++j) {
    IdentityPair const& p = pairs[idx[j]];
    const int32_t maxlen = (int32_t) std::max(p.a.size(), p.b.size());
    out[idx[j]] = (dist[j] >= 0) ? 1.0 - (double) dist[j] / (double) maxlen : ((kk[j] >= 0) ? 0.0 : -1.0);
  }
  return DGPU_OK;
}

// _bestSeqIdentity for every pair
inline int bestSeqIdentityBatch(dgpu_ctx* ctx, std::vector<IdentityPair> const& pairs, double minId, int32_t seqCutoff, std::vector<double>& best) {
  int rc = seqIdentityBatch(ctx, pairs, minId, best);
  if (rc) return rc;
  auto open = [&](std::size_t i) {  // still undecided after the attempts so far, and short enough for the rotations (:232)
    IdentityPair const& p = pairs[i];
    if (p.a.empty() || p.b.empty()) return false;
    if (minId > 0.0 && best[i] >= minId) return false;
    return ((int32_t) p.a.size() < seqCutoff) && ((int32_t) p.b.size() < seqCutoff);
  };
  for (int round = 0; round < 2; ++round) {
    std::vector<IdentityPair> sub;
    std::vector<std::size_t> who;
    for (std::size_t i = 0; i < pairs.size(); ++i) {
      if (!open(i)) continue;
      IdentityPair const& p = pairs[i];
      IdentityPair q;
      if (round == 0) {   // b rotated right by the offset between the two calls
        const int32_t f = p.posOff % (int32_t) p.b.size();
        if (f <= 0) continue;
        q.a = p.a; q.b = p.b.substr(p.b.size() - (std::size_t) f) + p.b.substr(0, p.b.size() - (std::size_t) f);
      } else { q.a = _minRotation(p.a); q.b = _minRotation(p.b); }
      sub.push_back(q); who.push_back(i);
    }
    std::vector<double> id;
    if ((rc = seqIdentityBatch(ctx, sub, minId, id))) return rc;
    for (std::size_t j = 0; j < who.size(); ++j) best[who[j]] = std::max(best[who[j]], id[j]);
  }
  return DGPU_OK;
}

}  // namespace dellyb200
