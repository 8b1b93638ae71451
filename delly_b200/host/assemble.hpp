// assemble.hpp — the split-read assembly stage of `delly sr` (assembleSplitReads, src/shortpe.h:48-282), batched:
// the reference collects the split reads of every SV from the BAM, then runs msa() + alignConsensus() per SV on its thread pool
// (:172-201, :246-271). Here the collection is the same host code over an in-memory record list, and ALL SVs go through
// ONE msaBatch (one dgpu_msa) and ONE alignConsensusBatch (one dgpu_long_needle + the splitAlign rounds for insertions).
#pragma once
#include <algorithm>
#include <map>
#include <string>
#include <unordered_set>
#include <vector>

#include "genotype.hpp"
#include "msa.hpp"
#include "shard.hpp"
#include "split.hpp"
#include "splitalign.hpp"

namespace dellyb200 {

typedef std::map<std::pair<int32_t, std::size_t>, int32_t> TPosReadSV;  // (position, read id) -> SV id, per contig (src/shortpe.h:462-476)

// read id as the reference derives it (hash_sr, src/util.h:519-527: read 2 of a pair = read 1 + 1)
// (a caller that already has the reference's id passes it in SrRecord::seed)
inline std::size_t srSeed(SrRecord const& r) { return r.seed ? r.seed : (std::size_t) r.name * 2 + ((r.flag & 0x80) ? 1 : 0); }

// samples: the alignment lists of all input files, in file order (the reference walks contig by contig and, inside a contig, file by
// file: src/shortpe.h:81-156), each sorted like a coordinate-sorted BAM.
inline int assembleSplitReadsBatch(dgpu_ctx* ctx, Config const& c, std::vector<uint32_t> const& target_len, std::vector<const char*> const& chrseq,
                                   std::vector<TPosReadSV> const& srStore, std::vector<StructuralVariantRecord>& svs,
                                   std::vector<std::vector<SrRecord> const*> const& samples, AssembleShard const* shard = nullptr) {
  // The reference keeps the reads of an SV in a std::unordered_set<std::string>; its iteration order is the order msa() sees
  // (and it de-duplicates identical reads). The same container with the same insertion sequence gives the same order.
  typedef std::unordered_set<std::string> TSequences;
  const std::size_t N = svs.size();
  std::vector<TSequences> traStore(N);
  std::vector<std::vector<uint8_t> > traQualStore(N);
  std::vector<uint32_t> queued;                          // SV ids that reach msa + alignConsensus
  std::vector<std::vector<std::string> > clusters;       // their reads, in the set's iteration order
  std::vector<std::vector<uint8_t> > quals;
  auto reset = [&](StructuralVariantRecord& sv) { sv.consensus = ""; sv.srSupport = 0; sv.srAlignQuality = 0; };
  std::vector<std::size_t> ri(samples.size(), 0), rlo(samples.size(), 0);
  for (int32_t refIndex = 0; refIndex < (int32_t) target_len.size(); ++refIndex) {
    for (std::size_t f = 0; f < samples.size(); ++f) {
      std::vector<SrRecord> const& recs = *samples[f];
      rlo[f] = ri[f];
      while (ri[f] < recs.size() && recs[ri[f]].tid == refIndex) ++ri[f];
    }
    if (srStore[refIndex].empty()) continue;
    std::vector<bool> hits(target_len[refIndex], false);
    for (auto const& kv : srStore[refIndex]) hits[(std::size_t) kv.first.first] = true;
    std::vector<TSequences> seqStore(N);
    std::vector<std::vector<uint8_t> > qualStore(N);
    for (std::size_t f = 0; f < samples.size(); ++f)
    for (std::size_t q = rlo[f]; q < ri[f]; ++q) {
      SrRecord const& rec = (*samples[f])[q];
      if (rec.flag & (BAMF_QCFAIL | BAMF_DUP | BAMF_UNMAP | BAMF_SECONDARY | BAMF_SUPPLEMENTARY)) continue;
      if (rec.mapq < c.minMapQual) continue;
      if (!hits[(std::size_t) rec.pos]) continue;
      auto it = srStore[refIndex].find(std::make_pair(rec.pos, srSeed(rec)));
      if (it == srStore[refIndex].end()) continue;
      const int32_t svid = it->second;
      StructuralVariantRecord const& sv = svs[svid];
      if (svid != sv.id) continue;
      std::string sequence = rec.seq;
      bool bpPoint = false;   // :115-131
      if (_translocation(sv.svt)) { if (rec.tid == sv.chr2) bpPoint = true; }
      else if (sv.svt == 0) bpPoint = (rec.pos + 25 > sv.svStart);
      else if (sv.svt == 1) bpPoint = (rec.pos + 25 > sv.svEnd);
      _adjustOrientation(sequence, bpPoint ? 1u : 0u, sv.svt);
      if (seqStore[svid].size() < c.maxReadPerSV) {   // (the cap looks at seqStore for translocations too, :140)
        if (_translocation(sv.svt)) { if (traStore[svid].insert(std::move(sequence)).second) traQualStore[svid].push_back(rec.mapq); }
        else { if (seqStore[svid].insert(std::move(sequence)).second) qualStore[svid].push_back(rec.mapq); }
      }
    }
    for (uint32_t svid = 0; svid < N; ++svid) {
      if (_translocation(svs[svid].svt)) continue;
      if (svs[svid].chr != refIndex) continue;
      if (seqStore[svid].size() <= 1) { reset(svs[svid]); continue; }
      queued.push_back(svid);
      clusters.push_back(std::vector<std::string>(seqStore[svid].begin(), seqStore[svid].end()));
      quals.push_back(qualStore[svid]);
    }
  }
  for (int32_t refIndex2 = 0; refIndex2 < (int32_t) target_len.size(); ++refIndex2)
    for (int32_t refIndex = refIndex2 + 1; refIndex < (int32_t) target_len.size(); ++refIndex)
      for (uint32_t svid = 0; svid < N; ++svid) {
        if (!_translocation(svs[svid].svt)) continue;
        if ((svs[svid].chr != refIndex) || (svs[svid].chr2 != refIndex2)) continue;
        if (traStore[svid].size() <= 1) { reset(svs[svid]); continue; }
        queued.push_back(svid);
        clusters.push_back(std::vector<std::string>(traStore[svid].begin(), traStore[svid].end()));
        quals.push_back(traQualStore[svid]);
      }
  if (queued.empty()) return DGPU_OK;
  // The two device batches — msa() and alignConsensus() of every queued SV — are what a rank shards: the read collection above is
  // host work every rank repeats; of the queue a rank processes one contiguous range, cut by cost (all-pairs LCS + progressive alignment grow with
  // reads^2 x length^2), and the finished records of all ranks are exchanged once.
  std::size_t lo = 0, hi = queued.size();
  std::vector<std::size_t> bounds;
  if (shard && shard->nranks > 1) {
    std::vector<uint64_t> cost(queued.size());
    for (std::size_t k = 0; k < queued.size(); ++k) {
      uint64_t L = 0;
      for (auto const& s : clusters[k]) L = std::max<uint64_t>(L, s.size());
      cost[k] = (uint64_t) clusters[k].size() * clusters[k].size() * L * L;
    }
    bounds = partitionByCost(cost, shard->nranks);
    lo = bounds[shard->rank]; hi = bounds[shard->rank + 1];
  }
  std::vector<std::string> consensus;
  std::vector<int> rows;
  int rc = DGPU_OK;
  std::vector<StructuralVariantRecord> work(hi - lo);
  std::vector<uint8_t> ok;
  if (hi > lo) {
    std::vector<std::vector<std::string> > mine(clusters.begin() + lo, clusters.begin() + hi);
    if ((rc = msaBatch(ctx, c, mine, consensus, rows))) return rc;
    for (std::size_t k = lo; k < hi; ++k) { work[k - lo] = svs[queued[k]]; work[k - lo].consensus = consensus[k - lo]; }
    if ((rc = alignConsensusBatch(ctx, c, target_len, chrseq, work, false, ok))) return rc;
  }
  if (shard && shard->nranks > 1) {
    std::vector<StructuralVariantRecord> allWork;
    std::vector<uint8_t> allOk;
    if ((rc = shard->exchange(work, ok, bounds, allWork, allOk))) return rc;
    if (allWork.size() != queued.size()) return DGPU_ERR_NCCL;
    work.swap(allWork); ok.swap(allOk);
  }
  for (std::size_t k = 0; k < queued.size(); ++k) {
    StructuralVariantRecord& sv = svs[queued[k]];
    sv = work[k];
    if (!ok[k]) { reset(sv); continue; }
    std::vector<uint8_t>& qv = quals[k];
    sv.mapq = 0;
    for (uint8_t q : qv) sv.mapq += q;
    if (!qv.empty()) { const std::size_t n = qv.size() / 2; std::nth_element(qv.begin(), qv.begin() + n, qv.end()); sv.srMapQuality = qv[n]; }   // medianVector, src/util.h:78-84
    else sv.srMapQuality = 0;
    sv.srSupport = (int32_t) clusters[k].size();
  }
  return DGPU_OK;
}

inline int assembleSplitReadsBatch(dgpu_ctx* ctx, Config const& c, std::vector<uint32_t> const& target_len, std::vector<const char*> const& chrseq,
                                   std::vector<TPosReadSV> const& srStore, std::vector<StructuralVariantRecord>& svs, std::vector<SrRecord> const& recs) {
  return assembleSplitReadsBatch(ctx, c, target_len, chrseq, srStore, svs, std::vector<std::vector<SrRecord> const*>(1, &recs));
}

}  // namespace dellyb200
