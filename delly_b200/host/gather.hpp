// gather.hpp — the sharded (one process per GPU) form of the call chains and its one exchange step.
//
// After the SV list exists every SV is independent for consensus building, consensus alignment and genotyping, so the list is cut into
// CONTIGUOUS ranges balanced by estimated cost; each rank runs the batched stage mirrors on its range only (its own GPU) and the finished
// records — the full StructuralVariantRecord (src/tags.h:82-130 incl. SVAnno, consensus, alleles) plus the per-sample count vectors
// (JunctionCount / SpanningCount, src/coverage.h:69-85; ReadCount, src/util.h:69-76; MethylInfo, src/methyl.h:35-55) — are exchanged with ONE
// all-gatherv (dgpu_gather_records over NCCL in the product; any GatherFn in the tests). Because the ranges are contiguous ranges of the
// reference's order (the list is sorted and renumbered BEFORE it is cut, exactly where the reference sorts: src/delly.h:155-158), restoring
// the reference's order is the concatenation of the ranks' records in rank order, and the id of a record is its local index plus the
// exclusive prefix sum of the range sizes.
//
// What is replicated on every rank: reading the alignments, the junction scan and the (sequential, order-dependent) clustering — host work
// whose result every rank needs in full. What is sharded: every device batch (msa, longNeedle / splitAlign, probe generation, the junction
// read realignment, read-depth and spanning-pair bookkeeping of the rank's own SVs).
//
// Exactness: the N-rank result equals the 1-rank result record for record, unconditionally. The only cross-SV coupling in the genotyping
// pass is the batching of the junction-read jobs (flushed every 131072 x threads jobs and at the end of a contig, src/coverage.h:271,446,671):
// the per-SV cap is consulted both when a job is queued and when its result is merged, results are merged in job order, and a job that is
// not queued because the cap was already reached would have been dropped at the merge anyway — so the final lists of an SV depend only on
// that SV's own jobs in order, not on where the flushes fall. Every SV's jobs stay on one rank, in the reference's order.
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "pipeline.hpp"
#include "shard.hpp"

namespace dellyb200 {

// all-gatherv of one byte string per rank: parts[r] = rank r's payload, same on every rank. Returns DGPU_OK or an error code.
typedef std::function<int(std::string const& local, std::vector<std::string>& parts)> GatherFn;

struct Shard {
  int rank = 0, nranks = 1;
  GatherFn gather;   // empty = single process
  bool active() const { return nranks > 1 && (bool) gather; }
};

// wall-clock of the chain's stages (for the pipeline-level benchmark): name -> milliseconds, in call order
struct StageClock {
  std::vector<std::pair<std::string, double> > ms;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void lap(const char* name) {
    const auto t1 = std::chrono::steady_clock::now();
    ms.push_back(std::make_pair(std::string(name), std::chrono::duration<double, std::milli>(t1 - t0).count()));
    t0 = t1;
  }
};

// ---- wire format (little endian, host order; all ranks are the same binary on the same box) ----------------------------------------
class ByteWriter {
 public:
  std::string buf;
  template <typename T> void pod(T const& v) { buf.append(reinterpret_cast<const char*>(&v), sizeof(T)); }
  void str(std::string const& s) { pod<uint32_t>((uint32_t) s.size()); buf.append(s); }
  void bytes(std::vector<uint8_t> const& v) { pod<uint32_t>((uint32_t) v.size()); if (!v.empty()) buf.append(reinterpret_cast<const char*>(v.data()), v.size()); }
};
class ByteReader {
 public:
  ByteReader(const char* p, std::size_t n) : p_(p), end_(p + n) {}
  bool ok() const { return ok_; }
  bool done() const { return p_ >= end_; }
  template <typename T> T pod() { T v = T(); if (p_ + sizeof(T) > end_) { ok_ = false; return v; } std::memcpy(&v, p_, sizeof(T)); p_ += sizeof(T); return v; }
  std::string str() { const uint32_t n = pod<uint32_t>(); if (!ok_ || p_ + n > end_) { ok_ = false; return std::string(); } std::string s(p_, n); p_ += n; return s; }
  std::vector<uint8_t> bytes() { const uint32_t n = pod<uint32_t>(); if (!ok_ || p_ + n > end_) { ok_ = false; return {}; } std::vector<uint8_t> v(p_, p_ + n); p_ += n; return v; }
 private:
  const char* p_; const char* end_; bool ok_ = true;
};

inline void putSv(ByteWriter& w, StructuralVariantRecord const& v) {
  const int32_t f[20] = {v.chr, v.svStart, v.chr2, v.svEnd, v.ciposlow, v.ciposhigh, v.ciendlow, v.ciendhigh, v.srSupport, v.srMapQuality,
                         v.mapq, v.insLen, v.svt, v.id, v.homLen, v.peSupport, v.peMapQuality, v.consBp, v.alleleid, v.nallele};
  for (int32_t x : f) w.pod(x);
  w.pod(v.srAlignQuality);
  w.pod<uint8_t>(v.precise ? 1 : 0);
  w.pod<uint8_t>(v.anno.isRC ? 1 : 0); w.pod(v.anno.seqType); w.pod(v.anno.homLen); w.pod(v.anno.trPeriod); w.pod(v.anno.trCopies);   // SVAnno (src/tags.h:82-91)
  w.str(v.alleles); w.str(v.consensus);
}
inline void getSv(ByteReader& r, StructuralVariantRecord& v) {
  int32_t f[20];
  for (int32_t& x : f) x = r.pod<int32_t>();
  v.chr = f[0]; v.svStart = f[1]; v.chr2 = f[2]; v.svEnd = f[3]; v.ciposlow = f[4]; v.ciposhigh = f[5]; v.ciendlow = f[6]; v.ciendhigh = f[7]; v.srSupport = f[8];
  v.srMapQuality = f[9]; v.mapq = f[10]; v.insLen = f[11]; v.svt = f[12]; v.id = f[13]; v.homLen = f[14]; v.peSupport = f[15]; v.peMapQuality = f[16]; v.consBp = f[17];
  v.alleleid = f[18]; v.nallele = f[19];
  v.srAlignQuality = r.pod<float>();
  v.precise = r.pod<uint8_t>() != 0;
  v.anno.isRC = r.pod<uint8_t>() != 0; v.anno.seqType = r.pod<int32_t>(); v.anno.homLen = r.pod<int32_t>(); v.anno.trPeriod = r.pod<int32_t>(); v.anno.trCopies = r.pod<float>();
  v.alleles = r.str(); v.consensus = r.str();
}
inline void putJct(ByteWriter& w, JunctionCount const& j) { w.bytes(j.ref); w.bytes(j.alt); w.bytes(j.hp1ref); w.bytes(j.hp1alt); w.bytes(j.hp2ref); w.bytes(j.hp2alt); w.pod(j.ps); }
inline void getJct(ByteReader& r, JunctionCount& j) { j.ref = r.bytes(); j.alt = r.bytes(); j.hp1ref = r.bytes(); j.hp1alt = r.bytes(); j.hp2ref = r.bytes(); j.hp2alt = r.bytes(); j.ps = r.pod<int32_t>(); }
inline void putMethyl(ByteWriter& w, MethylInfo const& m) { for (int k = 0; k < 4; ++k) { w.pod(m.alt[k]); w.pod(m.ref[k]); w.pod(m.mnc[k]); w.pod(m.mdp[k]); } }
inline void getMethyl(ByteReader& r, MethylInfo& m) { for (int k = 0; k < 4; ++k) { m.alt[k] = r.pod<int32_t>(); m.ref[k] = r.pod<int32_t>(); m.mnc[k] = r.pod<int32_t>(); m.mdp[k] = r.pod<int32_t>(); } }

// One rank's finished records: its SVs (ids already global) and, per sample, the count vectors of those SVs.
struct ShardRecords {
  std::vector<StructuralVariantRecord> svs;
  struct Sample { std::vector<JunctionCount> jct; std::vector<SpanningCount> span; std::vector<ReadCount> rc; std::vector<MethylInfo> methyl; };
  std::vector<Sample> sample;
};

inline std::string packShardRecords(ShardRecords const& s) {
  ByteWriter w;
  w.pod<uint32_t>(0x44475055u);   // "DGPU"
  w.pod<uint32_t>((uint32_t) s.svs.size());
  w.pod<uint32_t>((uint32_t) s.sample.size());
  for (auto const& v : s.svs) putSv(w, v);
  for (auto const& sm : s.sample) {
    const uint8_t has[4] = {(uint8_t) !sm.jct.empty(), (uint8_t) !sm.span.empty(), (uint8_t) !sm.rc.empty(), (uint8_t) !sm.methyl.empty()};
    for (uint8_t h : has) w.pod(h);
    if (has[0]) for (auto const& j : sm.jct) putJct(w, j);
    if (has[1]) for (auto const& sp : sm.span) { w.bytes(sp.ref); w.bytes(sp.alt); }
    if (has[2]) for (auto const& rc : sm.rc) { w.pod(rc.leftRC); w.pod(rc.rc); w.pod(rc.rightRC); }
    if (has[3]) for (auto const& m : sm.methyl) putMethyl(w, m);
  }
  return w.buf;
}

inline bool unpackShardRecords(std::string const& buf, ShardRecords& s) {
  ByteReader r(buf.data(), buf.size());
  if (r.pod<uint32_t>() != 0x44475055u) return false;
  const uint32_t n = r.pod<uint32_t>(), F = r.pod<uint32_t>();
  if (!r.ok()) return false;
  s.svs.assign(n, StructuralVariantRecord());
  for (auto& v : s.svs) getSv(r, v);
  s.sample.assign(F, ShardRecords::Sample());
  for (auto& sm : s.sample) {
    uint8_t has[4];
    for (uint8_t& h : has) h = r.pod<uint8_t>();
    if (has[0]) { sm.jct.resize(n); for (auto& j : sm.jct) getJct(r, j); }
    if (has[1]) { sm.span.resize(n); for (auto& sp : sm.span) { sp.ref = r.bytes(); sp.alt = r.bytes(); } }
    if (has[2]) { sm.rc.resize(n); for (auto& rc : sm.rc) { rc.leftRC = r.pod<int32_t>(); rc.rc = r.pod<int32_t>(); rc.rightRC = r.pod<int32_t>(); } }
    if (has[3]) { sm.methyl.resize(n); for (auto& m : sm.methyl) getMethyl(r, m); }
  }
  return r.ok() && r.done();
}

// Genotyping cost model of one SV, in DP cell units: the probe alignment of a precise SV (longNeedle of the consensus against the SV
// window: 3 matrices of |cons| x (|cons| + 2 x window)) plus the junction-read realignments of both breakpoints
// (≈ coverage x read length cells per probe; a constant per SV at a given depth).
inline uint64_t genotypingCost(StructuralVariantRecord const& sv, Config const& c) {
  uint64_t cost = 20000;
  if (sv.precise && !sv.consensus.empty()) cost += 3ull * sv.consensus.size() * (sv.consensus.size() + 2ull * (uint64_t) std::max(c.minConsWindow, 1));
  return cost;
}

// ---- sharded genotyping of a sorted, renumbered SV list (one or several samples) ---------------------------------------------------
// in/out: cs.svs = the COMPLETE list, sorted and renumbered, identical on every rank. On return every rank holds the complete result.
inline int genotypeSrSitesSharded(dgpu_ctx* ctx, Config const& c, std::vector<LibraryInfo> const& libs, std::vector<uint32_t> const& target_len,
                                  std::vector<std::string> const& target_name, std::vector<const char*> const& chrseq,
                                  std::vector<std::vector<SrRecord> const*> const& samples, Shard const& shard, SrMultiCallSet& cs, StageClock* clock = nullptr) {
  const std::size_t F = samples.size();
  cs.sample.assign(F, SrSampleCounts());
  if (cs.svs.empty()) return DGPU_OK;
  if (!shard.active()) {
    JunctionProbes probes;
    int rc = prepareJunctionProbes(ctx, c, target_len, target_name, chrseq, cs.svs, probes);
    if (rc) return rc;
    if (clock) clock->lap("probes");
    for (std::size_t f = 0; f < F; ++f) {
      SrCallSet one; one.svs = cs.svs;
      if ((rc = genotypeSrSites(ctx, c, libs[f], target_len, target_name, chrseq, *samples[f], one, &probes))) return rc;
      cs.sample[f].jctMap.swap(one.jctMap); cs.sample[f].spanMap.swap(one.spanMap); cs.sample[f].rcMap.swap(one.rcMap); cs.sample[f].format.swap(one.format);
    }
    if (clock) clock->lap("genotype");
    return DGPU_OK;
  }
  std::vector<uint64_t> cost(cs.svs.size());
  for (std::size_t i = 0; i < cs.svs.size(); ++i) cost[i] = genotypingCost(cs.svs[i], c);
  const std::vector<std::size_t> bounds = partitionByCost(cost, shard.nranks);
  const std::size_t lo = bounds[shard.rank], hi = bounds[shard.rank + 1];
  // this rank's range with rank-local ids 0..k-1 (the count maps are indexed by id)
  ShardRecords mine;
  mine.svs.assign(cs.svs.begin() + lo, cs.svs.begin() + hi);
  for (std::size_t i = 0; i < mine.svs.size(); ++i) mine.svs[i].id = (int32_t) i;
  mine.sample.assign(F, ShardRecords::Sample());
  int rc = DGPU_OK;
  if (!mine.svs.empty()) {
    JunctionProbes probes;
    if ((rc = prepareJunctionProbes(ctx, c, target_len, target_name, chrseq, mine.svs, probes))) return rc;
    if (clock) clock->lap("probes");
    for (std::size_t f = 0; f < F; ++f) {
      SrCallSet one; one.svs = mine.svs;
      if ((rc = genotypeSrSites(ctx, c, libs[f], target_len, target_name, chrseq, *samples[f], one, &probes))) return rc;
      mine.sample[f].jct.swap(one.jctMap); mine.sample[f].span.swap(one.spanMap); mine.sample[f].rc.swap(one.rcMap);
    }
  }
  if (clock) clock->lap("genotype");
  // ids become global: local index + exclusive prefix sum of the range sizes (= lo)
  for (std::size_t i = 0; i < mine.svs.size(); ++i) mine.svs[i].id = (int32_t) (lo + i);
  std::vector<std::string> parts;
  if ((rc = shard.gather(packShardRecords(mine), parts))) return rc;
  if ((int) parts.size() != shard.nranks) return DGPU_ERR_NCCL;
  if (clock) clock->lap("gather_records");
  // restore the reference's order: concatenation in rank order (the ranges are contiguous ranges of the sorted list)
  std::vector<StructuralVariantRecord> all;
  for (std::size_t f = 0; f < F; ++f) { cs.sample[f].jctMap.clear(); cs.sample[f].spanMap.clear(); cs.sample[f].rcMap.clear(); }
  for (int r = 0; r < shard.nranks; ++r) {
    ShardRecords part;
    if (!unpackShardRecords(parts[r], part)) return DGPU_ERR_NCCL;
    if (part.svs.size() != bounds[r + 1] - bounds[r] || (part.sample.size() != F && !part.svs.empty())) return DGPU_ERR_NCCL;
    for (std::size_t i = 0; i < part.svs.size(); ++i) {
      if (part.svs[i].id != (int32_t) (bounds[r] + i)) return DGPU_ERR_NCCL;
      all.push_back(part.svs[i]);
    }
    for (std::size_t f = 0; f < F && !part.svs.empty(); ++f) {
      cs.sample[f].jctMap.insert(cs.sample[f].jctMap.end(), part.sample[f].jct.begin(), part.sample[f].jct.end());
      cs.sample[f].spanMap.insert(cs.sample[f].spanMap.end(), part.sample[f].span.begin(), part.sample[f].span.end());
      cs.sample[f].rcMap.insert(cs.sample[f].rcMap.end(), part.sample[f].rc.begin(), part.sample[f].rc.end());
    }
  }
  if (all.size() != cs.svs.size()) return DGPU_ERR_NCCL;
  cs.svs.swap(all);
  static const BoLog bl;
  for (std::size_t f = 0; f < F; ++f) {
    SrSampleCounts& sc = cs.sample[f];
    sc.format.resize(cs.svs.size());
    for (std::size_t i = 0; i < cs.svs.size(); ++i) {
      JunctionCount const& j = sc.jctMap[i];
      std::vector<uint8_t> const& rr = cs.svs[i].precise ? j.ref : sc.spanMap[i].ref;
      std::vector<uint8_t> const& aa = cs.svs[i].precise ? j.alt : sc.spanMap[i].alt;
      sc.format[i] = sampleFormat(bl, rr, aa, j.ps, (int32_t) j.hp1alt.size(), (int32_t) j.hp2alt.size(), sc.rcMap[i].leftRC, sc.rcMap[i].rc, sc.rcMap[i].rightRC);
    }
  }
  return DGPU_OK;
}

// `delly sr -v sites.bcf` sharded over ranks (BASELINE configs[3]: genotype mode, cluster-sharded)
inline int dellySrGenotypeSharded(dgpu_ctx* ctx, Config const& c, std::vector<LibraryInfo> const& libs, std::vector<uint32_t> const& target_len,
                                  std::vector<std::string> const& target_name, std::vector<const char*> const& chrseq, std::vector<VcfSite> const& sites,
                                  bool headerHasConsBp, std::vector<std::vector<SrRecord> const*> const& samples, Shard const& shard, SrMultiCallSet& out,
                                  StageClock* clock = nullptr) {
  out = SrMultiCallSet();
  const bool ok = vcfParseSites(sites, headerHasConsBp, target_name, out.svs);
  for (auto const& sv : out.svs)
    if (sv.chr < 0 || sv.chr2 < 0) return DGPU_ERR_ARG;
  std::sort(out.svs.begin(), out.svs.end());   // src/delly.h:155-158, on every rank identically, BEFORE the list is cut
  for (std::size_t i = 0; i < out.svs.size(); ++i) out.svs[i].id = (int32_t) i;
  if (clock) clock->lap("parse_sort_sites");
  const int rc = genotypeSrSitesSharded(ctx, c, libs, target_len, target_name, chrseq, samples, shard, out, clock);
  if (rc) return rc;
  return ok ? DGPU_OK : DGPU_ERR_ARG;
}

inline AssembleShard makeAssembleShard(Shard const& shard);

// `delly sr` discovery + genotyping sharded over ranks: scan + clustering replicated (host), split-read assembly and genotyping sharded.
inline int dellySrCallSharded(dgpu_ctx* ctx, Config const& c, std::vector<LibraryInfo>& libs, std::vector<uint32_t> const& target_len,
                              std::vector<std::string> const& target_name, std::vector<const char*> const& chrseq,
                              std::vector<std::vector<SrRecord> const*> const& samples, Shard const& shard, SrMultiCallSet& out, StageClock* clock = nullptr,
                              std::vector<std::vector<SrRecord> const*> const* genoSamples = nullptr) {
  // samples: the records of the valid regions (scan, clustering, assembly); genoSamples: the records of the whole contigs for the genotyping pass
  // (src/coverage.h reads every contig end to end) — the same lists unless an exclude file is in play.
  std::vector<StructuralVariantRecord> srSVs;
  std::vector<TPosReadSV> srStore;
  out = SrMultiCallSet();
  int rc = scanPEandSRBatch(ctx, c, libs, target_len, samples, out.svs, srSVs, srStore);
  if (rc) return rc;
  if (clock) clock->lap("scan_cluster");
  if (shard.active()) {
    AssembleShard as = makeAssembleShard(shard);
    if ((rc = assembleSplitReadsBatch(ctx, c, target_len, chrseq, srStore, srSVs, samples, &as))) return rc;
  } else if ((rc = assembleSplitReadsBatch(ctx, c, target_len, chrseq, srStore, srSVs, samples))) return rc;
  if (clock) clock->lap("assemble");
  mergeSort(out.svs, srSVs);
  std::sort(out.svs.begin(), out.svs.end());
  for (std::size_t i = 0; i < out.svs.size(); ++i) out.svs[i].id = (int32_t) i;
  if (clock) clock->lap("merge_sort");
  return genotypeSrSitesSharded(ctx, c, libs, target_len, target_name, chrseq, genoSamples ? *genoSamples : samples, shard, out, clock);
}

// The exchange of an assembly queue cut by cost (AssembleShard): every rank sends the finished records of its range and receives all of them in
// queue order.
inline AssembleShard makeAssembleShard(Shard const& shard) {
  AssembleShard as;
  as.rank = shard.rank; as.nranks = shard.nranks;
  as.exchange = [shard](std::vector<StructuralVariantRecord> const& mineWork, std::vector<uint8_t> const& mineOk, std::vector<std::size_t> const& bounds,
                        std::vector<StructuralVariantRecord>& allWork, std::vector<uint8_t>& allOk) -> int {
    ByteWriter w;
    w.pod<uint32_t>((uint32_t) mineWork.size());
    for (std::size_t k = 0; k < mineWork.size(); ++k) { putSv(w, mineWork[k]); w.pod<uint8_t>(mineOk[k]); }
    std::vector<std::string> parts;
    int r2 = shard.gather(w.buf, parts);
    if (r2) return r2;
    if ((int) parts.size() != shard.nranks) return DGPU_ERR_NCCL;
    allWork.clear(); allOk.clear();
    for (int r = 0; r < shard.nranks; ++r) {
      ByteReader rd(parts[r].data(), parts[r].size());
      const uint32_t n = rd.pod<uint32_t>();
      if (!rd.ok() || n != bounds[r + 1] - bounds[r]) return DGPU_ERR_NCCL;
      for (uint32_t k = 0; k < n; ++k) { StructuralVariantRecord v; getSv(rd, v); allWork.push_back(v); allOk.push_back(rd.pod<uint8_t>()); }
      if (!rd.ok()) return DGPU_ERR_NCCL;
    }
    return DGPU_OK;
  };
  return as;
}

// ---- long reads (`delly lr`) ------------------------------------------------------------------------------------------------------------
// Sharded genotyping + annotation of an SV list whose ids are its indices (discovery: sorted and renumbered, src/tegua.h:150-156; -v: file order).
// A rank takes a contiguous range by cost, genotypes it in every sample with rank-local ids (genotypeLR fills the alleles, annotateSV reads them),
// ids become local index + range start, one all-gatherv of the complete records and count vectors, concatenation in rank order.
inline int genotypeLrSitesSharded(dgpu_ctx* ctx, Config const& c, std::vector<uint32_t> const& target_len, std::vector<std::string> const& target_name,
                                  std::vector<const char*> const& chrseq, std::vector<LrSample> const& samples, Shard const& shard, LrMultiCallSet& cs,
                                  MeiTemplates const* mei, AnnoConfig const& annoCfg, MethylConfig const* methylCfg, bool annotateBeforeFormat) {
  const std::size_t F = samples.size();
  cs.sample.assign(F, LrSampleCounts());
  int rc;
  auto runRange = [&](std::vector<StructuralVariantRecord>& svs, std::vector<LrSampleCounts>& out) -> int {
    out.assign(F, LrSampleCounts());
    for (std::size_t f = 0; f < F; ++f) {
      LrSampleCounts& sc = out[f];
      int r = genotypeLRBatch(ctx, c, target_len, target_name, chrseq, svs, *samples[f].recs, sc.jctMap, sc.rcMap, methylCfg, methylCfg ? &sc.methyl : nullptr);
      if (r) return r;
    }
    if (mei) { int r = annotateSVs(ctx, annoCfg, *mei, chrseq, target_len, svs); if (r) return r; }
    return DGPU_OK;
  };
  (void) annotateBeforeFormat;
  if (!shard.active()) {
    if ((rc = runRange(cs.svs, cs.sample))) return rc;
  } else {
    std::vector<uint64_t> cost(cs.svs.size());
    for (std::size_t i = 0; i < cs.svs.size(); ++i) cost[i] = genotypingCost(cs.svs[i], c);
    const std::vector<std::size_t> bounds = partitionByCost(cost, shard.nranks);
    const std::size_t lo = bounds[shard.rank], hi = bounds[shard.rank + 1];
    ShardRecords mine;
    mine.svs.assign(cs.svs.begin() + lo, cs.svs.begin() + hi);
    for (std::size_t i = 0; i < mine.svs.size(); ++i) mine.svs[i].id = (int32_t) i;
    std::vector<LrSampleCounts> local;
    if ((rc = runRange(mine.svs, local))) return rc;
    mine.sample.assign(F, ShardRecords::Sample());
    for (std::size_t f = 0; f < F; ++f) { mine.sample[f].jct.swap(local[f].jctMap); mine.sample[f].rc.swap(local[f].rcMap); mine.sample[f].methyl.swap(local[f].methyl); }
    for (std::size_t i = 0; i < mine.svs.size(); ++i) mine.svs[i].id = (int32_t) (lo + i);
    std::vector<std::string> parts;
    if ((rc = shard.gather(packShardRecords(mine), parts))) return rc;
    if ((int) parts.size() != shard.nranks) return DGPU_ERR_NCCL;
    std::vector<StructuralVariantRecord> all;
    for (int r = 0; r < shard.nranks; ++r) {
      ShardRecords part;
      if (!unpackShardRecords(parts[r], part)) return DGPU_ERR_NCCL;
      if (part.svs.size() != bounds[r + 1] - bounds[r] || (part.sample.size() != F && !part.svs.empty())) return DGPU_ERR_NCCL;
      for (std::size_t i = 0; i < part.svs.size(); ++i) {
        if (part.svs[i].id != (int32_t) (bounds[r] + i)) return DGPU_ERR_NCCL;
        all.push_back(part.svs[i]);
      }
      for (std::size_t f = 0; f < F && !part.svs.empty(); ++f) {
        LrSampleCounts& sc = cs.sample[f];
        sc.jctMap.insert(sc.jctMap.end(), part.sample[f].jct.begin(), part.sample[f].jct.end());
        sc.rcMap.insert(sc.rcMap.end(), part.sample[f].rc.begin(), part.sample[f].rc.end());
        sc.methyl.insert(sc.methyl.end(), part.sample[f].methyl.begin(), part.sample[f].methyl.end());
      }
    }
    if (all.size() != cs.svs.size()) return DGPU_ERR_NCCL;
    cs.svs.swap(all);
  }
  for (std::size_t f = 0; f < F; ++f) lrSampleFormat(cs.svs, cs.sample[f].jctMap, cs.sample[f].rcMap, cs.sample[f].format);
  return DGPU_OK;
}

// `delly lr` over ranks: the record stream, the junction clustering and the read collection replicated (host), the per-SV consensus + consensus
// alignment and the genotyping / annotation sharded by cost.
inline int dellyLrCallSharded(dgpu_ctx* ctx, Config const& c, float indelExtension, std::vector<uint32_t> const& target_len, std::vector<std::string> const& target_name,
                              std::vector<const char*> const& chrseq, std::vector<LrSample> const& samples, Shard const& shard, LrMultiCallSet& out,
                              MeiTemplates const* mei = nullptr, AnnoConfig const& annoCfg = AnnoConfig(), MethylConfig const* methylCfg = nullptr,
                              std::vector<LrSample> const* genoSamples = nullptr) {
  out = LrMultiCallSet();
  std::vector<LrRecord> stream;
  std::vector<std::size_t> streamIds;
  buildLrStream(target_len, samples, stream, streamIds);
  AssembleShard as;
  if (shard.active()) as = makeAssembleShard(shard);
  int rc = discoverLrSVs(ctx, c, indelExtension, target_len, chrseq, stream, streamIds, out.svs, shard.active() ? &as : nullptr);
  if (rc) return rc;
  return genotypeLrSitesSharded(ctx, c, target_len, target_name, chrseq, genoSamples ? *genoSamples : samples, shard, out, mei, annoCfg, methylCfg, true);
}

// `delly lr -v sites.bcf` over ranks (ids = file order, src/tegua.h:157)
inline int dellyLrGenotypeSharded(dgpu_ctx* ctx, Config const& c, std::vector<uint32_t> const& target_len, std::vector<std::string> const& target_name,
                                  std::vector<const char*> const& chrseq, std::vector<VcfSite> const& sites, bool headerHasConsBp, std::vector<LrSample> const& samples,
                                  Shard const& shard, LrMultiCallSet& out, MeiTemplates const* mei = nullptr, AnnoConfig const& annoCfg = AnnoConfig(),
                                  MethylConfig const* methylCfg = nullptr) {
  out = LrMultiCallSet();
  const bool ok = vcfParseSites(sites, headerHasConsBp, target_name, out.svs);
  for (auto const& sv : out.svs)
    if (sv.chr < 0 || sv.chr2 < 0) return DGPU_ERR_ARG;
  for (std::size_t i = 0; i < out.svs.size(); ++i)
    if (out.svs[i].id != (int32_t) i) return DGPU_ERR_ARG;   // the count vectors are indexed by id
  const int rc = genotypeLrSitesSharded(ctx, c, target_len, target_name, chrseq, samples, shard, out, mei, annoCfg, methylCfg, false);
  if (rc) return rc;
  return ok ? DGPU_OK : DGPU_ERR_ARG;
}

}  // namespace dellyb200
