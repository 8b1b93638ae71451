// pipeline.hpp — the stage sequence of `delly sr` discovery + genotyping (dellyRun, src/delly.h:86-196) over in-memory
// alignments of one sample: scanPEandSR -> assembleSplitReads -> mergeSort -> sort + renumber -> annotateCoverage ->
// per-sample genotype fields (what vcfOutput derives, src/modvcf.h:667-715). File handling, exclude intervals, library
// estimation (getLibraryParams) and the BCF writer stay with the caller; every stage is the batched mirror of this repository.
#pragma once
#include "assemble.hpp"
#include "genotype.hpp"
#include "gl.hpp"
#include "scan.hpp"

namespace dellyb200 {

struct SrCallSet {
  std::vector<StructuralVariantRecord> svs;   // sorted, ids renumbered (src/delly.h:155-158)
  std::vector<JunctionCount> jctMap;          // junction-read support per SV
  std::vector<SpanningCount> spanMap;         // spanning-pair support per SV
  std::vector<ReadCount> rcMap;               // read-depth left / inside / right of the SV
  std::vector<SampleFormat> format;           // GT / GL / GQ / PL / FT / RCN of the sample
};

inline int dellySrCall(dgpu_ctx* ctx, Config const& c, LibraryInfo& lib, std::vector<uint32_t> const& target_len, std::vector<std::string> const& target_name,
                       std::vector<const char*> const& chrseq, std::vector<SrRecord> const& recs, SrCallSet& out) {
  std::vector<StructuralVariantRecord> srSVs;
  std::vector<TPosReadSV> srStore;
  out = SrCallSet();
  int rc = scanPEandSRBatch(ctx, c, lib, target_len, recs, out.svs, srSVs, srStore);
  if (rc) return rc;
  if ((rc = assembleSplitReadsBatch(ctx, c, target_len, chrseq, srStore, srSVs, recs))) return rc;
  mergeSort(out.svs, srSVs);
  std::sort(out.svs.begin(), out.svs.end());
  for (std::size_t i = 0; i < out.svs.size(); ++i) out.svs[i].id = (int32_t) i;
  if (out.svs.empty()) return DGPU_OK;
  // annotateCoverage = junction-read half + spanning / read-depth half (independent of each other)
  if ((rc = annotateJunctionReadsBatch(ctx, c, target_len, target_name, chrseq, out.svs, recs, out.jctMap))) return rc;
  std::vector<bool> svOnChr(target_len.size(), false);
  for (auto const& sv : out.svs) { svOnChr[sv.chr] = true; svOnChr[sv.chr2] = true; }
  annotateSpanningAndDepth(c, lib, target_len, out.svs, svOnChr, recs, out.rcMap, out.spanMap);
  static const BoLog bl;
  out.format.resize(out.svs.size());
  for (std::size_t i = 0; i < out.svs.size(); ++i) {
    JunctionCount const& j = out.jctMap[i];
    // precise SVs are genotyped from junction reads, imprecise ones from spanning pairs (src/modvcf.h:668-669)
    std::vector<uint8_t> const& r = out.svs[i].precise ? j.ref : out.spanMap[i].ref;
    std::vector<uint8_t> const& a = out.svs[i].precise ? j.alt : out.spanMap[i].alt;
    out.format[i] = sampleFormat(bl, r, a, j.ps, (int32_t) j.hp1alt.size(), (int32_t) j.hp2alt.size(), out.rcMap[i].leftRC, out.rcMap[i].rc, out.rcMap[i].rightRC);
  }
  return DGPU_OK;
}

}  // namespace dellyb200
