// vcfparse.hpp — genotyping mode (`delly call -v sites.bcf`, BASELINE config 4): the site list a Delly BCF holds, turned
// back into StructuralVariantRecords. Host mirror of
//   src/modvcf.h:156-339  vcfParse (the per-record logic; reading the BCF itself stays with htslib in the binding)
//   src/util.h:218-235    _decodeOrientation(CT, SVTYPE)
// The binding fills one VcfSite per BCF record with what the reference's bcf_get_info_* calls would return (value and
// presence), in file order; vcfParseSites applies the reference's rules: the file counts as a Delly file from the first
// record whose SVMETHOD starts with "EMBL.DELLY" while the header defines CONSBP; records without SVTYPE / CT, and
// insertions without SVLEN, are skipped; sequence-resolved alleles fix the end position; missing CONSENSUS makes the site
// imprecise; a record seen before the file qualified as a Delly file stops the parse (the reference prints an error).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "types.hpp"

namespace dellyb200 {

template <typename T>
struct VcfOpt {  // an INFO value and whether the record carries it
  bool present = false;
  T value = T();
};

struct VcfSite {
  std::string chrom;   // bcf_hdr_id2name(hdr, rec->rid)
  int32_t pos0 = 0;    // rec->pos (0-based)
  float qual = 0;      // rec->qual
  std::string ref, alt;  // rec->d.allele[0], rec->d.allele[1]
  bool precise = false;  // INFO/PRECISE flag
  VcfOpt<std::string> svMethod, svType, ct, chr2, consensus;
  VcfOpt<int32_t> pe, insLen, svLen, homLen, sr, end, pos2, consBp, mapq, srMapq, alleleId, nAllele;
  VcfOpt<float> srq;
  bool hasCipos = false, hasCiend = false;
  int32_t cipos[2] = {0, 0}, ciend[2] = {0, 0};
};

inline int32_t _decodeOrientation(std::string const& value, std::string const& svt) {  // src/util.h:218-235
  const int32_t ct = (value == "3to3") ? 0 : (value == "5to5") ? 1 : (value == "3to5") ? 2 : (value == "5to3") ? 3 : 4;
  if (svt == "BND") return (ct < 4) ? DELLY_SVT_TRANS + ct : -1;
  if (svt == "CNV") return 9;
  return ct;
}

// Returns false when the parse stopped at a record that is not from a Delly (>= v1.1.7) file. target_name plays the role
// of the BAM header (bam_name2id: -1 for an unknown contig). headerHasConsBp = _isKeyPresent(hdr, "CONSBP").
inline bool vcfParseSites(std::vector<VcfSite> const& sites, bool headerHasConsBp, std::vector<std::string> const& target_name,
                          std::vector<StructuralVariantRecord>& svs) {
  auto name2id = [&](std::string const& nm) {
    for (std::size_t k = 0; k < target_name.size(); ++k)
      if (target_name[k] == nm) return (int32_t) k;
    return (int32_t) -1;
  };
  bool dellyVCF = false;
  for (VcfSite const& site : sites) {
    VcfSite s = site;
    // bcf_get_info_string(...) > 0: a string of length zero counts as absent
    for (VcfOpt<std::string>* o : {&s.svMethod, &s.svType, &s.ct, &s.chr2, &s.consensus}) o->present = o->present && !o->value.empty();
    if (!dellyVCF && s.svMethod.present) {
      std::string const& m = s.svMethod.value;
      if (m.size() >= 10 && m.compare(0, 10, "EMBL.DELLY") == 0 && headerHasConsBp) dellyVCF = true;
    }
    if (!dellyVCF) return false;
    StructuralVariantRecord sv;
    const int32_t tid = name2id(s.chrom);
    sv.chr = tid;
    sv.svStart = s.pos0 + 1;
    sv.id = (int32_t) svs.size();
    sv.mapq = (int32_t) s.qual;
    if (!s.ref.empty() && s.ref != "." && !s.alt.empty() && s.alt != ".") sv.alleles = s.ref + "," + s.alt;
    if (!(s.svType.present && s.ct.present)) continue;
    sv.svt = _decodeOrientation(s.ct.value, s.svType.value);
    sv.precise = s.precise;
    sv.peSupport = s.pe.present ? s.pe.value : (sv.precise ? 0 : 2);
    if (sv.svt != 4) sv.insLen = s.insLen.present ? s.insLen.value : 0;
    else {
      if (!s.svLen.present) continue;  // insertions must carry SVLEN
      sv.insLen = s.svLen.value;
    }
    sv.homLen = s.homLen.present ? s.homLen.value : 0;
    sv.srSupport = s.sr.present ? s.sr.value : 0;
    sv.chr2 = tid;
    sv.svEnd = s.pos0 + 1;
    if (sv.svt < DELLY_SVT_TRANS) {
      if (s.end.present) sv.svEnd = s.end.value;
      if (!sv.alleles.empty()) {  // a sequence-resolved ALT: the end follows from the REF allele's length
        const std::size_t comma = sv.alleles.find(',');
        bool symbolic = (comma == std::string::npos);
        if (!symbolic) {
          const std::string alt = sv.alleles.substr(comma + 1);
          symbolic = !alt.empty() && (alt[0] == '<' || alt.find('[') != std::string::npos || alt.find(']') != std::string::npos);
        }
        if (!symbolic) sv.svEnd = sv.svStart + (int32_t) comma;
      }
    } else {
      if (s.chr2.present) sv.chr2 = name2id(s.chr2.value);
      if (s.pos2.present) sv.svEnd = s.pos2.value;
    }
    if (s.consensus.present) {
      sv.consensus = s.consensus.value;
      if (s.consBp.present) sv.consBp = s.consBp.value;
    } else sv.precise = false;
    sv.ciposlow = s.hasCipos ? s.cipos[0] : -50;
    sv.ciposhigh = s.hasCipos ? s.cipos[1] : 50;
    sv.ciendlow = s.hasCiend ? s.ciend[0] : -50;
    sv.ciendhigh = s.hasCiend ? s.ciend[1] : 50;
    sv.peMapQuality = s.mapq.present ? (int32_t) (uint8_t) s.mapq.value : 0;
    sv.srMapQuality = s.srMapq.present ? (int32_t) (uint8_t) s.srMapq.value : 0;
    sv.srAlignQuality = s.srq.present ? (float) (double) s.srq.value : 0;
    if (s.alleleId.present) {
      sv.alleleid = s.alleleId.value;
      if (s.nAllele.present) sv.nallele = s.nAllele.value;
    }
    svs.push_back(sv);
  }
  return true;
}

}  // namespace dellyb200
