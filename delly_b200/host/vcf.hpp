// vcf.hpp — what vcfOutput (src/modvcf.h:344-791) hands to htslib, as a field-by-field record description: the header lines,
// and per SV record CHROM / POS / QUAL / ID / alleles / FILTER, every INFO key and every FORMAT key with its values, in the
// reference's call order. Serialising these to BCF bytes is htslib's job and stays with the caller; this is everything
// above that line. Text form (one line per header line / record; floats as bit patterns "f%08x") so that it can be
// compared against the reference's calls one to one (oracle/ref_wrap7.cpp records them the same way).
#pragma once
#include <cmath>
#include <cstdio>
#include <cstring>
#include <set>
#include <string>
#include <vector>

#include "genotype.hpp"
#include "gl.hpp"
#include "methyl.hpp"
#include "types.hpp"

namespace dellyb200 {

inline std::string _addOrientation(int32_t svt) {  // src/util.h:276-285
  switch (_getSpanOrientation(svt)) { case 0: return "3to3"; case 1: return "5to5"; case 2: return "3to5"; case 3: return "5to3"; default: return "NtoN"; }
}

inline double entropy(std::string const& st) {  // src/util.h:565-579 (std::set<char> order, double arithmetic)
  std::set<char> alphabet(st.begin(), st.end());
  double ent = 0;
  for (char c : alphabet) {
    int ctr = 0;
    for (char s : st) if (s == c) ++ctr;
    const double freq = (double) ctr / (double) st.size();
    ent += freq * std::log2(freq);
  }
  return -ent;
}

inline std::string _replaceIUPAC(std::string const& alleles) {  // src/modvcf.h:101-137: IUPAC codes in the ALT allele become a plain base
  std::string out(alleles.size(), 'N');
  int32_t inTag = 0;
  bool inRef = true;
  for (std::size_t i = 0; i < alleles.size(); ++i) {
    const char a = alleles[i];
    if (a == ',') inRef = false;
    if (inRef || inTag || std::strchr("ACGTNacgtn<>][,", a) != nullptr) {
      out[i] = a;
      // as in the reference, a bracket always (re)opens its tag: the closing tests for ']' / '[' come after these and never fire,
      // so everything behind a breakend bracket is copied through unchanged
      if (a == '<') inTag = 1;
      else if (a == ']') inTag = 2;
      else if (a == '[') inTag = 3;
      else if ((a == '>') && (inTag == 1)) inTag = 0;
    } else {
      switch (a) {
        case 'U': case 'u': out[i] = 'T'; break;
        case 'R': case 'r': case 'W': case 'w': case 'M': case 'm': case 'D': case 'd': case 'H': case 'h': case 'V': case 'v': out[i] = 'A'; break;
        case 'Y': case 'y': case 'S': case 's': case 'B': case 'b': out[i] = 'C'; break;
        case 'K': case 'k': out[i] = 'G'; break;
        default: out[i] = 'N';
      }
    }
  }
  return out;
}

class VcfLog {  // the recorder: same text as oracle/ref_wrap7.cpp's stand-ins produce
 public:
  std::string text, cur;
  void header(std::string const& line) { text += "H " + line + "\n"; }
  void sample(std::string const& s) { text += "S " + s + "\n"; }
  void headerWritten() { text += "HW\n"; }
  void begin(int32_t, int64_t) {}
  void str(const char* kind, const char* key, std::string const& v) { cur += std::string(kind) + key + "=" + v + ";"; }
  void ints(const char* kind, const char* key, const int32_t* v, int n) {
    cur += std::string(kind) + key + "=";
    char b[32];
    for (int i = 0; i < n; ++i) { std::snprintf(b, sizeof(b), "%d", v[i]); cur += (i ? "," : ""); cur += b; }
    cur += ";";
  }
  void flt(const char* kind, const char* key, float v) { uint32_t u; std::memcpy(&u, &v, 4); char b[32]; std::snprintf(b, sizeof(b), "f%08x", u); cur += std::string(kind) + key + "=" + b + ";"; }
  void write(int32_t rid, int64_t pos, float qual) {
    uint32_t q; std::memcpy(&q, &qual, 4);
    char b[96]; std::snprintf(b, sizeof(b), "R rid=%d;pos=%lld;qual=f%08x;", rid, (long long) pos, q);
    text += b + cur + "\n";
    cur.clear();
  }
};

// The count maps of one sample (indexed by sv.id), one column of the BCF
struct VcfSample {
  std::string name;
  std::vector<JunctionCount> const* jctMap = nullptr;
  std::vector<ReadCount> const* rcMap = nullptr;
  std::vector<SpanningCount> const* spanMap = nullptr;
  std::vector<MethylInfo> const* methylMap = nullptr;   // may be null / empty: no methylation calls
};

// All samples of a call set. hasVcfFile = genotyping mode (`-v`): SVs without ALT support are kept.
// Sink = where the calls go: VcfLog (the text description the parity tests compare) or the htslib writer of the binding
// (bindings/hts_io.hpp HtsVcfWriter, which turns each call into the bcf_hdr_append / bcf_update_* / bcf_write1 call vcfOutput makes).
template <typename Sink>
inline void vcfRecordsTo(Sink& o, std::vector<StructuralVariantRecord> const& svs, std::vector<VcfSample> const& samples, std::vector<std::string> const& target_name,
                         std::vector<uint32_t> const& target_len, std::string const& genome, std::string const& fileDate, bool hasVcfFile, uint32_t minCpgDepth = 0) {
  static const BoLog bl;
  const std::size_t F = samples.size();
  static const char* fixed1[] = {
      "##ALT=<ID=DEL,Description=\"Deletion\">", "##ALT=<ID=DUP,Description=\"Duplication\">", "##ALT=<ID=INV,Description=\"Inversion\">",
      "##ALT=<ID=BND,Description=\"Translocation\">", "##ALT=<ID=INS,Description=\"Insertion\">",
      "##FILTER=<ID=LowQual,Description=\"Poor quality and insufficient number of PEs and SRs.\">",
      "##INFO=<ID=CIEND,Number=2,Type=Integer,Description=\"PE confidence interval around END\">",
      "##INFO=<ID=CIPOS,Number=2,Type=Integer,Description=\"PE confidence interval around POS\">",
      "##INFO=<ID=CHR2,Number=1,Type=String,Description=\"Chromosome for POS2 coordinate in case of an inter-chromosomal translocation\">",
      "##INFO=<ID=POS2,Number=1,Type=Integer,Description=\"Genomic position for CHR2 in case of an inter-chromosomal translocation\">",
      "##INFO=<ID=END,Number=1,Type=Integer,Description=\"End position of the structural variant\">",
      "##INFO=<ID=PE,Number=1,Type=Integer,Description=\"Paired-end support of the structural variant\">",
      "##INFO=<ID=MAPQ,Number=1,Type=Integer,Description=\"Median mapping quality of paired-ends\">",
      "##INFO=<ID=SRMAPQ,Number=1,Type=Integer,Description=\"Median mapping quality of split-reads\">",
      "##INFO=<ID=SR,Number=1,Type=Integer,Description=\"Split-read support\">",
      "##INFO=<ID=SRQ,Number=1,Type=Float,Description=\"Split-read consensus alignment quality\">",
      "##INFO=<ID=CONSENSUS,Number=1,Type=String,Description=\"Split-read consensus sequence\">",
      "##INFO=<ID=CONSBP,Number=1,Type=Integer,Description=\"Consensus SV breakpoint position\">",
      "##INFO=<ID=CE,Number=1,Type=Float,Description=\"Consensus sequence entropy\">",
      "##INFO=<ID=CT,Number=1,Type=String,Description=\"Paired-end signature induced connection type\">",
      "##INFO=<ID=SVLEN,Number=1,Type=Integer,Description=\"SV length; negative for DEL, positive for DUP/INV/INS.\">",
      "##INFO=<ID=IMPRECISE,Number=0,Type=Flag,Description=\"Imprecise structural variation\">",
      "##INFO=<ID=PRECISE,Number=0,Type=Flag,Description=\"Precise structural variation\">",
      "##INFO=<ID=SVTYPE,Number=1,Type=String,Description=\"Type of structural variant\">",
      "##INFO=<ID=SVMETHOD,Number=1,Type=String,Description=\"Type of approach used to detect SV\">",
      "##INFO=<ID=INSLEN,Number=1,Type=Integer,Description=\"Predicted length of the insertion\">",
      "##INFO=<ID=HOMLEN,Number=1,Type=Integer,Description=\"Breakpoint homology length\">",
      "##INFO=<ID=SUBTYPE,Number=1,Type=String,Description=\"SV subtype: INS:ME:ALU, INS:ME:LINE1, INS:ME:SVA, INS:NUMT, INS:LTR, INS:HERVK, INS:TR, or DEL:TR\">",
      "##INFO=<ID=ALLELEID,Number=1,Type=Integer,Description=\"Identifier of the merged locus\">",
      "##INFO=<ID=NALLELE,Number=1,Type=Integer,Description=\"Number of distinct alleles at this locus\">",
      "##INFO=<ID=AC,Number=A,Type=Integer,Description=\"Allele count\">", "##INFO=<ID=AN,Number=1,Type=Integer,Description=\"Total number of alleles\">",
      "##INFO=<ID=INSSTRAND,Number=1,Type=String,Description=\"Insertion strand for MEIs\">",
      "##INFO=<ID=TRPERIOD,Number=1,Type=Integer,Description=\"Tandem repeat period in bp\">",
      "##INFO=<ID=TRCOPIES,Number=1,Type=Float,Description=\"Tandem repeat copy number\">",
      "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">",
      "##FORMAT=<ID=PL,Number=G,Type=Integer,Description=\"Phred-scaled genotype likelihoods for RR,RA,AA genotypes\">",
      "##FORMAT=<ID=GQ,Number=1,Type=Integer,Description=\"Genotype Quality\">",
      "##FORMAT=<ID=FT,Number=1,Type=String,Description=\"Per-sample genotype filter\">",
      "##FORMAT=<ID=RC,Number=1,Type=Integer,Description=\"Raw high-quality read counts or base counts for the SV\">",
      "##FORMAT=<ID=RCL,Number=1,Type=Integer,Description=\"Raw high-quality read counts or base counts for the left control region\">",
      "##FORMAT=<ID=RCR,Number=1,Type=Integer,Description=\"Raw high-quality read counts or base counts for the right control region\">",
      "##FORMAT=<ID=RDCN,Number=1,Type=Integer,Description=\"Read-depth based copy-number estimate for autosomal sites\">",
      "##FORMAT=<ID=DR,Number=1,Type=Integer,Description=\"# high-quality reference pairs\">",
      "##FORMAT=<ID=DV,Number=1,Type=Integer,Description=\"# high-quality variant pairs\">",
      "##FORMAT=<ID=RR,Number=1,Type=Integer,Description=\"# high-quality reference junction reads\">",
      "##FORMAT=<ID=RV,Number=1,Type=Integer,Description=\"# high-quality variant junction reads\">",
      "##FORMAT=<ID=HP,Number=4,Type=Integer,Description=\"Haplotype-specific junction read counts (HP1_ref,HP1_alt,HP2_ref,HP2_alt)\">",
      "##FORMAT=<ID=PS,Number=1,Type=Integer,Description=\"Phase set identifier from HP-tagged alignments\">",
      "##FORMAT=<ID=MR,Number=4,Type=Integer,Description=\"Methylation % for REF allele [SV start left/right, SV end left/right]\">",
      "##FORMAT=<ID=MA,Number=4,Type=Integer,Description=\"Methylation % for ALT allele [SV start left/right, SV end left/right]\">",
      "##FORMAT=<ID=MNC,Number=4,Type=Integer,Description=\"Unique CpG sites observed per window [SV start left/right, SV end left/right]\">",
      "##FORMAT=<ID=MDV,Number=4,Type=Integer,Description=\"Avg. read depth per CpG site per window [SV start left/right, SV end left/right]\">"};
  o.header("##fileDate=" + fileDate);
  for (const char* h : fixed1) o.header(h);
  o.header("##reference=" + genome);
  for (std::size_t i = 0; i < target_name.size(); ++i) o.header("##contig=<ID=" + target_name[i] + ",length=" + std::to_string(target_len[i]) + ">");
  for (VcfSample const& sm : samples) o.sample(sm.name);
  o.headerWritten();
  std::vector<int32_t> gt(2 * F), gq(F), pl(3 * F), rcl(F), rcc(F), rcr(F), rdcn(F), dr(F), dv(F), rr(F), rv(F), hp(4 * F), ps(F), mr(4 * F), ma(4 * F), mnc(4 * F), mdv(4 * F);
  // the genotype fields of every (SV, sample) up front and in parallel (likelihoods over up to maxGenoReadCount qualities per allele: the bulk of this
  // function's time); the records are then emitted in order
  std::vector<SampleFormat> fmAll(svs.size() * F);
  parallelFor(svs.size(), [&](std::size_t i) {
    StructuralVariantRecord const& sv = svs[i];
    if ((sv.srSupport == 0) && (sv.peSupport == 0)) return;
    for (std::size_t f = 0; f < F; ++f) {
      VcfSample const& sm = samples[f];
      JunctionCount const& jc = (*sm.jctMap)[sv.id];
      SpanningCount const& sc = (*sm.spanMap)[sv.id];
      ReadCount const& rcv = (*sm.rcMap)[sv.id];
      fmAll[i * F + f] = sampleFormat(bl, sv.precise ? jc.ref : sc.ref, sv.precise ? jc.alt : sc.alt, jc.ps, (int32_t) jc.hp1alt.size(), (int32_t) jc.hp2alt.size(),
                                      rcv.leftRC, rcv.rc, rcv.rightRC);
    }
  }, 32);
  for (std::size_t svIdx = 0; svIdx < svs.size(); ++svIdx) {
    StructuralVariantRecord const& sv = svs[svIdx];
    if ((sv.srSupport == 0) && (sv.peSupport == 0)) continue;
    if (!hasVcfFile) {   // discovery mode: at least two supporting reads over all samples after genotyping (:463-472)
      std::size_t totalGtSup = 0;
      for (VcfSample const& sm : samples) totalGtSup += (*sm.spanMap)[sv.id].alt.size() + (*sm.jctMap)[sv.id].alt.size();
      if (totalGtSup < 2) continue;
    }
    int32_t filter = 0;   // PASS
    const int32_t need = (sv.chr == sv.chr2) ? 3 : 5;
    if (((sv.peSupport < need) || (sv.peMapQuality < 20)) && ((sv.srSupport < need) || (sv.srMapQuality < 20))) filter = 1;   // LowQual
    int32_t svStartPos = std::max(sv.svStart - 1, 0);
    int32_t svEndPos = std::max(sv.svEnd, 1);
    if (svEndPos > (int32_t) target_len[sv.chr2]) svEndPos = (int32_t) target_len[sv.chr2];
    std::string pad = std::to_string(sv.id);
    pad.insert(pad.begin(), 8 - pad.length(), '0');
    o.begin(sv.chr, svStartPos);   // rec->rid / rec->pos are set before the first bcf_update_* (htslib derives rlen from pos when END is set)
    o.str("", "ID", _addID(sv.svt) + pad);
    const std::string alleles = _replaceIUPAC(sv.alleles);
    o.str("", "ALLELES", alleles);
    o.ints("", "FILTER", &filter, 1);
    o.str("I:", sv.precise ? "PRECISE" : "IMPRECISE", "1");
    o.str("I:", "SVTYPE", _addID(sv.svt));
    o.str("I:", "SVMETHOD", "EMBL.DELLYv2.5.1");   // src/version.h:8
    int32_t tmpi;
    if (sv.svt < DELLY_SVT_TRANS) {
      const std::size_t commaPos = alleles.find(',');
      bool isSymbolic = (commaPos == std::string::npos);
      if (!isSymbolic) {
        const std::string alt = alleles.substr(commaPos + 1);
        isSymbolic = (!alt.empty()) && ((alt[0] == '<') || (alt.find('[') != std::string::npos) || (alt.find(']') != std::string::npos));
      }
      if (!isSymbolic) tmpi = svStartPos + (int32_t) commaPos;
      else { if (svEndPos < svStartPos + 1) svEndPos = svStartPos + 1; tmpi = svEndPos; }
      o.ints("I:", "END", &tmpi, 1);
    } else {
      tmpi = svStartPos + 1; o.ints("I:", "END", &tmpi, 1);
      o.str("I:", "CHR2", target_name[sv.chr2]);
      tmpi = svEndPos; o.ints("I:", "POS2", &tmpi, 1);
    }
    if (sv.svt == 4) { tmpi = sv.insLen; o.ints("I:", "SVLEN", &tmpi, 1); }
    else if (sv.svt == 2) { tmpi = sv.svStart - sv.svEnd; o.ints("I:", "SVLEN", &tmpi, 1); }
    else if ((sv.svt == 3) || (sv.svt == 0) || (sv.svt == 1)) { tmpi = sv.svEnd - sv.svStart; o.ints("I:", "SVLEN", &tmpi, 1); }
    tmpi = sv.peSupport; o.ints("I:", "PE", &tmpi, 1);
    tmpi = sv.peMapQuality; o.ints("I:", "MAPQ", &tmpi, 1);
    o.str("I:", "CT", _addOrientation(sv.svt));
    const int32_t cipos[2] = {sv.ciposlow, sv.ciposhigh}, ciend[2] = {sv.ciendlow, sv.ciendhigh};
    o.ints("I:", "CIPOS", cipos, 2);
    o.ints("I:", "CIEND", ciend, 2);
    if (sv.alleleid >= 0) { tmpi = sv.alleleid; o.ints("I:", "ALLELEID", &tmpi, 1); tmpi = sv.nallele; o.ints("I:", "NALLELE", &tmpi, 1); }
    if (sv.precise) {
      tmpi = sv.srMapQuality; o.ints("I:", "SRMAPQ", &tmpi, 1);
      tmpi = sv.insLen; o.ints("I:", "INSLEN", &tmpi, 1);
      tmpi = sv.homLen; o.ints("I:", "HOMLEN", &tmpi, 1);
      tmpi = sv.srSupport; o.ints("I:", "SR", &tmpi, 1);
      o.flt("I:", "SRQ", sv.srAlignQuality);
      if (sv.consensus.size()) {
        o.str("I:", "CONSENSUS", sv.consensus);
        o.flt("I:", "CE", (float) entropy(sv.consensus));
        tmpi = sv.consBp; o.ints("I:", "CONSBP", &tmpi, 1);
      }
    }
    if (!_translocation(sv.svt)) {   // reference-based annotation (src/svanno.h results carried in sv.anno)
      if (sv.anno.homLen > 0) { tmpi = sv.anno.homLen; o.ints("I:", "HOMLEN", &tmpi, 1); }
      if (sv.anno.seqType > 0 && sv.anno.seqType < 7) {
        static const char* seqTypeStr[] = {"", "INS:ME:ALU", "INS:ME:LINE1", "INS:ME:SVA", "INS:NUMT", "INS:LTR", "INS:HERVK"};
        o.str("I:", "SUBTYPE", seqTypeStr[sv.anno.seqType]);
        o.str("I:", "INSSTRAND", sv.anno.isRC ? "-" : "+");
      } else if (sv.anno.seqType == 7) {
        o.str("I:", "SUBTYPE", (sv.svt == 4) ? "INS:TR" : "DEL:TR");
        tmpi = sv.anno.trPeriod; o.ints("I:", "TRPERIOD", &tmpi, 1);
        o.flt("I:", "TRCOPIES", sv.anno.trCopies);
      }
    }
    // the samples' FORMAT values (:596-715), one column per sample
    std::string ft;
    int32_t ac = 0, an = 0;
    for (std::size_t f = 0; f < F; ++f) {
      VcfSample const& sm = samples[f];
      JunctionCount const& jc = (*sm.jctMap)[sv.id];
      SpanningCount const& sc = (*sm.spanMap)[sv.id];
      ReadCount const& rcv = (*sm.rcMap)[sv.id];
      SampleFormat const& fm = fmAll[svIdx * F + f];
      for (int k = 0; k < 2; ++k) { gt[2 * f + k] = fm.gt[k]; if ((fm.gt[k] >> 1) == 0) continue; ++an; if (((fm.gt[k] >> 1) - 1) > 0) ++ac; }
      gq[f] = fm.gq;
      for (int k = 0; k < 3; ++k) pl[3 * f + k] = fm.pl[k];
      ft += (f ? "," : ""); ft += fm.pass ? "PASS" : "LowQual";
      rcl[f] = rcv.leftRC; rcc[f] = rcv.rc; rcr[f] = rcv.rightRC; rdcn[f] = fm.rcn;
      dr[f] = (int32_t) sc.ref.size(); dv[f] = (int32_t) sc.alt.size(); rr[f] = (int32_t) jc.ref.size(); rv[f] = (int32_t) jc.alt.size();
      hp[4 * f] = (int32_t) jc.hp1ref.size(); hp[4 * f + 1] = (int32_t) jc.hp1alt.size(); hp[4 * f + 2] = (int32_t) jc.hp2ref.size(); hp[4 * f + 3] = (int32_t) jc.hp2alt.size();
      ps[f] = jc.ps;
      // src/modvcf.h:622-665: missing unless this sample has methylation calls for the SV
      if (sm.methylMap && !sm.methylMap->empty() && sv.id < (int32_t) sm.methylMap->size())
        methylFormat((*sm.methylMap)[sv.id], sv.svt, minCpgDepth, &ma[4 * f], &mr[4 * f], &mnc[4 * f], &mdv[4 * f]);
      else for (int k = 0; k < 4; ++k) ma[4 * f + k] = mr[4 * f + k] = mnc[4 * f + k] = mdv[4 * f + k] = INT32_MISSING;
    }
    const int32_t qual = std::min(std::max(sv.mapq, 0), 10000);
    const int nF = (int) F;
    o.ints("I:", "AC", &ac, 1);
    o.ints("I:", "AN", &an, 1);
    o.ints("F:", "GT", gt.data(), 2 * nF);
    o.ints("F:", "GQ", gq.data(), nF);
    o.ints("F:", "PL", pl.data(), 3 * nF);
    o.str("F:", "FT", ft);
    o.ints("F:", "RCL", rcl.data(), nF);
    o.ints("F:", "RC", rcc.data(), nF);
    o.ints("F:", "RCR", rcr.data(), nF);
    o.ints("F:", "RDCN", rdcn.data(), nF);
    o.ints("F:", "DR", dr.data(), nF); o.ints("F:", "DV", dv.data(), nF); o.ints("F:", "RR", rr.data(), nF); o.ints("F:", "RV", rv.data(), nF);
    o.ints("F:", "HP", hp.data(), 4 * nF);
    o.ints("F:", "PS", ps.data(), nF);
    o.ints("F:", "MR", mr.data(), 4 * nF); o.ints("F:", "MA", ma.data(), 4 * nF); o.ints("F:", "MNC", mnc.data(), 4 * nF); o.ints("F:", "MDV", mdv.data(), 4 * nF);
    o.write(sv.chr, svStartPos, (float) qual);
  }
}

inline std::string vcfRecords(std::vector<StructuralVariantRecord> const& svs, std::vector<VcfSample> const& samples, std::vector<std::string> const& target_name,
                              std::vector<uint32_t> const& target_len, std::string const& genome, std::string const& fileDate, bool hasVcfFile,
                              uint32_t minCpgDepth = 0) {
  VcfLog o;
  vcfRecordsTo(o, svs, samples, target_name, target_len, genome, fileDate, hasVcfFile, minCpgDepth);
  return o.text;
}

// One sample (jctMap / rcMap / spanMap indexed by sv.id)
inline std::string vcfRecords(std::vector<StructuralVariantRecord> const& svs, std::vector<JunctionCount> const& jctMap, std::vector<ReadCount> const& rcMap,
                              std::vector<SpanningCount> const& spanMap, std::vector<std::string> const& target_name, std::vector<uint32_t> const& target_len,
                              std::string const& sampleName, std::string const& genome, std::string const& fileDate, bool hasVcfFile,
                              std::vector<MethylInfo> const* methylMap = nullptr, uint32_t minCpgDepth = 0) {
  VcfSample one;
  one.name = sampleName; one.jctMap = &jctMap; one.rcMap = &rcMap; one.spanMap = &spanMap; one.methylMap = methylMap;
  return vcfRecords(svs, std::vector<VcfSample>(1, one), target_name, target_len, genome, fileDate, hasVcfFile, minCpgDepth);
}

}  // namespace dellyb200
