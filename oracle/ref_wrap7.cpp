// oracle/_ref wrapper, part 7 (TEST INFRASTRUCTURE ONLY): the reference's BCF record construction (vcfOutput,
// src/modvcf.h:344-791) compiled VERBATIM, with util.h compiled itself (as in ref_wrap6.cpp). The htslib VCF/BCF calls it
// makes are replaced by a RECORDING stand-in: every header line, every field update (ID, alleles, filter, INFO, FORMAT,
// genotypes) and every record write is appended to a text log in call order (floats as their bit patterns), so what the
// reference would hand to htslib for serialisation can be compared field for field. Nothing from the reference is copied.
#define PANGENOME_H
#define ORACLE_REAL_UTIL_H
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <cstdlib>
#include <iostream>
#include <fstream>
#include <limits>
#include <numeric>
#include <map>
#include <set>
#include <sstream>
#include <unordered_map>
#include <unordered_set>
#include <boost/filesystem.hpp>
#include "shim/prelude.h"

namespace boost {
template <typename K, typename V> using unordered_map = std::unordered_map<K, V>;
namespace gregorian { struct date {}; inline std::string to_iso_string(date const&) { return "00000000"; } }
namespace posix_time {
struct ptime { gregorian::date date() const { return gregorian::date(); } };
struct second_clock { static ptime local_time() { return ptime(); } };
inline std::string to_simple_string(ptime const&) { return "now"; }
}  // namespace posix_time
}  // namespace boost

#include <htslib/faidx.h>
#include <htslib/vcf.h>
#include <htslib/sam.h>
#include "version.h"
#include "util.h"
#include "tags.h"
#include "threadpool.h"
#define MAX_CN 10
#include "coverage.h"
#include "modvcf.h"

namespace {
struct RefConfig7 {
  std::vector<boost::filesystem::path> files;
  boost::filesystem::path genome, outfile;
  std::vector<std::string> sampleName;
  bool hasVcfFile = false;
  uint32_t minCpgDepth = 1;
  boost::filesystem::path vcffile;
};
// ---- the in-memory site list vcfParse reads (ref_vcf_parse) ----
struct MemSite {
  int32_t rid = 0, pos0 = 0; float qual = 0; uint32_t mask = 0; bool precise = false;
  int32_t iv[21] = {0};          // the integer / float-bit INFO values, indexed like the hook's row
  std::string str[7];            // ref, alt, SVMETHOD, SVTYPE, CT, CHR2, CONSENSUS
};
std::vector<MemSite> g_sites;
std::size_t g_site_next = 0;
MemSite const* g_site_cur = nullptr;
bool g_hasConsBp = true;
char* g_allele_ptrs[2];
std::vector<uint32_t> g_tlen;
std::vector<std::string> g_names;
std::vector<char*> g_name_ptrs;
std::string g_log;          // everything vcfOutput hands to htslib, in call order
std::string g_cur;          // the record under construction
void put_vals(std::string& o, const void* values, int n, int type) {
  char buf[64];
  for (int i = 0; i < n; ++i) {
    if (type == BCF_HT_INT) snprintf(buf, sizeof(buf), "%d", ((const int32_t*) values)[i]);
    else if (type == BCF_HT_REAL) { uint32_t b; memcpy(&b, (const float*) values + i, 4); snprintf(buf, sizeof(buf), "f%08x", b); }
    else buf[0] = 0;
    o += (i ? "," : ""); o += buf;
  }
}
}  // namespace

extern "C" {

// ---- BAM side: only the header is read by vcfOutput ----
htsFile* hts_open(const char*, const char*) { htsFile* f = (htsFile*) calloc(1, sizeof(htsFile)); f->is_bgzf = 1; return f; }
int hts_close(htsFile* f) { free(f); return 0; }
int hts_set_fai_filename(htsFile*, const char*) { return 0; }
sam_hdr_t* sam_hdr_read(samFile*) {
  sam_hdr_t* h = (sam_hdr_t*) calloc(1, sizeof(sam_hdr_t));
  h->n_targets = (int32_t) g_tlen.size(); h->target_len = g_tlen.data(); h->target_name = g_name_ptrs.data();
  return h;
}
void sam_hdr_destroy(sam_hdr_t* h) { free(h); }
void hts_log(enum htsLogLevel, const char*, const char*, ...) {}
uint8_t* bam_aux_get(const bam1_t*, const char[2]) { return NULL; }
int64_t bam_aux2i(const uint8_t*) { return 0; }
hts_idx_t* sam_index_load(htsFile*, const char*) { return NULL; }
void hts_idx_destroy(hts_idx_t*) {}
int hts_idx_get_stat(const hts_idx_t*, int, uint64_t* m, uint64_t* u) { *m = 0; *u = 0; return 0; }
hts_itr_t* sam_itr_queryi(const hts_idx_t*, int, hts_pos_t, hts_pos_t) { return NULL; }
int hts_itr_next(BGZF*, hts_itr_t*, void*, void*) { return -1; }
int hts_itr_multi_next(htsFile*, hts_itr_t*, void*) { return -1; }
void hts_itr_destroy(hts_itr_t*) {}
bam1_t* bam_init1(void) { return (bam1_t*) calloc(1, sizeof(bam1_t)); }
void bam_destroy1(bam1_t* b) { free(b); }
faidx_t* fai_load(const char*) { return NULL; }
void fai_destroy(faidx_t*) {}
char* faidx_fetch_seq(const faidx_t*, const char*, int, int, int* len) { *len = 0; return NULL; }

// ---- the recording VCF/BCF stand-in ----
bcf_hdr_t* bcf_hdr_init(const char*) { return (bcf_hdr_t*) calloc(1, sizeof(bcf_hdr_t)); }
void bcf_hdr_destroy(bcf_hdr_t* h) { free(h); }
int bcf_hdr_append(bcf_hdr_t*, const char* line) { g_log += "H "; g_log += line; g_log += "\n"; return 0; }
int bcf_hdr_add_sample(bcf_hdr_t* h, const char* s) { if (s) { g_log += "S "; g_log += s; g_log += "\n"; h->n[BCF_DT_SAMPLE]++; } return 0; }
int bcf_hdr_write(htsFile*, bcf_hdr_t*) { g_log += "HW\n"; return 0; }
int bcf_hdr_id2int(const bcf_hdr_t*, int type, const char* id) {
  if (type == BCF_DT_CTG) { for (std::size_t k = 0; k < g_names.size(); ++k) if (g_names[k] == id) return (int) k; return -1; }
  if (!strcmp(id, "CONSBP") && !g_hasConsBp) return -1;
  if (!strcmp(id, "PASS")) return 0;
  if (!strcmp(id, "LowQual")) return 1;
  return 2;
}
bcf1_t* bcf_init(void) { return (bcf1_t*) calloc(1, sizeof(bcf1_t)); }
void bcf_destroy(bcf1_t* v) { free(v); }
void bcf_clear(bcf1_t* v) { memset(v, 0, sizeof(bcf1_t)); g_cur.clear(); }
int bcf_update_id(const bcf_hdr_t*, bcf1_t*, const char* id) { g_cur += "ID="; g_cur += id; g_cur += ";"; return 0; }
int bcf_update_alleles_str(const bcf_hdr_t*, bcf1_t*, const char* a) { g_cur += "ALLELES="; g_cur += a; g_cur += ";"; return 0; }
int bcf_update_filter(const bcf_hdr_t*, bcf1_t*, int* flt, int n) { g_cur += "FILTER="; put_vals(g_cur, flt, n, BCF_HT_INT); g_cur += ";"; return 0; }
int bcf_update_info(const bcf_hdr_t*, bcf1_t*, const char* key, const void* values, int n, int type) {
  g_cur += "I:"; g_cur += key; g_cur += "=";
  if (type == BCF_HT_STR) g_cur += (const char*) values;
  else if (type == BCF_HT_FLAG) g_cur += "1";
  else put_vals(g_cur, values, n, type);
  g_cur += ";";
  return 0;
}
int bcf_update_format(const bcf_hdr_t*, bcf1_t*, const char* key, const void* values, int n, int type) {
  g_cur += "F:"; g_cur += key; g_cur += "="; put_vals(g_cur, values, n, type); g_cur += ";";
  return 0;
}
int bcf_update_format_string(const bcf_hdr_t*, bcf1_t*, const char* key, const char** values, int n) {
  g_cur += "F:"; g_cur += key; g_cur += "=";
  for (int i = 0; i < n; ++i) { g_cur += (i ? "," : ""); g_cur += values[i]; }
  g_cur += ";";
  return 0;
}
int bcf_write(htsFile*, bcf_hdr_t*, bcf1_t* v) {
  char buf[96];
  uint32_t qb; memcpy(&qb, &v->qual, 4);
  snprintf(buf, sizeof(buf), "R rid=%d;pos=%lld;qual=f%08x;", v->rid, (long long) v->pos, qb);
  g_log += buf; g_log += g_cur; g_log += "\n";
  return 0;
}
int bcf_index_build(const char*, int) { return 0; }
int bcf_unpack(bcf1_t*, int) { return 0; }
uint32_t bcf_float_missing = 0x7F800001;      // htslib vcf.c
uint32_t bcf_float_vector_end = 0x7F800002;

// vcfOutput over in-memory SV records and count maps of ONE sample.
//   sv: n x 25 [chr,svStart,chr2,svEnd,ciposlow,ciposhigh,ciendlow,ciendhigh,peSupport,srSupport,peMapQuality,srMapQuality,mapq,insLen,homLen,svt,precise,consBp,id,
//              srAlignQuality bits, alleleid, nallele, anno.homLen, anno.seqType, anno.isRC] + anno.trPeriod / trCopies not used (0);
//   alleles / consensus as strings (stride + lengths); counts: per SV jct ref/alt, span ref/alt lists (qualities), hp1ref/hp1alt/hp2ref/hp2alt sizes, ps, rc triple.
// Returns the log length (copied into out, truncated to cap).
// The _ex form adds anno.trPeriod / anno.trCopies (anno_tr: n x 2, copies as float bits) and the per-SV MethylInfo of the sample
// (methyl16: n x 16 in field order alt x4, ref x4, mnc x4, mdp x4; NULL = the empty methylMap of the short-read path) with its depth gate.
static int run_vcf_output(const uint32_t* contig_len, int ncontig, const int32_t* sv25, int n, const char* alleles, int astride, const int32_t* alen, const char* cons,
                   int cstride, const int32_t* clen, const uint8_t* quals, const uint32_t* jr_off, const uint32_t* ja_off, const uint32_t* sr_off,
                   const uint32_t* sa_off, const int32_t* hp5, const int32_t* rc3, int hasVcfFile, char* out, int cap, const int32_t* anno_tr,
                   const int32_t* methyl16, int minCpgDepth, int nfile) {
  RefConfig7 c;
  c.minCpgDepth = (uint32_t) minCpgDepth; c.genome = boost::filesystem::path("in-memory.fa");
  for (int f = 0; f < nfile; ++f) { c.files.push_back(boost::filesystem::path("in-memory." + std::to_string(f) + ".bam")); c.sampleName.push_back("sample" + std::to_string(f + 1)); }
  c.outfile = boost::filesystem::path("-"); c.hasVcfFile = hasVcfFile != 0;
  g_names.clear(); g_tlen.clear(); g_name_ptrs.clear();
  for (int k = 0; k < ncontig; ++k) { g_names.push_back("chr" + std::to_string(k)); g_tlen.push_back(contig_len[k]); }
  for (auto& nm : g_names) g_name_ptrs.push_back((char*) nm.c_str());
  std::vector<torali::StructuralVariantRecord> svs(n);
  std::vector<std::vector<torali::JunctionCount> > jct((size_t) nfile, std::vector<torali::JunctionCount>(n));
  std::vector<std::vector<torali::SpanningCount> > span((size_t) nfile, std::vector<torali::SpanningCount>(n));
  std::vector<std::vector<torali::ReadCount> > rcm((size_t) nfile, std::vector<torali::ReadCount>(n));
  for (int i = 0; i < n; ++i) {
    const int32_t* r = sv25 + 25 * i; torali::StructuralVariantRecord& v = svs[i];
    v.chr = r[0]; v.svStart = r[1]; v.chr2 = r[2]; v.svEnd = r[3]; v.ciposlow = r[4]; v.ciposhigh = r[5]; v.ciendlow = r[6]; v.ciendhigh = r[7];
    v.peSupport = r[8]; v.srSupport = r[9]; v.peMapQuality = r[10]; v.srMapQuality = r[11]; v.mapq = r[12]; v.insLen = r[13]; v.homLen = r[14]; v.svt = r[15];
    v.precise = r[16] != 0; v.consBp = r[17]; v.id = r[18]; memcpy(&v.srAlignQuality, r + 19, 4); v.alleleid = r[20]; v.nallele = r[21];
    v.anno.homLen = r[22]; v.anno.seqType = r[23]; v.anno.isRC = r[24] != 0;
    if (anno_tr) { v.anno.trPeriod = anno_tr[2 * i]; memcpy(&v.anno.trCopies, anno_tr + 2 * i + 1, 4); }
    v.alleles = std::string(alleles + (size_t) i * astride, alen[i]); v.consensus = std::string(cons + (size_t) i * cstride, clen[i]);
    const int id = v.id;
    for (int f = 0; f < nfile; ++f) {   // per-sample arrays are file-major: offsets (n + 1) per file, hp5 5n, rc3 3n
      const uint32_t* jr = jr_off + (size_t) f * (n + 1); const uint32_t* ja = ja_off + (size_t) f * (n + 1);
      const uint32_t* sr = sr_off + (size_t) f * (n + 1); const uint32_t* sa = sa_off + (size_t) f * (n + 1);
      const int32_t* hp = hp5 + (size_t) f * 5 * n; const int32_t* rc = rc3 + (size_t) f * 3 * n;
      jct[f][id].ref.assign(quals + jr[i], quals + jr[i + 1]); jct[f][id].alt.assign(quals + ja[i], quals + ja[i + 1]);
      span[f][id].ref.assign(quals + sr[i], quals + sr[i + 1]); span[f][id].alt.assign(quals + sa[i], quals + sa[i + 1]);
      jct[f][id].hp1ref.assign(hp[5 * i], 30); jct[f][id].hp1alt.assign(hp[5 * i + 1], 30); jct[f][id].hp2ref.assign(hp[5 * i + 2], 30); jct[f][id].hp2alt.assign(hp[5 * i + 3], 30);
      jct[f][id].ps = hp[5 * i + 4];
      rcm[f][id] = torali::ReadCount(rc[3 * i], rc[3 * i + 1], rc[3 * i + 2]);
    }
  }
  g_log.clear(); g_cur.clear();
  std::streambuf* old = std::cerr.rdbuf(nullptr);
  if (methyl16) {
    std::vector<std::vector<torali::MethylInfo> > methylMap((size_t) nfile, std::vector<torali::MethylInfo>(n));
    for (int f = 0; f < nfile; ++f)
    for (int i = 0; i < n; ++i) {
      const int32_t* m = methyl16 + 16 * ((size_t) f * n + i); torali::MethylInfo& mi = methylMap[f][svs[i].id];
      mi.altSvStartL = m[0]; mi.altSvStartR = m[1]; mi.altSvRightL = m[2]; mi.altSvRightR = m[3];
      mi.refSvStartL = m[4]; mi.refSvStartR = m[5]; mi.refSvRightL = m[6]; mi.refSvRightR = m[7];
      mi.mncStartL = m[8]; mi.mncStartR = m[9]; mi.mncRightL = m[10]; mi.mncRightR = m[11];
      mi.mdpStartL = m[12]; mi.mdpStartR = m[13]; mi.mdpRightL = m[14]; mi.mdpRightR = m[15];
    }
    torali::vcfOutput(c, svs, jct, rcm, span, methylMap);
  } else {
    torali::vcfOutput(c, svs, jct, rcm, span);
  }
  std::cerr.rdbuf(old);
  const int L = (int) g_log.size();
  memcpy(out, g_log.data(), (size_t) std::min(L, cap));
  return L;
}

// ---- BCF reading side (vcfParse, src/modvcf.h:156-339): records served from g_sites ----
int sam_hdr_name2tid(sam_hdr_t*, const char* ref) {
  for (std::size_t k = 0; k < g_names.size(); ++k) if (g_names[k] == ref) return (int) k;
  return -1;
}
bcf_hdr_t* bcf_hdr_read(htsFile*) {
  bcf_hdr_t* h = (bcf_hdr_t*) calloc(1, sizeof(bcf_hdr_t));
  h->n[BCF_DT_CTG] = (int) g_names.size();
  h->id[BCF_DT_CTG] = (bcf_idpair_t*) calloc(g_names.size() + 1, sizeof(bcf_idpair_t));
  for (std::size_t k = 0; k < g_names.size(); ++k) h->id[BCF_DT_CTG][k].key = g_names[k].c_str();
  return h;
}
int bcf_read(htsFile*, const bcf_hdr_t*, bcf1_t* v) {
  if (g_site_next >= g_sites.size()) return -1;
  g_site_cur = &g_sites[g_site_next++];
  v->rid = g_site_cur->rid; v->pos = g_site_cur->pos0; v->qual = g_site_cur->qual; v->n_allele = 2;
  g_allele_ptrs[0] = (char*) g_site_cur->str[0].c_str(); g_allele_ptrs[1] = (char*) g_site_cur->str[1].c_str();
  v->d.allele = g_allele_ptrs;
  return 0;
}
// tag -> (presence bit, value slot / string slot); layout of the hook's site row (see ref_vcf_parse)
int bcf_get_info_values(const bcf_hdr_t*, bcf1_t*, const char* tag, void** dst, int* ndst, int type) {
  static const struct { const char* tag; int bit; int slot; int n; } ints[] = {
      {"PE", 3, 4, 1}, {"INSLEN", 4, 5, 1}, {"SVLEN", 5, 6, 1}, {"HOMLEN", 6, 7, 1}, {"SR", 7, 8, 1}, {"END", 8, 9, 1}, {"POS2", 10, 10, 1}, {"CONSBP", 12, 11, 1},
      {"CIPOS", 13, 12, 2}, {"CIEND", 14, 14, 2}, {"MAPQ", 15, 16, 1}, {"SRMAPQ", 16, 17, 1}, {"SRQ", 17, 18, 1}, {"ALLELEID", 18, 19, 1}, {"NALLELE", 19, 20, 1}};
  static const struct { const char* tag; int bit; int slot; } strs[] = {{"SVMETHOD", 0, 2}, {"SVTYPE", 1, 3}, {"CT", 2, 4}, {"CHR2", 9, 5}, {"CONSENSUS", 11, 6}};
  MemSite const& m = *g_site_cur;
  if (type == BCF_HT_FLAG) return (!strcmp(tag, "PRECISE") && m.precise) ? 1 : 0;
  if (type == BCF_HT_STR) {
    for (auto const& e : strs) if (!strcmp(tag, e.tag)) {
      if (!(m.mask & (1u << e.bit))) return -3;
      std::string const& v = m.str[e.slot];
      if (*ndst < (int) v.size() + 1) { *ndst = (int) v.size() + 1; *dst = realloc(*dst, (size_t) *ndst); }
      memcpy(*dst, v.c_str(), v.size() + 1);
      return (int) v.size();
    }
    return -1;
  }
  for (auto const& e : ints) if (!strcmp(tag, e.tag)) {
    if (!(m.mask & (1u << e.bit))) return -3;
    if (*ndst < e.n) { *ndst = e.n; *dst = realloc(*dst, (size_t) e.n * 4); }
    memcpy(*dst, m.iv + e.slot, (size_t) e.n * 4);   // int32 values and float bit patterns alike
    return e.n;
  }
  return -1;
}

// vcfParse over an in-memory site list.
//   site: nsite x 22 int32 [rid, pos0, qual bits, presence mask, PE, INSLEN, SVLEN, HOMLEN, SR, END, POS2, CONSBP, CIPOS x2, CIEND x2, MAPQ, SRMAPQ, SRQ bits, ALLELEID,
//         NALLELE, PRECISE flag]; presence bits: 0 SVMETHOD 1 SVTYPE 2 CT 3 PE 4 INSLEN 5 SVLEN 6 HOMLEN 7 SR 8 END 9 CHR2 10 POS2 11 CONSENSUS 12 CONSBP 13 CIPOS 14 CIEND
//         15 MAPQ 16 SRMAPQ 17 SRQ 18 ALLELEID 19 NALLELE; strings: 7 per site (ref, alt, SVMETHOD, SVTYPE, CT, CHR2, CONSENSUS) in one arena, str_off[7 * nsite + 1]
//   out: sv rows x 22 [chr, svStart, chr2, svEnd, ciposlow, ciposhigh, ciendlow, ciendhigh, peSupport, srSupport, peMapQuality, srMapQuality, mapq, insLen, homLen, svt,
//        precise, consBp, id, srAlignQuality bits, alleleid, nallele] + alleles / consensus (fixed strides). Returns the number of SV records.
int ref_vcf_parse(int ncontig, const int32_t* site22, int nsite, const char* strs, const uint32_t* str_off, int headerHasConsBp, int32_t* sv_out, int cap,
                  char* alleles_out, int astride, int32_t* alen, char* cons_out, int cstride, int32_t* clen) {
  RefConfig7 c; c.vcffile = boost::filesystem::path("in-memory.bcf");
  g_names.clear(); g_tlen.clear(); g_name_ptrs.clear();
  for (int k = 0; k < ncontig; ++k) { g_names.push_back("chr" + std::to_string(k)); g_tlen.push_back(1000000); }
  for (auto& nm : g_names) g_name_ptrs.push_back((char*) nm.c_str());
  g_sites.assign((size_t) nsite, MemSite());
  for (int i = 0; i < nsite; ++i) {
    const int32_t* r = site22 + 22 * i; MemSite& m = g_sites[i];
    m.rid = r[0]; m.pos0 = r[1]; memcpy(&m.qual, r + 2, 4); m.mask = (uint32_t) r[3]; m.precise = r[21] != 0;
    for (int k = 4; k <= 20; ++k) m.iv[k] = r[k];
    for (int k = 0; k < 7; ++k) m.str[k].assign(strs + str_off[7 * i + k], strs + str_off[7 * i + k + 1]);
  }
  g_site_next = 0; g_site_cur = nullptr; g_hasConsBp = headerHasConsBp != 0;
  sam_hdr_t hd; memset(&hd, 0, sizeof(hd));
  std::vector<torali::StructuralVariantRecord> svs;
  std::streambuf* old = std::cerr.rdbuf(nullptr);
  torali::vcfParse(c, &hd, svs);
  std::cerr.rdbuf(old);
  g_hasConsBp = true;
  const int n = (int) svs.size();
  if (n > cap) return -1;
  for (int i = 0; i < n; ++i) {
    torali::StructuralVariantRecord const& v = svs[i];
    int32_t* o = sv_out + 22 * i;
    o[0] = v.chr; o[1] = v.svStart; o[2] = v.chr2; o[3] = v.svEnd; o[4] = v.ciposlow; o[5] = v.ciposhigh; o[6] = v.ciendlow; o[7] = v.ciendhigh;
    o[8] = v.peSupport; o[9] = v.srSupport; o[10] = v.peMapQuality; o[11] = v.srMapQuality; o[12] = v.mapq; o[13] = v.insLen; o[14] = v.homLen; o[15] = v.svt;
    o[16] = v.precise ? 1 : 0; o[17] = v.consBp; o[18] = v.id; memcpy(o + 19, &v.srAlignQuality, 4); o[20] = v.alleleid; o[21] = v.nallele;
    alen[i] = (int32_t) v.alleles.size(); memcpy(alleles_out + (size_t) i * astride, v.alleles.data(), std::min<size_t>(v.alleles.size(), astride));
    clen[i] = (int32_t) v.consensus.size(); memcpy(cons_out + (size_t) i * cstride, v.consensus.data(), std::min<size_t>(v.consensus.size(), cstride));
  }
  return n;
}

int ref_vcf_output(const uint32_t* contig_len, int ncontig, const int32_t* sv25, int n, const char* alleles, int astride, const int32_t* alen, const char* cons,
                   int cstride, const int32_t* clen, const uint8_t* quals, const uint32_t* jr_off, const uint32_t* ja_off, const uint32_t* sr_off,
                   const uint32_t* sa_off, const int32_t* hp5, const int32_t* rc3, int hasVcfFile, char* out, int cap) {
  return run_vcf_output(contig_len, ncontig, sv25, n, alleles, astride, alen, cons, cstride, clen, quals, jr_off, ja_off, sr_off, sa_off, hp5, rc3, hasVcfFile, out,
                        cap, nullptr, nullptr, 1, 1);
}

int ref_vcf_output_ex(const uint32_t* contig_len, int ncontig, const int32_t* sv25, int n, const char* alleles, int astride, const int32_t* alen, const char* cons,
                      int cstride, const int32_t* clen, const uint8_t* quals, const uint32_t* jr_off, const uint32_t* ja_off, const uint32_t* sr_off,
                      const uint32_t* sa_off, const int32_t* hp5, const int32_t* rc3, int hasVcfFile, char* out, int cap, const int32_t* anno_tr,
                      const int32_t* methyl16, int minCpgDepth) {
  return run_vcf_output(contig_len, ncontig, sv25, n, alleles, astride, alen, cons, cstride, clen, quals, jr_off, ja_off, sr_off, sa_off, hp5, rc3, hasVcfFile, out,
                        cap, anno_tr, methyl16, minCpgDepth, 1);
}

// several samples: the per-sample arrays (offsets, hp5, rc3, methyl16) are file-major; sample names "sample1", "sample2", ...
int ref_vcf_output_multi(const uint32_t* contig_len, int ncontig, const int32_t* sv25, int n, const char* alleles, int astride, const int32_t* alen, const char* cons,
                         int cstride, const int32_t* clen, const uint8_t* quals, const uint32_t* jr_off, const uint32_t* ja_off, const uint32_t* sr_off,
                         const uint32_t* sa_off, const int32_t* hp5, const int32_t* rc3, int hasVcfFile, char* out, int cap, const int32_t* anno_tr,
                         const int32_t* methyl16, int minCpgDepth, int nfile) {
  return run_vcf_output(contig_len, ncontig, sv25, n, alleles, astride, alen, cons, cstride, clen, quals, jr_off, ja_off, sr_off, sa_off, hp5, rc3, hasVcfFile, out,
                        cap, anno_tr, methyl16, minCpgDepth, nfile);
}

}  // extern "C"
