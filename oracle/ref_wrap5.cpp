// oracle/_ref wrapper, part 5 (TEST INFRASTRUCTURE ONLY): the reference's split-read assembly stage
// (assembleSplitReads, src/shortpe.h:48-282: per-SV read collection, msa(), alignConsensus(), support / quality fields)
// compiled VERBATIM from /root/reference/src and run over an in-memory alignment list and in-memory contigs
// (htslib stand-ins as in ref_wrap3/4.cpp). util.h / pangenome.h are masked by their include guards; the util.h symbols
// the included headers name are restated with their reference lines. Nothing from the reference is copied into this repository.
#define UTIL_H
#define PANGENOME_H
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <iostream>
#include <fstream>
#include <limits>
#include <numeric>
#include <map>
#include <set>
#include <unordered_map>
#include <unordered_set>
#include "shim/prelude.h"

namespace boost {
template <typename K, typename V> using unordered_map = std::unordered_map<K, V>;
namespace posix_time {
struct ptime {};
struct second_clock { static ptime local_time() { return ptime(); } };
inline std::string to_simple_string(ptime const&) { return "now"; }
}  // namespace posix_time
namespace filesystem {
struct path {
  std::string s;
  path() {}
  path(std::string const& x) : s(x) {}
  std::string const& string() const { return s; }
  void clear() { s.clear(); }
};
inline std::ostream& operator<<(std::ostream& o, path const& p) { return o << p.s; }
}  // namespace filesystem
// boost::split(container, string, is_any_of(chars)) as src/methyl.h uses it on the MM tag (no token compression)
struct AnyOf { std::string chars; };
inline AnyOf is_any_of(const char* c) { return AnyOf{c}; }
template <typename TCont> inline void split(TCont& out, std::string const& in, AnyOf const& sep) {
  out.clear();
  std::string cur;
  for (char ch : in) { if (sep.chars.find(ch) != std::string::npos) { out.push_back(cur); cur.clear(); } else cur += ch; }
  out.push_back(cur);
}
}  // namespace boost

#include <htslib/faidx.h>
#include <htslib/vcf.h>
#include <htslib/sam.h>
#include "tags.h"
#include "threadpool.h"

namespace torali {
// util.h:69-76
struct ReadCount {
  int32_t leftRC, rc, rightRC;
  ReadCount() : leftRC(0), rc(0), rightRC(0) {}
  ReadCount(int32_t l, int32_t m, int32_t r) : leftRC(l), rc(m), rightRC(r) {}
};
// util.h:78-84
template <typename TValue> inline TValue medianVector(std::vector<TValue>& v) {
  std::size_t n = v.size() / 2;
  std::nth_element(v.begin(), v.begin() + n, v.end());
  return v[n];
}
// util.h:29-41
struct LibraryInfo {
  int32_t rs, median, mad, minNormalISize, minISizeCutoff, maxNormalISize, maxISizeCutoff;
  uint32_t abnormal_pairs;
  LibraryInfo() : rs(0), median(0), mad(0), minNormalISize(0), minISizeCutoff(0), maxNormalISize(0), maxISizeCutoff(0), abnormal_pairs(0) {}
};
// util.h:430-451
inline uint32_t readLength(bam1_t const* rec) {
  uint32_t const* cigar = bam_get_cigar(rec);
  uint32_t slen = 0;
  for (uint32_t i = 0; i < rec->core.n_cigar; ++i)
    if ((bam_cigar_op(cigar[i]) == BAM_CMATCH) || (bam_cigar_op(cigar[i]) == BAM_CEQUAL) || (bam_cigar_op(cigar[i]) == BAM_CDIFF) ||
        (bam_cigar_op(cigar[i]) == BAM_CINS) || (bam_cigar_op(cigar[i]) == BAM_CSOFT_CLIP) || (bam_cigar_op(cigar[i]) == BAM_CHARD_CLIP))
      slen += bam_cigar_oplen(cigar[i]);
  return slen;
}
inline uint32_t alignmentLength(bam1_t const* rec) {
  uint32_t const* cigar = bam_get_cigar(rec);
  uint32_t alen = 0;
  for (std::size_t i = 0; i < rec->core.n_cigar; ++i)
    if ((bam_cigar_op(cigar[i]) == BAM_CMATCH) || (bam_cigar_op(cigar[i]) == BAM_CEQUAL) || (bam_cigar_op(cigar[i]) == BAM_CDIFF) || (bam_cigar_op(cigar[i]) == BAM_CDEL) || (bam_cigar_op(cigar[i]) == BAM_CREF_SKIP)) alen += bam_cigar_oplen(cigar[i]);
  return alen;
}
inline uint32_t halfAlignmentLength(bam1_t const* rec) { return (alignmentLength(rec) / 2); }
// opaque read / fragment ids (util.h:501-542 use boost::hash; the stage tested here only needs them to be ids; the
// "read 2 = read 1 + 1" property of hash_sr, util.h:525, is kept)
inline std::size_t _mix(std::size_t seed, std::size_t v) { return seed ^ (v + 0x9e3779b97f4a7c15ull + (seed << 6) + (seed >> 2)); }
inline std::size_t hash_lr(bam1_t* rec) { return std::hash<std::string>()(bam_get_qname(rec)); }
inline std::size_t hash_sr(bam1_t* rec) { return std::hash<std::string>()(bam_get_qname(rec)) * 2 + ((rec->core.flag & BAM_FREAD2) ? 1 : 0); }
inline std::size_t hash_pair(bam1_t* rec) {
  std::size_t seed = hash_string(bam_get_qname(rec));
  seed = _mix(seed, (std::size_t) rec->core.tid); seed = _mix(seed, (std::size_t) rec->core.pos);
  seed = _mix(seed, (std::size_t) rec->core.mtid); seed = _mix(seed, (std::size_t) rec->core.mpos);
  return seed;
}
inline std::size_t hash_pair_mate(bam1_t* rec) {
  std::size_t seed = hash_string(bam_get_qname(rec));
  seed = _mix(seed, (std::size_t) rec->core.mtid); seed = _mix(seed, (std::size_t) rec->core.mpos);
  seed = _mix(seed, (std::size_t) rec->core.tid); seed = _mix(seed, (std::size_t) rec->core.pos);
  return seed;
}
inline bool isBamCram(std::string const&) { return true; }
// util.h:237-246
inline std::string _addID(int32_t const svt) {
  if (svt == 0) return "INV";
  else if (svt == 1) return "INV";
  else if (svt == 2) return "DEL";
  else if (svt == 3) return "DUP";
  else if (svt == 4) return "INS";
  else if (svt == 9) return "CNV";
  else return "BND";
}
inline std::string _addOrientation(int32_t const) { return "NtoN"; }
// util.h:253-272
inline std::string _addAlleles(std::string const& ref, std::string const& chr2, StructuralVariantRecord const& sv, int32_t const svt) {
  if (_translocation(svt)) {
    uint8_t ct = _getSpanOrientation(svt);
    if (ct == 0) return ref + "," + ref + "]" + chr2 + ":" + std::to_string(sv.svEnd) + "]";
    else if (ct == 1) return ref + "," + "[" + chr2 + ":" + std::to_string(sv.svEnd) + "[" + ref;
    else if (ct == 2) return ref + "," + ref + "[" + chr2 + ":" + std::to_string(sv.svEnd) + "[";
    else if (ct == 3) return ref + "," + "]" + chr2 + ":" + std::to_string(sv.svEnd) + "]" + ref;
    else return ref + ",<" + _addID(svt) + ">";
  } else return ref + ",<" + _addID(svt) + ">";
}
// util.h:759-768
template <typename TConfig> inline int32_t getVariability(TConfig const&, std::vector<LibraryInfo> const& lib) {
  int32_t overallVariability = 0;
  for (uint32_t libIdx = 0; libIdx < lib.size(); ++libIdx) {
    if (lib[libIdx].maxNormalISize > overallVariability) overallVariability = lib[libIdx].maxNormalISize;
    if (lib[libIdx].rs > overallVariability) overallVariability = lib[libIdx].rs;
  }
  return overallVariability;
}
template <typename TConfig, typename A, typename B> inline void _alternateAlignments(TConfig const&, A&, B&) {}
struct Graph { std::map<std::string, std::size_t> smap; };
template <typename TConfig> inline bool parseGfa(TConfig const&, Graph&) { return false; }
// named by the alternate-alignment branch of _clusterSRReads (src/junction.h:504-590), which the oracle never takes (hasAltFile = false)
template <typename TConfig, typename THdr, typename TRegions> inline bool _parseExcludeIntervals(TConfig const&, THdr*, TRegions&) { return true; }
template <typename TConfig, typename TRegions, typename TGraph, typename TSR> inline void _findGraphSRBreakpoints(TConfig const&, TRegions const&, TGraph const&, TSR&) {}
}  // namespace torali

#ifdef ORACLE_FULL_LR
// libdelly_ref9.so: the same wrapper with the annotation step real (src/svanno.h compiled verbatim, with the reference's edlib) and the
// MM / ML tags of the records visible to genotypeLR — the complete long-read chain (ref_delly_lr_call_ex)
#include "edlib.h"
#include "svanno.h"
#else
namespace torali {
// src/svanno.h annotateSV: a no-op in libdelly_ref5.so (the chain without annotation; libdelly_ref9.so has the real one)
template <typename TConfig> inline void annotateSV(TConfig const&, bam_hdr_t*, char const*, StructuralVariantRecord&) {}
}  // namespace torali
#endif

#define MAX_CN 10
#include "shortpe.h"
#include "genotype.h"

namespace {
struct RefConfig5 {   // the fields assembleSplitReads, msa and alignConsensus read from TConfig (src/delly.h:49-82)
  std::vector<boost::filesystem::path> files;
  boost::filesystem::path genome;
  uint32_t maxThreads = 1, maxReadPerSV = 20;
  uint16_t minMapQual = 1, minCliqueSize = 2;
  torali::DnaScore<int> aliscore;
  float flankQuality = 0.95f;
  int32_t minimumFlankSize = 13, indelsize = 1000, minConsWindow = 100;
  // scanPEandSR (src/shortpe.h:285-533)
  uint16_t minTraQual = 20;
  uint32_t minClip = 25, minRefSep = 25, maxReadSep = 40, graphPruning = 1000;
  int32_t nchr = 0;
  std::set<int32_t> svtset;
  // annotateCoverage (src/coverage.h:265-743)
  boost::filesystem::path dumpfile;
  bool hasDumpFile = false;
  uint32_t maxGenoReadCount = 250;
  uint16_t minGenoQual = 5;
  // _clusterSRReads / findJunctions / genotypeLR (src/tegua.h:39-74)
  bool hasAltFile = false, hasExcludeFile = false;
  boost::filesystem::path exclude;
  float indelExtension = 0.5f;
  int32_t genoCap = 25, methylWindow = 500;
  uint32_t methylProb = 128, minCpgDepth = 1;
  float meiMinFrac = 0.8f, trMinFrac = 0.85f;   // annotateSV (src/tegua.h:63-64)
};
struct MemRecord5 { bam1_core_t core; std::vector<uint8_t> data; };
std::vector<MemRecord5> g_bam;                       // the in-memory BAM of file 0 ("in-memory.bam")
std::vector<std::vector<MemRecord5> > g_more;        // files 1.. of a multi-sample call ("in-memory.<k>.bam")
std::vector<uint32_t> g_tlen;
std::vector<std::string> g_names;
std::vector<char*> g_name_ptrs;
std::vector<const char*> g_seq;
uint32_t g_svtmask = 0;   // `-t` restriction for the next chain call (bit svt set = type wanted; 0 = all), see ref_set_svtset
struct MemInterval { uint32_t lo, hi; uint32_t lower() const { return lo; } uint32_t upper() const { return hi; } };
}  // namespace

extern "C" {

// the file index travels in htsFile::lineno, the file's record list in the index handle and then in hts_itr_t::reg_list
htsFile* hts_open(const char* name, const char*) {
  htsFile* f = (htsFile*) calloc(1, sizeof(htsFile)); f->is_bgzf = 1;
  int k = 0;
  if (name && sscanf(name, "in-memory.%d.bam", &k) == 1) f->lineno = k;
  return f;
}
int hts_close(htsFile* f) { free(f); return 0; }
int hts_set_fai_filename(htsFile*, const char*) { return 0; }
hts_idx_t* sam_index_load(htsFile* f, const char*) {
  const std::size_t k = (std::size_t) f->lineno;
  return (hts_idx_t*) ((k >= 1 && k <= g_more.size()) ? &g_more[k - 1] : &g_bam);
}
void hts_idx_destroy(hts_idx_t*) {}
int hts_idx_get_stat(const hts_idx_t*, int, uint64_t* mapped, uint64_t* unmapped) { *mapped = 1; *unmapped = 0; return 0; }
sam_hdr_t* sam_hdr_read(samFile*) {
  sam_hdr_t* h = (sam_hdr_t*) calloc(1, sizeof(sam_hdr_t));
  h->n_targets = (int32_t) g_tlen.size(); h->target_len = g_tlen.data(); h->target_name = g_name_ptrs.data();
  return h;
}
void sam_hdr_destroy(sam_hdr_t* h) { free(h); }
hts_itr_t* sam_itr_queryi(const hts_idx_t* idx, int tid, hts_pos_t beg, hts_pos_t end) {
  hts_itr_t* it = (hts_itr_t*) calloc(1, sizeof(hts_itr_t));
  it->tid = tid; it->beg = beg; it->end = end; it->i = 0;
  it->reg_list = (hts_reglist_t*) idx;
  return it;
}
int hts_itr_next(BGZF*, hts_itr_t* it, void* r, void*) {
  bam1_t* b = (bam1_t*) r;
  std::vector<MemRecord5>& bam = it->reg_list ? *(std::vector<MemRecord5>*) it->reg_list : g_bam;
  while (it->i < (int) bam.size()) {
    MemRecord5& m = bam[it->i++];
    if (m.core.tid != it->tid || m.core.pos < it->beg || m.core.pos >= it->end) continue;
    b->core = m.core;
    b->data = m.data.data(); b->l_data = (int) m.data.size(); b->m_data = (uint32_t) m.data.size();
    return 0;
  }
  return -1;
}
int hts_itr_multi_next(htsFile*, hts_itr_t*, void*) { return -1; }
void hts_itr_destroy(hts_itr_t* it) { free(it); }
bam1_t* bam_init1(void) { return (bam1_t*) calloc(1, sizeof(bam1_t)); }
void bam_destroy1(bam1_t* b) { free(b); }
void hts_log(enum htsLogLevel, const char*, const char*, ...) {}
#ifdef ORACLE_FULL_LR
// aux tags as in ref_wrap4.cpp: 'C' uint8, 'i' int32, 'Z' string, 'B:C' byte array
uint8_t* bam_aux_get(const bam1_t* b, const char tag[2]) {
  uint8_t* p = bam_get_aux(b);
  uint8_t* end = b->data + b->l_data;
  while (p + 3 <= end) {
    const bool hit = (p[0] == (uint8_t) tag[0] && p[1] == (uint8_t) tag[1]);
    const uint8_t type = p[2];
    if (hit) return p + 2;
    if (type == 'Z') p += 3 + strlen((const char*) p + 3) + 1;
    else if (type == 'B') { int32_t n; memcpy(&n, p + 4, 4); p += 3 + 1 + 4 + (size_t) n; }
    else p += 3 + (type == 'C' ? 1 : 4);
  }
  return NULL;
}
int64_t bam_aux2i(const uint8_t* s) {
  if (s[0] == 'C') return s[1];
  int32_t v; memcpy(&v, s + 1, 4); return v;
}
#else
uint8_t* bam_aux_get(const bam1_t*, const char[2]) { return NULL; }
int64_t bam_aux2i(const uint8_t*) { return 0; }
#endif
faidx_t* fai_load(const char*) { return (faidx_t*) &g_names; }
void fai_destroy(faidx_t*) {}
char* faidx_fetch_seq(const faidx_t*, const char* name, int beg, int end, int* len) {
  std::size_t k = 0;
  while (k < g_names.size() && g_names[k] != name) ++k;
  if (end >= (int) g_tlen[k]) end = (int) g_tlen[k] - 1;
  const int n = end - beg + 1;
  char* out = (char*) malloc((size_t) n + 1);
  memcpy(out, g_seq[k] + beg, (size_t) n);
  out[n] = 0;
  *len = n;
  return out;
}

// assembleSplitReads (src/shortpe.h:48-282) over in-memory data.
//   contigs "chr0", "chr1", ...; rec: nrec x 12 as in ref_wrap3.cpp (name id in column 11; BAM_FREAD2 in the flag selects the mate);
//   store: nstore x [refIndex, pos, name id, read2 (0/1), svid] = the srStore entries (src/shortpe.h:462-476 builds them from the clusters);
//   sv: nsv x [chr, svStart, chr2, svEnd, svt, insLen, id];
//   out: sv_out nsv x 13 [svStart, svEnd, srSupport, mapq, srMapQuality, insLen, homLen, consBp, precise, ciposlow, ciposhigh, ciendlow, ciendhigh],
//        srq (float), consensus and alleles strings (stride bytes each + lengths).
int ref_assemble_split_reads(const char* contig_arena, const uint32_t* contig_off, const uint32_t* contig_len, int ncontig, const int32_t* rec12, int nrec,
                             const uint32_t* cigar, const char* reads, const int32_t* store5, int nstore, const int32_t* sv7, int nsv, int maxReadPerSV,
                             int minMapQual, int minCliqueSize, float flankQuality, int minimumFlankSize, int indelsize, int minConsWindow,
                             int32_t* sv_out, float* srq, char* cons_out, int cons_stride, int32_t* cons_len, char* alleles_out, int alleles_stride,
                             int32_t* alleles_len) {
  RefConfig5 c;
  c.files.push_back(boost::filesystem::path("in-memory.bam")); c.genome = boost::filesystem::path("in-memory.fa");
  c.maxReadPerSV = (uint32_t) maxReadPerSV; c.minMapQual = (uint16_t) minMapQual; c.minCliqueSize = (uint16_t) minCliqueSize;
  c.aliscore = torali::DnaScore<int>(5, -4, -10, -1);   // src/delly.h:222
  c.flankQuality = flankQuality; c.minimumFlankSize = minimumFlankSize; c.indelsize = indelsize; c.minConsWindow = minConsWindow;
  g_names.clear(); g_tlen.clear(); g_seq.clear(); g_name_ptrs.clear();
  for (int k = 0; k < ncontig; ++k) { g_names.push_back("chr" + std::to_string(k)); g_tlen.push_back(contig_len[k]); g_seq.push_back(contig_arena + contig_off[k]); }
  for (auto& nm : g_names) g_name_ptrs.push_back((char*) nm.c_str());
  g_bam.clear();
  for (int i = 0; i < nrec; ++i) {
    const int32_t* r = rec12 + 12 * i;
    MemRecord5 m; memset(&m.core, 0, sizeof(m.core));
    std::string qn = "q" + std::to_string(r[11]);
    m.core.tid = r[0]; m.core.pos = r[1]; m.core.flag = (uint16_t) r[2]; m.core.qual = (uint8_t) r[3]; m.core.l_qseq = r[4]; m.core.n_cigar = (uint32_t) r[6];
    m.core.mtid = r[8]; m.core.mpos = r[9]; m.core.isize = r[10];
    m.core.l_qname = (uint16_t) ((qn.size() + 1 + 3) & ~3u);
    const std::size_t lq = (std::size_t) r[4];
    m.data.assign(m.core.l_qname + 4 * (std::size_t) r[6] + (lq + 1) / 2 + lq, 0);
    memcpy(m.data.data(), qn.data(), qn.size());
    memcpy(m.data.data() + m.core.l_qname, cigar + r[5], 4 * (std::size_t) r[6]);
    uint8_t* sq = m.data.data() + m.core.l_qname + 4 * (std::size_t) r[6];
    for (std::size_t k = 0; k < lq; ++k) {
      const char* tab = "=ACMGRSVTWYHKDBN";
      const char* f = strchr(tab, reads[(std::size_t) r[7] + k]);
      const uint8_t code = f ? (uint8_t) (f - tab) : 15;
      sq[k >> 1] |= (k & 1) ? code : (uint8_t) (code << 4);
    }
    g_bam.push_back(m);
  }
  std::vector<torali::StructuralVariantRecord> svs(nsv);
  for (int i = 0; i < nsv; ++i) {
    const int32_t* s = sv7 + 7 * i;
    svs[i].chr = s[0]; svs[i].svStart = s[1]; svs[i].chr2 = s[2]; svs[i].svEnd = s[3]; svs[i].svt = s[4]; svs[i].insLen = s[5]; svs[i].id = s[6];
    svs[i].precise = false;
  }
  typedef std::map<std::pair<int32_t, std::size_t>, int32_t> TPosReadSV;
  std::vector<TPosReadSV> srStore(ncontig);
  for (int i = 0; i < nstore; ++i) {
    const int32_t* e = store5 + 5 * i;
    const std::size_t seed = std::hash<std::string>()("q" + std::to_string(e[2])) * 2 + (e[3] ? 1 : 0);
    srStore[e[0]].insert(std::make_pair(std::make_pair(e[1], seed), e[4]));
  }
  std::vector<std::vector<MemInterval> > validRegions(ncontig);
  for (int t = 0; t < ncontig; ++t) validRegions[t].push_back(MemInterval{0u, 0x7fffffffu});
  std::streambuf* old = std::cerr.rdbuf(nullptr);
  torali::assembleSplitReads(c, validRegions, srStore, svs);
  std::cerr.rdbuf(old);
  for (int i = 0; i < nsv; ++i) {
    torali::StructuralVariantRecord const& sv = svs[i];
    int32_t* o = sv_out + 13 * i;
    o[0] = sv.svStart; o[1] = sv.svEnd; o[2] = sv.srSupport; o[3] = sv.mapq; o[4] = sv.srMapQuality; o[5] = sv.insLen; o[6] = sv.homLen; o[7] = sv.consBp;
    o[8] = sv.precise ? 1 : 0; o[9] = sv.ciposlow; o[10] = sv.ciposhigh; o[11] = sv.ciendlow; o[12] = sv.ciendhigh;
    srq[i] = sv.srAlignQuality;
    cons_len[i] = (int32_t) sv.consensus.size();
    memcpy(cons_out + (size_t) i * cons_stride, sv.consensus.data(), std::min<size_t>(sv.consensus.size(), cons_stride));
    alleles_len[i] = (int32_t) sv.alleles.size();
    memcpy(alleles_out + (size_t) i * alleles_stride, sv.alleles.data(), std::min<size_t>(sv.alleles.size(), alleles_stride));
  }
  return 0;
}

// scanPEandSR (src/shortpe.h:285-533) over in-memory alignments: the discovery front end (CIGAR junction scan, abnormal-pair
// collection, junction selection, SR and PE clustering, the split-read store).
//   rec: nrec x 12 (as above); lib: [rs, median, mad, maxNormalISize, maxISizeCutoff]
//   out: pe_out cap x 12 [chr,svStart,chr2,svEnd,ciposlow,ciposhigh,ciendlow,ciendhigh,peSupport,peMapQuality,mapq,svt] (PE SVs);
//        sr_out cap x 14 [chr,svStart,chr2,svEnd,ciposlow,ciposhigh,ciendlow,ciendhigh,srSupport,srMapQuality,mapq,insLen,svt,id] (SR SVs);
//        store: per contig ascending (pos, read seed): [refIndex, pos, svid] + seed (uint64). counts in n_out[3] = {nPE, nSR, nStore}.
int ref_scan_pe_sr(const uint32_t* contig_len, int ncontig, const int32_t* rec12, int nrec, const uint32_t* cigar, const int32_t* lib5, int minMapQual,
                   int minTraQual, int minClip, int minRefSep, int maxReadSep, int minCliqueSize, int graphPruning, int32_t* pe_out, int32_t* sr_out,
                   int cap, int32_t* store_out, uint64_t* store_seed, int store_cap, int32_t* n_out, uint32_t* abnormal_pairs) {
  RefConfig5 c;
  c.files.push_back(boost::filesystem::path("in-memory.bam")); c.genome = boost::filesystem::path("in-memory.fa");
  c.minMapQual = (uint16_t) minMapQual; c.minTraQual = (uint16_t) minTraQual; c.minClip = (uint32_t) minClip; c.minRefSep = (uint32_t) minRefSep;
  c.maxReadSep = (uint32_t) maxReadSep; c.minCliqueSize = (uint16_t) minCliqueSize; c.graphPruning = (uint32_t) graphPruning; c.nchr = ncontig;
  g_names.clear(); g_tlen.clear(); g_seq.clear(); g_name_ptrs.clear();
  for (int k = 0; k < ncontig; ++k) { g_names.push_back("chr" + std::to_string(k)); g_tlen.push_back(contig_len[k]); g_seq.push_back(nullptr); }
  for (auto& nm : g_names) g_name_ptrs.push_back((char*) nm.c_str());
  g_bam.clear();
  for (int i = 0; i < nrec; ++i) {
    const int32_t* r = rec12 + 12 * i;
    MemRecord5 m; memset(&m.core, 0, sizeof(m.core));
    std::string qn = "q" + std::to_string(r[11]);
    m.core.tid = r[0]; m.core.pos = r[1]; m.core.flag = (uint16_t) r[2]; m.core.qual = (uint8_t) r[3]; m.core.l_qseq = r[4]; m.core.n_cigar = (uint32_t) r[6];
    m.core.mtid = r[8]; m.core.mpos = r[9]; m.core.isize = r[10];
    m.core.l_qname = (uint16_t) ((qn.size() + 1 + 3) & ~3u);
    const std::size_t lq = (std::size_t) r[4];
    m.data.assign(m.core.l_qname + 4 * (std::size_t) r[6] + (lq + 1) / 2 + lq, 0);
    memcpy(m.data.data(), qn.data(), qn.size());
    memcpy(m.data.data() + m.core.l_qname, cigar + r[5], 4 * (std::size_t) r[6]);
    g_bam.push_back(m);
  }
  std::vector<torali::LibraryInfo> sampleLib(1);
  sampleLib[0].rs = lib5[0]; sampleLib[0].median = lib5[1]; sampleLib[0].mad = lib5[2]; sampleLib[0].maxNormalISize = lib5[3]; sampleLib[0].maxISizeCutoff = lib5[4];
  std::vector<std::vector<MemInterval> > validRegions(ncontig);
  for (int t = 0; t < ncontig; ++t) validRegions[t].push_back(MemInterval{0u, 0x7fffffffu});
  typedef std::map<std::pair<int32_t, std::size_t>, int32_t> TPosReadSV;
  std::vector<TPosReadSV> srStore(ncontig);
  std::vector<torali::StructuralVariantRecord> svs, srSVs;
  std::streambuf* old = std::cerr.rdbuf(nullptr);
  torali::scanPEandSR(c, validRegions, svs, srSVs, srStore, sampleLib);
  std::cerr.rdbuf(old);
  if ((int) svs.size() > cap || (int) srSVs.size() > cap) return -1;
  for (std::size_t i = 0; i < svs.size(); ++i) {
    int32_t* o = pe_out + 12 * i; torali::StructuralVariantRecord const& v = svs[i];
    o[0] = v.chr; o[1] = v.svStart; o[2] = v.chr2; o[3] = v.svEnd; o[4] = v.ciposlow; o[5] = v.ciposhigh; o[6] = v.ciendlow; o[7] = v.ciendhigh;
    o[8] = v.peSupport; o[9] = v.peMapQuality; o[10] = v.mapq; o[11] = v.svt;
  }
  for (std::size_t i = 0; i < srSVs.size(); ++i) {
    int32_t* o = sr_out + 14 * i; torali::StructuralVariantRecord const& v = srSVs[i];
    o[0] = v.chr; o[1] = v.svStart; o[2] = v.chr2; o[3] = v.svEnd; o[4] = v.ciposlow; o[5] = v.ciposhigh; o[6] = v.ciendlow; o[7] = v.ciendhigh;
    o[8] = v.srSupport; o[9] = v.srMapQuality; o[10] = v.mapq; o[11] = v.insLen; o[12] = v.svt; o[13] = v.id;
  }
  int k = 0;
  for (int t = 0; t < ncontig; ++t)
    for (auto const& kv : srStore[t]) {
      if (k >= store_cap) return -1;
      store_out[3 * k] = t; store_out[3 * k + 1] = kv.first.first; store_out[3 * k + 2] = kv.second; store_seed[k] = (uint64_t) kv.first.second;
      ++k;
    }
  n_out[0] = (int32_t) svs.size(); n_out[1] = (int32_t) srSVs.size(); n_out[2] = k;
  *abnormal_pairs = sampleLib[0].abnormal_pairs;
  return 0;
}

// mergeSort (src/shortpe.h:536-621)
// sv rows: [chr, svStart, chr2, svEnd, ciposlow, ciposhigh, ciendlow, ciendhigh, peSupport, srSupport, peMapQuality, srMapQuality, mapq, insLen,
//           homLen, svt, precise, consBp, consensus id, srAlignQuality x 1000]  (the consensus travels as a decimal id string)
int ref_merge_sort(const int32_t* pe20, int npe, const int32_t* sr20, int nsr, int32_t* out20, int cap) {
  auto load = [](const int32_t* r) {
    torali::StructuralVariantRecord v;
    v.chr = r[0]; v.svStart = r[1]; v.chr2 = r[2]; v.svEnd = r[3]; v.ciposlow = r[4]; v.ciposhigh = r[5]; v.ciendlow = r[6]; v.ciendhigh = r[7];
    v.peSupport = r[8]; v.srSupport = r[9]; v.peMapQuality = r[10]; v.srMapQuality = r[11]; v.mapq = r[12]; v.insLen = r[13]; v.homLen = r[14]; v.svt = r[15];
    v.precise = r[16] != 0; v.consBp = r[17]; v.consensus = std::to_string(r[18]); v.srAlignQuality = (float) r[19] / 1000.0f;
    return v;
  };
  std::vector<torali::StructuralVariantRecord> pe, sr;
  for (int i = 0; i < npe; ++i) pe.push_back(load(pe20 + 20 * i));
  for (int i = 0; i < nsr; ++i) sr.push_back(load(sr20 + 20 * i));
  torali::mergeSort(pe, sr);
  if ((int) pe.size() > cap) return -1;
  for (std::size_t i = 0; i < pe.size(); ++i) {
    int32_t* o = out20 + 20 * i; torali::StructuralVariantRecord const& v = pe[i];
    o[0] = v.chr; o[1] = v.svStart; o[2] = v.chr2; o[3] = v.svEnd; o[4] = v.ciposlow; o[5] = v.ciposhigh; o[6] = v.ciendlow; o[7] = v.ciendhigh;
    o[8] = v.peSupport; o[9] = v.srSupport; o[10] = v.peMapQuality; o[11] = v.srMapQuality; o[12] = v.mapq; o[13] = v.insLen; o[14] = v.homLen; o[15] = v.svt;
    o[16] = v.precise ? 1 : 0; o[17] = v.consBp; o[18] = v.consensus.empty() ? -1 : std::stoi(v.consensus); o[19] = (int32_t) std::lround(v.srAlignQuality * 1000.0f);
  }
  return (int) pe.size();
}

// The stage sequence of dellyRun (src/delly.h:127-178) over in-memory alignments of one sample, every stage the reference's own
// function compiled verbatim: scanPEandSR -> assembleSplitReads -> mergeSort -> sort + renumber -> annotateCoverage -> _computeGLs.
// (The ten lines of glue between the stages follow src/delly.h:139-158; PL / RCN / FT follow src/modvcf.h:671-715.)
//   lib6: [rs, median, mad, minNormalISize, maxNormalISize, maxISizeCutoff]
//   sv_out n x 20 (as ref_merge_sort, [18] = id, [19] = srAlignQuality bits); fmt_out n x 14 [gt0, gt1, gq, pl0, pl1, pl2, rcn, pass, DR, DV, RR, RV, RC, RCL+RCR]
// Several input files: records grouped by file (file_off[nfile + 1], each group sorted by (tid, pos)), one lib6 row per file; fmt_out / gl_out
// are [file][sv] (file-major, `cap` rows per file). nfile = 1 is the single-sample call.
static int run_delly_sr_call(const char* contig_arena, const uint32_t* contig_off, const uint32_t* contig_len, int ncontig, const int32_t* rec12, const uint32_t* file_off,
                      int nfile, const uint32_t* cigar, const char* reads, const int32_t* lib6, int32_t* sv_out, int cap, int32_t* fmt_out, float* gl_out, char* cons_out,
                      int cons_stride, int32_t* cons_len) {
  RefConfig5 c;   // short-read defaults (src/delly.h:212-232)
  for (int b = 0; b < 10; ++b) if (g_svtmask & (1u << b)) c.svtset.insert(b);
  c.files.push_back(boost::filesystem::path("in-memory.bam")); c.genome = boost::filesystem::path("in-memory.fa");
  for (int f = 1; f < nfile; ++f) c.files.push_back(boost::filesystem::path("in-memory." + std::to_string(f) + ".bam"));
  const int nrec = (int) file_off[nfile];
  c.aliscore = torali::DnaScore<int>(5, -4, -10, -1);
  c.nchr = ncontig;
  g_names.clear(); g_tlen.clear(); g_seq.clear(); g_name_ptrs.clear();
  for (int k = 0; k < ncontig; ++k) { g_names.push_back("chr" + std::to_string(k)); g_tlen.push_back(contig_len[k]); g_seq.push_back(contig_arena + contig_off[k]); }
  for (auto& nm : g_names) g_name_ptrs.push_back((char*) nm.c_str());
  g_bam.clear(); g_more.assign((size_t) std::max(0, nfile - 1), std::vector<MemRecord5>());
  for (int i = 0; i < nrec; ++i) {
    const int32_t* r = rec12 + 12 * i;
    MemRecord5 m; memset(&m.core, 0, sizeof(m.core));
    std::string qn = "q" + std::to_string(r[11]);
    m.core.tid = r[0]; m.core.pos = r[1]; m.core.flag = (uint16_t) r[2]; m.core.qual = (uint8_t) r[3]; m.core.l_qseq = r[4]; m.core.n_cigar = (uint32_t) r[6];
    m.core.mtid = r[8]; m.core.mpos = r[9]; m.core.isize = r[10];
    m.core.l_qname = (uint16_t) ((qn.size() + 1 + 3) & ~3u);
    const std::size_t lq = (std::size_t) r[4];
    m.data.assign(m.core.l_qname + 4 * (std::size_t) r[6] + (lq + 1) / 2 + lq, 0);
    memcpy(m.data.data(), qn.data(), qn.size());
    memcpy(m.data.data() + m.core.l_qname, cigar + r[5], 4 * (std::size_t) r[6]);
    uint8_t* sq = m.data.data() + m.core.l_qname + 4 * (std::size_t) r[6];
    for (std::size_t k = 0; k < lq; ++k) {
      const char* tab = "=ACMGRSVTWYHKDBN";
      const char* f = strchr(tab, reads[(std::size_t) r[7] + k]);
      const uint8_t code = f ? (uint8_t) (f - tab) : 15;
      sq[k >> 1] |= (k & 1) ? code : (uint8_t) (code << 4);
    }
    int file = 0;
    while ((uint32_t) i >= file_off[file + 1]) ++file;
    (file == 0 ? g_bam : g_more[(size_t) file - 1]).push_back(m);
  }
  std::vector<torali::LibraryInfo> sampleLib((size_t) nfile);
  for (int f = 0; f < nfile; ++f) {
    const int32_t* l = lib6 + 6 * f;
    sampleLib[f].rs = l[0]; sampleLib[f].median = l[1]; sampleLib[f].mad = l[2]; sampleLib[f].minNormalISize = l[3]; sampleLib[f].maxNormalISize = l[4];
    sampleLib[f].maxISizeCutoff = l[5];
  }
  std::vector<std::vector<MemInterval> > validRegions(ncontig);
  for (int t = 0; t < ncontig; ++t) validRegions[t].push_back(MemInterval{0u, 0x7fffffffu});
  typedef std::vector<torali::StructuralVariantRecord> TVariants;
  TVariants svs;
  std::streambuf* old = std::cerr.rdbuf(nullptr);
  {
    TVariants srSVs;
    typedef std::map<std::pair<int32_t, std::size_t>, int32_t> TPosReadSV;
    std::vector<TPosReadSV> srStore(c.nchr, TPosReadSV());
    torali::scanPEandSR(c, validRegions, svs, srSVs, srStore, sampleLib);
    torali::assembleSplitReads(c, validRegions, srStore, srSVs);
    torali::mergeSort(svs, srSVs);
  }
  sort(svs.begin(), svs.end());
  uint32_t cliqueCount = 0;
  for (TVariants::iterator svIt = svs.begin(); svIt != svs.end(); ++svIt, ++cliqueCount) svIt->id = cliqueCount;
  std::vector<std::vector<torali::JunctionCount> > jctMap;
  std::vector<std::vector<torali::SpanningCount> > spanMap;
  std::vector<std::vector<torali::ReadCount> > rcMap;
  if (!svs.empty()) torali::annotateCoverage(c, sampleLib, svs, rcMap, jctMap, spanMap);
  std::cerr.rdbuf(old);
  const int n = (int) svs.size();
  if (n > cap) return -1;
  static torali::BoLog<double> bl;
  for (int i = 0; i < n; ++i) {
    torali::StructuralVariantRecord const& v = svs[i];
    int32_t* o = sv_out + 20 * i;
    o[0] = v.chr; o[1] = v.svStart; o[2] = v.chr2; o[3] = v.svEnd; o[4] = v.ciposlow; o[5] = v.ciposhigh; o[6] = v.ciendlow; o[7] = v.ciendhigh;
    o[8] = v.peSupport; o[9] = v.srSupport; o[10] = v.peMapQuality; o[11] = v.srMapQuality; o[12] = v.mapq; o[13] = v.insLen; o[14] = v.homLen; o[15] = v.svt;
    o[16] = v.precise ? 1 : 0; o[17] = v.consBp; o[18] = v.id;
    memcpy(o + 19, &v.srAlignQuality, 4);
    for (int f = 0; f < nfile; ++f) {
    float gls[3]; int32_t gq[1]; int32_t gts[2];
    if (v.precise) torali::_computeGLs(bl, jctMap[f][v.id].ref, jctMap[f][v.id].alt, gls, gq, gts, 0);
    else torali::_computeGLs(bl, spanMap[f][v.id].ref, spanMap[f][v.id].alt, gls, gq, gts, 0);
    int32_t* q = fmt_out + 14 * ((size_t) f * cap + i);
    q[0] = gts[0]; q[1] = gts[1]; q[2] = gq[0];
    for (int k = 0; k < 3; ++k) q[3 + k] = (gts[0] == bcf_gt_missing) ? bcf_int32_missing : (int32_t) std::max(0.0f, std::round(-10.0f * gls[k]));
    torali::ReadCount const& rcv = rcMap[f][v.id];
    int32_t cnest = -1;
    if ((rcv.leftRC + rcv.rightRC) > 0) {
      double cn = 2.0 * (double) rcv.rc / (double) (rcv.leftRC + rcv.rightRC);
      if (cn < 0) cn = 0;
      if (cn > 100000) cn = 100000;
      cnest = boost::math::iround(cn);
    }
    q[6] = cnest; q[7] = (gq[0] < 15) ? 0 : 1;
    q[8] = (int32_t) spanMap[f][v.id].ref.size(); q[9] = (int32_t) spanMap[f][v.id].alt.size(); q[10] = (int32_t) jctMap[f][v.id].ref.size(); q[11] = (int32_t) jctMap[f][v.id].alt.size();
    q[12] = rcv.rc; q[13] = rcv.leftRC + rcv.rightRC;
    for (int k = 0; k < 3; ++k) gl_out[3 * ((size_t) f * cap + i) + k] = gls[k];
    }
    cons_len[i] = (int32_t) v.consensus.size();
    memcpy(cons_out + (size_t) i * cons_stride, v.consensus.data(), std::min<size_t>(v.consensus.size(), cons_stride));
  }
  g_more.clear();
  return n;
}

// `-t`: restrict the SV types of the following chain calls (src/util.h:370-395 fills c.svtset from the option string)
void ref_set_svtset(uint32_t mask) { g_svtmask = mask; }

int ref_delly_sr_call(const char* contig_arena, const uint32_t* contig_off, const uint32_t* contig_len, int ncontig, const int32_t* rec12, int nrec,
                      const uint32_t* cigar, const char* reads, const int32_t* lib6, int32_t* sv_out, int cap, int32_t* fmt_out, float* gl_out, char* cons_out,
                      int cons_stride, int32_t* cons_len) {
  const uint32_t file_off[2] = {0u, (uint32_t) nrec};
  return run_delly_sr_call(contig_arena, contig_off, contig_len, ncontig, rec12, file_off, 1, cigar, reads, lib6, sv_out, cap, fmt_out, gl_out, cons_out, cons_stride,
                           cons_len);
}

int ref_delly_sr_call_multi(const char* contig_arena, const uint32_t* contig_off, const uint32_t* contig_len, int ncontig, const int32_t* rec12, const uint32_t* file_off,
                            int nfile, const uint32_t* cigar, const char* reads, const int32_t* lib6, int32_t* sv_out, int cap, int32_t* fmt_out, float* gl_out,
                            char* cons_out, int cons_stride, int32_t* cons_len) {
  return run_delly_sr_call(contig_arena, contig_off, contig_len, ncontig, rec12, file_off, nfile, cigar, reads, lib6, sv_out, cap, fmt_out, gl_out, cons_out, cons_stride,
                           cons_len);
}

// assemble (src/assemble.h:736-964, the long-read assembly stage) over in-memory data.
//   rec: nrec x 12 (read id = hash_lr of the query name "q<name id>"); store: nstore x [refIndex, pos, name id, svid, sstart, inslen];
//   sv: nsv x [chr, svStart, chr2, svEnd, svt, insLen, id, srSupport, ciposlow, ciposhigh, ciendlow, ciendhigh];
//   outputs as ref_assemble_split_reads.
int ref_assemble_lr(const char* contig_arena, const uint32_t* contig_off, const uint32_t* contig_len, int ncontig, const int32_t* rec12, int nrec,
                    const uint32_t* cigar, const char* reads, const int32_t* store6, int nstore, const int32_t* sv12, int nsv, int maxReadPerSV, int minCliqueSize,
                    float flankQuality, int minimumFlankSize, int indelsize, int minConsWindow, int32_t* sv_out, float* srq, char* cons_out, int cons_stride,
                    int32_t* cons_len, char* alleles_out, int alleles_stride, int32_t* alleles_len) {
  RefConfig5 c;
  c.files.push_back(boost::filesystem::path("in-memory.bam")); c.genome = boost::filesystem::path("in-memory.fa");
  c.maxReadPerSV = (uint32_t) maxReadPerSV; c.minCliqueSize = (uint16_t) minCliqueSize;
  c.flankQuality = flankQuality; c.minimumFlankSize = minimumFlankSize; c.indelsize = indelsize; c.minConsWindow = minConsWindow;
  g_names.clear(); g_tlen.clear(); g_seq.clear(); g_name_ptrs.clear();
  for (int k = 0; k < ncontig; ++k) { g_names.push_back("chr" + std::to_string(k)); g_tlen.push_back(contig_len[k]); g_seq.push_back(contig_arena + contig_off[k]); }
  for (auto& nm : g_names) g_name_ptrs.push_back((char*) nm.c_str());
  g_bam.clear();
  for (int i = 0; i < nrec; ++i) {
    const int32_t* r = rec12 + 12 * i;
    MemRecord5 m; memset(&m.core, 0, sizeof(m.core));
    std::string qn = "q" + std::to_string(r[11]);
    m.core.tid = r[0]; m.core.pos = r[1]; m.core.flag = (uint16_t) r[2]; m.core.qual = (uint8_t) r[3]; m.core.l_qseq = r[4]; m.core.n_cigar = (uint32_t) r[6];
    m.core.mtid = r[8]; m.core.mpos = r[9]; m.core.isize = r[10];
    m.core.l_qname = (uint16_t) ((qn.size() + 1 + 3) & ~3u);
    const std::size_t lq = (std::size_t) r[4];
    m.data.assign(m.core.l_qname + 4 * (std::size_t) r[6] + (lq + 1) / 2 + lq, 0);
    memcpy(m.data.data(), qn.data(), qn.size());
    memcpy(m.data.data() + m.core.l_qname, cigar + r[5], 4 * (std::size_t) r[6]);
    uint8_t* sq = m.data.data() + m.core.l_qname + 4 * (std::size_t) r[6];
    for (std::size_t k = 0; k < lq; ++k) {
      const char* tab = "=ACMGRSVTWYHKDBN";
      const char* f = strchr(tab, reads[(std::size_t) r[7] + k]);
      const uint8_t code = f ? (uint8_t) (f - tab) : 15;
      sq[k >> 1] |= (k & 1) ? code : (uint8_t) (code << 4);
    }
    g_bam.push_back(m);
  }
  std::vector<torali::StructuralVariantRecord> svs(nsv);
  for (int i = 0; i < nsv; ++i) {
    const int32_t* s = sv12 + 12 * i;
    svs[i].chr = s[0]; svs[i].svStart = s[1]; svs[i].chr2 = s[2]; svs[i].svEnd = s[3]; svs[i].svt = s[4]; svs[i].insLen = s[5]; svs[i].id = s[6];
    svs[i].srSupport = s[7]; svs[i].ciposlow = s[8]; svs[i].ciposhigh = s[9]; svs[i].ciendlow = s[10]; svs[i].ciendhigh = s[11];
    svs[i].precise = false;
  }
  typedef std::map<std::pair<int32_t, std::size_t>, std::vector<torali::SeqSlice> > TPosReadSV;
  std::vector<TPosReadSV> srStore(ncontig);
  for (int i = 0; i < nstore; ++i) {
    const int32_t* e = store6 + 6 * i;
    const std::size_t seed = std::hash<std::string>()("q" + std::to_string(e[2]));
    srStore[e[0]][std::make_pair(e[1], seed)].push_back(torali::SeqSlice(e[3], e[4], e[5], 60));
  }
  std::vector<std::vector<MemInterval> > validRegions(ncontig);
  for (int t = 0; t < ncontig; ++t) validRegions[t].push_back(MemInterval{0u, 0x7fffffffu});
  std::streambuf* old = std::cerr.rdbuf(nullptr);
  torali::assemble(c, validRegions, svs, srStore);
  std::cerr.rdbuf(old);
  for (int i = 0; i < nsv; ++i) {
    torali::StructuralVariantRecord const& sv = svs[i];
    int32_t* o = sv_out + 13 * i;
    o[0] = sv.svStart; o[1] = sv.svEnd; o[2] = sv.srSupport; o[3] = sv.mapq; o[4] = sv.srMapQuality; o[5] = sv.insLen; o[6] = sv.homLen; o[7] = sv.consBp;
    o[8] = sv.precise ? 1 : 0; o[9] = sv.ciposlow; o[10] = sv.ciposhigh; o[11] = sv.ciendlow; o[12] = sv.ciendhigh;
    srq[i] = sv.srAlignQuality;
    cons_len[i] = (int32_t) sv.consensus.size();
    memcpy(cons_out + (size_t) i * cons_stride, sv.consensus.data(), std::min<size_t>(sv.consensus.size(), cons_stride));
    alleles_len[i] = (int32_t) sv.alleles.size();
    memcpy(alleles_out + (size_t) i * alleles_stride, sv.alleles.data(), std::min<size_t>(sv.alleles.size(), alleles_stride));
  }
  return 0;
}

// runTegua's stage sequence for one sample (src/tegua.h:104-193) over in-memory alignments, every stage the reference's own function:
// _clusterSRReads -> assemble -> sort -> neighbour de-duplication -> sort + renumber -> genotypeLR -> _computeGLs.
// (The glue between the stages follows src/tegua.h:118-146; PL / RCN / FT follow src/modvcf.h:671-715.)
//   cfg: [minMapQual, minClip, minRefSep, maxReadSep, minCliqueSize, graphPruning, maxReadPerSV, minimumFlankSize, indelsize, minConsWindow, maxGenoReadCount, genoCap]
// The _ex form (meaningful in libdelly_ref9.so): records with MM / ML tags (tagflags bit 0 / 1, texts / bytes by offsets as in
// ref_wrap4.cpp::ref_genotype_lr_methyl), methylation and annotation thresholds; extra outputs per SV: anno5 [isRC, seqType, homLen, trPeriod,
// trCopies bits], methyl16 (MethylInfo in field order), alleles (fixed stride, lengths).
static int run_delly_lr_call(const char* contig_arena, const uint32_t* contig_off, const uint32_t* contig_len, int ncontig, const int32_t* rec12, const uint32_t* file_off,
                      int nfile, const uint32_t* cigar, const char* reads, const int32_t* cfg12, float flankQuality, float indelExtension, int32_t* sv_out, int cap,
                      int32_t* fmt_out, float* gl_out, char* cons_out, int cons_stride, int32_t* cons_len, const uint8_t* tagflags, const char* mm_arena,
                      const uint32_t* mm_off, const uint8_t* ml_arena, const uint32_t* ml_off, int methylWindow, int methylProb, int minCpgDepth, float meiMinFrac,
                      float trMinFrac, int32_t* anno_out, int32_t* methyl_out, char* alleles_out, int alleles_stride, int32_t* alleles_len) {
  RefConfig5 c;
  for (int b = 0; b < 10; ++b) if (g_svtmask & (1u << b)) c.svtset.insert(b);
  const int nrec = (int) file_off[nfile];
  if (tagflags) {
    c.methylWindow = methylWindow; c.methylProb = (uint32_t) methylProb; c.minCpgDepth = (uint32_t) minCpgDepth; c.meiMinFrac = meiMinFrac; c.trMinFrac = trMinFrac;
  }
  c.files.push_back(boost::filesystem::path("in-memory.bam")); c.genome = boost::filesystem::path("in-memory.fa");
  for (int f = 1; f < nfile; ++f) c.files.push_back(boost::filesystem::path("in-memory." + std::to_string(f) + ".bam"));
  c.minMapQual = (uint16_t) cfg12[0]; c.minClip = (uint32_t) cfg12[1]; c.minRefSep = (uint32_t) cfg12[2]; c.maxReadSep = (uint32_t) cfg12[3];
  c.minCliqueSize = (uint16_t) cfg12[4]; c.graphPruning = (uint32_t) cfg12[5]; c.maxReadPerSV = (uint32_t) cfg12[6]; c.minimumFlankSize = cfg12[7];
  c.indelsize = cfg12[8]; c.minConsWindow = cfg12[9]; c.maxGenoReadCount = (uint32_t) cfg12[10]; c.genoCap = cfg12[11];
  c.flankQuality = flankQuality; c.indelExtension = indelExtension; c.nchr = ncontig;
  g_names.clear(); g_tlen.clear(); g_seq.clear(); g_name_ptrs.clear();
  for (int k = 0; k < ncontig; ++k) { g_names.push_back("chr" + std::to_string(k)); g_tlen.push_back(contig_len[k]); g_seq.push_back(contig_arena + contig_off[k]); }
  for (auto& nm : g_names) g_name_ptrs.push_back((char*) nm.c_str());
  g_bam.clear(); g_more.assign((size_t) std::max(0, nfile - 1), std::vector<MemRecord5>());
  for (int i = 0; i < nrec; ++i) {
    const int32_t* r = rec12 + 12 * i;
    MemRecord5 m; memset(&m.core, 0, sizeof(m.core));
    std::string qn = "q" + std::to_string(r[11]);
    m.core.tid = r[0]; m.core.pos = r[1]; m.core.flag = (uint16_t) r[2]; m.core.qual = (uint8_t) r[3]; m.core.l_qseq = r[4]; m.core.n_cigar = (uint32_t) r[6];
    m.core.mtid = r[8]; m.core.mpos = r[9]; m.core.isize = r[10];
    m.core.l_qname = (uint16_t) ((qn.size() + 1 + 3) & ~3u);
    const std::size_t lq = (std::size_t) r[4];
    m.data.assign(m.core.l_qname + 4 * (std::size_t) r[6] + (lq + 1) / 2 + lq, 0);
    memcpy(m.data.data(), qn.data(), qn.size());
    memcpy(m.data.data() + m.core.l_qname, cigar + r[5], 4 * (std::size_t) r[6]);
    uint8_t* sq = m.data.data() + m.core.l_qname + 4 * (std::size_t) r[6];
    for (std::size_t k = 0; k < lq; ++k) {
      const char* tab = "=ACMGRSVTWYHKDBN";
      const char* f = strchr(tab, reads[(std::size_t) r[7] + k]);
      const uint8_t code = f ? (uint8_t) (f - tab) : 15;
      sq[k >> 1] |= (k & 1) ? code : (uint8_t) (code << 4);
    }
    if (tagflags && (tagflags[i] & 1)) {
      m.data.push_back('M'); m.data.push_back('M'); m.data.push_back('Z');
      m.data.insert(m.data.end(), (const uint8_t*) mm_arena + mm_off[i], (const uint8_t*) mm_arena + mm_off[i + 1]);
      m.data.push_back(0);
    }
    if (tagflags && (tagflags[i] & 2)) {
      m.data.push_back('M'); m.data.push_back('L'); m.data.push_back('B'); m.data.push_back('C');
      int32_t nml = (int32_t) (ml_off[i + 1] - ml_off[i]); uint8_t b4[4]; memcpy(b4, &nml, 4); m.data.insert(m.data.end(), b4, b4 + 4);
      m.data.insert(m.data.end(), ml_arena + ml_off[i], ml_arena + ml_off[i + 1]);
    }
    int file = 0;
    while ((uint32_t) i >= file_off[file + 1]) ++file;
    (file == 0 ? g_bam : g_more[(size_t) file - 1]).push_back(m);
  }
  std::vector<std::vector<MemInterval> > validRegions(ncontig);
  for (int t = 0; t < ncontig; ++t) validRegions[t].push_back(MemInterval{0u, 0x7fffffffu});
  typedef std::vector<torali::StructuralVariantRecord> TVariants;
  TVariants svs;
  std::streambuf* old = std::cerr.rdbuf(nullptr);
  {
    TVariants svc;
    typedef std::map<std::pair<int32_t, std::size_t>, std::vector<torali::SeqSlice> > TPosReadSV;
    std::vector<TPosReadSV> srStore(c.nchr, TPosReadSV());
    torali::_clusterSRReads(c, validRegions, svc, srStore);
    torali::assemble(c, validRegions, svc, srStore);
    sort(svc.begin(), svc.end());
    std::map<int32_t, torali::StructuralVariantRecord> lastSVperType;
    for (TVariants::iterator svIter = svc.begin(); svIter != svc.end(); ++svIter) {
      if ((svIter->srSupport == 0) && (svIter->peSupport == 0)) continue;
      if (!svs.empty()) {
        std::map<int32_t, torali::StructuralVariantRecord>::const_iterator ltIt = lastSVperType.find(svIter->svt);
        if (ltIt != lastSVperType.end()) {
          torali::StructuralVariantRecord const& lastSV = ltIt->second;
          if ((lastSV.chr == svIter->chr) && (lastSV.chr2 == svIter->chr2) && (std::abs(svIter->svStart - lastSV.svStart) < c.minRefSep) && (std::abs(svIter->svEnd - lastSV.svEnd) < c.minRefSep)) {
            int32_t len1 = (svIter->svt == 4) ? svIter->insLen : (svIter->svEnd - svIter->svStart);
            int32_t len2 = (lastSV.svt == 4) ? lastSV.insLen : (lastSV.svEnd - lastSV.svStart);
            int32_t lengthvar = std::min(0.1 * len1, 0.1 * len2);
            int32_t lengthdiff = std::abs(len1 - len2);
            if (lengthvar < 15) lengthvar = 15;
            if (lengthdiff < lengthvar) continue;
          }
        }
      }
      lastSVperType[svIter->svt] = *svIter;
      svs.push_back(*svIter);
    }
    sort(svs.begin(), svs.end());
    uint32_t cliqueCount = 0;
    for (TVariants::iterator svIt = svs.begin(); svIt != svs.end(); ++svIt, ++cliqueCount) svIt->id = cliqueCount;
  }
  std::vector<std::vector<torali::JunctionCount> > jctMap((size_t) nfile);
  std::vector<std::vector<torali::ReadCount> > rcMap((size_t) nfile);
  std::vector<std::vector<torali::MethylInfo> > methylMap((size_t) nfile);
  for (int f = 0; f < nfile; ++f) { jctMap[f].resize(svs.size(), torali::JunctionCount()); rcMap[f].resize(svs.size()); methylMap[f].resize(svs.size(), torali::MethylInfo()); }
  torali::genotypeLR(c, svs, jctMap, rcMap, methylMap);
  std::cerr.rdbuf(old);
  const int n = (int) svs.size();
  if (n > cap) return -1;
  static torali::BoLog<double> bl;
  for (int i = 0; i < n; ++i) {
    torali::StructuralVariantRecord const& v = svs[i];
    int32_t* o = sv_out + 20 * i;
    o[0] = v.chr; o[1] = v.svStart; o[2] = v.chr2; o[3] = v.svEnd; o[4] = v.ciposlow; o[5] = v.ciposhigh; o[6] = v.ciendlow; o[7] = v.ciendhigh;
    o[8] = v.peSupport; o[9] = v.srSupport; o[10] = v.peMapQuality; o[11] = v.srMapQuality; o[12] = v.mapq; o[13] = v.insLen; o[14] = v.homLen; o[15] = v.svt;
    o[16] = v.precise ? 1 : 0; o[17] = v.consBp; o[18] = v.id;
    memcpy(o + 19, &v.srAlignQuality, 4);
    for (int f = 0; f < nfile; ++f) {
    float gls[3]; int32_t gq[1]; int32_t gts[2];
    std::vector<uint8_t> none;
    if (v.precise) torali::_computeGLs(bl, jctMap[f][v.id].ref, jctMap[f][v.id].alt, gls, gq, gts, 0);
    else torali::_computeGLs(bl, none, none, gls, gq, gts, 0);
    int32_t* q = fmt_out + 14 * ((size_t) f * cap + i);
    q[0] = gts[0]; q[1] = gts[1]; q[2] = gq[0];
    for (int k = 0; k < 3; ++k) q[3 + k] = (gts[0] == bcf_gt_missing) ? bcf_int32_missing : (int32_t) std::max(0.0f, std::round(-10.0f * gls[k]));
    torali::ReadCount const& rcv = rcMap[f][v.id];
    int32_t cnest = -1;
    if ((rcv.leftRC + rcv.rightRC) > 0) {
      double cn = 2.0 * (double) rcv.rc / (double) (rcv.leftRC + rcv.rightRC);
      if (cn < 0) cn = 0;
      if (cn > 100000) cn = 100000;
      cnest = boost::math::iround(cn);
    }
    q[6] = cnest; q[7] = (gq[0] < 15) ? 0 : 1;
    q[8] = 0; q[9] = 0; q[10] = (int32_t) jctMap[f][v.id].ref.size(); q[11] = (int32_t) jctMap[f][v.id].alt.size();
    q[12] = rcv.rc; q[13] = rcv.leftRC + rcv.rightRC;
    for (int k = 0; k < 3; ++k) gl_out[3 * ((size_t) f * cap + i) + k] = gls[k];
    }
    cons_len[i] = (int32_t) v.consensus.size();
    memcpy(cons_out + (size_t) i * cons_stride, v.consensus.data(), std::min<size_t>(v.consensus.size(), cons_stride));
    if (anno_out) {
      int32_t* a = anno_out + 5 * i;
      a[0] = v.anno.isRC ? 1 : 0; a[1] = v.anno.seqType; a[2] = v.anno.homLen; a[3] = v.anno.trPeriod; memcpy(a + 4, &v.anno.trCopies, 4);
    }
    if (methyl_out) {
      torali::MethylInfo const& mi = methylMap[0][v.id];
      const int32_t mv[16] = {mi.altSvStartL, mi.altSvStartR, mi.altSvRightL, mi.altSvRightR, mi.refSvStartL, mi.refSvStartR, mi.refSvRightL, mi.refSvRightR,
                              mi.mncStartL, mi.mncStartR, mi.mncRightL, mi.mncRightR, mi.mdpStartL, mi.mdpStartR, mi.mdpRightL, mi.mdpRightR};
      memcpy(methyl_out + 16 * i, mv, sizeof(mv));
    }
    if (alleles_out) {
      alleles_len[i] = (int32_t) v.alleles.size();
      memcpy(alleles_out + (size_t) i * alleles_stride, v.alleles.data(), std::min<size_t>(v.alleles.size(), alleles_stride));
    }
  }
  g_more.clear();
  return n;
}

int ref_delly_lr_call(const char* contig_arena, const uint32_t* contig_off, const uint32_t* contig_len, int ncontig, const int32_t* rec12, int nrec,
                      const uint32_t* cigar, const char* reads, const int32_t* cfg12, float flankQuality, float indelExtension, int32_t* sv_out, int cap,
                      int32_t* fmt_out, float* gl_out, char* cons_out, int cons_stride, int32_t* cons_len) {
  const uint32_t file_off[2] = {0u, (uint32_t) nrec};
  return run_delly_lr_call(contig_arena, contig_off, contig_len, ncontig, rec12, file_off, 1, cigar, reads, cfg12, flankQuality, indelExtension, sv_out, cap, fmt_out, gl_out,
                           cons_out, cons_stride, cons_len, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0.f, 0.f, nullptr, nullptr, nullptr, 0, nullptr);
}

// several long-read files in one call set: records grouped by file (file_off[nfile + 1]); fmt_out / gl_out file-major (`cap` rows per file)
int ref_delly_lr_call_multi(const char* contig_arena, const uint32_t* contig_off, const uint32_t* contig_len, int ncontig, const int32_t* rec12, const uint32_t* file_off,
                            int nfile, const uint32_t* cigar, const char* reads, const int32_t* cfg12, float flankQuality, float indelExtension, int32_t* sv_out, int cap,
                            int32_t* fmt_out, float* gl_out, char* cons_out, int cons_stride, int32_t* cons_len) {
  return run_delly_lr_call(contig_arena, contig_off, contig_len, ncontig, rec12, file_off, nfile, cigar, reads, cfg12, flankQuality, indelExtension, sv_out, cap, fmt_out,
                           gl_out, cons_out, cons_stride, cons_len, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0.f, 0.f, nullptr, nullptr, nullptr, 0, nullptr);
}

int ref_delly_lr_call_ex(const char* contig_arena, const uint32_t* contig_off, const uint32_t* contig_len, int ncontig, const int32_t* rec12, int nrec,
                         const uint32_t* cigar, const char* reads, const int32_t* cfg12, float flankQuality, float indelExtension, int32_t* sv_out, int cap,
                         int32_t* fmt_out, float* gl_out, char* cons_out, int cons_stride, int32_t* cons_len, const uint8_t* tagflags, const char* mm_arena,
                         const uint32_t* mm_off, const uint8_t* ml_arena, const uint32_t* ml_off, int methylWindow, int methylProb, int minCpgDepth, float meiMinFrac,
                         float trMinFrac, int32_t* anno_out, int32_t* methyl_out, char* alleles_out, int alleles_stride, int32_t* alleles_len) {
  const uint32_t file_off[2] = {0u, (uint32_t) nrec};
  return run_delly_lr_call(contig_arena, contig_off, contig_len, ncontig, rec12, file_off, 1, cigar, reads, cfg12, flankQuality, indelExtension, sv_out, cap, fmt_out, gl_out,
                           cons_out, cons_stride, cons_len, tagflags, mm_arena, mm_off, ml_arena, ml_off, methylWindow, methylProb, minCpgDepth, meiMinFrac, trMinFrac,
                           anno_out, methyl_out, alleles_out, alleles_stride, alleles_len);
}

// _clusterSRReads (src/junction.h:495-623, no alternate alignments) over in-memory alignments: clustered SVs and the read store.
//   out: sv_out cap x 14 (as ref_scan_pe_sr's sr_out); store: per contig ascending (pos, read id), one row per SeqSlice:
//        [refIndex, pos, svid, sstart, inslen, qual] + read id (uint64)
int ref_cluster_sr_reads(const uint32_t* contig_len, int ncontig, const int32_t* rec12, int nrec, const uint32_t* cigar, const int32_t* cfg12, float indelExtension,
                         int32_t* sv_out, int cap, int32_t* store_out, uint64_t* store_seed, int store_cap, int32_t* n_out) {
  RefConfig5 c;
  c.files.push_back(boost::filesystem::path("in-memory.bam")); c.genome = boost::filesystem::path("in-memory.fa");
  c.minMapQual = (uint16_t) cfg12[0]; c.minClip = (uint32_t) cfg12[1]; c.minRefSep = (uint32_t) cfg12[2]; c.maxReadSep = (uint32_t) cfg12[3];
  c.minCliqueSize = (uint16_t) cfg12[4]; c.graphPruning = (uint32_t) cfg12[5]; c.indelExtension = indelExtension; c.nchr = ncontig;
  g_names.clear(); g_tlen.clear(); g_seq.clear(); g_name_ptrs.clear();
  for (int k = 0; k < ncontig; ++k) { g_names.push_back("chr" + std::to_string(k)); g_tlen.push_back(contig_len[k]); g_seq.push_back(nullptr); }
  for (auto& nm : g_names) g_name_ptrs.push_back((char*) nm.c_str());
  g_bam.clear();
  for (int i = 0; i < nrec; ++i) {
    const int32_t* r = rec12 + 12 * i;
    MemRecord5 m; memset(&m.core, 0, sizeof(m.core));
    std::string qn = "q" + std::to_string(r[11]);
    m.core.tid = r[0]; m.core.pos = r[1]; m.core.flag = (uint16_t) r[2]; m.core.qual = (uint8_t) r[3]; m.core.l_qseq = r[4]; m.core.n_cigar = (uint32_t) r[6];
    m.core.l_qname = (uint16_t) ((qn.size() + 1 + 3) & ~3u);
    m.data.assign(m.core.l_qname + 4 * (std::size_t) r[6], 0);
    memcpy(m.data.data(), qn.data(), qn.size());
    memcpy(m.data.data() + m.core.l_qname, cigar + r[5], 4 * (std::size_t) r[6]);
    g_bam.push_back(m);
  }
  std::vector<std::vector<MemInterval> > validRegions(ncontig);
  for (int t = 0; t < ncontig; ++t) validRegions[t].push_back(MemInterval{0u, 0x7fffffffu});
  std::vector<torali::StructuralVariantRecord> svc;
  typedef std::map<std::pair<int32_t, std::size_t>, std::vector<torali::SeqSlice> > TPosReadSV;
  std::vector<TPosReadSV> srStore(c.nchr, TPosReadSV());
  std::streambuf* old = std::cerr.rdbuf(nullptr);
  torali::_clusterSRReads(c, validRegions, svc, srStore);
  std::cerr.rdbuf(old);
  if ((int) svc.size() > cap) return -1;
  for (std::size_t i = 0; i < svc.size(); ++i) {
    int32_t* o = sv_out + 14 * i; torali::StructuralVariantRecord const& v = svc[i];
    o[0] = v.chr; o[1] = v.svStart; o[2] = v.chr2; o[3] = v.svEnd; o[4] = v.ciposlow; o[5] = v.ciposhigh; o[6] = v.ciendlow; o[7] = v.ciendhigh;
    o[8] = v.srSupport; o[9] = v.srMapQuality; o[10] = v.mapq; o[11] = v.insLen; o[12] = v.svt; o[13] = v.id;
  }
  int k = 0;
  for (int t = 0; t < ncontig; ++t)
    for (auto const& kv : srStore[t])
      for (auto const& sl : kv.second) {
        if (k >= store_cap) return -1;
        int32_t* o = store_out + 6 * k;
        o[0] = t; o[1] = kv.first.first; o[2] = sl.svid; o[3] = sl.sstart; o[4] = sl.inslen; o[5] = sl.qual;
        store_seed[k++] = (uint64_t) kv.first.second;
      }
  n_out[0] = (int32_t) svc.size(); n_out[1] = k;
  return 0;
}

// read ids as restated above
uint64_t ref_hash_lr_name5(const char* qname) { return (uint64_t) std::hash<std::string>()(qname); }
// the read id scanPEandSR derives from a query name and the mate flag (hash_sr as restated above)
uint64_t ref_hash_sr_name(const char* qname, int read2) { return (uint64_t) (std::hash<std::string>()(qname) * 2 + (read2 ? 1 : 0)); }

}  // extern "C"
