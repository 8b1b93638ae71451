// oracle/_ref wrapper, part 3 (TEST INFRASTRUCTURE ONLY): the reference's genotype-likelihood code (src/bolog.h) and
// the probe generation of its genotyping pass (src/coverage.h:117-263) compiled VERBATIM from /root/reference/src.
// util.h / pangenome.h are masked by their include guards (as in ref_wrap2.cpp); the util.h symbols coverage.h names
// are restated below (or in shim/prelude.h) with their reference lines; the three faidx calls _generateProbes makes are served from memory
// (no htslib library is linked). Nothing from the reference is copied into this repository.
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
};
}  // namespace filesystem
}  // namespace boost

#include <htslib/faidx.h>
#include <htslib/vcf.h>
#include <htslib/sam.h>
#include "tags.h"
#include "threadpool.h"

namespace torali {
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
// util.h:253-272 (boost::lexical_cast<std::string>(int) == std::to_string for int32)
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
// util.h:69-76
struct ReadCount {
  int32_t leftRC, rc, rightRC;
  ReadCount() : leftRC(0), rc(0), rightRC(0) {}
  ReadCount(int32_t l, int32_t m, int32_t r) : leftRC(l), rc(m), rightRC(r) {}
};
// util.h:440-451
inline uint32_t alignmentLength(bam1_t const* rec) {
  uint32_t const* cigar = bam_get_cigar(rec);
  uint32_t alen = 0;
  for (std::size_t i = 0; i < rec->core.n_cigar; ++i)
    if ((bam_cigar_op(cigar[i]) == BAM_CMATCH) || (bam_cigar_op(cigar[i]) == BAM_CEQUAL) || (bam_cigar_op(cigar[i]) == BAM_CDIFF) || (bam_cigar_op(cigar[i]) == BAM_CDEL) || (bam_cigar_op(cigar[i]) == BAM_CREF_SKIP)) alen += bam_cigar_oplen(cigar[i]);
  return alen;
}
inline uint32_t halfAlignmentLength(bam1_t const* rec) { return (alignmentLength(rec) / 2); }
// util.h:29-41
struct LibraryInfo {
  int32_t rs, median, mad, minNormalISize, minISizeCutoff, maxNormalISize, maxISizeCutoff;
  uint32_t abnormal_pairs;
  LibraryInfo() : rs(0), median(0), mad(0), minNormalISize(0), minISizeCutoff(0), maxNormalISize(0), maxISizeCutoff(0), abnormal_pairs(0) {}
};
// util.h:501-517 — opaque fragment ids of the spanning-pair bookkeeping (boost::hash_combine there; any mixing works for the
// junction-read counts checked against this wrapper, which do not depend on them)
inline std::size_t _mix(std::size_t seed, std::size_t v) { return seed ^ (v + 0x9e3779b97f4a7c15ull + (seed << 6) + (seed >> 2)); }
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
}  // namespace torali

#define MAX_CN 10
#include "bolog.h"
#include "coverage.h"

namespace {
struct RefConfig3 {
  torali::DnaScore<int> aliscore;
  uint16_t minCliqueSize;
  float flankQuality;
  int32_t minimumFlankSize, indelsize, minConsWindow;
  boost::filesystem::path genome;
  // annotateCoverage (src/coverage.h:265-743)
  std::vector<boost::filesystem::path> files;
  boost::filesystem::path dumpfile;
  bool hasDumpFile = false;
  uint32_t maxThreads = 1, maxGenoReadCount = 250;
  uint16_t minGenoQual = 5;
};
struct MemRecord3 { bam1_core_t core; std::vector<uint8_t> data; };
std::vector<MemRecord3> g_bam;   // in-memory alignments, sorted by (tid, pos)
uint32_t g_tlen[2];
char* g_tname[2];
// in-memory FASTA for the three faidx calls
const char* g_seq[2];
int g_len[2];
const char* g_names[2] = {"chrA", "chrB"};
}  // namespace

extern "C" {

faidx_t* fai_load(const char*) { return (faidx_t*) g_names; }
void fai_destroy(faidx_t*) {}
char* faidx_fetch_seq(const faidx_t*, const char* name, int beg, int end, int* len) {
  const int k = strcmp(name, g_names[0]) == 0 ? 0 : 1;
  if (end >= g_len[k]) end = g_len[k] - 1;
  const int n = end - beg + 1;
  char* out = (char*) malloc((size_t) n + 1);
  memcpy(out, g_seq[k] + beg, (size_t) n);
  out[n] = 0;
  *len = n;
  return out;
}

// _computeGLs (src/bolog.h:25-85) with BoLog<double> (src/modvcf.h:373), one sample
void ref_compute_gls(const uint8_t* refq, int nref, const uint8_t* altq, int nalt, float* gls3, int32_t* gq, int32_t* gts2) {
  static torali::BoLog<double> bl;
  std::vector<uint8_t> r(refq, refq + nref), a(altq, altq + nalt);
  torali::_computeGLs(bl, r, a, gls3, gq, gts2, 0);
}

// _generateProbes (src/coverage.h:164-263) on the two-contig toy header of ref_wrap.cpp.
// sv_in: n x [chr, svStart, chr2, svEnd, svt, insLen, precise, id]; cons arena/off/len.
// out: probe strings into `arena` in the order (bp0 cons, bp0 ref, bp1 cons, bp1 ref) per SV id -> p_off/p_len [n][4];
//      regions [cap][9] = chr, regionStart, regionEnd, bppos, homLeft, homRight, svt, id, bpPoint (per contig, sorted as the reference sorts);
//      alleles [n][256] + lengths; svOnChr[2]. Returns the number of regions, -1 on capacity, -2 if the reference threw.
int ref_generate_probes(const char* seq0, int len0, const char* seq1, int len1, int n, const int32_t* sv_in, const uint8_t* cons_arena,
                        const uint32_t* cons_off, const uint32_t* cons_len, float flankQuality, int minimumFlankSize, int indelsize,
                        int minConsWindow, uint8_t* arena, uint64_t arena_cap, uint64_t* p_off, uint32_t* p_len, int32_t* regions, int cap,
                        uint8_t* alleles, int32_t* alleles_len, uint8_t* svOnChrOut) {
  RefConfig3 c; c.flankQuality = flankQuality; c.minimumFlankSize = minimumFlankSize; c.indelsize = indelsize; c.minConsWindow = minConsWindow;
  c.genome = boost::filesystem::path("in-memory");
  g_seq[0] = seq0; g_seq[1] = seq1; g_len[0] = len0; g_len[1] = len1;
  bam_hdr_t hdr; memset(&hdr, 0, sizeof(hdr));
  uint32_t tlen[2] = {(uint32_t) len0, (uint32_t) len1};
  char n0[] = "chrA"; char n1[] = "chrB"; char* names[2] = {n0, n1};
  hdr.n_targets = 2; hdr.target_len = tlen; hdr.target_name = names;
  std::vector<torali::StructuralVariantRecord> svs(n);
  for (int i = 0; i < n; ++i) {
    const int32_t* s = sv_in + 8 * i;
    svs[i].chr = s[0]; svs[i].svStart = s[1]; svs[i].chr2 = s[2]; svs[i].svEnd = s[3]; svs[i].svt = s[4]; svs[i].insLen = s[5];
    svs[i].precise = s[6] != 0; svs[i].id = s[7];
    svs[i].consensus = std::string((const char*) cons_arena + cons_off[i], cons_len[i]);
  }
  typedef std::vector<std::string> TProbes;
  std::vector<TProbes> refProbeArr(2, TProbes(n)), consProbeArr(2, TProbes(n));
  std::vector<std::vector<torali::BpRegion> > bpRegion(2);
  std::vector<bool> svOnChr(2, false);
  std::streambuf* old = std::cerr.rdbuf(nullptr);  // the reference logs a timestamp line
  try {
    torali::_generateProbes(c, &hdr, svs, refProbeArr, consProbeArr, bpRegion, svOnChr);
  } catch (std::exception const&) {
    std::cerr.rdbuf(old);
    return -2;
  }
  std::cerr.rdbuf(old);
  uint64_t pos = 0;
  for (int i = 0; i < n; ++i) {
    const std::string* p[4] = {&consProbeArr[0][i], &refProbeArr[0][i], &consProbeArr[1][i], &refProbeArr[1][i]};
    for (int k = 0; k < 4; ++k) {
      if (pos + p[k]->size() > arena_cap) return -1;
      p_off[4 * i + k] = pos; p_len[4 * i + k] = (uint32_t) p[k]->size();
      memcpy(arena + pos, p[k]->data(), p[k]->size());
      pos += p[k]->size();
    }
    alleles_len[i] = (int32_t) svs[i].alleles.size();
    memcpy(alleles + 256 * (size_t) i, svs[i].alleles.data(), std::min<std::size_t>(256, svs[i].alleles.size()));
  }
  int r = 0;
  for (int chr = 0; chr < 2; ++chr) {
    svOnChrOut[chr] = svOnChr[chr] ? 1 : 0;
    for (auto const& b : bpRegion[chr]) {
      if (r >= cap) return -1;
      int32_t* o = regions + 9 * r++;
      o[0] = chr; o[1] = b.regionStart; o[2] = b.regionEnd; o[3] = b.bppos; o[4] = b.homLeft; o[5] = b.homRight; o[6] = b.svt; o[7] = (int32_t) b.id; o[8] = b.bpPoint;
    }
  }
  return r;
}

// ---- htslib stand-ins for annotateCoverage's BAM scan (no htslib library is linked) --------------------------------------
htsFile* hts_open(const char*, const char*) { htsFile* f = (htsFile*) calloc(1, sizeof(htsFile)); f->is_bgzf = 1; return f; }
int hts_close(htsFile* f) { free(f); return 0; }
int hts_set_fai_filename(htsFile*, const char*) { return 0; }
hts_idx_t* sam_index_load(htsFile*, const char*) { return (hts_idx_t*) &g_bam; }
void hts_idx_destroy(hts_idx_t*) {}
int hts_idx_get_stat(const hts_idx_t*, int, uint64_t* mapped, uint64_t* unmapped) { *mapped = 1; *unmapped = 0; return 0; }
sam_hdr_t* sam_hdr_read(samFile*) {
  sam_hdr_t* h = (sam_hdr_t*) calloc(1, sizeof(sam_hdr_t));
  h->n_targets = 2; h->target_len = g_tlen; h->target_name = g_tname;
  return h;
}
void sam_hdr_destroy(sam_hdr_t* h) { free(h); }
hts_itr_t* sam_itr_queryi(const hts_idx_t*, int tid, hts_pos_t beg, hts_pos_t end) {
  hts_itr_t* it = (hts_itr_t*) calloc(1, sizeof(hts_itr_t));
  it->tid = tid; it->beg = beg; it->end = end; it->i = 0;
  return it;
}
int hts_itr_next(BGZF*, hts_itr_t* it, void* r, void*) {
  bam1_t* b = (bam1_t*) r;
  while (it->i < (int) g_bam.size()) {
    MemRecord3& m = g_bam[it->i++];
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

// annotateCoverage (src/coverage.h:265-743) over in-memory alignments; reports the junction-read counts (countMap).
//   rec: nrec x [tid, pos, flag, mapq, l_qseq, cigar_off, n_cigar, seq_off, mtid, mpos, isize, name id], sorted by (tid, pos); cigar BAM-encoded;
//        records with the same name id are mates (same query name);
//   sv:  nsv x [chr, svStart, chr2, svEnd, svt, insLen, precise, id, peSupport] + consensus arena
//   lib: [median, minNormalISize, maxNormalISize, maxISizeCutoff] of the sample's library (src/util.h:29-41)
//   out: per SV the junction REF and ALT quality lists (concatenated; offsets nsv+1 each), then likewise the spanning-pair REF and ALT
//        lists (span_out / sref_off / salt_off) and the read-depth triple (leftRC, rc, rightRC). Returns the number of junction qualities.
int ref_annotate_junction_reads(const char* seq0, int len0, const char* seq1, int len1, const int32_t* rec10, int nrec, const uint32_t* cigar, const char* reads,
                                const int32_t* sv9, int nsv, const uint8_t* cons_arena, const uint32_t* cons_off, const uint32_t* cons_len, float flankQuality,
                                int minimumFlankSize, int indelsize, int minConsWindow, int minGenoQual, int maxGenoReadCount, int maxThreads,
                                uint8_t* qual_out, int qual_cap, uint32_t* ref_off, uint32_t* alt_off, const int32_t* lib4, uint8_t* span_out, int span_cap,
                                uint32_t* sref_off, uint32_t* salt_off, int32_t* rc_out) {
  RefConfig3 c; c.flankQuality = flankQuality; c.minimumFlankSize = minimumFlankSize; c.indelsize = indelsize; c.minConsWindow = minConsWindow;
  c.genome = boost::filesystem::path("in-memory"); c.files.push_back(boost::filesystem::path("in-memory.bam"));
  c.minGenoQual = (uint16_t) minGenoQual; c.maxGenoReadCount = (uint32_t) maxGenoReadCount; c.maxThreads = (uint32_t) maxThreads;
  g_seq[0] = seq0; g_seq[1] = seq1; g_len[0] = len0; g_len[1] = len1;
  g_tlen[0] = (uint32_t) len0; g_tlen[1] = (uint32_t) len1;
  static char n0[] = "chrA"; static char n1[] = "chrB"; g_tname[0] = n0; g_tname[1] = n1;
  g_bam.clear();
  for (int i = 0; i < nrec; ++i) {
    const int32_t* r = rec10 + 12 * i;
    MemRecord3 m; memset(&m.core, 0, sizeof(m.core));
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
    const int32_t* s = sv9 + 9 * i;
    svs[i].chr = s[0]; svs[i].svStart = s[1]; svs[i].chr2 = s[2]; svs[i].svEnd = s[3]; svs[i].svt = s[4]; svs[i].insLen = s[5];
    svs[i].precise = s[6] != 0; svs[i].id = s[7]; svs[i].peSupport = s[8];
    svs[i].consensus = std::string((const char*) cons_arena + cons_off[i], cons_len[i]);
  }
  std::vector<torali::LibraryInfo> sampleLib(1);
  sampleLib[0].rs = 150; sampleLib[0].median = lib4[0]; sampleLib[0].mad = 20; sampleLib[0].minNormalISize = lib4[1]; sampleLib[0].maxNormalISize = lib4[2];
  sampleLib[0].minISizeCutoff = 50; sampleLib[0].maxISizeCutoff = lib4[3];
  std::vector<std::vector<torali::ReadCount> > covCount;
  std::vector<std::vector<torali::JunctionCount> > countMap;
  std::vector<std::vector<torali::SpanningCount> > spanMap;
  std::streambuf* old = std::cerr.rdbuf(nullptr);
  try {
    torali::annotateCoverage(c, sampleLib, svs, covCount, countMap, spanMap);
  } catch (std::exception const&) {
    std::cerr.rdbuf(old);
    return -2;
  }
  std::cerr.rdbuf(old);
  int pos = 0;
  for (int pass = 0; pass < 2; ++pass) {
    uint32_t* off = pass ? alt_off : ref_off;
    for (int i = 0; i < nsv; ++i) {
      off[i] = (uint32_t) pos;
      std::vector<uint8_t> const& v = pass ? countMap[0][i].alt : countMap[0][i].ref;
      if (pos + (int) v.size() > qual_cap) return -1;
      for (uint8_t q : v) qual_out[pos++] = q;
    }
    off[nsv] = (uint32_t) pos;
  }
  int sp = 0;
  for (int pass = 0; pass < 2; ++pass) {
    uint32_t* off = pass ? salt_off : sref_off;
    for (int i = 0; i < nsv; ++i) {
      off[i] = (uint32_t) sp;
      std::vector<uint8_t> const& v = pass ? spanMap[0][i].alt : spanMap[0][i].ref;
      if (sp + (int) v.size() > span_cap) return -1;
      for (uint8_t q : v) span_out[sp++] = q;
    }
    off[nsv] = (uint32_t) sp;
  }
  for (int i = 0; i < nsv; ++i) { rc_out[3 * i] = covCount[0][i].leftRC; rc_out[3 * i + 1] = covCount[0][i].rc; rc_out[3 * i + 2] = covCount[0][i].rightRC; }
  return pos;
}

}  // extern "C"
