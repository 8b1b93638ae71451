// oracle/_ref wrapper, part 4 (TEST INFRASTRUCTURE ONLY): the reference's long-read genotyping pass
// (src/genotype.h:22-397: _editDistanceNW, _readStart/_readEnd/_findSeqBp, genotypeLR) compiled VERBATIM from
// /root/reference/src and run over an IN-MEMORY alignment list: the htslib calls genotypeLR makes (open / index / header /
// iterator / record / aux tags / faidx) are served by the small stand-ins below, so no htslib library is linked and no
// BAM file is needed. util.h / pangenome.h are masked by their include guards as in ref_wrap2/3.cpp; annotateSV
// (src/svanno.h, mobile-element annotation: SURVEY §8f row 4, not on the path yet) is a no-op here.
// Nothing from the reference is copied into this repository.
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
// boost::split(container, string, is_any_of(chars)) as src/methyl.h uses it on the MM tag (no token compression: n separators give n + 1 tokens)
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
// src/svanno.h annotateSV: mobile-element annotation, not on the path (see the header comment)
template <typename TConfig> inline void annotateSV(TConfig const&, bam_hdr_t*, char const*, StructuralVariantRecord&) {}
}  // namespace torali

#define MAX_CN 10
#include "bolog.h"
#include "coverage.h"
#include "genotype.h"

namespace {
struct RefConfig4 {   // the fields genotypeLR and its callees read from TConfig (src/tegua.h:39-74)
  std::vector<boost::filesystem::path> files;
  boost::filesystem::path genome, dumpfile;
  bool hasDumpFile = false;
  uint16_t minMapQual = 1;
  int32_t minimumFlankSize = 100, minConsWindow = 1000, genoCap = 25;
  uint32_t maxGenoReadCount = 250;
  float flankQuality = 0.9f;
  uint32_t methylProb = 128, minCpgDepth = 1;
  int32_t methylWindow = 500;
  float indelExtension = 0.5f;
  int32_t indelsize = 10000;
};

// ---- the in-memory "BAM" and "FASTA" -------------------------------------------------------------------------------
struct MemRecord { bam1_core_t core; std::vector<uint8_t> data; };
std::vector<MemRecord> g_records;        // file order (sorted by tid, pos like a coordinate-sorted BAM)
std::vector<std::string> g_names;        // contig names
std::vector<uint32_t> g_tlen;
std::vector<const char*> g_seq;
std::vector<char*> g_name_ptrs;
}  // namespace

extern "C" {

// htslib stand-ins (declared in htslib/sam.h, hts.h, faidx.h)
htsFile* hts_open(const char*, const char*) { htsFile* f = (htsFile*) calloc(1, sizeof(htsFile)); f->is_bgzf = 1; return f; }
int hts_close(htsFile* f) { free(f); return 0; }
int hts_set_fai_filename(htsFile*, const char*) { return 0; }
hts_idx_t* sam_index_load(htsFile*, const char*) { return (hts_idx_t*) &g_records; }
void hts_idx_destroy(hts_idx_t*) {}
sam_hdr_t* sam_hdr_read(samFile*) {
  sam_hdr_t* h = (sam_hdr_t*) calloc(1, sizeof(sam_hdr_t));
  h->n_targets = (int32_t) g_names.size(); h->target_len = g_tlen.data(); h->target_name = g_name_ptrs.data();
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
  while (it->i < (int) g_records.size()) {
    MemRecord& m = g_records[it->i++];
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
void bam_destroy1(bam1_t* b) { free(b); }   // the record data belongs to g_records
void hts_log(enum htsLogLevel, const char*, const char*, ...) {}
// aux tags: the encodings the test records use ('C' uint8, 'i' int32, 'Z' string, 'B:C' byte array)
uint8_t* bam_aux_get(const bam1_t* b, const char tag[2]) {
  uint8_t* p = bam_get_aux(b);
  uint8_t* end = b->data + b->l_data;
  while (p + 3 <= end) {
    const bool hit = (p[0] == (uint8_t) tag[0] && p[1] == (uint8_t) tag[1]);
    const uint8_t type = p[2];
    if (hit) return p + 2;
    if (type == 'Z') p += 3 + strlen((const char*) p + 3) + 1;
    else if (type == 'B') { int32_t n; memcpy(&n, p + 4, 4); p += 3 + 1 + 4 + (size_t) n; }   // subtype C only (1 byte per element)
    else p += 3 + (type == 'C' ? 1 : 4);
  }
  return NULL;
}
int64_t bam_aux2i(const uint8_t* s) {
  if (s[0] == 'C') return s[1];
  int32_t v; memcpy(&v, s + 1, 4); return v;
}
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

// genotypeLR (src/genotype.h:93-397) over in-memory data.
//   contigs: ncontig sequences (seq arena + offsets/lengths), named "chr0", "chr1", ...
//   rec: nrec x [tid, pos, flag, mapq, l_qseq, cigar_off, n_cigar, seq_off, hp (0 = no tag), ps (-1 = no tag)], sorted by (tid, pos);
//        cigar BAM-encoded; read bases as ASCII in `reads` at seq_off
//   sv:  nsv x [chr, svStart, chr2, svEnd, svt, insLen, consBp, id] + consensus arena
//   out: per SV: ref / alt quality lists (concatenated, with offsets), hp1/hp2 counts, ps, (leftRC, rc, rightRC)
// methylation (optional, tagflags != NULL): per record bit 0 = MM:Z present (text at mm_arena[mm_off[i] .. mm_off[i+1])), bit 1 = ML:B:C present
// (bytes at ml_arena[ml_off[i] .. ml_off[i+1])); methyl_out: nsv x 16 = MethylInfo in field order (alt x4, ref x4, mnc x4, mdp x4)
static int run_genotype_lr(const char* contig_arena, const uint32_t* contig_off, const uint32_t* contig_len, int ncontig, const int32_t* rec10, int nrec,
                    const uint32_t* cigar, const char* reads, const int32_t* sv8, int nsv, const char* cons_arena, const uint32_t* cons_off,
                    const uint32_t* cons_len, int minMapQual, int minimumFlankSize, int minConsWindow, int maxGenoReadCount, float flankQuality, int genoCap,
                    uint8_t* qual_out, int qual_cap, uint32_t* ref_off /* nsv+1 */, uint32_t* alt_off /* nsv+1 */, int32_t* hp_out /* nsv x 5: hp1ref hp1alt hp2ref hp2alt ps */,
                    int32_t* rc_out /* nsv x 3 */, const uint8_t* tagflags, const char* mm_arena, const uint32_t* mm_off, const uint8_t* ml_arena,
                    const uint32_t* ml_off, int methylWindow, int methylProb, int minCpgDepth, int32_t* methyl_out) {
  RefConfig4 c;
  if (tagflags) { c.methylWindow = methylWindow; c.methylProb = (uint32_t) methylProb; c.minCpgDepth = (uint32_t) minCpgDepth; }
  c.files.push_back(boost::filesystem::path("in-memory.bam")); c.genome = boost::filesystem::path("in-memory.fa");
  c.minMapQual = (uint16_t) minMapQual; c.minimumFlankSize = minimumFlankSize; c.minConsWindow = minConsWindow;
  c.maxGenoReadCount = (uint32_t) maxGenoReadCount; c.flankQuality = flankQuality; c.genoCap = genoCap;
  g_names.clear(); g_tlen.clear(); g_seq.clear(); g_name_ptrs.clear();
  for (int k = 0; k < ncontig; ++k) { g_names.push_back("chr" + std::to_string(k)); g_tlen.push_back(contig_len[k]); g_seq.push_back(contig_arena + contig_off[k]); }
  for (auto& nm : g_names) g_name_ptrs.push_back((char*) nm.c_str());
  g_records.clear();
  for (int i = 0; i < nrec; ++i) {
    const int32_t* r = rec10 + 10 * i;
    MemRecord m; memset(&m.core, 0, sizeof(m.core));
    std::string qn = "r" + std::to_string(i);
    m.core.tid = r[0]; m.core.pos = r[1]; m.core.flag = (uint16_t) r[2]; m.core.qual = (uint8_t) r[3]; m.core.l_qseq = r[4]; m.core.n_cigar = (uint32_t) r[6];
    m.core.l_qname = (uint16_t) ((qn.size() + 1 + 3) & ~3u);
    const std::size_t lq = (std::size_t) r[4];
    m.data.assign(m.core.l_qname + 4 * (std::size_t) r[6] + (lq + 1) / 2 + lq, 0);
    memcpy(m.data.data(), qn.data(), qn.size());
    memcpy(m.data.data() + m.core.l_qname, cigar + r[5], 4 * (std::size_t) r[6]);
    uint8_t* sq = m.data.data() + m.core.l_qname + 4 * (std::size_t) r[6];
    for (std::size_t k = 0; k < lq; ++k) {
      const char ch = reads[(std::size_t) r[7] + k];
      const char* tab = "=ACMGRSVTWYHKDBN";
      const char* f = strchr(tab, ch);
      const uint8_t code = f ? (uint8_t) (f - tab) : 15;
      sq[k >> 1] |= (k & 1) ? code : (uint8_t) (code << 4);
    }
    if (r[8] > 0) { m.data.push_back('H'); m.data.push_back('P'); m.data.push_back('C'); m.data.push_back((uint8_t) r[8]); }
    if (r[9] >= 0) { m.data.push_back('P'); m.data.push_back('S'); m.data.push_back('i'); int32_t v = r[9]; uint8_t b4[4]; memcpy(b4, &v, 4); m.data.insert(m.data.end(), b4, b4 + 4); }
    if (tagflags && (tagflags[i] & 1)) {
      m.data.push_back('M'); m.data.push_back('M'); m.data.push_back('Z');
      m.data.insert(m.data.end(), (const uint8_t*) mm_arena + mm_off[i], (const uint8_t*) mm_arena + mm_off[i + 1]);
      m.data.push_back(0);
    }
    if (tagflags && (tagflags[i] & 2)) {
      m.data.push_back('M'); m.data.push_back('L'); m.data.push_back('B'); m.data.push_back('C');
      int32_t n = (int32_t) (ml_off[i + 1] - ml_off[i]); uint8_t b4[4]; memcpy(b4, &n, 4); m.data.insert(m.data.end(), b4, b4 + 4);
      m.data.insert(m.data.end(), ml_arena + ml_off[i], ml_arena + ml_off[i + 1]);
    }
    g_records.push_back(m);
  }
  std::vector<torali::StructuralVariantRecord> svs(nsv);
  for (int i = 0; i < nsv; ++i) {
    const int32_t* s = sv8 + 8 * i;
    svs[i].chr = s[0]; svs[i].svStart = s[1]; svs[i].chr2 = s[2]; svs[i].svEnd = s[3]; svs[i].svt = s[4]; svs[i].insLen = s[5]; svs[i].consBp = s[6]; svs[i].id = s[7];
    svs[i].consensus = std::string(cons_arena + cons_off[i], cons_len[i]);
    svs[i].precise = true;
  }
  std::vector<std::vector<torali::JunctionCount> > jctMap(1, std::vector<torali::JunctionCount>(nsv));
  std::vector<std::vector<torali::ReadCount> > covMap(1, std::vector<torali::ReadCount>(nsv));
  std::vector<std::vector<torali::MethylInfo> > methylMap(1, std::vector<torali::MethylInfo>(nsv));
  std::streambuf* old = std::cerr.rdbuf(nullptr);
  torali::genotypeLR(c, svs, jctMap, covMap, methylMap);
  std::cerr.rdbuf(old);
  int pos = 0;
  for (int pass = 0; pass < 2; ++pass) {
    uint32_t* off = pass ? alt_off : ref_off;
    for (int i = 0; i < nsv; ++i) {
      off[i] = (uint32_t) pos;
      std::vector<uint8_t> const& v = pass ? jctMap[0][i].alt : jctMap[0][i].ref;
      if (pos + (int) v.size() > qual_cap) return -1;
      for (uint8_t q : v) qual_out[pos++] = q;
    }
    off[nsv] = (uint32_t) pos;
  }
  for (int i = 0; i < nsv; ++i) {
    torali::JunctionCount const& j = jctMap[0][i];
    hp_out[5 * i] = (int32_t) j.hp1ref.size(); hp_out[5 * i + 1] = (int32_t) j.hp1alt.size(); hp_out[5 * i + 2] = (int32_t) j.hp2ref.size();
    hp_out[5 * i + 3] = (int32_t) j.hp2alt.size(); hp_out[5 * i + 4] = j.ps;
    rc_out[3 * i] = covMap[0][i].leftRC; rc_out[3 * i + 1] = covMap[0][i].rc; rc_out[3 * i + 2] = covMap[0][i].rightRC;
    if (methyl_out) {
      torali::MethylInfo const& mi = methylMap[0][i];
      const int32_t v[16] = {mi.altSvStartL, mi.altSvStartR, mi.altSvRightL, mi.altSvRightR, mi.refSvStartL, mi.refSvStartR, mi.refSvRightL, mi.refSvRightR,
                             mi.mncStartL, mi.mncStartR, mi.mncRightL, mi.mncRightR, mi.mdpStartL, mi.mdpStartR, mi.mdpRightL, mi.mdpRightR};
      memcpy(methyl_out + 16 * i, v, sizeof(v));
    }
  }
  return pos;
}

int ref_genotype_lr(const char* contig_arena, const uint32_t* contig_off, const uint32_t* contig_len, int ncontig, const int32_t* rec10, int nrec,
                    const uint32_t* cigar, const char* reads, const int32_t* sv8, int nsv, const char* cons_arena, const uint32_t* cons_off,
                    const uint32_t* cons_len, int minMapQual, int minimumFlankSize, int minConsWindow, int maxGenoReadCount, float flankQuality, int genoCap,
                    uint8_t* qual_out, int qual_cap, uint32_t* ref_off, uint32_t* alt_off, int32_t* hp_out, int32_t* rc_out) {
  return run_genotype_lr(contig_arena, contig_off, contig_len, ncontig, rec10, nrec, cigar, reads, sv8, nsv, cons_arena, cons_off, cons_len, minMapQual,
                         minimumFlankSize, minConsWindow, maxGenoReadCount, flankQuality, genoCap, qual_out, qual_cap, ref_off, alt_off, hp_out, rc_out, nullptr,
                         nullptr, nullptr, nullptr, nullptr, 0, 0, 0, nullptr);
}

// genotypeLR with MM / ML tags on the records: the same outputs plus the per-SV MethylInfo
int ref_genotype_lr_methyl(const char* contig_arena, const uint32_t* contig_off, const uint32_t* contig_len, int ncontig, const int32_t* rec10, int nrec,
                           const uint32_t* cigar, const char* reads, const int32_t* sv8, int nsv, const char* cons_arena, const uint32_t* cons_off,
                           const uint32_t* cons_len, int minMapQual, int minimumFlankSize, int minConsWindow, int maxGenoReadCount, float flankQuality, int genoCap,
                           uint8_t* qual_out, int qual_cap, uint32_t* ref_off, uint32_t* alt_off, int32_t* hp_out, int32_t* rc_out, const uint8_t* tagflags,
                           const char* mm_arena, const uint32_t* mm_off, const uint8_t* ml_arena, const uint32_t* ml_off, int methylWindow, int methylProb,
                           int minCpgDepth, int32_t* methyl_out) {
  return run_genotype_lr(contig_arena, contig_off, contig_len, ncontig, rec10, nrec, cigar, reads, sv8, nsv, cons_arena, cons_off, cons_len, minMapQual,
                         minimumFlankSize, minConsWindow, maxGenoReadCount, flankQuality, genoCap, qual_out, qual_cap, ref_off, alt_off, hp_out, rc_out, tagflags,
                         mm_arena, mm_off, ml_arena, ml_off, methylWindow, methylProb, minCpgDepth, methyl_out);
}

}  // extern "C"
