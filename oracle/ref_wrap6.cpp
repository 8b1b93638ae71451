// oracle/_ref wrapper, part 6 (TEST INFRASTRUCTURE ONLY): the reference's util.h compiled ITSELF (not masked), for the
// library-parameter estimation that precedes everything else in `delly sr` (getLibraryParams, src/util.h:771-891), run
// over in-memory alignments (htslib stand-ins as in ref_wrap3/4/5.cpp). Boost names util.h uses are served by oracle/shim6.
// Nothing from the reference is copied into this repository.
#define PANGENOME_H
#define ORACLE_REAL_UTIL_H
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
#include <boost/filesystem.hpp>
#include "shim/prelude.h"

namespace boost {
template <typename K, typename V> using unordered_map = std::unordered_map<K, V>;
namespace posix_time {
struct ptime {};
struct second_clock { static ptime local_time() { return ptime(); } };
inline std::string to_simple_string(ptime const&) { return "now"; }
}  // namespace posix_time
}  // namespace boost

#include <htslib/faidx.h>
#include <htslib/vcf.h>
#include <htslib/sam.h>
#include "util.h"

namespace {
struct RefConfig6 {
  std::vector<boost::filesystem::path> files;
  boost::filesystem::path genome;
  uint16_t madCutoff = 9, madNormalCutoff = 5;   // src/delly.h (defaults 9 / 5)
};
struct MemRecord6 { bam1_core_t core; std::vector<uint8_t> data; };
std::vector<MemRecord6> g_bam;
std::vector<uint32_t> g_tlen;
std::vector<std::string> g_names;
std::vector<char*> g_name_ptrs;
struct MemInterval { uint32_t lo, hi; uint32_t lower() const { return lo; } uint32_t upper() const { return hi; } };
}  // namespace

extern "C" {

htsFile* hts_open(const char*, const char*) { htsFile* f = (htsFile*) calloc(1, sizeof(htsFile)); f->is_bgzf = 1; return f; }
int hts_close(htsFile* f) { free(f); return 0; }
int hts_set_fai_filename(htsFile*, const char*) { return 0; }
hts_idx_t* sam_index_load(htsFile*, const char*) { return (hts_idx_t*) &g_bam; }
void hts_idx_destroy(hts_idx_t*) {}
sam_hdr_t* sam_hdr_read(samFile*) {
  sam_hdr_t* h = (sam_hdr_t*) calloc(1, sizeof(sam_hdr_t));
  h->n_targets = (int32_t) g_tlen.size(); h->target_len = g_tlen.data(); h->target_name = g_name_ptrs.data();
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
    MemRecord6& m = g_bam[it->i++];
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

// getLibraryParams (src/util.h:771-891) over in-memory alignments (rec: nrec x 12 as in the other wrappers).
// out: [rs, median, mad, minNormalISize, minISizeCutoff, maxNormalISize, maxISizeCutoff]
int ref_get_library_params(const uint32_t* contig_len, int ncontig, const int32_t* rec12, int nrec, const uint32_t* cigar, int madCutoff, int madNormalCutoff,
                           int32_t* out7) {
  RefConfig6 c; c.madCutoff = (uint16_t) madCutoff; c.madNormalCutoff = (uint16_t) madNormalCutoff;
  c.files.push_back(boost::filesystem::path("in-memory.bam")); c.genome = boost::filesystem::path("in-memory.fa");
  g_names.clear(); g_tlen.clear(); g_name_ptrs.clear();
  for (int k = 0; k < ncontig; ++k) { g_names.push_back("chr" + std::to_string(k)); g_tlen.push_back(contig_len[k]); }
  for (auto& nm : g_names) g_name_ptrs.push_back((char*) nm.c_str());
  g_bam.clear();
  for (int i = 0; i < nrec; ++i) {
    const int32_t* r = rec12 + 12 * i;
    MemRecord6 m; memset(&m.core, 0, sizeof(m.core));
    std::string qn = "q" + std::to_string(r[11]);
    m.core.tid = r[0]; m.core.pos = r[1]; m.core.flag = (uint16_t) r[2]; m.core.qual = (uint8_t) r[3]; m.core.l_qseq = r[4]; m.core.n_cigar = (uint32_t) r[6];
    m.core.mtid = r[8]; m.core.mpos = r[9]; m.core.isize = r[10];
    m.core.l_qname = (uint16_t) ((qn.size() + 1 + 3) & ~3u);
    m.data.assign(m.core.l_qname + 4 * (std::size_t) r[6], 0);
    memcpy(m.data.data(), qn.data(), qn.size());
    memcpy(m.data.data() + m.core.l_qname, cigar + r[5], 4 * (std::size_t) r[6]);
    g_bam.push_back(m);
  }
  std::vector<std::vector<MemInterval> > validRegions(ncontig);
  for (int t = 0; t < ncontig; ++t) validRegions[t].push_back(MemInterval{0u, 0x7fffffffu});
  std::vector<torali::LibraryInfo> sampleLib(1);
  std::streambuf* old = std::cerr.rdbuf(nullptr);
  torali::getLibraryParams(c, validRegions, sampleLib);
  std::cerr.rdbuf(old);
  torali::LibraryInfo const& l = sampleLib[0];
  out7[0] = l.rs; out7[1] = l.median; out7[2] = l.mad; out7[3] = l.minNormalISize; out7[4] = l.minISizeCutoff; out7[5] = l.maxNormalISize; out7[6] = l.maxISizeCutoff;
  return 0;
}

}  // extern "C"
