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
};
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

}  // extern "C"
