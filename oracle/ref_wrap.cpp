// oracle/_ref wrapper (TEST INFRASTRUCTURE ONLY).
// Compiles the reference's own hot-path headers VERBATIM from /root/reference/src
// (edlib.cpp, align.h, gotoh.h, needle.h, msa.h, tags.h, split.h) behind a flat C ABI so
// tests and bench.py's cpu_baseline / --impl reference leg can call the real reference
// functions through ctypes. Nothing from the reference is copied into this repository:
// this file only #includes the sources where they lie.
#include "shim/prelude.h"
#include "tags.h"
#include "align.h"
#include "gotoh.h"
#include "needle.h"
#include "msa.h"
#include "split.h"

#include <atomic>
#include <cstring>
#include <thread>

namespace {
struct RefConfig {  // the fields the templated reference code reads from TConfig
  torali::DnaScore<int> aliscore;
  uint16_t minCliqueSize;
  float flankQuality;
  int32_t minimumFlankSize;
  int32_t indelsize;
  int32_t minConsWindow;
};
typedef boost::multi_array<char, 2> TAlign;

inline void to_align(const char* rows, int r, int L, TAlign& a) {
  a.resize(boost::extents[r][L]);
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < L; ++j) a[i][j] = rows[(size_t) i * L + j];
}
inline int from_align(TAlign const& a, char* out, long cap) {
  long r = a.shape()[0], L = a.shape()[1];
  if (r * L > cap) return -1;
  for (long i = 0; i < r; ++i)
    for (long j = 0; j < L; ++j) out[i * L + j] = a[i][j];
  return 0;
}
}  // namespace

extern "C" {

// ---- edlib (src/edlib.cpp verbatim) ------------------------------------------------
// mode: 0 NW, 1 SHW, 2 HW; task: 0 DISTANCE, 1 LOC, 2 PATH (edlib.h enums).
// eq: additional equality pairs as 2*neq chars (may be NULL).
int ref_edlib(const char* q, int ql, const char* t, int tl, int k, int mode, int task,
              const char* eq, int neq,
              int* dist, int* end0, int* start0, int* numloc,
              unsigned char* aln, int aln_cap, int* aln_len) {
  std::vector<EdlibEqualityPair> pairs(neq);
  for (int i = 0; i < neq; ++i) { pairs[i].first = eq[2 * i]; pairs[i].second = eq[2 * i + 1]; }
  EdlibAlignResult r = edlibAlign(q, ql, t, tl,
      edlibNewAlignConfig(k, (EdlibAlignMode) mode, (EdlibAlignTask) task, neq ? pairs.data() : NULL, neq));
  *dist = r.editDistance;
  *numloc = r.numLocations;
  *end0 = (r.endLocations != NULL) ? r.endLocations[0] : -2;
  *start0 = (r.startLocations != NULL) ? r.startLocations[0] : -2;
  *aln_len = r.alignmentLength;
  int rc = r.status;
  if (r.alignment != NULL) {
    if (r.alignmentLength > aln_cap) rc = -1;
    else memcpy(aln, r.alignment, r.alignmentLength);
  }
  edlibFreeAlignResult(r);
  return rc;
}

// edlibAlignmentToCigar (src/edlib.cpp:296-345): returns the string length, -1 for NULL
int ref_edlib_cigar(const unsigned char* aln, int n, int format, char* out, int cap) {
  char* c = edlibAlignmentToCigar(aln, n, (EdlibCigarFormat) format);
  if (!c) return -1;
  const int L = (int) strlen(c);
  if (L + 1 <= cap) memcpy(out, c, (size_t) L + 1);
  free(c);
  return L;
}

// Batched distance over a packed arena with the reference's threading scheme
// (fixed workers pulling an atomic index: src/coverage.h:412-441).
void ref_edlib_distance_batch(const char* arena, const uint64_t* q_off, const uint32_t* q_len,
                              const uint64_t* t_off, const uint32_t* t_len, const int32_t* k,
                              int mode, uint64_t n, int32_t* dist, int threads) {
  std::atomic<uint64_t> next(0);
  auto work = [&]() {
    for (;;) {
      uint64_t i = next.fetch_add(1, std::memory_order_relaxed);
      if (i >= n) break;
      EdlibAlignResult r = edlibAlign(arena + q_off[i], q_len[i], arena + t_off[i], t_len[i],
          edlibNewAlignConfig(k[i], (EdlibAlignMode) mode, EDLIB_TASK_DISTANCE, NULL, 0));
      dist[i] = r.editDistance;
      edlibFreeAlignResult(r);
    }
  };
  if (threads <= 1) { work(); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < threads; ++t) th.emplace_back(work);
  for (auto& x : th) x.join();
}

// ---- longNeedle (src/needle.h:45-222) via the instantiation of src/split.h:540-558 ----
int ref_long_needle(const char* s1, int m, const char* s2, int n, char* rows, long cap, int* alilen) {
  torali::AlignConfig<true, false> semiglobal;
  torali::DnaScore<int> lnsc(1, -1, -1, -1);
  TAlign aln;
  bool ok = torali::longNeedle(std::string(s1, m), std::string(s2, n), aln, semiglobal, lnsc);
  *alilen = 0;
  if (!ok) return 0;
  *alilen = (int) aln.shape()[1];
  if (from_align(aln, rows, cap) != 0) return -1;
  return 1;
}

void ref_long_needle_batch(const char* arena, const uint64_t* c_off, const uint32_t* c_len,
                           const uint64_t* r_off, const uint32_t* r_len, uint64_t n,
                           uint8_t* ok, uint32_t* alilen, int threads) {
  std::atomic<uint64_t> next(0);
  auto work = [&]() {
    torali::AlignConfig<true, false> semiglobal;
    torali::DnaScore<int> lnsc(1, -1, -1, -1);
    for (;;) {
      uint64_t i = next.fetch_add(1, std::memory_order_relaxed);
      if (i >= n) break;
      TAlign aln;
      bool r = torali::longNeedle(std::string(arena + c_off[i], c_len[i]), std::string(arena + r_off[i], r_len[i]), aln, semiglobal, lnsc);
      ok[i] = r ? 1 : 0;
      alilen[i] = r ? (uint32_t) aln.shape()[1] : 0;
    }
  };
  if (threads <= 1) { work(); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < threads; ++t) th.emplace_back(work);
  for (auto& x : th) x.join();
}

int ref_longest_homology(const char* s1, int m, const char* s2, int n, int thr) {
  return torali::longestHomology(std::string(s1, m), std::string(s2, n), thr);
}

// ---- gotoh (src/gotoh.h:71-174), AlignConfig<true,true> as in src/msa.h:106-107 ----
int ref_gotoh(const char* rows1, int r1, int L1, const char* rows2, int r2, int L2,
              int match, int mismatch, int go, int ge, char* out, long cap, int* outL, int* score) {
  TAlign a1, a2, aln;
  to_align(rows1, r1, L1, a1);
  to_align(rows2, r2, L2, a2);
  torali::AlignConfig<true, true> endFree;
  torali::DnaScore<int> sc(match, mismatch, go, ge);
  *score = torali::gotoh(a1, a2, aln, endFree, sc);
  *outL = (int) aln.shape()[1];
  return from_align(aln, out, cap);
}

int ref_lcs(const char* s1, int m, const char* s2, int n) {
  return torali::lcs(std::string(s1, m), std::string(s2, n));
}

// ---- msa (src/msa.h:185-239); reads given in the caller's iteration order ----
// Returns number of alignment rows; writes ungapped consensus and (optionally) the full
// progressive alignment (rows x L) for white-box tests.
int ref_msa(const char* arena, const uint32_t* off, const uint32_t* len, int nreads, int minClique,
            int match, int mismatch, int go, int ge, char* cons, int cons_cap, int* cons_len,
            char* aln_out, long aln_cap, int* alnL) {
  RefConfig c;
  c.aliscore = torali::DnaScore<int>(match, mismatch, go, ge);
  c.minCliqueSize = (uint16_t) minClique;
  std::vector<std::string> sps;
  for (int i = 0; i < nreads; ++i) sps.push_back(std::string(arena + off[i], len[i]));
  if (aln_out != NULL) {
    // White-box: repeat msa()'s steps to expose the alignment (same calls as msa.h:185-239).
    typedef boost::multi_array<int, 2> TDistArray;
    typedef TDistArray::index TDIndex;
    TDIndex num = sps.size();
    TDistArray d(boost::extents[2 * num + 1][2 * num + 1]);
    for (TDIndex i = 0; i < (2 * num + 1); ++i)
      for (TDIndex j = i + 1; j < (2 * num + 1); ++j) d[i][j] = -1;
    torali::distanceMatrix(sps, d);
    typedef boost::multi_array<int, 2> TPhylogeny;
    TPhylogeny p(boost::extents[2 * num + 1][3]);
    for (TDIndex i = 0; i < (2 * num + 1); ++i)
      for (TDIndex j = 0; j < 3; ++j) p[i][j] = -1;
    TDIndex root = torali::upgma(d, p, num);
    TAlign align;
    torali::palign(c, sps, p, root, align);
    *alnL = (int) align.shape()[1];
    if (from_align(align, aln_out, aln_cap) != 0) return -1;
  }
  std::string cs;
  int rows = torali::msa(c, sps, cs);
  *cons_len = (int) cs.size();
  if ((int) cs.size() > cons_cap) return -1;
  memcpy(cons, cs.data(), cs.size());
  return rows;
}

void ref_msa_batch(const char* arena, const uint64_t* read_off, const uint32_t* read_len,
                   const uint32_t* cluster_off, uint32_t nclusters, int minClique,
                   int match, int mismatch, int go, int ge,
                   char* cons, const uint64_t* cons_off, uint32_t* cons_len, int threads) {
  std::atomic<uint32_t> next(0);
  auto work = [&]() {
    RefConfig c;
    c.aliscore = torali::DnaScore<int>(match, mismatch, go, ge);
    c.minCliqueSize = (uint16_t) minClique;
    for (;;) {
      uint32_t i = next.fetch_add(1, std::memory_order_relaxed);
      if (i >= nclusters) break;
      std::vector<std::string> sps;
      for (uint32_t r = cluster_off[i]; r < cluster_off[i + 1]; ++r) sps.push_back(std::string(arena + read_off[r], read_len[r]));
      std::string cs;
      torali::msa(c, sps, cs);
      cons_len[i] = (uint32_t) cs.size();
      memcpy(cons + cons_off[i], cs.data(), cs.size());
    }
  };
  if (threads <= 1) { work(); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < threads; ++t) th.emplace_back(work);
  for (auto& x : th) x.join();
}

// ---- consensus (src/msa.h:111-173) ----
int ref_consensus(const char* rows, int r, int L, int minClique, char* gapped, char* cs, int* cs_len) {
  RefConfig c; c.minCliqueSize = (uint16_t) minClique;
  TAlign a; to_align(rows, r, L, a);
  std::string g, s;
  torali::consensus(c, a, g, s);
  memcpy(gapped, g.data(), g.size());
  memcpy(cs, s.data(), s.size());
  *cs_len = (int) s.size();
  return 0;
}

// ---- _consRefAlignment (src/split.h:540-558): longNeedle for svt != 4, splitAlign for svt == 4 ----
int ref_cons_ref_alignment(const char* cons, int m, const char* ref, int n, int svt, char* rows, long cap, int* alilen) {
  TAlign aln;
  bool ok = torali::_consRefAlignment(std::string(cons, m), std::string(ref, n), aln, svt);
  *alilen = (int) aln.shape()[1];
  if (aln.shape()[0] == 2) { if (from_align(aln, rows, cap) != 0) return -1; }
  return ok ? 1 : 0;
}

// ---- _findSplit (src/split.h:319-375). ad = {cStart,cEnd,rStart,rEnd,homLeft,homRight}, percId separately ----
int ref_find_split(const char* cons, int m, const char* ref, int n, const char* rows, int L, int svt,
                   float flankQuality, int minimumFlankSize, int32_t* ad6, float* percId) {
  RefConfig c; c.flankQuality = flankQuality; c.minimumFlankSize = minimumFlankSize;
  TAlign a; to_align(rows, 2, L, a);
  torali::AlignDescriptor ad;
  bool ok = torali::_findSplit(c, std::string(cons, m), std::string(ref, n), a, ad, svt);
  ad6[0] = ad.cStart; ad6[1] = ad.cEnd; ad6[2] = ad.rStart; ad6[3] = ad.rEnd; ad6[4] = ad.homLeft; ad6[5] = ad.homRight;
  *percId = ad.percId;
  return ok ? 1 : 0;
}

// ---- alignConsensus (src/split.h:646-672) on a single-contig or two-contig toy header ----
// sv_io: [chr, svStart, chr2, svEnd, svt, insLen] in; out: [svStart, svEnd, insLen, consBp, homLen,
// ciposlow, ciposhigh, ciendlow, ciendhigh, precise]; alleles/consensus returned as strings.
int ref_align_consensus(const char* seq, int seqlen, const char* sndSeq, int sndlen,
                        const int32_t* sv_in, const char* consensus, int conslen, int realign,
                        float flankQuality, int minimumFlankSize, int indelsize, int minConsWindow,
                        int32_t* sv_out, float* srAlignQuality, char* alleles, int alleles_cap, int* alleles_len,
                        char* cons_out, int* cons_out_len) {
  RefConfig c; c.flankQuality = flankQuality; c.minimumFlankSize = minimumFlankSize; c.indelsize = indelsize; c.minConsWindow = minConsWindow;
  bam_hdr_t hdr; memset(&hdr, 0, sizeof(hdr));
  uint32_t tlen[2] = {(uint32_t) seqlen, (uint32_t) sndlen};
  char n0[] = "chrA"; char n1[] = "chrB"; char* names[2] = {n0, n1};
  hdr.n_targets = 2; hdr.target_len = tlen; hdr.target_name = names;
  torali::StructuralVariantRecord sv;
  sv.chr = sv_in[0]; sv.svStart = sv_in[1]; sv.chr2 = sv_in[2]; sv.svEnd = sv_in[3]; sv.svt = sv_in[4]; sv.insLen = sv_in[5];
  sv.consensus = std::string(consensus, conslen);
  // reference call site: seq = sequence of sv.chr, sndSeq = sequence of sv.chr2 (src/shortpe.h:186)
  bool ok = torali::alignConsensus(c, &hdr, seq, sndSeq, sv, realign != 0);
  sv_out[0] = sv.svStart; sv_out[1] = sv.svEnd; sv_out[2] = sv.insLen; sv_out[3] = sv.consBp; sv_out[4] = sv.homLen;
  sv_out[5] = sv.ciposlow; sv_out[6] = sv.ciposhigh; sv_out[7] = sv.ciendlow; sv_out[8] = sv.ciendhigh; sv_out[9] = sv.precise ? 1 : 0;
  *srAlignQuality = sv.srAlignQuality;
  *alleles_len = (int) sv.alleles.size();
  if ((int) sv.alleles.size() <= alleles_cap) memcpy(alleles, sv.alleles.data(), sv.alleles.size());
  *cons_out_len = (int) sv.consensus.size();
  memcpy(cons_out, sv.consensus.data(), sv.consensus.size());
  return ok ? 1 : 0;
}

// _getSVRef on the same toy header, so tests can build the exact window the reference builds.
int ref_get_sv_ref(const char* seq, int seqlen, const char* sndSeq, int sndlen, const int32_t* sv_in, int conslen,
                   int minimumFlankSize, int indelsize, int minConsWindow, char* out, int cap, int* outlen) {
  RefConfig c; c.minimumFlankSize = minimumFlankSize; c.indelsize = indelsize; c.minConsWindow = minConsWindow;
  bam_hdr_t hdr; memset(&hdr, 0, sizeof(hdr));
  uint32_t tlen[2] = {(uint32_t) seqlen, (uint32_t) sndlen};
  hdr.n_targets = 2; hdr.target_len = tlen;
  torali::StructuralVariantRecord sv;
  sv.chr = sv_in[0]; sv.svStart = sv_in[1]; sv.chr2 = sv_in[2]; sv.svEnd = sv_in[3]; sv.svt = sv_in[4]; sv.insLen = sv_in[5];
  torali::Breakpoint bp(sv);
  if (sv.svt == 4) {
    int32_t bufferSpace = std::max((int32_t) ((conslen - sv.insLen) / 3), c.minimumFlankSize);
    torali::_initBreakpoint(&hdr, bp, bufferSpace, sv.svt);
  } else torali::_initBreakpoint(&hdr, bp, conslen, sv.svt);
  if (bp.chr != bp.chr2) bp.part1 = torali::_getSVRef(c, sndSeq, bp, bp.chr2, sv.svt);
  std::string s = torali::_getSVRef(c, seq, bp, bp.chr, sv.svt);
  *outlen = (int) s.size();
  if ((int) s.size() > cap) return -1;
  memcpy(out, s.data(), s.size());
  return 0;
}

int ref_hardware_threads() { return (int) std::thread::hardware_concurrency(); }

}  // extern "C"
