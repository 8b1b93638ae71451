// long_needle.cu — batched consensus-vs-SV-reference split alignment, the device replacement for
// longNeedle() (src/needle.h:45-222) as instantiated by _consRefAlignment (src/split.h:540-558):
// AlignConfig<true,false>, DnaScore(1,-1,-1,-1) => linear gaps, row 0 free, last row moves right free.
//
// Reference semantics reproduced bit-for-bit:
//   * forward DP on (s1,s2) and a second DP on (revcomp s1, revcomp s2) (needle.h:59-81, revcomp per
//     src/util.h:549-563 incl. its handling of non-ACGTN bytes);
//   * false if mat[m][n] != rev[m][n] (:83-86);
//   * best join: first strict maximum in row-major order of prefmax(mat[row])[col] +
//     prefmax(rev[m-row])[n-col] (:88-115); refRight = LAST right in [0,n-refLeft] with
//     mat[consLeft][refLeft] + rev[consRight][right] == bestScore (:116-123); false if no gain (:152);
//   * tracebacks with priority vertical > horizontal > diagonal (:155-192) and the stitched
//     2 x alilen alignment (:196-219), bytes other than ACGTN- in the reverse part left as '\0'.
//
// Device design (see wavefront.cuh): ONE WARP (short-read shapes) or one CTA (long-read shapes) PER
// ALIGNMENT, anti-diagonal wavefront with C columns per thread held in registers. The reverse pass stores,
// per cell, the running prefix maximum of its row (int16, column-mirrored) plus a 4-bit nibble (traceback
// direction + "cell equals the row maximum so far"); the forward pass never materialises mat: it carries the
// running maximum of its own row across threads and evaluates the join of needle.h:104-115 on the fly
// against the stored reverse maxima, keeping a per-thread arg-max in row-major first-max order that is
// reduced once at the end. refRight (needle.h:116-123) is recovered from the stored maxima and the
// "equals maximum" bits. Workspace ~3 B per DP cell (0.8 MB for 300x1100; a batch in flight exceeds L2 and streams
// through HBM at about 1 TB/s, far below the roof).
// Job geometry (ln_class / ln_cols / ln_threads): up to 1273 reference columns one warp with C = 8..40 columns per lane;
// up to 2041 two warps x C=32; up to 4089 four warps x C=32; up to 16377 eight warps x C = 24..64 (C >= 48 keeps the
// previous-row state in shared memory), chosen so that at least 3/4 of the lanes own columns. Warps of one CTA hand the
// boundary column over through a shared-memory ring (wavefront.cuh), not a block barrier. The two tracebacks are
// warp-cooperative (a diagonal window of direction words per round of loads). Job classes of one call run concurrently on
// forked streams. Measured: 885 GCUPS on short-read shapes, 810 GCUPS on 3 kb x 10 kb (profiles/r1n_bench.json).
#include "common.cuh"
#include "wavefront.cuh"
#include "wavefront2.cuh"
#include <algorithm>
#include <string>
#include <vector>

namespace {

constexpr int LN_NCLS = 16;  // 0 trivial; 1..8 one warp, C = 8*cls columns per lane; 9..15 multi-warp CTA (see ln_class)
constexpr int LN_GEN = 16;   // class + LN_GEN: the same geometry for jobs with a byte outside ACGTN (scalar engine, byte compares)

struct LnArgs {
  const uint8_t* seqs;
  const uint32_t* c_off;
  const uint32_t* c_len;
  const uint32_t* r_off;
  const uint32_t* r_len;
  uint32_t n;
  uint8_t* aln;             // output arena
  const uint64_t* aln_off;  // per job: offset of row 0; row 1 at +(m+n)
  uint32_t* aln_len;
  uint8_t* ok;
  int32_t* info;            // optional per job [consLeft, refLeft, refRight, bestScore] (may be NULL)
  uint32_t* perm;
  uint32_t* counts;              // [0..31] class counts (16.. = generic), [32..63] starts, [64..95] cursors
  unsigned long long* maxcells;  // [0..31] per class max (m+1)*bstride ; [32..63] per class max m+n ; [64..95] per class max m
  uint8_t* jobcls;               // per job: its class (0..31), written by the count kernel
  uint8_t* work;                 // per-CTA workspace slabs
  size_t work_stride;
  size_t off_dirsR, off_dirsF, off_trace;  // offsets inside a slab (reverse prefix maxima at 0)
};

// columns per thread of a class
// classes 1..5: one warp, C = 8*cls; 6..8: two warps x C=32 (one warp with C = 48..64 spills: 772 vs 876 GCUPS measured);
// 9: four warps x C=32; 10..15: eight warps x C = 24..64
__host__ __device__ inline int ln_cols(int cls) { return cls <= 5 ? 8 * cls : (cls <= 9 ? 32 : 8 * (cls - 7)); }
__host__ __device__ inline int ln_threads(int cls) { return cls <= 5 ? 32 : (cls <= 8 ? 64 : (cls == 9 ? 128 : 256)); }
// row geometry for n reference columns with C columns per thread (see wavefront.cuh)
__host__ __device__ inline uint32_t ln_nact(uint32_t n, uint32_t C) { return (n + C - 1) / C; }
__host__ __device__ inline uint32_t ln_P(uint32_t n, uint32_t C) { return 1 + ln_nact(n, C) * C; }
__host__ __device__ inline uint32_t ln_bstride(uint32_t n, uint32_t C) { return ln_nact(n, C) * C + C + 16; }  // P-1 .. + pad for overhanging forward blocks
__host__ __device__ inline uint32_t ln_dstride(uint32_t n, uint32_t C) { return (ln_nact(n, C) + 1) * (C / 8); }

// class by number of reference columns n: -1 unsupported
__host__ __device__ inline int ln_class(uint32_t m, uint32_t n) {
  if (m == 0 || n == 0) return 0;
  if ((uint64_t) m + n > 32000u) return -1;  // int16 score storage
  if (n + 7 <= 2048) return (int) ((n + 7 + 255) / 256);  // 1..8 by 256 columns (7 columns of slack for the aligned forward blocks); geometry: ln_cols / ln_threads
  // multi-warp CTAs: the CTA is sized so that at least 3/4 of its lanes own columns (idle lanes are idle issue slots)
  if (n + 7 <= 4096) return 9;     // 4 warps x 32 lanes x C=32
  if (n + 7 <= 16384) return 10 + (int) ((n + 7 - 4097) / 2048);  // 8 warps x 32 lanes x C = 24, 32, 40, 48, 56, 64
  return -1;
}

// one WARP per job: class by shape, and whether every byte of both sequences is one of ACGTN (then the packed engine applies)
__global__ void ln_count_kernel(LnArgs a, int* unsupported) {
  const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= a.n) return;
  const uint32_t m = a.c_len[i], n = a.r_len[i];
  int c = ln_class(m, n);
  if (c < 0) {   // beyond the kernel's shapes: this ONE job is reported as not aligned (ok = 0, info[2] = -1) and counted; the batch goes on
    if (lane == 0) {
      atomicAdd(unsupported, 1);
      a.jobcls[i] = 0; atomicAdd(&a.counts[0], 1u);
      a.ok[i] = 0; a.aln_len[i] = 0;
      if (a.info) { a.info[4 * i] = 0; a.info[4 * i + 1] = 0; a.info[4 * i + 2] = -1; a.info[4 * i + 3] = 0; }
    }
    return;
  }
  if (c > 0) {
    bool bad = false;
    const uint8_t* s1 = a.seqs + a.c_off[i];
    const uint8_t* s2 = a.seqs + a.r_off[i];
    for (uint32_t k = lane; k < m; k += 32) bad |= !wf2::hot_ok(__ldg(s1 + k));
    for (uint32_t k = lane; k < n; k += 32) bad |= !wf2::hot_ok(__ldg(s2 + k));
    if (__any_sync(0xffffffffu, bad)) c += LN_GEN;
  }
  if (lane != 0) return;
  a.jobcls[i] = (uint8_t) c;
  atomicAdd(&a.counts[c], 1u);
  if (c == 0) { a.ok[i] = 0; a.aln_len[i] = 0; return; }
  const int g = c & (LN_GEN - 1);
  atomicMax(&a.maxcells[c], (unsigned long long) (m + 1) * ln_bstride(n, ln_cols(g)));
  atomicMax(&a.maxcells[32 + c], (unsigned long long) (m + n));
  atomicMax(&a.maxcells[64 + c], (unsigned long long) m);
}

__global__ void ln_offsets_kernel(uint32_t* counts) {
  uint32_t s = 0;
  for (int c = 0; c < 32; ++c) { counts[32 + c] = s; counts[64 + c] = s; s += counts[c]; }
}

__global__ void ln_scatter_kernel(LnArgs a) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const int c = a.jobcls[i];
  uint32_t p = atomicAdd(&a.counts[64 + c], 1u);
  a.perm[p] = i;
}

// src/util.h:549-563 applied to position i of the output
__device__ __forceinline__ uint8_t revcomp_at(const uint8_t* s, uint32_t L, uint32_t i) {
  uint8_t c = s[L - 1 - i];
  if (c >= 'a' && c <= 'z') c -= 32;  // boost::to_upper_copy, "C" locale
  switch (c) {
    case 'A': return 'T';
    case 'C': return 'G';
    case 'G': return 'C';
    case 'T': return 'A';
    case 'N': return 'N';
    default: return s[i];  // the reference leaves the original (un-reversed) byte in place
  }
}

// needle.h:207-216: complement of an alignment character of the reverse part; other bytes stay unwritten (0)
__device__ __forceinline__ uint8_t comp_aln(uint8_t c) {
  switch (c) {
    case 'A': return 'T';
    case 'C': return 'G';
    case 'G': return 'C';
    case 'T': return 'A';
    case 'N': return 'N';
    case '-': return '-';
    default: return 0;
  }
}

// Warp-cooperative traceback over the direction nibbles from (rr,cc) to (0,0); emits alignment characters in
// traceback order and returns the number of columns (same value in every lane). Executed by one full warp.
//   * the warp stages a window of direction words that follows the diagonal (lane i: row rr-i, the two words around
//     column cc-i) with ONE round of loads, so the walk pays one memory latency per <= 32 steps instead of one per step;
//   * lane 0 walks the window (shared memory) and records which steps consumed a row / a column as two bit masks;
//   * all lanes then expand the masks in parallel: lane t derives the coordinates of step t by popcount, fetches the two
//     sequence bytes and stores the alignment characters (coalesced), so no load sits on the walk's dependency chain;
//   * once row 0 or column 0 is reached the rest is a pure gap run, filled in by all lanes.
// REVSTR: sequences are the reverse complements (the reverse matrix of src/needle.h:74-82).
template <bool REVSTR>
__device__ __forceinline__ uint32_t ln_traceback_warp(const uint32_t* __restrict__ dirs, uint32_t dstride, int cstart, const uint8_t* s1, uint32_t m,
                                                      const uint8_t* s2, uint32_t n, uint32_t rr0, uint32_t cc0, uint8_t* tA, uint8_t* tB,
                                                      uint32_t* win /* 64 words of shared memory owned by this warp */, int lane) {
  int rr = (int) rr0, cc = (int) cc0;
  uint32_t k = 0;
  auto chA = [&](int r) -> uint8_t { uint8_t a = REVSTR ? revcomp_at(s1, m, (uint32_t) r) : s1[r]; return REVSTR ? comp_aln(a) : a; };
  auto chB = [&](int c) -> uint8_t { uint8_t b = REVSTR ? revcomp_at(s2, n, (uint32_t) c) : s2[c]; return REVSTR ? comp_aln(b) : b; };
  const uint8_t gap = '-';
  while (rr > 0 && cc > 0) {
    {
      const int ri = rr - lane;
      const int wi = (max(cc - lane, 1) - cstart) >> 3;
      uint32_t w1 = 0, w0 = 0;
      if (ri >= 1) {
        const uint32_t* drow = dirs + (size_t) ri * dstride;
        w1 = __ldcg(drow + wi);
        if (wi > 0) w0 = __ldcg(drow + wi - 1);
      }
      win[2 * lane] = w0; win[2 * lane + 1] = w1;
    }
    __syncwarp();
    uint32_t maskR = 0, maskC = 0;
    int steps = 0;
    const int row0 = rr, col0 = cc;
    if (lane == 0) {
      int r = rr, c = cc;
      while (r > 0 && c > 0 && steps < 32) {
        const int i = row0 - r;
        if (i >= 32) break;
        const int k2 = c - cstart;
        const int wsel = (k2 >> 3) - ((max(col0 - i, 1) - cstart) >> 3) + 1;
        if (wsel < 0 || wsel > 1) break;
        const uint32_t code = (win[2 * i + wsel] >> ((k2 & 7) * 4)) & 3u;
        const uint32_t dr = (code != 2u), dc = (code != 1u);
        maskR |= dr << steps; maskC |= dc << steps;
        r -= (int) dr; c -= (int) dc;
        ++steps;
      }
    }
    maskR = __shfl_sync(0xffffffffu, maskR, 0);
    maskC = __shfl_sync(0xffffffffu, maskC, 0);
    steps = __shfl_sync(0xffffffffu, steps, 0);
    if (lane < steps) {
      const uint32_t below = (1u << lane) - 1u;
      const int r = row0 - __popc(maskR & below), c = col0 - __popc(maskC & below);   // coordinates before step `lane`
      tA[k + lane] = ((maskR >> lane) & 1u) ? chA(r - 1) : gap;
      tB[k + lane] = ((maskC >> lane) & 1u) ? chB(c - 1) : gap;
    }
    rr = row0 - __popc(maskR); cc = col0 - __popc(maskC);
    k += (uint32_t) steps;
    __syncwarp();
  }
  // row 0: horizontal run; column 0: vertical run (needle.h:159-171 with one index exhausted)
  for (int t = lane; t < cc; t += 32) { tA[k + t] = gap; tB[k + t] = chB(cc - 1 - t); }
  for (int t = lane; t < rr; t += 32) { tA[k + t] = chA(rr - 1 - t); tB[k + t] = gap; }
  return k + (uint32_t) cc + (uint32_t) rr;
}

// multi-warp classes whose C columns of row state do not fit the register file keep it in shared memory
#ifndef LN_UPS_COLS
#define LN_UPS_COLS 48
#endif
constexpr int LN_UPS_FROM = 7 + LN_UPS_COLS / 8;  // first class with C >= LN_UPS_COLS
// (register caps that would allow 10-16 single-warp CTAs per SM were measured and lose 15-25 %: spills cost more than occupancy gains)
// MAXT = threads per CTA the register allocation is sized for: a multi-warp CTA of 8 warps may use 255 registers per
// thread (the DP state of C columns lives in registers), 16 warps would be capped at 128 and spill.
template <int C, bool MULTI, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) ln_kernel(LnArgs a, int cls) {
  extern __shared__ uint8_t sm_dyn[];
  constexpr int UPS = (MULTI && C >= LN_UPS_COLS) ? MAXT : 0;   // wide column blocks: previous-row state in shared memory (wavefront.cuh)
  int* sm_up = (int*) sm_dyn;                         // [C][UPS]
  uint8_t* sm_rows = sm_dyn + (size_t) C * UPS * sizeof(int);  // row string of the current pass (m bytes)
  __shared__ int sm_x[wf::WF_SMX];
  __shared__ int sm_pub[8];
  __shared__ uint32_t sm_win[128];  // traceback windows (one per tracing warp)
  __shared__ wf::Best sm_best[16];
  const int tid = threadIdx.x;
  const int T = blockDim.x;
  const int lane = tid & 31;
  const uint32_t cnt = a.counts[cls], start = a.counts[32 + cls];
  uint8_t* slab = a.work + (size_t) blockIdx.x * a.work_stride;
  int16_t* brev = (int16_t*) slab;
  uint32_t* dirsR = (uint32_t*) (slab + a.off_dirsR);
  uint32_t* dirsF = (uint32_t*) (slab + a.off_dirsF);
  uint8_t* trace = slab + a.off_trace;
  const wf::Scoring sc = {1, -1, true, true};  // DnaScore(1,-1,-1,-1), AlignConfig<true,false> (src/split.h:541-542)

  for (uint32_t idx = blockIdx.x; idx < cnt; idx += gridDim.x) {
    const uint32_t job = a.perm[start + idx];
    const uint32_t m = a.c_len[job], n = a.r_len[job];
    const uint8_t* s1 = a.seqs + a.c_off[job];
    const uint8_t* s2 = a.seqs + a.r_off[job];
    const uint32_t bstride = ln_bstride(n, C), dstride = ln_dstride(n, C);
    const int P = (int) ln_P(n, C);
    const int delta = P - 1 - (int) n;
    const int cstartF = ((delta + 1) & ~7) - delta;  // forward blocks are aligned to the mirrored storage
    __syncthreads();
    // ---- reverse pass on (revcomp s1, revcomp s2) ----
    for (uint32_t i = tid; i < m; i += T) sm_rows[i] = revcomp_at(s1, m, i);
    __syncthreads();
    wf::Best dummy;
    int revCorner, matCorner;
    wf::pass<C, MULTI, wf::REV, false, UPS>([&](int i) { return sm_rows[i]; }, [&](int i) { return revcomp_at(s2, n, (uint32_t) i); }, (int) m, (int) n, sc,
                                            1, dirsR, dstride, brev, bstride, P, sm_x, dummy, revCorner, nullptr, nullptr, sm_up);
    __syncthreads();
    // ---- forward pass with the fused join ----
    for (uint32_t i = tid; i < m; i += T) sm_rows[i] = s1[i];
    __syncthreads();
    wf::Best best;
    wf::pass<C, MULTI, wf::FWD, false, UPS>([&](int i) { return sm_rows[i]; }, [&](int i) { return s2[i]; }, (int) m, (int) n, sc, cstartF, dirsF, dstride,
                                            brev, bstride, P, sm_x, best, matCorner, nullptr, nullptr, sm_up);
    // reduce the per-thread arg-max (row-major first max)
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      wf::Best o;
      o.val = __shfl_xor_sync(0xffffffffu, best.val, d);
      o.row = __shfl_xor_sync(0xffffffffu, best.row, d);
      o.col = __shfl_xor_sync(0xffffffffu, best.col, d);
      o.bm = __shfl_xor_sync(0xffffffffu, best.bm, d);
      if (wf::best_before(o, best)) best = o;
    }
    if (MULTI) {
      if (lane == 0) sm_best[tid >> 5] = best;
      __syncthreads();
      best = sm_best[0];
      for (int w = 1; w < (T >> 5); ++w)
        if (wf::best_before(sm_best[w], best)) best = sm_best[w];
      __syncthreads();
    }
    const int gbest = best.val, consLeft = best.row, refLeft = best.col, matv = best.bm;
    const bool ok = (matCorner == revCorner) && (gbest != matCorner);  // needle.h:83-86, :152
    if (!ok) {
      if (tid == 0) {
        a.ok[job] = 0; a.aln_len[job] = 0;
        if (a.info) { a.info[4 * job] = consLeft; a.info[4 * job + 1] = refLeft; a.info[4 * job + 2] = 0; a.info[4 * job + 3] = gbest; }
      }
      continue;
    }
    // refRight: last right in [0, n-refLeft] with matv + rev[consRight][right] == gbest (needle.h:116-123);
    // rev[cR][x] == target  <=>  prefixmax[cR][x] == target and the cell carries the "equals maximum" bit
    const uint32_t consRight = m - (uint32_t) consLeft;
    {
      const int16_t* brow = brev + (size_t) consRight * bstride;
      const uint32_t* drow = dirsR + (size_t) consRight * dstride;
      const int target = gbest - matv + (int) consRight;   // stored prefix maxima carry their row index as a shift (wavefront.cuh)
      int bestRight = 0;
      for (uint32_t x = tid; x <= n - (uint32_t) refLeft; x += T) {
        if ((int) __ldcg(brow + (P - 1 - (int) x)) != target) continue;
        bool eq = true;
        if (x > 0) { uint32_t w = __ldcg(drow + ((x - 1) >> 3)); eq = ((w >> (((x - 1) & 7) * 4)) & 4u) != 0; }
        if (eq) bestRight = max(bestRight, (int) x);
      }
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) bestRight = max(bestRight, __shfl_xor_sync(0xffffffffu, bestRight, d));
      if (MULTI) {
        if (lane == 0) sm_x[tid >> 5] = bestRight;
        __syncthreads();
        for (int w = 0; w < (T >> 5); ++w) bestRight = max(bestRight, sm_x[w]);
        __syncthreads();
      }
      if (tid == 0) sm_pub[2] = bestRight;
      __syncthreads();
    }
    const uint32_t refRight = (uint32_t) sm_pub[2];
    uint8_t* tFA = trace;
    uint8_t* tFB = trace + (m + n);
    uint8_t* tRA = trace + 2 * (size_t) (m + n);
    uint8_t* tRB = trace + 3 * (size_t) (m + n);
    // two warp-cooperative tracebacks: warps 0 and 1 of a CTA side by side, or one after the other in the single warp
    if (!MULTI || tid < 32) {
      const uint32_t v = ln_traceback_warp<false>(dirsF, dstride, cstartF, s1, m, s2, n, (uint32_t) consLeft, (uint32_t) refLeft, tFA, tFB, sm_win, lane);
      if (lane == 0) sm_pub[3] = (int) v;
    }
    if (!MULTI || (tid >= 32 && tid < 64)) {
      const uint32_t v = ln_traceback_warp<true>(dirsR, dstride, 1, s1, m, s2, n, consRight, refRight, tRA, tRB, sm_win + (MULTI ? 64 : 0), lane);
      if (lane == 0) sm_pub[4] = (int) v;
    }
    __syncthreads();
    const uint32_t Lf = (uint32_t) sm_pub[3], Lr = (uint32_t) sm_pub[4];
    const uint32_t gapref = (n - refRight) - (uint32_t) refLeft;
    const uint32_t L = Lf + gapref + Lr;
    uint8_t* o0 = a.aln + a.aln_off[job];
    uint8_t* o1 = o0 + (m + n);
    for (uint32_t i = tid; i < L; i += T) {
      uint8_t x, y;
      if (i < Lf) { x = tFA[Lf - 1 - i]; y = tFB[Lf - 1 - i]; }
      else if (i < Lf + gapref) { x = '-'; y = s2[refLeft + (i - Lf)]; }
      else { x = tRA[i - Lf - gapref]; y = tRB[i - Lf - gapref]; }
      o0[i] = x; o1[i] = y;
    }
    if (tid == 0) {
      a.ok[job] = 1; a.aln_len[job] = L;
      if (a.info) { a.info[4 * job] = consLeft; a.info[4 * job + 1] = refLeft; a.info[4 * job + 2] = (int) refRight; a.info[4 * job + 3] = gbest; }
    }
  }
}


// ---- second-generation kernel on the packed two-row engine (wavefront2.cuh) ------------------------------------------------------
// Same job anatomy as ln_kernel (reverse pass storing mirrored prefix maxima + nibbles, forward pass with the fused join, refRight,
// two warp-cooperative tracebacks, stitched output); what differs: nibbles hold the differences (X - up, X - left) and even rows are
// rotated by one column inside a thread's group (wf2::nib_index), there is no "equals the row maximum" bit — refRight (needle.h:116-123)
// rebuilds the one row it needs from the horizontal differences with a block-wide scan — and the partner rows of the forward pass are staged
// through shared memory with cp.async.
template <int C, bool REVSTR>
__device__ __forceinline__ uint32_t ln_traceback_warp2(const uint32_t* __restrict__ dirs, int T, int cstart, const uint8_t* s1, uint32_t m,
                                                       const uint8_t* s2, uint32_t n, uint32_t rr0, uint32_t cc0, uint8_t* tA, uint8_t* tB,
                                                       uint32_t* win /* 64 words of shared memory owned by this warp */, int lane) {
  int rr = (int) rr0, cc = (int) cc0;
  uint32_t k = 0;
  auto chA = [&](int r) -> uint8_t { uint8_t a = REVSTR ? revcomp_at(s1, m, (uint32_t) r) : s1[r]; return REVSTR ? comp_aln(a) : a; };
  auto chB = [&](int c) -> uint8_t { uint8_t b = REVSTR ? revcomp_at(s2, n, (uint32_t) c) : s2[c]; return REVSTR ? comp_aln(b) : b; };
  const uint8_t gap = '-';
  while (rr > 0 && cc > 0) {
    {
      const int ri = rr - lane;
      uint32_t w1 = 0, w0 = 0;
      if (ri >= 1) {
        const int wi = wf2::nib_index<C>(ri, max(cc - lane, 1), cstart) >> 3;
        w1 = wf2::nib_word<C>(dirs, ri, wi, T);
        if (wi > 0) w0 = wf2::nib_word<C>(dirs, ri, wi - 1, T);
      }
      win[2 * lane] = w0; win[2 * lane + 1] = w1;
    }
    __syncwarp();
    uint32_t maskR = 0, maskC = 0;
    int steps = 0;
    const int row0 = rr, col0 = cc;
    if (lane == 0) {
      int r = rr, c = cc;
      while (r > 0 && c > 0 && steps < 32) {
        const int i = row0 - r;
        if (i >= 32) break;
        const int idx = wf2::nib_index<C>(r, c, cstart);
        const int wsel = (idx >> 3) - (wf2::nib_index<C>(r, max(col0 - i, 1), cstart) >> 3) + 1;
        if (wsel < 0 || wsel > 1) break;
        const uint32_t code = wf2::nib_dir((win[2 * i + wsel] >> ((idx & 7) * 4)) & 15u);
        const uint32_t dr = (code != 2u), dc = (code != 1u);
        maskR |= dr << steps; maskC |= dc << steps;
        r -= (int) dr; c -= (int) dc;
        ++steps;
      }
    }
    maskR = __shfl_sync(0xffffffffu, maskR, 0);
    maskC = __shfl_sync(0xffffffffu, maskC, 0);
    steps = __shfl_sync(0xffffffffu, steps, 0);
    if (steps == 0) {   // the staged window missed the very first cell (block wrap of an even row): take one step straight from memory
      uint32_t code = 0;
      if (lane == 0) {
        const int idx = wf2::nib_index<C>(rr, cc, cstart);
        code = wf2::nib_dir((wf2::nib_word<C>(dirs, rr, idx >> 3, T) >> ((idx & 7) * 4)) & 15u);
      }
      code = __shfl_sync(0xffffffffu, code, 0);
      maskR = (code != 2u); maskC = (code != 1u); steps = 1;
    }
    if (lane < steps) {
      const uint32_t below = (1u << lane) - 1u;
      const int r = row0 - __popc(maskR & below), c = col0 - __popc(maskC & below);
      tA[k + lane] = ((maskR >> lane) & 1u) ? chA(r - 1) : gap;
      tB[k + lane] = ((maskC >> lane) & 1u) ? chB(c - 1) : gap;
    }
    rr = row0 - __popc(maskR); cc = col0 - __popc(maskC);
    k += (uint32_t) steps;
    __syncwarp();
  }
  for (int t = lane; t < cc; t += 32) { tA[k + t] = gap; tB[k + t] = chB(cc - 1 - t); }
  for (int t = lane; t < rr; t += 32) { tA[k + t] = chA(rr - 1 - t); tB[k + t] = gap; }
  return k + (uint32_t) cc + (uint32_t) rr;
}

#ifndef LN2_OCC
#define LN2_OCC 12   // single-warp CTAs per SM the register allocation is sized for
#endif
template <int C, bool MULTI, int MAXT>
__global__ void __launch_bounds__(MAXT, (MAXT == 32 ? LN2_OCC : (MAXT == 64 ? LN2_OCC / 2 : (MAXT == 128 ? LN2_OCC / 4 : 1)))) ln_kernel2(LnArgs a, int cls) {
  extern __shared__ uint8_t sm_dyn[];
  const int tid = threadIdx.x;
  const int T = blockDim.x;
  const int lane = tid & 31;
  uint4* sm_part = (uint4*) sm_dyn;                                                    // partner ring of the forward pass
#ifndef LN2_PD_SMALL
#define LN2_PD_SMALL 2
#endif
#ifndef LN2_PD_C40
#define LN2_PD_C40 2
#endif
  constexpr int PD = (MAXT == 256) ? 2 : ((C >= 40) ? LN2_PD_C40 : LN2_PD_SMALL);         // partner ring depth (wavefront2.cuh)
  uint32_t* sm_scr = (uint32_t*) (sm_dyn + (size_t) wf2::part_slots<C, PD>(T) * sizeof(uint4));   // row scratch of the join's arg-max search
  uint8_t* sm_rows = (uint8_t*) (sm_scr + (size_t) (C + 1) * T);                         // one-hot codes of the row string of the current pass
  __shared__ int sm_x[wf::WF_SMX];
  __shared__ int sm_pub[8];
  __shared__ uint32_t sm_win[128];
  __shared__ wf::Best sm_best[16];
  __shared__ int sm_scan[16];
  const uint32_t cnt = a.counts[cls], start = a.counts[32 + cls];
  uint8_t* slab = a.work + (size_t) blockIdx.x * a.work_stride;
  uint4* brev4 = (uint4*) slab;                       // step-major workspace (wavefront2.cuh)
  uint32_t* dirsR = (uint32_t*) (slab + a.off_dirsR);
  uint32_t* dirsF = (uint32_t*) (slab + a.off_dirsF);
  uint8_t* trace = slab + a.off_trace;

  for (uint32_t idx = blockIdx.x; idx < cnt; idx += gridDim.x) {
    const uint32_t job = a.perm[start + idx];
    const uint32_t m = a.c_len[job], n = a.r_len[job];
    const uint8_t* s1 = a.seqs + a.c_off[job];
    const uint8_t* s2 = a.seqs + a.r_off[job];
    const int P = (int) ln_P(n, C);
    const int delta = P - 1 - (int) n;
    const int cstartF = ((delta + 1) & ~7) - delta;
    __syncthreads();
    // ---- reverse pass on (revcomp s1, revcomp s2) ----
    for (uint32_t i = tid; i < m; i += T) sm_rows[i] = (uint8_t) wf2::hot(revcomp_at(s1, m, i));
    __syncthreads();
    wf::Best dummy;
    int revCorner, matCorner;
#ifndef LN2_SKIP_REV   // (development switches LN2_SKIP_*: timing experiments only, results are wrong with them)
    wf2::pass<C, MULTI, wf2::REV, PD>(sm_rows, [&](int i) { return revcomp_at(s2, n, (uint32_t) i); }, (int) m, (int) n, 1, dirsR, brev4, P, sm_x, sm_part, sm_scr, dummy, revCorner);
#else
    revCorner = 0;
#endif
    __syncthreads();
#ifdef LN2_SKIP_FWD
    if (tid == 0) { a.ok[job] = 0; a.aln_len[job] = (uint32_t) revCorner; }
    continue;
#endif
    // ---- forward pass with the fused join ----
    for (uint32_t i = tid; i < m; i += T) sm_rows[i] = (uint8_t) wf2::hot(s1[i]);
    wf2::synth_groups<C>(brev4, (int) m);
    if (cstartF < 0) {
      // The forward blocks start at column cstartF <= 1 (aligned to the mirrored groups); its dummy columns c < 0 would pair with the elements
      // x = n - c > n of the reverse rows — prefix maxima of columns that do not exist. They are no join candidates: mark them once per job.
      const int nd = -cstartF, nactR = (P - 1) / C;
      for (int i = tid; i < (int) (m + 1) * nd; i += T) {
        const int r = i / nd, idxm = P - 1 - ((int) n + 1 + i % nd), g = idxm >> 3;
        ((int16_t*) (brev4 + wf2::grp_index<C>(r, nactR - 1 - g / (C / 8), g % (C / 8), T)))[idxm & 7] = (int16_t) -32768;
      }
    }
    __syncthreads();
    wf::Best best;
    wf2::pass<C, MULTI, wf2::FWD, PD>(sm_rows, [&](int i) { return s2[i]; }, (int) m, (int) n, cstartF, dirsF, brev4, P, sm_x, sm_part, sm_scr, best, matCorner);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      wf::Best o;
      o.val = __shfl_xor_sync(0xffffffffu, best.val, d);
      o.row = __shfl_xor_sync(0xffffffffu, best.row, d);
      o.col = __shfl_xor_sync(0xffffffffu, best.col, d);
      o.bm = __shfl_xor_sync(0xffffffffu, best.bm, d);
      if (wf::best_before(o, best)) best = o;
    }
    if (MULTI) {
      if (lane == 0) sm_best[tid >> 5] = best;
      __syncthreads();
      best = sm_best[0];
      for (int w = 1; w < (T >> 5); ++w)
        if (wf::best_before(sm_best[w], best)) best = sm_best[w];
      __syncthreads();
    }
    const int gbest = best.val, consLeft = best.row, refLeft = best.col, matv = best.bm;
    const bool ok = (matCorner == revCorner) && (gbest != matCorner);  // needle.h:83-86, :152
    if (!ok) {
      if (tid == 0) {
        a.ok[job] = 0; a.aln_len[job] = 0;
        if (a.info) { a.info[4 * job] = consLeft; a.info[4 * job + 1] = refLeft; a.info[4 * job + 2] = 0; a.info[4 * job + 3] = gbest; }
      }
      continue;
    }
    // refRight: last right in [0, n-refLeft] with matv + rev[consRight][right] == gbest (needle.h:116-123). Row consRight of the reverse matrix
    // is rebuilt from its horizontal differences: H(cR, x) = H(cR, x-1) + (nibble >> 2) - g, H(cR, 0) = -cR (block-wide prefix scan).
    const uint32_t consRight = m - (uint32_t) consLeft;
    {
      const int target = gbest - matv;
      const int xmax = (int) n - refLeft;
      int bestRight = 0;
      if (consRight == 0) { if (target == 0) bestRight = xmax; }
      else {
        const int g = (consRight == m) ? 0 : 1;
        const int per = (xmax + T) / T;                       // columns per thread (x = 1 .. xmax)
        const int xa = 1 + tid * per, xb = min(xmax, xa + per - 1);
        int sum = 0;
        for (int x = xa; x <= xb; ++x) {
          const int ni = wf2::nib_index<C>((int) consRight, x, 1);
          sum += (int) ((wf2::nib_word<C>(dirsR, (int) consRight, ni >> 3, T) >> ((ni & 7) * 4 + 2)) & 3u) - g;
        }
        int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
        int base = incl - sum;
        if (MULTI) {
          if (lane == 31) sm_scan[tid >> 5] = incl;
          __syncthreads();
          for (int w = 0; w < (tid >> 5); ++w) base += sm_scan[w];
          __syncthreads();
        }
        int h = -(int) consRight + base;                       // H(cR, xa - 1)
        if (tid == 0 && h == target) bestRight = 0;            // x = 0 (always within range)
        for (int x = xa; x <= xb; ++x) {
          const int ni = wf2::nib_index<C>((int) consRight, x, 1);
          h += (int) ((wf2::nib_word<C>(dirsR, (int) consRight, ni >> 3, T) >> ((ni & 7) * 4 + 2)) & 3u) - g;
          if (h == target) bestRight = x;
        }
      }
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) bestRight = max(bestRight, __shfl_xor_sync(0xffffffffu, bestRight, d));
      if (MULTI) {
        if (lane == 0) sm_x[tid >> 5] = bestRight;
        __syncthreads();
        for (int w = 0; w < (T >> 5); ++w) bestRight = max(bestRight, sm_x[w]);
        __syncthreads();
      }
      if (tid == 0) sm_pub[2] = bestRight;
      __syncthreads();
    }
    const uint32_t refRight = (uint32_t) sm_pub[2];
    uint8_t* tFA = trace;
    uint8_t* tFB = trace + (m + n);
    uint8_t* tRA = trace + 2 * (size_t) (m + n);
    uint8_t* tRB = trace + 3 * (size_t) (m + n);
    if (!MULTI || tid < 32) {
      const uint32_t v = ln_traceback_warp2<C, false>(dirsF, T, cstartF, s1, m, s2, n, (uint32_t) consLeft, (uint32_t) refLeft, tFA, tFB, sm_win, lane);
      if (lane == 0) sm_pub[3] = (int) v;
    }
    if (!MULTI || (tid >= 32 && tid < 64)) {
      const uint32_t v = ln_traceback_warp2<C, true>(dirsR, T, 1, s1, m, s2, n, consRight, refRight, tRA, tRB, sm_win + (MULTI ? 64 : 0), lane);
      if (lane == 0) sm_pub[4] = (int) v;
    }
    __syncthreads();
    const uint32_t Lf = (uint32_t) sm_pub[3], Lr = (uint32_t) sm_pub[4];
    const uint32_t gapref = (n - refRight) - (uint32_t) refLeft;
    const uint32_t L = Lf + gapref + Lr;
    uint8_t* o0 = a.aln + a.aln_off[job];
    uint8_t* o1 = o0 + (m + n);
    for (uint32_t i = tid; i < L; i += T) {
      uint8_t x, y;
      if (i < Lf) { x = tFA[Lf - 1 - i]; y = tFB[Lf - 1 - i]; }
      else if (i < Lf + gapref) { x = '-'; y = s2[refLeft + (i - Lf)]; }
      else { x = tRA[i - Lf - gapref]; y = tRB[i - Lf - gapref]; }
      o0[i] = x; o1[i] = y;
    }
    if (tid == 0) {
      a.ok[job] = 1; a.aln_len[job] = L;
      if (a.info) { a.info[4 * job] = consLeft; a.info[4 * job + 1] = refLeft; a.info[4 * job + 2] = (int) refRight; a.info[4 * job + 3] = gbest; }
    }
  }
}

template <int C, bool MULTI, int MAXT = 32>
int ln_launch2(dgpu_ctx* ctx, LnArgs& a, int cls, unsigned grid, unsigned threads, size_t smem, cudaStream_t st) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(ln_kernel2<C, MULTI, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return dgpu_set_cuda_error(ctx, e, "cudaFuncSetAttribute(ln_kernel2)");
  }
  ln_kernel2<C, MULTI, MAXT><<<grid, threads, smem, st>>>(a, cls);
  DGPU_LAUNCH_CHECK(ctx, "ln_kernel2");
  return DGPU_OK;
}

template <int C, bool MULTI, int MAXT = 32>
int ln_launch(dgpu_ctx* ctx, LnArgs& a, int cls, unsigned grid, unsigned threads, size_t smem, cudaStream_t st) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(ln_kernel<C, MULTI, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return dgpu_set_cuda_error(ctx, e, "cudaFuncSetAttribute(ln_kernel)");
  }
  ln_kernel<C, MULTI, MAXT><<<grid, threads, smem, st>>>(a, cls);
  DGPU_LAUNCH_CHECK(ctx, "ln_kernel");
  return DGPU_OK;
}

}  // namespace

extern "C" {

int dgpu_long_needle_dev(dgpu_ctx* ctx, const uint8_t* seqs, uint64_t seqs_bytes,
                         const uint32_t* c_off, const uint32_t* c_len, const uint32_t* r_off, const uint32_t* r_len,
                         uint64_t n, uint8_t* aln, const uint64_t* aln_off, uint32_t* aln_len, uint8_t* ok,
                         int32_t* info, void* stream) {
  if (!ctx) return DGPU_ERR_ARG;
  if (n == 0) return DGPU_OK;
  if (n >= (1ull << 31) || seqs_bytes >= (1ull << 32)) return DGPU_ERR_ARG;
  if (!seqs || !c_off || !c_len || !r_off || !r_len || !aln || !aln_off || !aln_len || !ok) return DGPU_ERR_ARG;
  DGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t) stream : ctx->stream;
  LnArgs a;
  a.seqs = seqs;
  a.c_off = c_off; a.c_len = c_len; a.r_off = r_off; a.r_len = r_len; a.n = (uint32_t) n;
  a.aln = aln; a.aln_off = aln_off; a.aln_len = aln_len; a.ok = ok; a.info = info;
  void* p;
  int rc;
  if ((rc = dgpu_reserve(ctx, SLOT_PERM, n * sizeof(uint32_t), &p))) return rc;
  a.perm = (uint32_t*) p;
  if ((rc = dgpu_reserve(ctx, SLOT_COUNTS, 2048, &p))) return rc;
  a.counts = (uint32_t*) p;                                   // 96 uint32 = 384 B
  a.maxcells = (unsigned long long*) ((uint8_t*) p + 512);     // 96 uint64 = 768 B
  int* d_unsupported = (int*) ((uint8_t*) p + 1280);
  DGPU_CUDA(ctx, cudaMemsetAsync(p, 0, 2048, st));
  void* pc;
  if ((rc = dgpu_reserve(ctx, SLOT_K, n + 64, &pc))) return rc;
  a.jobcls = (uint8_t*) pc;
  const uint32_t cb = (uint32_t) ((n + 255) / 256);
  ln_count_kernel<<<(uint32_t) ((n * 32 + 255) / 256), 256, 0, st>>>(a, d_unsupported);
  DGPU_LAUNCH_CHECK(ctx, "ln_count");
  ln_offsets_kernel<<<1, 1, 0, st>>>(a.counts);
  DGPU_LAUNCH_CHECK(ctx, "ln_offsets");
  ln_scatter_kernel<<<cb, 256, 0, st>>>(a);
  DGPU_LAUNCH_CHECK(ctx, "ln_scatter");
  struct { uint32_t counts[128]; unsigned long long maxc[96]; int unsupported; } h;
  DGPU_CUDA(ctx, cudaMemcpyAsync(&h, p, sizeof(h), cudaMemcpyDeviceToHost, st));
  DGPU_CUDA(ctx, cudaStreamSynchronize(st));
  if (h.unsupported) {   // per-item refusal: those jobs come back with ok = 0 (and info[2] = -1); never a silent approximation, never a failed batch
    ctx->unsupported += (uint64_t) h.unsupported;
    ctx->last_error = "dgpu_long_needle: " + std::to_string(h.unsupported) + " job(s) with |cons|+|ref| > 32000 or |ref| > 16377 were not aligned (ok = 0)";
  }
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  // Plan every non-empty class first: workspace geometry and grid. The classes then run CONCURRENTLY on a small pool of
  // streams forked from `st` (each class has its own region of the workspace): a batch of a few hundred SVs spread over
  // several classes would otherwise run as a sequence of under-filled launches.
  // Classes 1..15: jobs over ACGTN only -> packed two-row engine (ln_kernel2); 17..31: the same geometries for jobs with other bytes -> scalar engine.
  struct Plan { int c; LnArgs a; size_t grid, off; unsigned threads; size_t smem; };
  std::vector<Plan> plans;
  size_t total = 0;
  for (int c = 1; c < 2 * LN_NCLS; ++c) {
    if (c == LN_GEN || !h.counts[c]) continue;
    const int g = c & (LN_GEN - 1);
    const bool packed = c < LN_GEN;
    Plan pl;
    pl.c = c; pl.a = a;
    const size_t cells = (size_t) h.maxc[c];
    const size_t mn = (size_t) h.maxc[32 + c];
    const size_t mmax = (size_t) h.maxc[64 + c];
    auto al = [](size_t x) { return (x + 255) & ~(size_t) 255; };
    size_t b_rev = al(cells * 2 + 1024);
    size_t b_dirs = al(cells / 2 + 64 * mmax + 1024);  // (m+1) * dstride words, dstride <= bstride/8 + C/8
    if (packed) {   // step-major workspace of the packed engine (wavefront2.cuh): sized by the row pairs, not by the columns
      const size_t steps = (mmax + 1) / 2 + ln_threads(g) + 1;
      b_rev = al((steps * 2 * (ln_cols(g) / 8) + ln_cols(g) / 8 + 1) * ln_threads(g) * 16 + 1024);   // rows + the constant groups of synth_groups()
      b_dirs = al(steps * 2 * (ln_cols(g) / 8) * ln_threads(g) * 4 + 1024);
    }
    const size_t b_trace = al(4 * mn + 64);
    pl.a.off_dirsR = b_rev;
    pl.a.off_dirsF = b_rev + b_dirs;
    pl.a.off_trace = b_rev + 2 * b_dirs;
    pl.a.work_stride = pl.a.off_trace + b_trace;
    pl.threads = (unsigned) ln_threads(g);
    int per_sm;
    if (packed) {
      pl.smem = ((mmax + 15) & ~(size_t) 15) + (size_t) (ln_threads(g) == 256 ? 2 : (ln_cols(g) >= 40 ? LN2_PD_C40 : LN2_PD_SMALL)) * 2 * (ln_cols(g) / 8) * ln_threads(g) * sizeof(uint4) +
                (size_t) (ln_cols(g) + 1) * ln_threads(g) * sizeof(uint32_t);
      per_sm = ln_threads(g) == 32 ? LN2_OCC : (ln_threads(g) == 64 ? LN2_OCC / 2 : (ln_threads(g) == 128 ? LN2_OCC / 4 : 1));
      per_sm = std::max(1, std::min<int>(per_sm, (int) ((200 * 1024) / (pl.smem + 2048))));
    } else {
      pl.smem = ((mmax + 15) & ~(size_t) 15) + (g >= LN_UPS_FROM ? (size_t) ln_cols(g) * 256 * sizeof(int) : 0);
      per_sm = ln_threads(g) == 32 ? 16 : 256 / ln_threads(g);  // multi-warp CTAs: 8 warps per SM at up to 255 registers per thread
    }
    pl.grid = std::min<size_t>(h.counts[c], (size_t) ctx->num_sms * per_sm);
    pl.off = 0;
    total += pl.grid * pl.a.work_stride;
    plans.push_back(pl);
  }
  const size_t budget = (size_t) ((double) free_b * 0.6) + ctx->bufs[SLOT_WORK1].cap;
  if (total > budget) {  // shrink every class's grid by the same factor (at least one CTA each)
    const double f = (double) budget / (double) total;
    total = 0;
    for (auto& pl : plans) { pl.grid = std::max<size_t>(1, (size_t) ((double) pl.grid * f)); total += pl.grid * pl.a.work_stride; }
  }
  if ((rc = dgpu_reserve(ctx, SLOT_WORK1, total, &p))) return rc;
  { size_t off = 0; for (auto& pl : plans) { pl.a.work = (uint8_t*) p + off; off += pl.grid * pl.a.work_stride; } }
  const bool fork = plans.size() > 1;
  if (fork && (rc = dgpu_fork_init(ctx))) return rc;
  dgpu_prof_begin(ctx, st);
  if (fork) DGPU_CUDA(ctx, cudaEventRecord(ctx->fork_events[0], st));
  for (size_t k = 0; k < plans.size(); ++k) {
    Plan& pl = plans[k];
    const int c = pl.c;
    cudaStream_t cs = st;
    if (fork) {
      cs = ctx->fork_streams[k % DGPU_FORK_STREAMS];
      if (k < DGPU_FORK_STREAMS) DGPU_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->fork_events[0], 0));
    }
    const unsigned grid = (unsigned) pl.grid, threads = pl.threads;
    const size_t smem = pl.smem;
    switch (c) {
      case 1: rc = ln_launch2<8, false>(ctx, pl.a, c, grid, threads, smem, cs); break;
      case 2: rc = ln_launch2<16, false>(ctx, pl.a, c, grid, threads, smem, cs); break;
      case 3: rc = ln_launch2<24, false>(ctx, pl.a, c, grid, threads, smem, cs); break;
      case 4: rc = ln_launch2<32, false>(ctx, pl.a, c, grid, threads, smem, cs); break;
      case 5: rc = ln_launch2<40, false>(ctx, pl.a, c, grid, threads, smem, cs); break;
      case 6: case 7: case 8: rc = ln_launch2<32, true, 64>(ctx, pl.a, c, grid, threads, smem, cs); break;
      case 9: rc = ln_launch2<32, true, 128>(ctx, pl.a, c, grid, threads, smem, cs); break;
      case 10: rc = ln_launch2<24, true, 256>(ctx, pl.a, c, grid, threads, smem, cs); break;
      case 11: rc = ln_launch2<32, true, 256>(ctx, pl.a, c, grid, threads, smem, cs); break;
      case 12: rc = ln_launch2<40, true, 256>(ctx, pl.a, c, grid, threads, smem, cs); break;
      case 13: rc = ln_launch2<48, true, 256>(ctx, pl.a, c, grid, threads, smem, cs); break;
      case 14: rc = ln_launch2<56, true, 256>(ctx, pl.a, c, grid, threads, smem, cs); break;
      case 15: rc = ln_launch2<64, true, 256>(ctx, pl.a, c, grid, threads, smem, cs); break;
      case 17: rc = ln_launch<8, false>(ctx, pl.a, c, grid, threads, smem, cs); break;
      case 18: rc = ln_launch<16, false>(ctx, pl.a, c, grid, threads, smem, cs); break;
      case 19: rc = ln_launch<24, false>(ctx, pl.a, c, grid, threads, smem, cs); break;
      case 20: rc = ln_launch<32, false>(ctx, pl.a, c, grid, threads, smem, cs); break;
      case 21: rc = ln_launch<40, false>(ctx, pl.a, c, grid, threads, smem, cs); break;
      case 22: case 23: case 24: rc = ln_launch<32, true, 64>(ctx, pl.a, c, grid, threads, smem, cs); break;  // 2 warps x C=32: n <= 2041 without spills
      case 25: rc = ln_launch<32, true, 128>(ctx, pl.a, c, grid, threads, smem, cs); break;
      case 26: rc = ln_launch<24, true, 256>(ctx, pl.a, c, grid, threads, smem, cs); break;
      case 27: rc = ln_launch<32, true, 256>(ctx, pl.a, c, grid, threads, smem, cs); break;
      case 28: rc = ln_launch<40, true, 256>(ctx, pl.a, c, grid, threads, smem, cs); break;
      case 29: rc = ln_launch<48, true, 256>(ctx, pl.a, c, grid, threads, smem, cs); break;
      case 30: rc = ln_launch<56, true, 256>(ctx, pl.a, c, grid, threads, smem, cs); break;
      default: rc = ln_launch<64, true, 256>(ctx, pl.a, c, grid, threads, smem, cs); break;
    }
    if (rc) return rc;
  }
  if (fork) {  // join
    const size_t used = std::min<size_t>(plans.size(), DGPU_FORK_STREAMS);
    for (size_t k = 0; k < used; ++k) {
      DGPU_CUDA(ctx, cudaEventRecord(ctx->fork_events[1 + k], ctx->fork_streams[k]));
      DGPU_CUDA(ctx, cudaStreamWaitEvent(st, ctx->fork_events[1 + k], 0));
    }
  }
  dgpu_prof_end(ctx, st);
  return DGPU_OK;
}

int dgpu_long_needle(dgpu_ctx* ctx, const uint8_t* seqs, uint64_t seqs_bytes,
                     const uint32_t* c_off, const uint32_t* c_len, const uint32_t* r_off, const uint32_t* r_len,
                     uint64_t n, uint8_t* aln, const uint64_t* aln_off, uint64_t aln_bytes, uint32_t* aln_len, uint8_t* ok,
                     int32_t* info) {
  DgpuCallTrace trace_("dgpu_long_needle", n);
  if (!ctx) return DGPU_ERR_ARG;
  if (n == 0) return DGPU_OK;
  if (!seqs || !c_off || !c_len || !r_off || !r_len || !aln || !aln_off || !aln_len || !ok) return DGPU_ERR_ARG;
  for (uint64_t i = 0; i < n; ++i) {   // caller buffers: sequences inside the arena, alignment slots (2 rows of |cons| + |ref|) inside aln_bytes
    if ((uint64_t) c_off[i] + c_len[i] > seqs_bytes || (uint64_t) r_off[i] + r_len[i] > seqs_bytes) { ctx->last_error = "dgpu_long_needle: a sequence lies outside the arena"; return DGPU_ERR_ARG; }
    if (aln_off[i] + 2ull * ((uint64_t) c_len[i] + r_len[i]) > aln_bytes) { ctx->last_error = "dgpu_long_needle: alignment slot beyond aln_bytes"; return DGPU_ERR_CAPACITY; }
  }
  DGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  void *d_seqs, *d_coff, *d_clen, *d_roff, *d_rlen, *d_aln, *d_aoff, *d_alen, *d_ok, *d_info = nullptr;
  int rc;
  if ((rc = dgpu_reserve(ctx, SLOT_SEQS, seqs_bytes + 64, &d_seqs))) return rc;
  if ((rc = dgpu_reserve(ctx, SLOT_QOFF, n * 4, &d_coff))) return rc;
  if ((rc = dgpu_reserve(ctx, SLOT_QLEN, n * 4, &d_clen))) return rc;
  if ((rc = dgpu_reserve(ctx, SLOT_TOFF, n * 4, &d_roff))) return rc;
  if ((rc = dgpu_reserve(ctx, SLOT_TLEN, n * 4, &d_rlen))) return rc;
  if ((rc = dgpu_reserve(ctx, SLOT_A0, aln_bytes + 64, &d_aln))) return rc;
  if ((rc = dgpu_reserve(ctx, SLOT_A1, n * 8, &d_aoff))) return rc;
  if ((rc = dgpu_reserve(ctx, SLOT_A2, n * 4, &d_alen))) return rc;
  if ((rc = dgpu_reserve(ctx, SLOT_A3, n, &d_ok))) return rc;
  if (info && (rc = dgpu_reserve(ctx, SLOT_A4, n * 16, &d_info))) return rc;
  DGPU_CUDA(ctx, cudaMemcpyAsync(d_seqs, seqs, seqs_bytes, cudaMemcpyHostToDevice, st));
  DGPU_CUDA(ctx, cudaMemcpyAsync(d_coff, c_off, n * 4, cudaMemcpyHostToDevice, st));
  DGPU_CUDA(ctx, cudaMemcpyAsync(d_clen, c_len, n * 4, cudaMemcpyHostToDevice, st));
  DGPU_CUDA(ctx, cudaMemcpyAsync(d_roff, r_off, n * 4, cudaMemcpyHostToDevice, st));
  DGPU_CUDA(ctx, cudaMemcpyAsync(d_rlen, r_len, n * 4, cudaMemcpyHostToDevice, st));
  DGPU_CUDA(ctx, cudaMemcpyAsync(d_aoff, aln_off, n * 8, cudaMemcpyHostToDevice, st));
  rc = dgpu_long_needle_dev(ctx, (const uint8_t*) d_seqs, seqs_bytes, (const uint32_t*) d_coff, (const uint32_t*) d_clen,
                            (const uint32_t*) d_roff, (const uint32_t*) d_rlen, n, (uint8_t*) d_aln, (const uint64_t*) d_aoff,
                            (uint32_t*) d_alen, (uint8_t*) d_ok, (int32_t*) d_info, st);
  if (rc) return rc;
  DGPU_CUDA(ctx, cudaMemcpyAsync(aln, d_aln, aln_bytes, cudaMemcpyDeviceToHost, st));
  DGPU_CUDA(ctx, cudaMemcpyAsync(aln_len, d_alen, n * 4, cudaMemcpyDeviceToHost, st));
  DGPU_CUDA(ctx, cudaMemcpyAsync(ok, d_ok, n, cudaMemcpyDeviceToHost, st));
  if (info) DGPU_CUDA(ctx, cudaMemcpyAsync(info, d_info, n * 16, cudaMemcpyDeviceToHost, st));
  DGPU_CUDA(ctx, cudaStreamSynchronize(st));
  return DGPU_OK;
}

}  // extern "C"
