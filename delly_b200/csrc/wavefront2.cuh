// wavefront2.cuh — the longNeedle DP engine, second generation: TWO DP ROWS PER INSTRUCTION.
//
// Same decomposition as wavefront.cuh (thread t owns C consecutive DP columns, anti-diagonal wavefront across the threads, one shuffle
// hand-off per step), but a thread now advances a PAIR of rows (rA = 2k-1, rB = 2k) in one sweep over its columns, with the two cells of an
// iteration packed as int16x2 in one register: lane A (low half) computes cell (rA, c0+j), lane B (high half) cell (rB, c0+j-1) — the two
// cells of one anti-diagonal, which are independent. Every arithmetic instruction of the recurrence is a 16x2 SIMD-in-register DPX op
// (VIMNMX3.S16x2, VIMNMX.S16x2, VIADDMNMX.S16x2) or a plain 32-bit op that acts on both halves without cross-talk:
//
//     mm   = min.u16x2(colpair[j] & rowpair, 1)            match flags of both cells (one-hot base codes)
//     T    = diag + 2*mm                                    diagonal + (match ? 2 : 0)        [scores are row-shifted: U = H + r + 8]
//     X    = max3.s16x2(T, up, left)                        both cells
//     acc += (5X - up - 4left) << 4(j&3)                    both direction nibbles: (X-up) | (X-left) << 2, three IMADs on the FMA pipe
//     run  = max.s16x2(run, X)                              running row maxima (bestMat / bestRev of src/needle.h:88-103)
//     REV: every second iteration one PRMT packs two prefix maxima of a row for the mirrored int16 store
//     FWD: vmax = max(vmax, run + partner)                  one PRMT + one VIADDMNMX.S16x2: the fused join of src/needle.h:104-115
//     up'  = prmt(S[j+1], X) ; diag' = up ; left' = X - g   operands of the next iteration (lane B's `up` is lane A's result)
//
// i.e. ~10 instructions per TWO cells where the scalar engine spends ~12 per cell, and the per-thread state is C packed registers (previous
// row) + C+1 column-pair registers, so twice as many warps fit on an SM.
//
// Exactness notes:
//  * U = H + r + 8 >= 8 for every cell of a semiglobal matrix with a free first row (H >= -r), and <= 2m + 8: all packed values are positive
//    15-bit numbers; plain 32-bit add / subtract / multiply-add never carries between the halves, and the linear form 5X - up - 4left
//    yields both nibbles exactly (the differences X - up and X - left are in 0..3 by the DP's Lipschitz bounds: match +1, gaps -1).
//  * The direction nibble stores the DIFFERENCES, not a decoded direction: vertical iff (X - up) == 0, else horizontal iff (X - left) == 0,
//    else diagonal — longNeedle's priority (src/needle.h:159-171). Decoding happens in the traceback (O(m + n) cells).
//  * Lane B lags one column: its nibble for column c0+j-1 is produced in iteration j. Within a thread's C-nibble group the nibbles of an
//    EVEN row are therefore rotated by one (slot k holds column c0 + k - 1, slot 0 holds column c0 + C - 1); nib_index() maps (row, column).
//  * Lane B of iteration 0 and lane A of iteration C are not cells of this thread; their operands are arranged so that they reproduce the
//    neighbour's boundary value (B) / contribute a zero nibble (A) — see the operand set-up below.
//  * Bases are one-hot coded (A C G T N); a job with any other byte goes through the scalar engine (wavefront.cuh), which compares bytes.
#pragma once
#include "common.cuh"
#include "wavefront.cuh"

namespace wf2 {

using wf::Best;
using wf::best_before;

constexpr int PART_D = 3;        // partner-row ring depth in steps (FWD): rows are fetched PART_D - 1 steps ahead with cp.async
constexpr int BIAS = 8;          // U = H + row + BIAS
constexpr uint32_t NEG16x2 = 0x80008000u;

enum Mode { REV = 1, FWD = 2 };

__device__ __forceinline__ uint32_t hot(uint32_t ch) {
  return ch == 'A' ? 1u : ch == 'C' ? 2u : ch == 'G' ? 4u : ch == 'T' ? 8u : ch == 'N' ? 16u : 0u;
}
__host__ __device__ __forceinline__ bool hot_ok(uint8_t ch) { return ch == 'A' || ch == 'C' || ch == 'G' || ch == 'T' || ch == 'N'; }

// position of cell (r, c) inside the nibble row of r (cstart = first column of thread 0); see the header for the rotation of even rows
template <int C>
__device__ __forceinline__ int nib_index(int r, int c, int cstart) {
  const int k = c - cstart;
  const int t = k / C;
  int o = k - t * C;
  if (!(r & 1)) o = (o + 1 == C) ? 0 : o + 1;
  return t * C + o;
}
// decoded direction of a nibble: 1 vertical, 2 horizontal, 0 diagonal
__device__ __forceinline__ uint32_t nib_dir(uint32_t nib) { return ((nib & 3u) == 0u) ? 1u : (((nib >> 2) & 3u) == 0u ? 2u : 0u); }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const uint32_t s = (uint32_t) __cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// shared memory a pass needs for the partner ring (FWD), in uint4
template <int C> __host__ __device__ constexpr int part_slots(int T) { return PART_D * 2 * (C / 8) * T; }

// rowHot: one-hot codes of the row string (shared memory, m bytes); colChar(c-1) = column byte.
// Scoring is longNeedle's: match 1, mismatch -1, gap 1, first row free, last row horizontally free (src/needle.h:59-66).
// REV (cstart = 1): stores per row the running prefix maxima (int16, mirrored: element x at index P-1-x, value U-shifted) and the nibbles.
// FWD (cstart = idx0 - delta, see wavefront.cuh): nibbles + fused join against the stored maxima of REV; best = per-thread arg-max.
template <int C, bool MULTI, int MODE, typename TB>
__device__ __forceinline__ void pass(const uint8_t* rowHot, TB colChar, const int m, const int n, const int cstart, uint32_t* __restrict__ dirs,
                                     const uint32_t dstride, int16_t* __restrict__ brev, const uint32_t bstride, const int P, int* sm_x /* MULTI: WF_SMX ints */,
                                     uint4* sm_part /* FWD: part_slots<C>(T) uint4 */, Best& best, int& corner) {
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int T = blockDim.x;
  const int c0 = cstart + tid * C;
  constexpr int WPT = C / 8;
  const int delta = P - 1 - n;
  const bool first = (tid == 0);
  const bool owns = (c0 <= n) && (c0 + C - 1 >= 1);
  const int nvalid = min(C, n - c0 + 1);
  const int K = (m + 1) >> 1;                       // row pairs
  const int nact = (n - cstart + C) / C;
  const int nsteps = K + min(T, nact) - 1;
  constexpr uint32_t B2 = (uint32_t) BIAS | ((uint32_t) BIAS << 16);

  // column pairs: low half = one-hot of column c0+j (lane A), high half = one-hot of column c0+j-1 (lane B); 0 = matches nothing
  uint32_t colp[C + 1];
#pragma unroll
  for (int j = 0; j <= C; ++j) {
    const int ca = c0 + j, cb = c0 + j - 1;
    const uint32_t a = (j < C && ca >= 1 && ca <= n) ? hot((uint32_t) colChar(ca - 1)) : 0u;
    const uint32_t b = (j > 0 && cb >= 1 && cb <= n) ? hot((uint32_t) colChar(cb - 1)) : 0u;
    colp[j] = a | (b << 16);
  }
  uint32_t S[C];   // S[j]: high half = U(row above the current pair, column c0+j); low half = U(rA of the last pair, column c0+j+1)
#pragma unroll
  for (int j = 0; j < C; ++j) S[j] = B2;
  uint32_t X0 = B2;            // iteration-0 result of the last pair (lane A: column c0)
  uint32_t handH = B2, handX = B2, prevRecvHi = BIAS;
  best.val = wf::NEG; best.row = 0; best.col = 0; best.bm = 0;

  // ---- row 0 ------------------------------------------------------------------------------------------------------------
  if (MODE == REV) {
    if (owns) {
      uint4* o = (uint4*) (brev + (P - c0 - C));
      const uint32_t z = B2;   // prefix maxima of row 0: U = 0 + 0 + BIAS
#pragma unroll
      for (int w = 0; w < WPT; ++w) o[w] = make_uint4(z, z, z, z);
      if (nvalid < C)
        for (int j = nvalid; j < C; ++j) brev[P - 1 - (c0 + j)] = (int16_t) -32768;
    }
    const uint32_t padlen = bstride - (uint32_t) P;
    for (uint32_t i = tid; i < (uint32_t) (m + 1) * padlen; i += T) brev[(size_t) (i / padlen) * bstride + P + (i % padlen)] = (int16_t) -32768;
    if (first) brev[P - 1] = (int16_t) BIAS;
  }
  if (MODE == FWD) {
    // join candidates of row 0: bestMat[0][c] = 0, partner = element (m, n-c) at index c + delta of row m (stored with shift m + BIAS)
    const int16_t* brow = brev + (size_t) m * bstride;
    if (first && c0 == 1) { const int v = (int) __ldcg(brow + delta) - m - BIAS; if (v > best.val) { best.val = v; best.row = 0; best.col = 0; best.bm = 0; } }
    if (owns) {
#pragma unroll
      for (int j = 0; j < C; ++j) {
        const int c = c0 + j;
        if (c >= 0 && c <= n) {
          const int v = (int) __ldcg(brow + c + delta) - m - BIAS;
          if (v > best.val) { best.val = v; best.row = 0; best.col = c; best.bm = 0; }
        }
      }
    }
  }

  // ---- FWD: partner rows through a shared-memory ring, PART_D - 1 steps ahead ------------------------------------------
  auto part_at = [&](int slot, int row, int w) -> uint4* { return sm_part + ((size_t) ((slot * 2 + row) * WPT + w) * T + tid); };
  auto prefetch = [&](int sp) {   // partner rows of step sp (pair kp = sp - tid): rows m - rA, m - rB, this thread's columns
    if (MODE == FWD) {
      const int kp = sp - tid;
      if (owns && kp >= 1 && kp <= K) {
        const int slot = sp % PART_D;
        const int rA = 2 * kp - 1, rB = 2 * kp;
        const int16_t* srcA = brev + (size_t) (m - rA) * bstride + (c0 + delta);
#pragma unroll
        for (int w = 0; w < WPT; ++w) cp_async16(part_at(slot, 0, w), srcA + 8 * w);
        if (rB <= m) {
          const int16_t* srcB = brev + (size_t) (m - rB) * bstride + (c0 + delta);
#pragma unroll
          for (int w = 0; w < WPT; ++w) cp_async16(part_at(slot, 1, w), srcB + 8 * w);
        }
      }
      cp_async_commit();
    }
  };
  if (MODE == FWD) {
#pragma unroll
    for (int sp = 1; sp < PART_D; ++sp) prefetch(sp);
  }
#if WF_RING
  if (MULTI) {
    __syncthreads();
    if (tid < 2 * wf::WF_MAXW) sm_x[wf::WF_MAXW * wf::RING_D * 2 + tid] = 0;
    __syncthreads();
  }
#endif

  for (int s = 1; s <= nsteps; ++s) {
    // ---- hand-off from the left neighbour (what it produced in the previous step, for the same row pair) -------------------
    uint32_t recvH = __shfl_up_sync(0xffffffffu, handH, 1);
    uint32_t recvX = __shfl_up_sync(0xffffffffu, handX, 1);
    if (MULTI) {
      const int w = tid >> 5;
      volatile int* ring = sm_x;
      volatile int* prod = sm_x + wf::WF_MAXW * wf::RING_D * 2;
      volatile int* cons = prod + wf::WF_MAXW;
      const int nw = T >> 5;
      if (lane == 31 && w + 1 < nw) {
        while (cons[w + 1] < s - wf::RING_D) { }
        ring[(w * wf::RING_D + (s % wf::RING_D)) * 2] = (int) handH;
        ring[(w * wf::RING_D + (s % wf::RING_D)) * 2 + 1] = (int) handX;
        __threadfence_block();
        prod[w] = s;
      }
      if (lane == 0 && w > 0) {
        while (prod[w - 1] < s) { }
        recvH = (uint32_t) ring[((w - 1) * wf::RING_D + (s % wf::RING_D)) * 2];
        recvX = (uint32_t) ring[((w - 1) * wf::RING_D + (s % wf::RING_D)) * 2 + 1];
        __threadfence_block();
        cons[w] = s;
      }
      __syncwarp();
    }
    if (MODE == FWD) { prefetch(s + PART_D - 1); cp_async_wait<PART_D - 1>(); }
    const int k = s - tid;
    if (k >= 1 && k <= K && owns) {
      const int rA = 2 * k - 1, rB = 2 * k;
      const bool hasB = (rB <= m);
      const uint32_t gvec = ((rA == m) ? 0u : 1u) | (((rB == m) ? 0u : 1u) << 16);        // horizontal gap: free in the last row
      const uint32_t rowpair = (uint32_t) rowHot[rA - 1] | (hasB ? ((uint32_t) rowHot[rB - 1] << 16) : 0u);
      const uint32_t rH = first ? B2 : recvH;       // (U(rA, c0-1), U(rB, c0-1)); column 0 / the dummy columns of thread 0 are U = BIAS
      const uint32_t rX = first ? B2 : recvX;
      // iteration 0 operands. Lane B has no cell here: up.hi = U(rA,c0-1), left.hi = U(rB,c0-1) (no gap), T.hi = 0 < both, so X.hi = U(rB,c0-1),
      // which is exactly lane B's `left` source for iteration 1, and up.hi its diagonal.
      uint32_t diag = (first ? (uint32_t) BIAS : prevRecvHi);                                // lo = U(rA-1, c0-1); hi = 0
      uint32_t up = __byte_perm(S[0], rH, 0x5432);                                           // lo = S[0].hi = U(rA-1, c0); hi = rH.lo
      uint32_t left = rH - (gvec & 0xffffu);
      uint32_t run = rX;
      uint32_t acc[C / 4];
#pragma unroll
      for (int q = 0; q < C / 4; ++q) acc[q] = 0;
      acc[0] = 0u - (((rH >> 16) - (rH & 0xffffu)) << 16);   // cancels lane B's iteration-0 term (X - up).hi = U(rB,c0-1) - U(rA,c0-1)
      uint32_t runPrev = 0, keepA = 0, vmax = NEG16x2;
      uint32_t wA[4], wB[4];                         // REV: packed prefix maxima of the two rows, one 16-byte group at a time (mirrored word order)
      uint32_t pav[4], pbv[4], pbLast = NEG16x2;     // FWD: partner words of the current 8-column block
      const int slot = s % PART_D;
      if (MODE == FWD && first && c0 == 1) {
        // column 0 is not inside thread 0's block: candidates (r, 0) = H[r][0] + element (m-r, n)
        const int vA = (int) __ldcg(brev + (size_t) (m - rA) * bstride + delta) - m - BIAS;
        if (vA > best.val) { best.val = vA; best.row = rA; best.col = 0; best.bm = -rA; }
        if (hasB) {
          const int vB = (int) __ldcg(brev + (size_t) (m - rB) * bstride + delta) - m - BIAS;
          if (vB > best.val) { best.val = vB; best.row = rB; best.col = 0; best.bm = -rB; }
        }
      }
#pragma unroll
      for (int j = 0; j <= C; ++j) {
        const uint32_t mm = __vminu2(colp[j] & rowpair, 0x00010001u);
        const uint32_t Tv = mm * 2u + diag;
        const uint32_t X = __vimax3_s16x2(Tv, up, left);
        {
          constexpr int unused = 0; (void) unused;
          const int q = (j == C) ? 0 : (j >> 2);
          const int sh = (j == C) ? 0 : 4 * (j & 3);
          acc[q] += X * (5u << sh);
          acc[q] -= up * (1u << sh);
          acc[q] -= left * (4u << sh);
        }
        run = __vmaxs2(run, X);
        if (MODE == REV) {
          // a row's prefix maxima leave in 16-byte groups as soon as the four words of a group exist (word index descends with j)
          if (j < C && (j & 1)) {
            const int q = (C - 1 - j) >> 1;
            wA[q & 3] = __byte_perm(run, runPrev, 0x5410);           // columns (j, j-1) of row rA
            if ((q & 3) == 0) {
              if (nvalid < C) {
#pragma unroll
                for (int e = 0; e < 8; ++e) { const int slot8 = 2 * q + e; if (C - 1 - slot8 >= nvalid) wA[e >> 1] = (wA[e >> 1] & ~(0xffffu << ((e & 1) * 16))) | (0x8000u << ((e & 1) * 16)); }
              }
#ifndef LN2_NO_BREV_STORE
              ((uint4*) (brev + (size_t) rA * bstride + (P - c0 - C)))[q >> 2] = make_uint4(wA[0], wA[1], wA[2], wA[3]);
#else
              if (wA[0] == 0x12345678u) brev[0] = 1;
#endif
            }
          }
          if (j >= 2 && !(j & 1)) {
            const int q = (C - j) >> 1;
            wB[q & 3] = __byte_perm(run, runPrev, 0x7632);            // columns (j-1, j-2) of row rB
            if ((q & 3) == 0 && hasB) {
              if (nvalid < C) {
#pragma unroll
                for (int e = 0; e < 8; ++e) { const int slot8 = 2 * q + e; if (C - 1 - slot8 >= nvalid) wB[e >> 1] = (wB[e >> 1] & ~(0xffffu << ((e & 1) * 16))) | (0x8000u << ((e & 1) * 16)); }
              }
#ifndef LN2_NO_BREV_STORE
              ((uint4*) (brev + (size_t) rB * bstride + (P - c0 - C)))[q >> 2] = make_uint4(wB[0], wB[1], wB[2], wB[3]);
#else
              if (wB[0] == 0x12345678u) brev[0] = 1;
#endif
            }
          }
          runPrev = run;
        }
        if (MODE == FWD) {
          if (j < C && (j & 7) == 0) {
            pbLast = (j == 0) ? NEG16x2 : pbv[3];
            const uint4 a4 = *part_at(slot, 0, j >> 3);
            pav[0] = a4.x; pav[1] = a4.y; pav[2] = a4.z; pav[3] = a4.w;
            const uint4 b4 = *part_at(slot, 1, j >> 3);
            pbv[0] = b4.x; pbv[1] = b4.y; pbv[2] = b4.z; pbv[3] = b4.w;
          }
          uint32_t pv;
          if (j == C) pv = __byte_perm(NEG16x2, pbv[3], 0x7610);                                          // A: none; B: column C-1
          else if ((j & 1) == 0) pv = __byte_perm(pav[(j & 7) >> 1], ((j & 7) == 0) ? pbLast : pbv[((j & 7) >> 1) - 1], 0x7610);   // A: lo of word e; B: hi of the word before
          else pv = __byte_perm(pav[(j & 7) >> 1], pbv[(j & 7) >> 1], 0x5432);                          // A: hi of word e; B: lo of word e
          vmax = __viaddmax_s16x2(run, pv, vmax);
        }
        if (j == 0) X0 = X;
        if (j > 0) S[j - 1] = X;
        if (j == C - 1) keepA = X;
        if (j < C) {
          const uint32_t nl = X - gvec;
          diag = up;
          if (j + 1 == C) { diag = up & 0xffff0000u; up = __byte_perm(nl, X, 0x5410); }      // lane A idles: up.lo = left.lo, T.lo = 0 -> zero nibble
          else up = __byte_perm(S[j + 1], X, 0x5432);
          left = nl;
        } else {
          handH = __byte_perm(keepA, X, 0x7610);       // (U(rA, c0+C-1), U(rB, c0+C-1))
          handX = run;
        }
      }
      prevRecvHi = rH >> 16;
      // ---- nibble rows --------------------------------------------------------------------------------------------------
      {
#ifndef LN2_NO_NIB_STORE
        uint32_t* dA = dirs + (size_t) rA * dstride + (c0 - cstart) / 8;
#pragma unroll
        for (int w = 0; w < WPT; ++w) dA[w] = __byte_perm(acc[2 * w], acc[2 * w + 1], 0x5410);
        if (hasB) {
          uint32_t* dB = dirs + (size_t) rB * dstride + (c0 - cstart) / 8;
#pragma unroll
          for (int w = 0; w < WPT; ++w) dB[w] = __byte_perm(acc[2 * w], acc[2 * w + 1], 0x7632);
        }
#else
        uint32_t z = 0;
#pragma unroll
        for (int w = 0; w < C / 4; ++w) z ^= acc[w];
        if (z == 0x12345678u) dirs[0] = 1;
#endif
      }
      if (MODE == REV && first) {
        brev[(size_t) rA * bstride + (P - 1)] = (int16_t) BIAS;   // x = 0: H = -r
        if (hasB) brev[(size_t) rB * bstride + (P - 1)] = (int16_t) BIAS;
      }
      if (MODE == FWD) {
        // rare: a row of this pair improves the thread's best -> find the first column that attains the row's maximum (row rA before rB)
        const int off = m + 2 * BIAS;
        const int vA = (int) (int16_t) (vmax & 0xffffu) - off, vB = (int) (int16_t) (vmax >> 16) - off;
        if (vA > best.val) {
          int rn = (int) (rX & 0xffffu);
#pragma unroll
          for (int j = 0; j < C; ++j) {
            const int x = (int) (((j == 0) ? X0 : S[j - 1]) & 0xffffu);
            rn = max(rn, x);
            const int pvv = (int) ((const int16_t*) part_at(slot, 0, j >> 3))[j & 7];
            if (rn + pvv - off == vA && vA > best.val) { best.val = vA; best.row = rA; best.col = c0 + j; best.bm = rn - rA - BIAS; }
          }
        }
        if (hasB && vB > best.val) {
          int rn = (int) (rX >> 16);
#pragma unroll
          for (int j = 0; j < C; ++j) {
            const int x = (int) (S[j] >> 16);
            rn = max(rn, x);
            const int pvv = (int) ((const int16_t*) part_at(slot, 1, j >> 3))[j & 7];
            if (rn + pvv - off == vB && vB > best.val) { best.val = vB; best.row = rB; best.col = c0 + j; best.bm = rn - rB - BIAS; }
          }
        }
      }
    }
  }
  if (MODE == FWD) cp_async_wait<0>();
  // corner H[m][n]: lane A of the last pair if m is odd, lane B if even
  {
    const int owner = (n - cstart) / C;
    __shared__ int sm_corner2;
    if (tid == owner) {
      const int jn = (n - cstart) % C;
      uint32_t v = (m & 1) ? (X0 & 0xffffu) : (S[0] >> 16);
#pragma unroll
      for (int j = 1; j < C; ++j)
        if (j == jn) v = (m & 1) ? (S[j - 1] & 0xffffu) : (S[j] >> 16);
      sm_corner2 = (int) v - m - BIAS;
    }
    __syncthreads();
    corner = sm_corner2;
    __syncthreads();
  }
}

}  // namespace wf2
