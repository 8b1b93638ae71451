// wavefront.cuh — linear-gap global/semiglobal DP with traceback directions, anti-diagonal wavefront.
//
// One alignment is spread over T = 32*G threads: thread t owns the C consecutive DP columns
// c = 1 + t*C .. (t+1)*C (column 0 is a closed-form boundary) and, at step s, computes row r = s - t for
// them, entirely in registers. The only value that crosses threads is the last column's H (and, for the
// fused join, a running row maximum): one warp shuffle per step (plus a shared-memory hand-off between
// warps when G > 1). There is no scan and no second pass over the row: per cell the work is
//     D = max(diag + sub, up - 1);  H = max(D, left - g);  dir = (H == up-1) ? V : (H == left-g) ? Hz : S
// i.e. two DPX viaddmax-style ops, the match select and the 2-bit direction.
//
// Scoring is (match, mismatch) with unit gap cost; `row0_free` makes row 0 all zeros (else -c), `last_free`
// makes horizontal moves in the last row free (longNeedle's AlignConfig<true,false>, src/needle.h:59-66);
// column 0 is always -r. Direction priority is vertical > horizontal > diagonal, which is both longNeedle's
// traceback rule (src/needle.h:159-171) and edlib's (up > left > diagonal, src/edlib.cpp:1021-1131).
#pragma once
#include "common.cuh"

namespace wf {

constexpr int NEG = -(1 << 28);

struct Scoring {
  int match, mismatch;
  bool row0_free, last_free;
};

// What a pass produces besides the 4-bit direction nibbles.
enum Mode {
  PLAIN = 0,  // directions only
  REV = 1,    // + prefix-max of every row (int16, stored column-mirrored) and an "is row maximum so far" bit
  FWD = 2     // + fused best-join search against the stored prefix maxima of the other matrix
};

struct Best {  // per-thread running arg-max of the join, row-major first-max order
  int val, row, col, bm;
};

__device__ __forceinline__ bool best_before(const Best& a, const Best& b) {  // a wins over b?
  if (a.val != b.val) return a.val > b.val;
  if (a.row != b.row) return a.row < b.row;
  return a.col < b.col;
}

// nibble layout: bits 0-1 direction (0 diagonal, 1 vertical, 2 horizontal), bit 2 = "H equals the running row maximum"
// dirs: row-major, (n+7)/8 words per row (dstride), cell (r, c>=1) at word (c-1)/8, nibble (c-1)%8.
// brev: row-major int16, bstride elements per row, element (r, x) stored at index n - x (mirror).
template <int C, bool MULTI, int MODE, typename TA, typename TB>
__device__ __forceinline__ void pass(TA rowChar /* r-1 -> char, shared memory */, TB colChar /* c-1 -> char */, const int m, const int n,
                                     const Scoring sc, uint32_t* __restrict__ dirs, const uint32_t dstride, int16_t* __restrict__ brev,
                                     const uint32_t bstride, int* sm_x /* MULTI: 4 * 32 ints */, Best& best, int& corner) {
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int T = blockDim.x;
  const int c0 = 1 + tid * C;
  constexpr int WPT = C / 8;

  uint32_t bw[C / 4];
#pragma unroll
  for (int j = 0; j < C / 4; ++j) bw[j] = 0;
#pragma unroll
  for (int j = 0; j < C; ++j) {
    const int c = c0 + j;
    const uint32_t ch = (c <= n) ? (uint32_t) colChar(c - 1) : 0u;
    bw[j >> 2] |= ch << ((j & 3) * 8);
  }
  int up[C];
#pragma unroll
  for (int j = 0; j < C; ++j) up[j] = sc.row0_free ? 0 : -(c0 + j);
  int lastH = up[C - 1];   // "row 0" value of this thread's last column
  int lastX = sc.row0_free ? 0 : 0;  // running row maximum handed to the right neighbour (row 0: max of zeros / of -c is H[0][0] = 0)
  int prevRecvH = sc.row0_free ? 0 : -(c0 - 1);  // H[0][c0-1]
  best.val = NEG; best.row = 0; best.col = 0; best.bm = 0;

  // ---- row 0 bookkeeping -----------------------------------------------------------------------------
  if (MODE == REV) {
    // brev row 0: prefix maxima of row 0 (all zeros when row0_free); "is maximum" bit set everywhere
#pragma unroll
    for (int j = 0; j < C; ++j) {
      const int c = c0 + j;
      if (c <= n) brev[(n - c)] = (int16_t) 0;
    }
    if (tid == 0) brev[n] = 0;
    if (c0 <= n) {
#pragma unroll
      for (int w = 0; w < WPT; ++w)
        if (c0 + 8 * w <= n) dirs[(c0 - 1) / 8 + w] = 0x44444444u;
    }
  }
  if (MODE == FWD) {
    // join candidates of row 0: bestMat[0][c] = 0, partner = brev[m][c]  (row m of the other matrix)
    const int16_t* brow = brev + (size_t) m * bstride;
    if (tid == 0) { int v = (int) __ldcg(brow + 0); if (v > best.val) { best.val = v; best.row = 0; best.col = 0; best.bm = 0; } }
#pragma unroll
    for (int j = 0; j < C; ++j) {
      const int c = c0 + j;
      if (c <= n) {
        int v = (int) __ldcg(brow + c);
        if (v > best.val) { best.val = v; best.row = 0; best.col = c; best.bm = 0; }
      }
    }
  }

  const int nact = (n + C - 1) / C;  // threads that own at least one column
  const int nsteps = m + min(T, nact) - 1;
  for (int s = 1; s <= nsteps; ++s) {
    // ---- hand-off from the left neighbour (value it produced in the previous step) -------------------
    int recvH = __shfl_up_sync(0xffffffffu, lastH, 1);
    int recvX = __shfl_up_sync(0xffffffffu, lastX, 1);
    if (MULTI) {
      const int w = tid >> 5;
      int* slot = sm_x + ((s & 1) ? 64 : 0);
      if (lane == 31) { slot[2 * w] = lastH; slot[2 * w + 1] = lastX; }
      __syncthreads();
      if (lane == 0 && w > 0) { recvH = slot[2 * (w - 1)]; recvX = slot[2 * (w - 1) + 1]; }
    }
    const int r = s - tid;
    if (r >= 1 && r <= m) {
      const int g = (sc.last_free && r == m) ? 0 : 1;
      const uint32_t a = (uint32_t) rowChar(r - 1);
      int left = (tid == 0) ? -r : recvH;         // H[r][c0-1]
      int diag = (tid == 0) ? -(r - 1) : prevRecvH;  // H[r-1][c0-1]
      int runX = (tid == 0) ? -r : recvX;         // running maximum of row r up to column c0-1
      const int16_t* brow = (MODE == FWD) ? brev + (size_t) (m - r) * bstride : nullptr;
      int16_t* bout = (MODE == REV) ? brev + (size_t) r * bstride : nullptr;
      if (MODE == FWD && tid == 0) {
        int v = -r + (int) __ldcg(brow + 0);      // candidate (r, 0): bestMat[r][0] = H[r][0]
        if (v > best.val) { best.val = v; best.row = r; best.col = 0; best.bm = -r; }
      }
      if (MODE == REV && tid == 0) bout[n] = (int16_t) (-r);  // x = 0
      uint32_t dw[WPT];
#pragma unroll
      for (int w = 0; w < WPT; ++w) dw[w] = 0;
#pragma unroll
      for (int j = 0; j < C; ++j) {
        const int c = c0 + j;
        if (c <= n) {
          const uint32_t b = (bw[j >> 2] >> ((j & 3) * 8)) & 0xffu;
          const int sub = (a == b) ? sc.match : sc.mismatch;
          const int u = up[j] - 1;
          const int l = left - g;
          const int D = max(diag + sub, u);
          const int H = max(D, l);
          uint32_t code = (H == u) ? 1u : ((H == l) ? 2u : 0u);
          diag = up[j];
          up[j] = H;
          left = H;
          if (MODE == REV) {
            runX = max(runX, H);
            if (H == runX) code |= 4u;
            bout[n - c] = (int16_t) runX;
          }
          if (MODE == FWD) {
            runX = max(runX, H);
            const int v = runX + (int) __ldcg(brow + c);
            if (v > best.val) { best.val = v; best.row = r; best.col = c; best.bm = runX; }
          }
          dw[j >> 3] |= code << ((j & 7) * 4);
        }
      }
      lastH = left;
      lastX = runX;
      if (c0 <= n) {
        uint32_t* drow = dirs + (size_t) r * dstride + (c0 - 1) / 8;
#pragma unroll
        for (int w = 0; w < WPT; ++w)
          if (c0 + 8 * w <= n) drow[w] = dw[w];
      }
    }
    prevRecvH = recvH;
  }
  // corner H[m][n]
  {
    const int owner = (n - 1) / C;
    __shared__ int sm_corner;
    if (tid == owner) {
      int v = up[0];
#pragma unroll
      for (int j = 1; j < C; ++j) if (j == (n - 1) % C) v = up[j];
      sm_corner = v;
    }
    __syncthreads();
    corner = sm_corner;
    __syncthreads();
  }
}

}  // namespace wf
