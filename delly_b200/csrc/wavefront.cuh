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
#include <type_traits>

#ifndef WF_RING
#define WF_RING 1   // multi-warp passes: ring-buffer hand-off between neighbouring warps instead of a block barrier per step
#endif

namespace wf {

constexpr int NEG = -(1 << 28);
constexpr int WF_MAXW = 16;   // warps per multi-warp pass
constexpr int RING_D = 8;     // hand-off ring depth (steps a warp may run ahead of its right neighbour)
constexpr int WF_SMX = WF_MAXW * RING_D * 2 + 2 * WF_MAXW;  // ints of shared memory a pass needs (sm_x)

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

// Optional generalised equality (edlib's additionalEqualities, src/edlib.cpp:58-79): byte a equals byte b iff a == b or
// some pair p has {a,b} = {first_p, second_p}. f[x] / s[x] = bitmask of the pairs in which x is the first / second member.
struct EqTables {
  const uint32_t* f;
  const uint32_t* s;
  __device__ __forceinline__ bool equal(uint32_t a, uint32_t b) const { return a == b || (f[a] & s[b]) || (s[a] & f[b]); }
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
// dirs: row-major, dstride words per row, cell (r, c >= max(1,cstart)) at word (c-cstart)/8, nibble (c-cstart)%8.
// Thread t owns columns cstart + t*C .. +C-1 (columns < 1 or > n are masked), so its nibble words and its
// slice of the prefix-maximum row are always whole, 8-element aligned vectors:
//   REV  (cstart = 1): element (r, x) — the row's prefix maximum PLUS r, see the row shift in the cell loop — is stored MIRRORED at index P-1-x with P = 1 + nact*C, so a thread's
//        C values are one aligned, reversed run -> C/8 STG.128 per row chunk;
//   FWD  (cstart = idx0 - delta, delta = P-1-n, idx0 = (delta+1) & ~7): the partner of column c is element
//        (m-r, n-c) = index c + delta -> a thread's C partners are one aligned run -> C/8 LDG.128.
// EQ: equality through EqTables (column masks over the row symbol classes A,C,G,T,N + exact fallback for other row bytes).
// colout (PLAIN only, may be NULL): receives H[r][n] for r = 0..m (the last DP column) — Hirschberg's half columns.
// dirs may be NULL in PLAIN mode (score-only pass).
// UPS > 0: the previous-row values of a thread's C columns live in shared memory (element j of thread t at
// sm_up[j * UPS + t], conflict-free) instead of registers — for C = 64, where the register file cannot hold them.
template <int C, int UPS>
struct UpStore {
  int v[C];
  __device__ __forceinline__ explicit UpStore(int*) {}
  __device__ __forceinline__ int& operator[](int j) { return v[j]; }
};
template <int C>
struct UpStore<C, 0> {
  int v[C];
  __device__ __forceinline__ explicit UpStore(int*) {}
  __device__ __forceinline__ int& operator[](int j) { return v[j]; }
};
template <int C, int UPS>
struct UpShared {
  int* p;
  __device__ __forceinline__ explicit UpShared(int* base) : p(base + threadIdx.x) {}
  __device__ __forceinline__ int& operator[](int j) { return p[j * UPS]; }
};

template <int C, bool MULTI, int MODE, bool EQ = false, int UPS = 0, typename TA, typename TB>
__device__ __forceinline__ void pass(TA rowChar /* r-1 -> char, shared memory */, TB colChar /* c-1 -> char */, const int m, const int n,
                                     const Scoring sc, const int cstart, uint32_t* __restrict__ dirs, const uint32_t dstride,
                                     int16_t* __restrict__ brev, const uint32_t bstride, const int P, int* sm_x /* MULTI: WF_SMX ints */,
                                     Best& best, int& corner, const EqTables* eqt = nullptr, int* __restrict__ colout = nullptr,
                                     int* sm_up = nullptr /* UPS > 0: C * UPS ints of shared memory */) {
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int T = blockDim.x;
  const int c0 = cstart + tid * C;
  constexpr int WPT = C / 8;
  constexpr uint32_t NEG16x2 = 0x80008000u;  // two int16 minima: "no partner here"
  const int delta = P - 1 - n;

  // Column characters, 4 per register. Columns outside 1..n get a byte no row character can equal... they are
  // computed like real cells (no per-cell bounds checks); their results never reach a valid cell or the join.
  uint32_t bw[C / 4];
#pragma unroll
  for (int j = 0; j < C / 4; ++j) bw[j] = 0;
#pragma unroll
  for (int j = 0; j < C; ++j) {
    const int c = c0 + j;
    const uint32_t ch = (c >= 1 && c <= n) ? (uint32_t) colChar(c - 1) : 0u;
    bw[j >> 2] |= ch << ((j & 3) * 8);
  }
  uint32_t cm[EQ ? C / 4 : 1];  // EQ: per column, bit k = "row symbol class k (A,C,G,T,N) equals this column's byte"
  if (EQ) {
#pragma unroll
    for (int j = 0; j < C / 4; ++j) cm[j] = 0;
#pragma unroll
    for (int j = 0; j < C; ++j) {
      const int c = c0 + j;
      uint32_t mk = 0;
      if (c >= 1 && c <= n) {
        const uint32_t ch = (uint32_t) colChar(c - 1);
        const uint8_t sym[5] = {'A', 'C', 'G', 'T', 'N'};
#pragma unroll
        for (int k = 0; k < 5; ++k) mk |= eqt->equal(sym[k], ch) ? (1u << k) : 0u;
      }
      cm[j >> 2] |= mk << ((j & 3) * 8);
    }
  }
  const int ownerN = (n - cstart) / C;  // thread that owns column n
  if (MODE == PLAIN && colout != nullptr && tid == ownerN) colout[0] = sc.row0_free ? 0 : -n;
  typename std::conditional<(UPS > 0), UpShared<C, UPS>, UpStore<C, 0> >::type up(sm_up);
#pragma unroll
  for (int j = 0; j < C; ++j) up[j] = sc.row0_free ? 0 : -max(c0 + j, 0);
  int lastH = sc.row0_free ? 0 : -min(c0 + C - 1, n);
  int lastX = 0;
  int prevRecvH = sc.row0_free ? 0 : -(c0 - 1);  // H[0][c0-1]
  best.val = NEG; best.row = 0; best.col = 0; best.bm = 0;
  const bool owns = (c0 <= n) && (c0 + C - 1 >= 1);
  const bool first = (tid == 0);
  const int jforce = first ? max(0, 1 - c0) : 0;   // thread 0: columns <= 0 are pinned to the column-0 boundary value -r
  const int nvalid = min(C, n - c0 + 1);            // columns of this thread that are <= n

  // ---- row 0 bookkeeping -----------------------------------------------------------------------------
  if (MODE == REV) {
    if (owns) {
      uint4* o = (uint4*) (brev + (P - c0 - C));
#pragma unroll
      for (int w = 0; w < WPT; ++w) o[w] = make_uint4(0, 0, 0, 0);   // row 0 prefix maxima are 0
      if (nvalid < C) {  // columns beyond n: no partner
        for (int j = nvalid; j < C; ++j) brev[P - 1 - (c0 + j)] = (int16_t) -32768;
      }
#pragma unroll
      for (int w = 0; w < WPT; ++w) dirs[(c0 - cstart) / 8 + w] = 0x44444444u;
    }
    {
      // the pad between the end of a row (index P-1 = column x=0) and the row stride carries "no partner" for the
      // forward blocks that overhang the row; written cooperatively once per pass
      const uint32_t padlen = bstride - (uint32_t) P;
      for (uint32_t i = tid; i < (uint32_t) (m + 1) * padlen; i += T) brev[(size_t) (i / padlen) * bstride + P + (i % padlen)] = (int16_t) -32768;
    }
    if (first) brev[P - 1] = 0;
  }
  if (MODE == FWD) {
    // join candidates of row 0: bestMat[0][c] = 0, partner = element (m, n-c) at index c + delta
    const int16_t* brow = brev + (size_t) m * bstride;
    // (stored prefix maxima carry their row index as a shift: row m here)
    if (first && c0 == 1) { int v = (int) __ldcg(brow + delta) - m; if (v > best.val) { best.val = v; best.row = 0; best.col = 0; best.bm = 0; } }
    if (owns) {
#pragma unroll
      for (int j = 0; j < C; ++j) {
        const int c = c0 + j;
        if (c >= 0 && c <= n) {
          int v = (int) __ldcg(brow + c + delta) - m;
          if (v > best.val) { best.val = v; best.row = 0; best.col = c; best.bm = 0; }
        }
      }
    }
  }

  const int nact = (n - cstart + C) / C;  // threads that own at least one column
  const int nsteps = m + min(T, nact) - 1;
  uint32_t nextPart[C / 2];               // FWD: partner row prefetched one step ahead
  if (MODE == FWD) {
#pragma unroll
    for (int q = 0; q < C / 2; ++q) nextPart[q] = NEG16x2;
    const int r1 = 1 - tid;               // row of step 1
    if (owns && r1 >= 1 && r1 <= m) {
      const uint4* src = (const uint4*) (brev + (size_t) (m - r1) * bstride + (c0 + delta));
#pragma unroll
      for (int w = 0; w < WPT; ++w) { uint4 v = __ldcg(src + w); nextPart[4 * w] = v.x; nextPart[4 * w + 1] = v.y; nextPart[4 * w + 2] = v.z; nextPart[4 * w + 3] = v.w; }
    }
  }
#if WF_RING
  if (MULTI) {
    __syncthreads();                       // nobody is still spinning on the previous pass's counters
    if (tid < 2 * WF_MAXW) sm_x[WF_MAXW * RING_D * 2 + tid] = 0;
    __syncthreads();
  }
#endif
  for (int s = 1; s <= nsteps; ++s) {
    // ---- hand-off from the left neighbour (value it produced in the previous step) -------------------
    int recvH = __shfl_up_sync(0xffffffffu, lastH, 1);
    int recvX = __shfl_up_sync(0xffffffffu, lastX, 1);
    if (MULTI) {
      const int w = tid >> 5;
#if WF_RING
      // Warp-to-warp hand-off without a block barrier: warp w publishes the boundary values of step s-1 into a ring of
      // RING_D entries and bumps its step counter; warp w+1 spins on that counter, copies the entry and bumps its own
      // "consumed" counter, which is what lets warp w reuse the slot. Warps drift up to RING_D steps apart, so a stall
      // in one warp (a late partner row, a store burst) no longer stops the other 15.
      volatile int* ring = sm_x;                                   // [warp][RING_D][2]
      volatile int* prod = sm_x + WF_MAXW * RING_D * 2;            // [warp] steps published
      volatile int* cons = prod + WF_MAXW;                         // [warp] steps consumed from the left neighbour
      const int nw = T >> 5;
      if (lane == 31 && w + 1 < nw) {
        while (cons[w + 1] < s - RING_D) { }
        ring[(w * RING_D + (s % RING_D)) * 2] = lastH;
        ring[(w * RING_D + (s % RING_D)) * 2 + 1] = lastX;
        __threadfence_block();
        prod[w] = s;
      }
      if (lane == 0 && w > 0) {
        while (prod[w - 1] < s) { }
        recvH = ring[((w - 1) * RING_D + (s % RING_D)) * 2];
        recvX = ring[((w - 1) * RING_D + (s % RING_D)) * 2 + 1];
        __threadfence_block();
        cons[w] = s;
      }
      __syncwarp();
#else
      int* slot = sm_x + ((s & 1) ? 64 : 0);
      if (lane == 31) { slot[2 * w] = lastH; slot[2 * w + 1] = lastX; }
      __syncthreads();
      if (lane == 0 && w > 0) { recvH = slot[2 * (w - 1)]; recvX = slot[2 * (w - 1) + 1]; }
#endif
    }
    const int r = s - tid;
    uint32_t part[C / 2];
    if (MODE == FWD) {
#pragma unroll
      for (int q = 0; q < C / 2; ++q) part[q] = nextPart[q];
      if (owns && r + 1 >= 1 && r + 1 <= m) {  // prefetch the partner row of the next step (row r+1 pairs with m-r-1)
        const uint4* src = (const uint4*) (brev + (size_t) (m - r - 1) * bstride + (c0 + delta));
#pragma unroll
        for (int w = 0; w < WPT; ++w) { uint4 v = __ldcg(src + w); nextPart[4 * w] = v.x; nextPart[4 * w + 1] = v.y; nextPart[4 * w + 2] = v.z; nextPart[4 * w + 3] = v.w; }
      }
    }
    if (r >= 1 && r <= m && owns) {
      // Row-shifted scores: the registers hold U[r][c] = H[r][c] + r. With a unit vertical gap the recurrence
      //   H = max(diag + sub, up - 1, left - g)   becomes   U = max3(Ud + (sub + 1), Uu, Ul - g)
      // (Ud, Uu carry the shift r-1, Ul the shift r): no per-cell decrement of the upper neighbour, one 3-input
      // maximum, and the direction tests compare against the operands themselves. Column 0 (H = -r) is U = 0.
      // Everything that leaves the registers is converted back: H = U - r.
      const int g = (sc.last_free && r == m) ? 0 : 1;
      const int M1 = sc.match + 1, M0 = sc.mismatch + 1;
      const uint32_t a = (uint32_t) rowChar(r - 1);
      const uint32_t acode = EQ ? dna_code(a) : 0u;
      int lg = (first ? 0 : recvH) - g;         // U[r][c0-1] - g
      int diag = first ? 0 : prevRecvH;          // U[r-1][c0-1]
      const int runIn = first ? 0 : recvX;       // running maximum of row r up to column c0-1 (shifted)
      int runX = runIn;
      if (MODE == FWD && first && c0 == 1) {
        // column 0 is not inside thread 0's block: candidate (r, 0) = H[r][0] + element (m-r, n); stored rows are shifted too
        const int v = (int) __ldcg(brev + (size_t) (m - r) * bstride + delta) - m;
        if (v > best.val) { best.val = v; best.row = r; best.col = 0; best.bm = -r; }
      }
      uint32_t dw[WPT];
#pragma unroll
      for (int w = 0; w < WPT; ++w) dw[w] = 0;
      int vmax = NEG;
      int runPrev = 0;  // REV: shifted running maximum of the previous (even) column, packed together with the odd one
#pragma unroll
      for (int j = 0; j < C; ++j) {
        const uint32_t b = (bw[j >> 2] >> ((j & 3) * 8)) & 0xffu;
        bool same;
        if (EQ) {
          if (acode < 5) same = ((cm[j >> 2] >> ((j & 3) * 8 + acode)) & 1u) != 0;
          else same = eqt->equal(a, b);
        } else same = (a == b);
        const int uu = up[j];
        int U = __vimax3_s32(diag + (same ? M1 : M0), uu, lg);   // DPX three-input maximum
        uint32_t code = (U == uu) ? 1u : ((U == lg) ? 2u : 0u);
        if (j < 8) { if (j < jforce) U = 0; }  // only thread 0 of an aligned forward pass has such columns (at most 7)
        diag = uu;
        up[j] = U;
        lg = U - g;
        runX = max(runX, U);
        if (MODE == REV) {
          if (U == runX) code |= 4u;
          // mirrored slot q = C-1-j inside this thread's run: even j -> high half of word q>>1, odd j -> low half
          if (j & 1) part[(C - 1 - j) >> 1] = __byte_perm((uint32_t) runX, (uint32_t) runPrev, 0x5410);
          else runPrev = runX;
        }
        if (MODE == FWD) {
          const int pv = (int) (int16_t) ((part[j >> 1] >> ((j & 1) * 16)) & 0xffffu);
          vmax = max(vmax, runX + pv);
        }
        dw[j >> 3] |= code << ((j & 7) * 4);
      }
      lastH = lg + g;
      lastX = runX;
      if (MODE == FWD && vmax - m > best.val) {
        // rare: this row chunk improves the thread's best -> locate the first column that attains vmax
        int run = runIn;
#pragma unroll
        for (int j = 0; j < C; ++j) {
          run = max(run, up[j]);
          const int pv = (int) (int16_t) ((part[j >> 1] >> ((j & 1) * 16)) & 0xffffu);
          if (run + pv == vmax && vmax - m > best.val) { best.val = vmax - m; best.row = r; best.col = c0 + j; best.bm = run - r; }
        }
      }
      if (MODE == REV) {
        if (nvalid < C) {  // columns beyond n carry "no partner"
#pragma unroll
          for (int j = 0; j < C; ++j) {
            if (j >= nvalid) {
              const int q = C - 1 - j;
              part[q >> 1] = (part[q >> 1] & ~(0xffffu << ((q & 1) * 16))) | (0x8000u << ((q & 1) * 16));
            }
          }
        }
        uint4* o = (uint4*) (brev + (size_t) r * bstride + (P - c0 - C));
#pragma unroll
        for (int w = 0; w < WPT; ++w) o[w] = make_uint4(part[4 * w], part[4 * w + 1], part[4 * w + 2], part[4 * w + 3]);
        if (first) brev[(size_t) r * bstride + (P - 1)] = (int16_t) 0;  // x = 0: H = -r, shifted 0
      }
      if (MODE != PLAIN || dirs != nullptr) {
        uint32_t* drow = dirs + (size_t) r * dstride + (c0 - cstart) / 8;
#pragma unroll
        for (int w = 0; w < WPT; ++w) drow[w] = dw[w];
      }
      if (MODE == PLAIN && colout != nullptr && tid == ownerN) {
        int v = up[0];
#pragma unroll
        for (int j = 1; j < C; ++j) if (j == (n - cstart) % C) v = up[j];
        colout[r] = v - r;
      }
    }
    prevRecvH = recvH;
  }
  // corner H[m][n]
  {
    const int owner = (n - cstart) / C;
    __shared__ int sm_corner;
    if (tid == owner) {
      int v = up[0];
#pragma unroll
      for (int j = 1; j < C; ++j) if (j == (n - cstart) % C) v = up[j];
      sm_corner = v - m;   // registers hold H + row
    }
    __syncthreads();
    corner = sm_corner;
    __syncthreads();
  }
}

}  // namespace wf
