// myers.cuh — pieces shared by the bit-parallel kernels (edit_distance.cu, edit_path.cu): Myers' 64-row block update and the geometry of
// the banded passes (a staircase of G 64-row blocks per job on a sub-warp lane group; see the comment above ed_band_kernel).
#pragma once
#include "common.cuh"

// 64-row block update with signed hin/hout (src/edlib.cpp:407-442, Myers' Advance_Block).
__device__ __forceinline__ int block64(uint64_t& Pv, uint64_t& Mv, uint64_t Eq, int hin, uint64_t& PhOut, uint64_t& MhOut) {
  uint64_t hinNeg = (uint64_t) ((uint32_t) hin >> 31);
  uint64_t Xv = Eq | Mv;
  Eq |= hinNeg;
  uint64_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
  uint64_t Ph = Mv | ~(Xh | Pv);
  uint64_t Mh = Pv & Xh;
  PhOut = Ph;  // pre-shift horizontal deltas: bit r = delta of row r of this block
  MhOut = Mh;
  int hout = (int) (Ph >> 63) - (int) (Mh >> 63);
  Ph <<= 1;
  Mh <<= 1;
  Mh |= hinNeg;
  Ph |= (uint64_t) ((uint32_t) (hin + 1) >> 1);
  Pv = Mh | ~(Xv | Ph);
  Mv = Ph & Xv;
  return hout;
}


constexpr int EDB_WARPS = 4;
constexpr int EDB_CLASSES = 5;          // G = 2, 4, 8, 16, 32
#ifndef EDB_PEQ_WORDS_V
#define EDB_PEQ_WORDS_V 1280
#endif
constexpr int EDB_PEQ_WORDS = EDB_PEQ_WORDS_V;   // 64-bit Peq words per warp, split over the warp's 32/G jobs
constexpr int EDB_BLOCK_WORDS = 6;      // per 64-row block: match masks of A, C, G, T, N + the rows holding any other byte
__host__ __device__ constexpr int edb_lanes(int c) { return 2 << c; }
__host__ __device__ constexpr int edb_block_cap(int c) { return EDB_PEQ_WORDS / EDB_BLOCK_WORDS / (32 / edb_lanes(c)); }  // 13, 26, 53, 106, 213 blocks

struct BandPlan { int A; int kvalid; bool whole; };

__host__ __device__ inline BandPlan band_plan(int G, int m, int n) {
  const int d = n - m, up = d > 0 ? d : 0, lo = d < 0 ? -d : 0;
  BandPlan b; b.A = 0; int best = -0x7fffffff;
  for (int A = 0; A < G; ++A) {
    const int h1 = 64 * A - up, h2 = 64 * (G - 1 - A) - lo;
    const int h = h1 < h2 ? h1 : h2;
    if (h > best) { best = h; b.A = A; }
  }
  b.kvalid = best >= 0 ? 2 * best + up + lo : -1;
  b.whole = ((n - 1) >> 6) <= b.A && 64 * (G - b.A) >= m;
  return b;
}


// generalised equality (edlib's additionalEqualities, src/edlib.cpp:58-79) as three 256-entry tables; cls == nullptr: plain byte equality
struct EqTabs {
  const uint32_t* f;   // bitmask of the pairs in which the byte is the first member
  const uint32_t* s;   // ... the second member
  const uint8_t* cls;  // bitmask over {A,C,G,T,N} of the class symbols the byte is equal to
};

// Match masks of the job's 64-row blocks: lane p of the G-lane group builds blocks p, p+G, ... (rows read back to front when rev).
template <int G>
__device__ __forceinline__ void band_build_peq(uint64_t* peq, const uint8_t* q, int m, bool rev, int p) {
  const int nblk = (m + 63) >> 6;
  for (int bb = p; bb < nblk; bb += G) {
    uint32_t pl[EDB_BLOCK_WORDS] = {0, 0, 0, 0, 0, 0}, ph[EDB_BLOCK_WORDS] = {0, 0, 0, 0, 0, 0};
    const int r0 = bb * 64, rows = min(64, m - r0);
    for (int i = 0; i < rows; ++i) {
      const uint32_t code = dna_code(__ldg(rev ? q + (m - 1 - r0 - i) : q + r0 + i));
      const uint32_t bit = 1u << (i & 31);
#pragma unroll
      for (int sy = 0; sy < EDB_BLOCK_WORDS; ++sy) {
        const uint32_t v = (code == (uint32_t) sy) ? bit : 0u;
        if (i < 32) pl[sy] |= v; else ph[sy] |= v;
      }
    }
#pragma unroll
    for (int sy = 0; sy < EDB_BLOCK_WORDS; ++sy) peq[bb * EDB_BLOCK_WORDS + sy] = ((uint64_t) ph[sy] << 32) | pl[sy];
  }
}

// Rows of block blk equal to the target byte ch.
template <bool EQ>
__device__ __forceinline__ uint64_t band_eq(const uint64_t* peq, int blk, uint32_t ch, const uint8_t* q, int m, bool rev, const EqTabs& tabs) {
  const uint64_t* w = peq + blk * EDB_BLOCK_WORDS;
  uint64_t Eq = 0;
  uint64_t rest = w[5];   // rows whose byte is outside ACGTN: compared byte by byte (exact equality for IUPAC, lower case, gaps)
  if (EQ) {
    const uint32_t cm = tabs.cls[ch & 0xffu];
#pragma unroll
    for (int sy = 0; sy < 5; ++sy)
      if ((cm >> sy) & 1u) Eq |= w[sy];
    if (rest) {
      const uint32_t cf = tabs.f[ch & 0xffu], cs = tabs.s[ch & 0xffu];
      while (rest) {
        const int i = __ffsll((long long) rest) - 1;
        rest &= rest - 1;
        const int row = blk * 64 + i;
        const uint32_t qb = __ldg(rev ? q + (m - 1 - row) : q + row);
        if (qb == (ch & 0xffu) || (tabs.f[qb] & cs) || (tabs.s[qb] & cf)) Eq |= 1ull << i;
      }
    }
  } else {
    const uint32_t code = dna_code(ch);
    if (code < 5) Eq = w[code];
    else
      while (rest) {
        const int i = __ffsll((long long) rest) - 1;
        rest &= rest - 1;
        const int row = blk * 64 + i;
        if (__ldg(rev ? q + (m - 1 - row) : q + row) == (uint8_t) ch) Eq |= 1ull << i;
      }
  }
  return Eq;
}
