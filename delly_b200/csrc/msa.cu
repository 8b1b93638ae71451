// msa.cu — batched per-cluster multiple sequence alignment + consensus, the device replacement for
// msa() (src/msa.h:185-239) as called per SV by assembleSplitReads (src/shortpe.h:185,252).
//
// Reference semantics reproduced bit-for-bit, ONE WARP PER CLUSTER (4 independent clusters per CTA):
//   * distanceMatrix: d[i][j] = lcs(si,sj)*100 / min(|si|,|sj|) (src/msa.h:10-44). LCS length is
//     computed bit-parallel (Allison-Dix/Hyyro: V' = (V + (V & M)) | (V & ~M), LCS = #zeros),
//     one lane per read pair; the value is the same LCS length the reference's DP returns.
//   * upgma: first strict maximum in row-major order, integer-average update, -1 masking (:46-89).
//   * palign: progressive gotoh(left,right) over the guide tree (:91-109). Nodes are evaluated in
//     creation order (children always precede parents), which yields the same alignments as the
//     reference's post-order recursion because each node is a pure function of its children.
//   * gotoh with AlignConfig<true,true> (free end gaps both ways), DnaScore(match,mismatch,go,ge)
//     (src/gotoh.h:71-174): exact s/v/h values, the 4 trace bitsets with the reference's tie rules
//     (bit3 before bit4; bit1/bit2 = "gap was opened here"), the h/v/s traceback state machine, and
//     _createAlignment's row order (src/align.h:202-229).
//   * _score: char compare when both sides have one row, else the float profile product summed
//     k1-outer/k2-inner in IEEE single without contraction and truncated (src/align.h:96-111);
//     _createProfile incl. the first..last aligned span rule (src/align.h:128-171).
//   * consensus: coverage threshold max(2,min(minCliqueSize,rows)), first-max vote over A,C,G,T,other,
//     gaps dropped (src/msa.h:111-173).
//
// DP parallelisation: anti-diagonal wavefront inside the warp — lane l owns C consecutive DP columns in
// registers and computes row (step - l) for them; the affine state (s, h) of its last column reaches lane l+1
// by shuffle, the vertical state v stays in registers. No block barriers anywhere; a CTA is just four
// independent warps sharing an SM slot. The 25-term float _score is evaluated once per pair of DISTINCT column
// profiles (columns whose integer counts agree have bit-identical float profiles) into an int8 table in shared
// memory and looked up per cell. Trace nibbles go to an L2-resident per-warp slab.
#include "common.cuh"
#include <algorithm>

namespace {

constexpr int MSA_WARPS = 4;        // independent clusters per CTA
#ifndef MSA_PER_SM
#define MSA_PER_SM 5
#endif
constexpr int MSA_T = 32 * MSA_WARPS;
constexpr int MSA_MAXR = 32;        // reads per cluster (reference default maxReadPerSV = 20)
constexpr int MSA_LCAP = 1023;      // max alignment columns (DP has LCAP+1 columns)
constexpr int MSA_N = 2 * MSA_MAXR + 1;
constexpr int MSA_NEGINF = 1000000; // DnaScore::inf (src/align.h:21)
constexpr int LCS_W = 8;            // 32-bit words per read in the bit-parallel LCS (<= 256 bp)
constexpr int MSA_UMAX = 64;        // distinct column profiles per side the de-duplication tracks (more: every column is its own profile)
constexpr int MSA_SSMEM = MSA_UMAX * MSA_UMAX;  // score-table entries that fit shared memory; larger tables live in the warp's slab
constexpr int MSA_HASH = 128;       // open-addressing slots for the profile de-duplication

constexpr int ST_OK = 0, ST_TOO_MANY = 1, ST_TOO_LONG = 2, ST_BAD_CHAR = 3;

struct MsaArgs {
  const uint8_t* seqs;
  const uint32_t* read_off;
  const uint32_t* read_len;
  const uint32_t* cluster_off;  // nclusters+1 indices into the read arrays
  uint32_t nclusters;
  int match, mismatch, go, ge, min_clique;
  uint8_t* cons;
  const uint64_t* cons_off;
  uint32_t* cons_len;
  uint32_t* n_rows;
  uint32_t* status;
  uint8_t* aln;             // optional root alignment dump
  const uint64_t* aln_off;
  uint32_t* aln_cols;
  uint8_t* work;            // one slab per warp
  size_t work_stride;
  size_t off_prof, off_trace, off_lcs, off_keys, off_src, off_bnd, off_tab, off_spans;
  size_t aln_cap;           // bytes of node alignment storage per warp (at slab offset 0)
};

struct NodeAln {
  const uint8_t* p;  // rows x L, row stride L
  int R, L;
};

// Per-warp shared state (about 10 KB, so that five 4-warp CTAs fit one SM)
constexpr int MSA_TRI = MSA_N * (MSA_N - 1) / 2;
struct WarpSm {
  int8_t d[MSA_TRI];                       // guide-tree similarity matrix, upper triangle (values -2..100)
  int8_t par[MSA_N], lc[MSA_N], rc[MSA_N];
  uint8_t nodeR[MSA_N];
  uint16_t nodeL[MSA_N];
  uint32_t nodeP[MSA_N];                   // leaf: offset into the read arena; inner node: offset into the warp's slab
  int16_t first1[MSA_MAXR], last1[MSA_MAXR], first2[MSA_MAXR], last2[MSA_MAXR];
  uint32_t win[64];                        // traceback window: 32 rows x 2 trace words
  union {
    uint32_t peq[MSA_MAXR * 5 * LCS_W];    // LCS phase
    struct {
      union {
        int8_t S[MSA_UMAX * MSA_UMAX];     // exact _score per pair of distinct column profiles
        struct {                           // profile de-duplication (finished before S is built)
          unsigned long long hkey[MSA_HASH];
          int hval[MSA_HASH];
          int hcount;
        } h;
      } t;
      unsigned long long uniq1[MSA_UMAX], uniq2[MSA_UMAX];
    } g;
  } u;
};

__device__ __forceinline__ int tri(int i, int j) {  // i < j < MSA_N
  return i * (2 * MSA_N - i - 1) / 2 + (j - i - 1);
}

// _createProfile (src/align.h:128-171) in integer form: key[j] = the six integers column j's profile is made of (counts of
// A,C,G,T,N and the covering-row count over the first..last aligned span of each row, 6 bits each). The float profile
// p[k][j] = count_k / rows is a pure function of the key, so two columns with equal keys have bit-identical profiles —
// which is what the score table exploits. The '-' row of the reference's profile is never read by _score.
__device__ void make_profile_keys(const NodeAln& a, const int16_t* first, const int16_t* last, unsigned long long* key, int lane) {
  for (int j = lane; j < a.L; j += 32) {
    int cnt[5] = {0, 0, 0, 0, 0};
    int sum = 0;
    for (int r = 0; r < a.R; ++r) {
      int f = first[r], l = last[r];
      if (f < 0) { f = -1; l = a.L; }  // all-gap row: the reference's initial values cover everything
      if (f <= j && j <= l) {
        ++sum;
        uint8_t c = a.p[(size_t) r * a.L + j];
        if (c == 'A' || c == 'a') ++cnt[0];
        else if (c == 'C' || c == 'c') ++cnt[1];
        else if (c == 'G' || c == 'g') ++cnt[2];
        else if (c == 'T' || c == 't') ++cnt[3];
        else if (c == 'N' || c == 'n') ++cnt[4];
        else if (c == '-') {}
        else --sum;
      }
    }
    unsigned long long kk = (unsigned long long) (uint32_t) sum;
#pragma unroll
    for (int k = 0; k < 5; ++k) kk = (kk << 6) | (unsigned long long) (uint32_t) cnt[k];
    key[j] = kk + 1ull;  // never 0: 0 marks an empty hash slot
  }
  __syncwarp();
}

// Map every column to the index of its distinct profile through a small open-addressing table.
// Returns the number of distinct profiles, or -1 if there are more than MSA_UMAX (the caller then treats every column as its own profile).
__device__ int dedup_profiles(WarpSm& w, const unsigned long long* key, int L, uint16_t* ids, unsigned long long* uniq, int lane) {
  for (int i = lane; i < MSA_HASH; i += 32) { w.u.g.t.h.hkey[i] = 0ull; w.u.g.t.h.hval[i] = -1; }
  if (lane == 0) w.u.g.t.h.hcount = 0;
  __syncwarp();
  bool overflow = false;
  for (int j0 = 0; j0 < L; j0 += 32) {
    const int j = j0 + lane;
    if (j < L) {
      const unsigned long long k = key[j];
      uint32_t h = (uint32_t) ((k * 0x9E3779B97F4A7C15ull) >> 57);  // 7 bits
      int id = -1;
      for (int probe = 0; probe < MSA_HASH; ++probe) {
        const unsigned long long prev = atomicCAS(&w.u.g.t.h.hkey[h], 0ull, k);
        if (prev == 0ull) {  // this lane created the slot
          const int nid = atomicAdd(&w.u.g.t.h.hcount, 1);
          if (nid < MSA_UMAX) uniq[nid] = k;
          __threadfence_block();
          atomicExch(&w.u.g.t.h.hval[h], nid);
          id = nid;
          break;
        }
        if (prev == k) {  // somebody (maybe in this very instruction) owns the slot: wait for its id
          int v;
          do { v = atomicAdd(&w.u.g.t.h.hval[h], 0); } while (v < 0);
          id = v;
          break;
        }
        h = (h + 1) & (MSA_HASH - 1);
      }
      if (id < 0 || id >= MSA_UMAX) overflow = true; else ids[j] = (uint16_t) id;
    }
    __syncwarp();
  }
  overflow = __any_sync(0xffffffffu, overflow);
  const int U = w.u.g.t.h.hcount;
  __syncwarp();
  return (overflow || U > MSA_UMAX) ? -1 : U;
}

// Exact _score (src/align.h:104-109) for every pair of distinct profiles: S[u1 * U2 + u2] (fits int8: |score| <= max(|match|,|mismatch|)).
// The 25 products are summed k1-outer / k2-inner in IEEE single without contraction, then truncated, as the reference does per cell.
__device__ void build_score_table(const unsigned long long* uniq1, int U1, const unsigned long long* uniq2, int U2, int match, int mismatch, int8_t* S, int lane) {
  const float fm = (float) match, fx = (float) mismatch;
  for (int x = lane; x < U1 * U2; x += 32) {
    const int u1 = x / U2, u2 = x % U2;
    float p1[5], p2[5];
    unsigned long long k1 = uniq1[u1] - 1ull, k2 = uniq2[u2] - 1ull;
    const float s1 = (float) (uint32_t) (k1 >> 30), s2 = (float) (uint32_t) (k2 >> 30);
#pragma unroll
    for (int k = 4; k >= 0; --k) {
      p1[k] = __fdiv_rn((float) (uint32_t) (k1 & 63ull), s1); k1 >>= 6;
      p2[k] = __fdiv_rn((float) (uint32_t) (k2 & 63ull), s2); k2 >>= 6;
    }
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 5; ++a)
#pragma unroll
      for (int b = 0; b < 5; ++b) acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(p1[a], p2[b]), (a == b) ? fm : fx));
    S[x] = (int8_t) (int) acc;
  }
  __syncwarp();
}

// The DP of gotoh() as a warp wavefront over one strip of 32*C columns starting at cbase: lane l owns columns
// cbase + l*C .. cbase + l*C + C-1 (column 0 of the matrix is the fixed boundary s = v = 0, h = -inf). Strips after the
// first take the (s, h) of the column left of them from bnd_in[row] and every strip that has a successor leaves its last
// column in bnd_out[row]. Writes one trace nibble per cell: bit0 = bit1, bit1 = bit2, bit2 = bit3, bit3 = bit4 of
// src/gotoh.h:135-138. tr: row-major, rowwords words per row, cell (r,c) at word c/8, nibble c%8.
// The substitution score of cell (r,c) is S[id1[r-1] * sstride + id2[c-1]].
#define MSA_SETBIT_GT(dw, a, b, imm) asm("{.reg .pred p; setp.gt.s32 p, %1, %2; @p or.b32 %0, %0, %3;}" : "+r"(dw) : "r"(a), "r"(b), "r"(imm))
#define MSA_SETBIT_EQ(dw, a, b, imm) asm("{.reg .pred p; setp.eq.s32 p, %1, %2; @p or.b32 %0, %0, %3;}" : "+r"(dw) : "r"(a), "r"(b), "r"(imm))

// BIG = false: the score table is in shared memory (32-bit addressing, one add + one load per cell);
// BIG = true: it is in the warp's global slab (more than MSA_SSMEM distinct profile pairs; rare).
// Trace bit 3 is stored as plain (s == v); the traceback tests bit 2 (s == h) first, as the reference's else-if does.
template <int C, bool BIG>
__device__ __noinline__ void gotoh_wave(const int m, const int n, const int o, const int e, uint32_t* __restrict__ tr, const int rowwords,
                                        const uint16_t* __restrict__ id1, const uint16_t* __restrict__ id2, const int8_t* S, const int sstride,
                                        const int cbase, const int2* __restrict__ bnd_in, int2* __restrict__ bnd_out, const int lane) {
  const int c0 = cbase + lane * C;
  constexpr int WPT = (C + 7) / 8;
  const bool owns = c0 <= n;
  const bool has_next = cbase + 32 * C <= n;
  const int nact = min(32, (n - cbase) / C + 1);  // lanes that own a column <= n
  const uint32_t sbase = BIG ? 0u : (uint32_t) __cvta_generic_to_shared(S);

  uint32_t cx[C];   // the column's distinct-profile index (plus the table's shared-memory address when !BIG)
  int vgo[C], vge[C];  // vertical gap costs: zero in the last column (free end gaps, AlignConfig<true,true>)
#pragma unroll
  for (int j = 0; j < C; ++j) {
    const int c = c0 + j;
    cx[j] = sbase + ((c >= 1 && c <= n) ? (uint32_t) id2[c - 1] : 0u);
    vgo[j] = (c == n) ? 0 : o;
    vge[j] = (c == n) ? 0 : e;
  }
  int sprev[C], vprev[C];
#pragma unroll
  for (int j = 0; j < C; ++j) { sprev[j] = 0; vprev[j] = -MSA_NEGINF; }
  // row 0: bit3 for every column >= 1 (src/gotoh.h:113-117)
  if (owns) {
    if (C >= 8) {
#pragma unroll
      for (int w = 0; w < WPT; ++w) {
        uint32_t bits = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) { const int c = c0 + 8 * w + q; if (c >= 1 && c <= n) bits |= 4u << (4 * q); }
        if (c0 + 8 * w <= n) tr[(c0 >> 3) + w] = bits;
      }
    } else {  // C == 4: half a trace word per lane
      uint32_t bits = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) { const int c = c0 + q; if (c >= 1 && c <= n) bits |= 4u << (4 * q); }
      ((uint16_t*) tr)[c0 >> 2] = (uint16_t) bits;
    }
  }
  int lastS = 0, lastH = 0, prevRecvS = 0;
  const int nsteps = m + nact - 1;
  int2 nextb = make_int2(0, 0);
  if (cbase > 0 && lane == 0) nextb = bnd_in[1];
  // the row profile index is fetched one step ahead of its use
  int nextid = 0;
  if (owns && lane == 0) nextid = id1[0];
  for (int st = 1; st <= nsteps; ++st) {
    int recvS = __shfl_up_sync(0xffffffffu, lastS, 1);
    int recvH = __shfl_up_sync(0xffffffffu, lastH, 1);
    if (cbase > 0 && lane == 0) {
      recvS = nextb.x; recvH = nextb.y;
      if (st < m) nextb = bnd_in[st + 1];
    }
    const int r = st - lane;
    const int roff = nextid * sstride;
    if (owns && r >= 0 && r < m) nextid = id1[r];  // for row r + 1
    if (r >= 1 && r <= m && owns) {
      const int er = (r == m) ? 0 : e;
      const int orr = (r == m) ? 0 : o;
      int sc[C];
#pragma unroll
      for (int j = 0; j < C; ++j) {
        if (BIG) sc[j] = (int) S[roff + (int) cx[j]];
        else asm("ld.shared.s8 %0, [%1];" : "=r"(sc[j]) : "r"(cx[j] + (uint32_t) roff));
      }
      int sleft = recvS, hleft = recvH, diag = prevRecvS;
      uint32_t dw[WPT];
#pragma unroll
      for (int w = 0; w < WPT; ++w) dw[w] = 0;
#pragma unroll
      for (int j = 0; j < C; ++j) {
        const int vext = vprev[j] + vge[j];
        const int vopen = sprev[j] + vgo[j];
        int v = max(vopen, vext);
        const int hext = hleft + er;
        const int hopen = sleft + orr;
        int h = max(hopen, hext);
        int s = max(max(diag + sc[j], h), v);
        if (j == 0) {  // column 0 of the matrix (src/gotoh.h:118-123) sits in lane 0 of the first strip
          uint32_t b0 = 0;
          MSA_SETBIT_GT(b0, hopen, hext, 1);
          MSA_SETBIT_GT(b0, vopen, vext, 2);
          MSA_SETBIT_EQ(b0, s, h, 4);
          MSA_SETBIT_EQ(b0, s, v, 8);
          const bool z = (c0 == 0);
          s = z ? 0 : s; v = z ? 0 : v; h = z ? -MSA_NEGINF : h;
          dw[0] = z ? 8u : b0;
        } else {
          MSA_SETBIT_GT(dw[j >> 3], hopen, hext, 1 << ((j & 7) * 4));
          MSA_SETBIT_GT(dw[j >> 3], vopen, vext, 2 << ((j & 7) * 4));
          MSA_SETBIT_EQ(dw[j >> 3], s, h, 4 << ((j & 7) * 4));
          MSA_SETBIT_EQ(dw[j >> 3], s, v, 8 << ((j & 7) * 4));
        }
        diag = sprev[j];
        sprev[j] = s; vprev[j] = v;
        sleft = s; hleft = h;
      }
      lastS = sleft; lastH = hleft;
      if (C >= 8) {
        uint32_t* drow = tr + (size_t) r * rowwords + (c0 >> 3);
#pragma unroll
        for (int w = 0; w < WPT; ++w)
          if (c0 + 8 * w <= n) drow[w] = dw[w];
      } else {
        ((uint16_t*) (tr + (size_t) r * rowwords))[c0 >> 2] = (uint16_t) dw[0];
      }
      if (has_next && lane == 31) bnd_out[r] = make_int2(lastS, lastH);
    }
    prevRecvS = recvS;
  }
  __syncwarp();
}

__device__ __forceinline__ void gotoh_dp(const int m, const int n, const MsaArgs& A, uint32_t* tr, const int rowwords,
                                         const uint16_t* id1, const uint16_t* id2, const int8_t* S, const int sstride, const bool big, int2* bnd, const int lane) {
  const int o = A.go + A.ge, e = A.ge;
  if (n + 1 <= 32 * 4) {
    if (!big) gotoh_wave<4, false>(m, n, o, e, tr, rowwords, id1, id2, S, sstride, 0, nullptr, nullptr, lane);
    else gotoh_wave<4, true>(m, n, o, e, tr, rowwords, id1, id2, S, sstride, 0, nullptr, nullptr, lane);
  } else {
    int2* bin = bnd;
    int2* bout = bnd + (MSA_LCAP + 1);
    for (int cbase = 0; cbase <= n; cbase += 32 * 8) {
      if (!big) gotoh_wave<8, false>(m, n, o, e, tr, rowwords, id1, id2, S, sstride, cbase, bin, bout, lane);
      else gotoh_wave<8, true>(m, n, o, e, tr, rowwords, id1, id2, S, sstride, cbase, bin, bout, lane);
      int2* t = bin; bin = bout; bout = t;
    }
  }
}

// bit-parallel LCS length of reads a (Peq masks in pa[5][LCS_W]) and b (codes 0..4)
__device__ int lcs_bitpar(const uint32_t* pa, int la, const uint8_t* b, int lb) {
  const int nw = (la + 31) >> 5;
  uint32_t V[LCS_W];
#pragma unroll
  for (int w = 0; w < LCS_W; ++w) V[w] = 0xffffffffu;
  for (int j = 0; j < lb; ++j) {
    const uint32_t code = dna_code(b[j]);
    uint32_t carry = 0;
#pragma unroll
    for (int w = 0; w < LCS_W; ++w) {
      if (w < nw) {
        const uint32_t M = (code < 5) ? pa[code * LCS_W + w] : 0u;
        const uint32_t u = V[w] & M;
        const uint32_t s1 = V[w] + u;
        const uint32_t c1 = s1 < u;
        const uint32_t s2 = s1 + carry;
        const uint32_t c2 = s2 < carry;
        V[w] = s2 | (V[w] & ~M);
        carry = c1 | c2;
      }
    }
  }
  int zeros = 0;
#pragma unroll
  for (int w = 0; w < LCS_W; ++w) {
    if (w < nw) {
      uint32_t x = ~V[w];
      if (w == nw - 1 && (la & 31)) x &= (1u << (la & 31)) - 1u;
      zeros += __popc(x);
    }
  }
  return zeros;
}

__global__ void __launch_bounds__(MSA_T, MSA_PER_SM) msa_kernel(MsaArgs A) {
  extern __shared__ __align__(16) uint8_t msa_smem[];
  WarpSm* wsm = reinterpret_cast<WarpSm*>(msa_smem);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  WarpSm& w = wsm[wib];
  const uint32_t gwarp = blockIdx.x * MSA_WARPS + wib, nwarps = gridDim.x * MSA_WARPS;
  uint8_t* slab = A.work + (size_t) gwarp * A.work_stride;
  uint8_t* alnbuf = slab;
  uint8_t* gapped = slab + A.off_prof;
  uint16_t* ids1 = (uint16_t*) (slab + A.off_prof + (MSA_LCAP + 1));
  uint16_t* ids2 = ids1 + (MSA_LCAP + 1);
  int8_t* bigtab = (int8_t*) (slab + A.off_tab);
  uint32_t* tr = (uint32_t*) (slab + A.off_trace);
  int* lcsrow = (int*) (slab + A.off_lcs);                          // 32 DP rows (one per lane) for the long-pair LCS
  int8_t* lcstmp = (int8_t*) (lcsrow + 32 * (MSA_LCAP + 2));      // its results, before they enter the similarity matrix
  unsigned long long* keys1 = (unsigned long long*) (slab + A.off_keys);
  unsigned long long* keys2 = keys1 + (MSA_LCAP + 1);
  int16_t* src1 = (int16_t*) (slab + A.off_src);
  int16_t* src2 = src1 + (2 * MSA_LCAP + 4);
  int2* bnd = (int2*) (slab + A.off_bnd);
  int16_t* spans = (int16_t*) (slab + A.off_spans);  // per node: first[MSA_MAXR], last[MSA_MAXR]

  for (uint32_t cl = gwarp; cl < A.nclusters; cl += nwarps) {
    __syncwarp();
    const uint32_t rbeg = A.cluster_off[cl], rend = A.cluster_off[cl + 1];
    const int num = (int) (rend - rbeg);
    int status = ST_OK;
    if (num > MSA_MAXR) status = ST_TOO_MANY;
    if (status == ST_OK) {
      // validate: lengths and alphabet (bytes outside ACGTN make the reference's float profile NaN)
      int bad = 0;
      for (int i = 0; i < num; ++i) {
        const uint32_t L = A.read_len[rbeg + i];
        if (L > MSA_LCAP || L == 0) bad |= 2;
        const uint8_t* s = A.seqs + A.read_off[rbeg + i];
        for (uint32_t j = lane; j < L; j += 32)
          if (dna_code(s[j]) > 4) bad |= 1;
      }
      bad = __reduce_or_sync(0xffffffffu, (unsigned) bad);
      if (bad & 2) status = ST_TOO_LONG; else if (bad & 1) status = ST_BAD_CHAR;
    }
    if (status != ST_OK || num == 0) {
      if (lane == 0) { A.status[cl] = status; A.cons_len[cl] = 0; A.n_rows[cl] = 0; if (A.aln_cols) A.aln_cols[cl] = 0; }
      continue;
    }
    const int N = 2 * num + 1;

    // ---- distance matrix (src/msa.h:32-44, :190-195) ---------------------------------
    for (int x = lane; x < MSA_TRI; x += 32) w.d[x] = -1;
    for (int x = lane; x < N; x += 32) { w.par[x] = -1; w.lc[x] = -1; w.rc[x] = -1; }
    for (int x = lane; x < num * 5 * LCS_W; x += 32) w.u.peq[x] = 0;
    __syncwarp();
    for (int i = 0; i < num; ++i) {
      const uint32_t L = A.read_len[rbeg + i];
      const uint8_t* s = A.seqs + A.read_off[rbeg + i];
      if (L <= 32 * LCS_W) {
        for (uint32_t j0 = 0; j0 < L; j0 += 32) {
          uint32_t j = j0 + lane;
          uint32_t code = (j < L) ? dna_code(s[j]) : 7u;
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            unsigned b = __ballot_sync(0xffffffffu, code == (uint32_t) k);
            if (lane == 0) w.u.peq[(i * 5 + k) * LCS_W + (j0 >> 5)] = b;
          }
        }
      }
    }
    __syncwarp();
    {
      const int npairs = num * (num - 1) / 2;
      for (int pidx = lane; pidx < npairs; pidx += 32) {
        int i = 0, rem = pidx;  // unrank (i<j) in row-major order
        while (rem >= num - 1 - i) { rem -= num - 1 - i; ++i; }
        const int j = i + 1 + rem;
        const int li = (int) A.read_len[rbeg + i], lj = (int) A.read_len[rbeg + j];
        int l;
        if (li <= 32 * LCS_W) l = lcs_bitpar(&w.u.peq[i * 5 * LCS_W], li, A.seqs + A.read_off[rbeg + j], lj);
        else if (lj <= 32 * LCS_W) l = lcs_bitpar(&w.u.peq[j * 5 * LCS_W], lj, A.seqs + A.read_off[rbeg + i], li);
        else l = -2;  // long pair: plain DP below
        w.d[tri(i, j)] = (l >= 0) ? (int8_t) ((l * 100) / min(li, lj)) : (int8_t) -2;
      }
      __syncwarp();
      {  // plain LCS DP for pairs of two long reads (reads beyond the 256-column bit vectors, e.g. 2 x 300 bp libraries): one pair per lane,
         // each lane with its own DP row in the warp's slab
        int* myrow = lcsrow + (size_t) lane * (MSA_LCAP + 2);
        int seen = 0;
        for (int i = 0; i < num; ++i)
          for (int j = i + 1; j < num; ++j)
            if (w.d[tri(i, j)] == -2) {
              if ((seen++ & 31) != lane) continue;
              const int li = (int) A.read_len[rbeg + i], lj = (int) A.read_len[rbeg + j];
              const uint8_t* si = A.seqs + A.read_off[rbeg + i];
              const uint8_t* sj = A.seqs + A.read_off[rbeg + j];
              for (int c = 0; c <= lj; ++c) myrow[c] = 0;
              for (int a = 1; a <= li; ++a) {
                int diag = 0;
                const uint8_t ca = si[a - 1];
                for (int c = 1; c <= lj; ++c) {
                  const int up = myrow[c];
                  myrow[c] = (ca == sj[c - 1]) ? diag + 1 : max(up, myrow[c - 1]);
                  diag = up;
                }
              }
              lcstmp[tri(i, j)] = (int8_t) ((myrow[lj] * 100) / min(li, lj));   // into the matrix after the scan: the other lanes still count the -2 marks
            }
        __syncwarp();
        for (int x = lane; x < MSA_TRI; x += 32)
          if (w.d[x] == -2) w.d[x] = lcstmp[x];
      }
      __syncwarp();
    }

    // ---- UPGMA (src/msa.h:46-89) ---------------------------------------------------------
    int nn = num;
    for (; nn < N; ++nn) {
      unsigned long long key = 0;
      for (int x = lane; x < nn * nn; x += 32) {
        int i = x / nn, j = x % nn;
        if (j > i) {
          int v = w.d[tri(i, j)];
          unsigned long long k2 = ((unsigned long long) (uint32_t) (v + 2) << 32) | (uint32_t) (0x7fffffff - (i * MSA_N + j));
          key = k2 > key ? k2 : key;
        }
      }
#pragma unroll
      for (int dd = 16; dd >= 1; dd >>= 1) {
        unsigned long long y = __shfl_xor_sync(0xffffffffu, key, dd);
        key = y > key ? y : key;
      }
      const int dMax = (int) (uint32_t) (key >> 32) - 2;
      if (key == 0 || dMax == -1) break;
      const int code = 0x7fffffff - (int) (uint32_t) (key & 0xffffffffu);
      const int dI = code / MSA_N, dJ = code % MSA_N;
      if (lane == 0) { w.par[dI] = (int8_t) nn; w.par[dJ] = (int8_t) nn; w.lc[nn] = (int8_t) dI; w.rc[nn] = (int8_t) dJ; }
      __syncwarp();
      for (int i = lane; i < nn; i += 32)
        if (w.par[i] == -1)
          w.d[tri(i, nn)] = (int8_t) ((((dI < i) ? w.d[tri(dI, i)] : w.d[tri(i, dI)]) + ((dJ < i) ? w.d[tri(dJ, i)] : w.d[tri(i, dJ)])) / 2);
      __syncwarp();
      for (int i = lane; i < nn + 1; i += 32) {
        if (i < dI) w.d[tri(i, dI)] = -1;
        if (i > dI) w.d[tri(dI, i)] = -1;
        if (i < dJ) w.d[tri(i, dJ)] = -1;
        if (i > dJ) w.d[tri(dJ, i)] = -1;
      }
      __syncwarp();
    }
    const int root = (nn > 0) ? nn - 1 : 0;

    // ---- progressive alignment in node-creation order ---------------------------------------
    if (lane < num) {
      w.nodeR[lane] = 1;
      w.nodeL[lane] = (uint16_t) A.read_len[rbeg + lane];
      w.nodeP[lane] = A.read_off[rbeg + lane];
      spans[2 * MSA_MAXR * lane] = 0;  // a read has no gaps: aligned span = the whole read
      spans[2 * MSA_MAXR * lane + MSA_MAXR] = (int16_t) (A.read_len[rbeg + lane] - 1);
    }
    __syncwarp();
    size_t bump = 0;
    for (int node = num; node <= root && status == ST_OK; ++node) {
      NodeAln a1, a2;
      const int c1 = w.lc[node], c2 = w.rc[node];
      {
        a1.p = (c1 < num ? A.seqs : alnbuf) + w.nodeP[c1]; a1.R = w.nodeR[c1]; a1.L = w.nodeL[c1];
        a2.p = (c2 < num ? A.seqs : alnbuf) + w.nodeP[c2]; a2.R = w.nodeR[c2]; a2.L = w.nodeL[c2];
      }
      const int m = a1.L, n = a2.L;
      // Substitution scores: _score (src/align.h:96-111) compares raw characters when both sides hold one row and otherwise
      // multiplies the float column profiles. Either way the score depends only on (kind of row column, kind of column column):
      // number the kinds and tabulate.
      const int8_t* Sptr;
      int sstride;
      if (a1.R == 1 && a2.R == 1) {
        for (int j = lane; j < m; j += 32) { const uint8_t ch = a1.p[j]; ids1[j] = (uint16_t) (dna_code(ch) + ((ch & 0x20) ? 5u : 0u)); }
        for (int j = lane; j < n; j += 32) { const uint8_t ch = a2.p[j]; ids2[j] = (uint16_t) (dna_code(ch) + ((ch & 0x20) ? 5u : 0u)); }
        for (int x = lane; x < 100; x += 32) w.u.g.t.S[x] = (int8_t) ((x / 10 == x % 10) ? A.match : A.mismatch);
        Sptr = w.u.g.t.S; sstride = 10;
        __syncwarp();
      } else {
        if (lane < a1.R) { w.first1[lane] = spans[2 * MSA_MAXR * c1 + lane]; w.last1[lane] = spans[2 * MSA_MAXR * c1 + MSA_MAXR + lane]; }
        if (lane < a2.R) { w.first2[lane] = spans[2 * MSA_MAXR * c2 + lane]; w.last2[lane] = spans[2 * MSA_MAXR * c2 + MSA_MAXR + lane]; }
        __syncwarp();
        make_profile_keys(a1, w.first1, w.last1, keys1, lane);
        make_profile_keys(a2, w.first2, w.last2, keys2, lane);
        const unsigned long long* u1 = w.u.g.uniq1;
        const unsigned long long* u2 = w.u.g.uniq2;
        int U1 = dedup_profiles(w, keys1, m, ids1, w.u.g.uniq1, lane);
        if (U1 < 0) { for (int j = lane; j < m; j += 32) ids1[j] = (uint16_t) j; U1 = m; u1 = keys1; }
        int U2 = dedup_profiles(w, keys2, n, ids2, w.u.g.uniq2, lane);
        if (U2 < 0) { for (int j = lane; j < n; j += 32) ids2[j] = (uint16_t) j; U2 = n; u2 = keys2; }
        __syncwarp();
        int8_t* Sw = (U1 * U2 <= MSA_SSMEM) ? w.u.g.t.S : bigtab;
        build_score_table(u1, U1, u2, U2, A.match, A.mismatch, Sw, lane);
        Sptr = Sw; sstride = U2;
      }
      const int rowwords = (n + 1 + 7) >> 3;
      gotoh_dp(m, n, A, tr, rowwords, ids1, ids2, Sptr, sstride, Sptr != w.u.g.t.S, bnd, lane);

      // traceback (src/gotoh.h:141-167): lane 0 runs the state machine over a window of trace words the warp stages in
      // shared memory (32 rows x 2 words following the diagonal); once the path reaches row 0 or column 0 the rest is a
      // pure gap run (row 0 holds only bit3, column 0 only bit4) and is filled in by all lanes.
      int row = m, col = n, k = 0, tst = 0;  // tst: 0 = 's', 1 = 'h', 2 = 'v'
      while (row > 0 && col > 0) {
        {
          const int ri = row - lane;
          const int wi = max(col - lane, 0) >> 3;
          uint32_t w1 = 0, w0 = 0;
          if (ri >= 0) {
            const uint32_t* trow = tr + (size_t) ri * rowwords;
            w1 = __ldcg(trow + wi);
            if (wi > 0) w0 = __ldcg(trow + wi - 1);
          }
          w.win[2 * lane] = w0; w.win[2 * lane + 1] = w1;
        }
        __syncwarp();
        if (lane == 0) {
          const int row0 = row, col0 = col;
          while (row > 0 && col > 0) {
            const int i = row0 - row;
            if (i >= 32) break;
            const int wsel = (col >> 3) - (max(col0 - i, 0) >> 3) + 1;
            if (wsel < 0 || wsel > 1) break;
            const uint32_t nb = (w.win[2 * i + wsel] >> ((col & 7) * 4)) & 0xfu;
            // branch-free form of the reference's loop: an 's' state that sees bit3/bit4 switches to 'h'/'v' and that
            // state consumes the same cell, so every iteration emits exactly one alignment column
            const int mode = (tst != 0) ? tst : ((nb & 4u) ? 1 : ((nb & 8u) ? 2 : 0));
            const int drow = (mode != 1), dcol = (mode != 2);
            row -= drow; col -= dcol;
            src1[k] = (int16_t) (drow ? row : -1);
            src2[k] = (int16_t) (dcol ? col : -1);
            ++k;
            tst = (mode == 0 || ((nb >> (mode - 1)) & 1u)) ? 0 : mode;
          }
        }
        __syncwarp();   // lane 0's reads of the staged window are done before the next refill (racecheck: write-after-read on w.win)
        row = __shfl_sync(0xffffffffu, row, 0);
        col = __shfl_sync(0xffffffffu, col, 0);
        k = __shfl_sync(0xffffffffu, k, 0);
      }
      for (int t = lane; t < col; t += 32) { src1[k + t] = -1; src2[k + t] = (int16_t) (col - 1 - t); }
      for (int t = lane; t < row; t += 32) { src1[k + t] = (int16_t) (row - 1 - t); src2[k + t] = -1; }
      const int L = k + row + col;
      __syncwarp();
      const int R = a1.R + a2.R;
      if (L > MSA_LCAP || bump + (size_t) R * L > A.aln_cap) { status = ST_TOO_LONG; break; }
      uint8_t* out = alnbuf + bump;
      bump += ((size_t) R * L + 15) & ~(size_t) 15;
      // _createAlignment (src/align.h:202-229): rows of a1, then rows of a2
      int16_t* inv1 = (int16_t*) keys1;  // child column -> column of the new alignment (the keys are dead after the DP)
      int16_t* inv2 = (int16_t*) keys2;
      for (int ai = lane; ai < L; ai += 32) {
        const int kk = L - 1 - ai;
        const int s1 = src1[kk], s2 = src2[kk];
        for (int i = 0; i < a1.R; ++i) out[(size_t) i * L + ai] = (s1 >= 0) ? a1.p[(size_t) i * a1.L + s1] : (uint8_t) '-';
        for (int i = 0; i < a2.R; ++i) out[(size_t) (a1.R + i) * L + ai] = (s2 >= 0) ? a2.p[(size_t) i * a2.L + s2] : (uint8_t) '-';
        if (s1 >= 0) inv1[s1] = (int16_t) ai;
        if (s2 >= 0) inv2[s2] = (int16_t) ai;
      }
      __syncwarp();
      // first/last aligned column of every row (src/align.h:137-150), carried over from the children
      if (lane < R) {
        const bool left = lane < a1.R;
        const int16_t* cs = spans + 2 * MSA_MAXR * (left ? c1 : c2);
        const int rr = left ? lane : lane - a1.R;
        const int16_t* inv = left ? inv1 : inv2;
        spans[2 * MSA_MAXR * node + lane] = inv[cs[rr]];
        spans[2 * MSA_MAXR * node + MSA_MAXR + lane] = inv[cs[MSA_MAXR + rr]];
      }
      if (lane == 0) { w.nodeR[node] = (uint8_t) R; w.nodeL[node] = (uint16_t) L; w.nodeP[node] = (uint32_t) (out - alnbuf); }
      __syncwarp();
    }
    if (status != ST_OK) {
      if (lane == 0) { A.status[cl] = status; A.cons_len[cl] = 0; A.n_rows[cl] = 0; if (A.aln_cols) A.aln_cols[cl] = 0; }
      continue;
    }

    // ---- consensus (src/msa.h:111-173) -------------------------------------------------------
    NodeAln ra;
    ra.p = (root < num ? A.seqs : alnbuf) + w.nodeP[root]; ra.R = w.nodeR[root]; ra.L = w.nodeL[root];
    if (lane < ra.R) { w.first1[lane] = spans[2 * MSA_MAXR * root + lane]; w.last1[lane] = spans[2 * MSA_MAXR * root + MSA_MAXR + lane]; }
    __syncwarp();
    const int thr = max(2, min(A.min_clique, ra.R));
    for (int j = lane; j < ra.L; j += 32) {
      int cov = 0;
      int cnt[5] = {0, 0, 0, 0, 0};
      for (int r = 0; r < ra.R; ++r) {
        if (w.first1[r] >= 0 && w.first1[r] <= j && j <= w.last1[r]) {
          ++cov;
          const uint8_t c = ra.p[(size_t) r * ra.L + j];
          if (c == 'A' || c == 'a') ++cnt[0];
          else if (c == 'C' || c == 'c') ++cnt[1];
          else if (c == 'G' || c == 'g') ++cnt[2];
          else if (c == 'T' || c == 't') ++cnt[3];
          else ++cnt[4];
        }
      }
      uint8_t g = '-';
      if (cov >= thr) {
        int mi = 0;
#pragma unroll
        for (int x = 1; x < 5; ++x) if (cnt[x] > cnt[mi]) mi = x;
        if (mi == 0) g = 'A'; else if (mi == 1) g = 'C'; else if (mi == 2) g = 'G'; else if (mi == 3) g = 'T';
      }
      gapped[j] = g;
    }
    __syncwarp();
    {
      // ordered compaction of the non-gap consensus letters (ballot prefix)
      uint8_t* o = A.cons + A.cons_off[cl];
      uint32_t k = 0;
      for (int j0 = 0; j0 < ra.L; j0 += 32) {
        const int j = j0 + lane;
        const uint8_t g = (j < ra.L) ? gapped[j] : (uint8_t) '-';
        const unsigned b = __ballot_sync(0xffffffffu, g != '-');
        if (g != '-') o[k + __popc(b & ((1u << lane) - 1u))] = g;
        k += __popc(b);
      }
      if (lane == 0) {
        A.cons_len[cl] = k;
        A.n_rows[cl] = (uint32_t) ra.R;
        A.status[cl] = ST_OK;
        if (A.aln_cols) A.aln_cols[cl] = (uint32_t) ra.L;
      }
    }
    if (A.aln) {
      uint8_t* o = A.aln + A.aln_off[cl];
      for (int x = lane; x < ra.R * ra.L; x += 32) o[x] = ra.p[x];
    }
  }
}

}  // namespace

extern "C" {

int dgpu_msa_dev(dgpu_ctx* ctx, const uint8_t* seqs, uint64_t seqs_bytes,
                 const uint32_t* read_off, const uint32_t* read_len, const uint32_t* cluster_off, uint32_t nclusters,
                 int match, int mismatch, int go, int ge, int min_clique,
                 uint8_t* cons, const uint64_t* cons_off, uint32_t* cons_len, uint32_t* n_rows, uint32_t* status,
                 uint8_t* aln, const uint64_t* aln_off, uint32_t* aln_cols, void* stream) {
  if (!ctx) return DGPU_ERR_ARG;
  if (nclusters == 0) return DGPU_OK;
  if (!seqs || !read_off || !read_len || !cluster_off || !cons || !cons_off || !cons_len || !n_rows || !status) return DGPU_ERR_ARG;
  if (seqs_bytes >= (1ull << 32)) return DGPU_ERR_ARG;
  if (match > 127 || match < -127 || mismatch > 127 || mismatch < -127) return DGPU_ERR_UNSUPPORTED;  // the score table is int8
  DGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t) stream : ctx->stream;
  MsaArgs A;
  A.seqs = seqs; A.read_off = read_off; A.read_len = read_len; A.cluster_off = cluster_off; A.nclusters = nclusters;
  A.match = match; A.mismatch = mismatch; A.go = go; A.ge = ge; A.min_clique = min_clique;
  A.cons = cons; A.cons_off = cons_off; A.cons_len = cons_len; A.n_rows = n_rows; A.status = status;
  A.aln = aln; A.aln_off = aln_off; A.aln_cols = aln_cols;
  auto al = [](size_t x) { return (x + 255) & ~(size_t) 255; };
  A.aln_cap = al((size_t) (MSA_MAXR * (MSA_MAXR + 1) / 2) * 384);  // node alignments of one cluster
  A.off_prof = A.aln_cap;
  const size_t b_prof = al((size_t) (MSA_LCAP + 1) * 5);  // gapped consensus + the two column-kind arrays
  A.off_trace = A.off_prof + b_prof;
  const size_t b_trace = al((size_t) (MSA_LCAP + 1) * ((MSA_LCAP + 8) / 8) * 4);
  A.off_lcs = A.off_trace + b_trace;
  A.off_keys = A.off_lcs + al((size_t) 32 * (MSA_LCAP + 2) * sizeof(int) + MSA_TRI);
  A.off_src = A.off_keys + al((size_t) (MSA_LCAP + 1) * 16);
  A.off_bnd = A.off_src + al((size_t) (2 * MSA_LCAP + 4) * 4);
  A.off_tab = A.off_bnd + al((size_t) 2 * (MSA_LCAP + 1) * 8);
  A.off_spans = A.off_tab + al((size_t) MSA_LCAP * MSA_LCAP);
  A.work_stride = A.off_spans + al((size_t) MSA_N * 2 * MSA_MAXR * 2);
  const int per_sm = MSA_PER_SM;  // CTAs of four independent warps
  const size_t smem = sizeof(WarpSm) * MSA_WARPS;
  DGPU_CUDA(ctx, cudaFuncSetAttribute(msa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));   // per device and cheap: set on every call
  size_t grid = std::min<size_t>((nclusters + MSA_WARPS - 1) / MSA_WARPS, (size_t) ctx->num_sms * per_sm);
  void* p;
  int rc = dgpu_reserve(ctx, SLOT_WORK2, grid * MSA_WARPS * A.work_stride, &p);
  if (rc) return rc;
  A.work = (uint8_t*) p;
  dgpu_prof_begin(ctx, st);
  msa_kernel<<<(unsigned) grid, MSA_T, smem, st>>>(A);
  DGPU_LAUNCH_CHECK(ctx, "msa_kernel");
  dgpu_prof_end(ctx, st);
  return DGPU_OK;
}

int dgpu_msa(dgpu_ctx* ctx, const uint8_t* seqs, uint64_t seqs_bytes,
             const uint32_t* read_off, const uint32_t* read_len, uint32_t nreads,
             const uint32_t* cluster_off, uint32_t nclusters,
             int match, int mismatch, int go, int ge, int min_clique,
             uint8_t* cons, const uint64_t* cons_off, uint64_t cons_bytes, uint32_t* cons_len, uint32_t* n_rows,
             uint32_t* status, uint8_t* aln, const uint64_t* aln_off, uint64_t aln_bytes, uint32_t* aln_cols) {
  DgpuCallTrace trace_("dgpu_msa", nclusters);
  if (!ctx) return DGPU_ERR_ARG;
  if (nclusters == 0) return DGPU_OK;
  if (!seqs || !read_off || !read_len || !cluster_off || !cons || !cons_off || !cons_len || !n_rows || !status) return DGPU_ERR_ARG;
  {   // caller buffers: every read inside the arena, every consensus slot (capacity = the cluster's total read length) inside cons_bytes
    for (uint32_t r = 0; r < nreads; ++r)
      if ((uint64_t) read_off[r] + read_len[r] > seqs_bytes) { ctx->last_error = "dgpu_msa: a read lies outside the arena"; return DGPU_ERR_ARG; }
    for (uint32_t i = 0; i < nclusters; ++i) {
      if (cluster_off[i] > cluster_off[i + 1] || cluster_off[i + 1] > nreads) { ctx->last_error = "dgpu_msa: cluster_off is not a partition of the reads"; return DGPU_ERR_ARG; }
      uint64_t cap = 0;
      for (uint32_t r = cluster_off[i]; r < cluster_off[i + 1]; ++r) cap += read_len[r];
      if (cons_off[i] + cap > cons_bytes) { ctx->last_error = "dgpu_msa: consensus slot beyond cons_bytes"; return DGPU_ERR_CAPACITY; }
    }
  }
  DGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  void *d_seqs, *d_roff, *d_rlen, *d_coff, *d_cons, *d_consoff, *d_conslen, *d_nrows, *d_status, *d_aln = nullptr, *d_alnoff = nullptr, *d_alncols = nullptr;
  int rc;
  if ((rc = dgpu_reserve(ctx, SLOT_SEQS, seqs_bytes + 64, &d_seqs))) return rc;
  if ((rc = dgpu_reserve(ctx, SLOT_QOFF, (size_t) nreads * 4, &d_roff))) return rc;
  if ((rc = dgpu_reserve(ctx, SLOT_QLEN, (size_t) nreads * 4, &d_rlen))) return rc;
  if ((rc = dgpu_reserve(ctx, SLOT_TOFF, ((size_t) nclusters + 1) * 4, &d_coff))) return rc;
  if ((rc = dgpu_reserve(ctx, SLOT_A0, cons_bytes + 64, &d_cons))) return rc;
  if ((rc = dgpu_reserve(ctx, SLOT_A1, (size_t) nclusters * 8, &d_consoff))) return rc;
  if ((rc = dgpu_reserve(ctx, SLOT_A2, (size_t) nclusters * 4, &d_conslen))) return rc;
  if ((rc = dgpu_reserve(ctx, SLOT_A3, (size_t) nclusters * 4, &d_nrows))) return rc;
  if ((rc = dgpu_reserve(ctx, SLOT_A4, (size_t) nclusters * 4, &d_status))) return rc;
  if (aln) {
    if (!aln_off || !aln_cols) return DGPU_ERR_ARG;
    if ((rc = dgpu_reserve(ctx, SLOT_A5, aln_bytes + 64, &d_aln))) return rc;
    if ((rc = dgpu_reserve(ctx, SLOT_A6, (size_t) nclusters * 8, &d_alnoff))) return rc;
    if ((rc = dgpu_reserve(ctx, SLOT_A7, (size_t) nclusters * 4, &d_alncols))) return rc;
    DGPU_CUDA(ctx, cudaMemcpyAsync(d_alnoff, aln_off, (size_t) nclusters * 8, cudaMemcpyHostToDevice, st));
  }
  DGPU_CUDA(ctx, cudaMemcpyAsync(d_seqs, seqs, seqs_bytes, cudaMemcpyHostToDevice, st));
  DGPU_CUDA(ctx, cudaMemcpyAsync(d_roff, read_off, (size_t) nreads * 4, cudaMemcpyHostToDevice, st));
  DGPU_CUDA(ctx, cudaMemcpyAsync(d_rlen, read_len, (size_t) nreads * 4, cudaMemcpyHostToDevice, st));
  DGPU_CUDA(ctx, cudaMemcpyAsync(d_coff, cluster_off, ((size_t) nclusters + 1) * 4, cudaMemcpyHostToDevice, st));
  DGPU_CUDA(ctx, cudaMemcpyAsync(d_consoff, cons_off, (size_t) nclusters * 8, cudaMemcpyHostToDevice, st));
  rc = dgpu_msa_dev(ctx, (const uint8_t*) d_seqs, seqs_bytes, (const uint32_t*) d_roff, (const uint32_t*) d_rlen,
                    (const uint32_t*) d_coff, nclusters, match, mismatch, go, ge, min_clique, (uint8_t*) d_cons,
                    (const uint64_t*) d_consoff, (uint32_t*) d_conslen, (uint32_t*) d_nrows, (uint32_t*) d_status,
                    (uint8_t*) d_aln, (const uint64_t*) d_alnoff, (uint32_t*) d_alncols, st);
  if (rc) return rc;
  DGPU_CUDA(ctx, cudaMemcpyAsync(cons, d_cons, cons_bytes, cudaMemcpyDeviceToHost, st));
  DGPU_CUDA(ctx, cudaMemcpyAsync(cons_len, d_conslen, (size_t) nclusters * 4, cudaMemcpyDeviceToHost, st));
  DGPU_CUDA(ctx, cudaMemcpyAsync(n_rows, d_nrows, (size_t) nclusters * 4, cudaMemcpyDeviceToHost, st));
  DGPU_CUDA(ctx, cudaMemcpyAsync(status, d_status, (size_t) nclusters * 4, cudaMemcpyDeviceToHost, st));
  if (aln) {
    DGPU_CUDA(ctx, cudaMemcpyAsync(aln, d_aln, aln_bytes, cudaMemcpyDeviceToHost, st));
    DGPU_CUDA(ctx, cudaMemcpyAsync(aln_cols, d_alncols, (size_t) nclusters * 4, cudaMemcpyDeviceToHost, st));
  }
  DGPU_CUDA(ctx, cudaStreamSynchronize(st));
  return DGPU_OK;
}

}  // extern "C"
