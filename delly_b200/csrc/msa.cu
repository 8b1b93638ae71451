// msa.cu — batched per-cluster multiple sequence alignment + consensus, the device replacement for
// msa() (src/msa.h:185-239) as called per SV by assembleSplitReads (src/shortpe.h:185,252).
//
// Reference semantics reproduced bit-for-bit, ONE CTA PER CLUSTER:
//   * distanceMatrix: d[i][j] = lcs(si,sj)*100 / min(|si|,|sj|) (src/msa.h:10-44). LCS length is
//     computed bit-parallel (Allison-Dix/Hyyro: V' = (V + (V & M)) | (V & ~M), LCS = #zeros),
//     one thread per read pair; the value is the same LCS length the reference's DP returns.
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
// DP parallelisation: a DP row is spread over the CTA (CPT columns per thread). The affine horizontal
// recurrence h[c] = max(s[c-1]+o, h[c-1]+e) collapses (o <= e) to an exclusive max-plus prefix scan of
// F[c'] = T[c'] - e*c' with T = max(diag+score, v): h[c] = o + (c-1)e + max_{c'<c} F[c'], so each row is
// elementwise work + ONE block scan. bit1 of column c ("h opened at c") equals "F[c-1] is a strict new
// prefix maximum" and is therefore produced by the thread that owns column c-1 and stored there.
// Trace nibbles live in shared memory when the matrix fits (typical short-read shapes), else in L2.
#include "common.cuh"
#include <algorithm>

namespace {

constexpr int MSA_T = 128;          // threads per CTA
constexpr int MSA_MAXR = 32;        // reads per cluster (reference default maxReadPerSV = 20)
constexpr int MSA_LCAP = 1023;      // max alignment columns (DP has LCAP+1 columns)
constexpr int MSA_N = 2 * MSA_MAXR + 1;
constexpr int MSA_TRACE_SMEM = 4 * 1024;
constexpr int MSA_NEGINF = 1000000; // DnaScore::inf (src/align.h:21)
constexpr int MSA_NEG = -(1 << 28);
constexpr int LCS_W = 8;            // 32-bit words per read in the bit-parallel LCS (<= 256 bp)

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
  uint8_t* work;
  size_t work_stride;
  size_t off_prof, off_trace, off_lcs, off_keys;
  size_t aln_cap;           // bytes of node alignment storage per CTA (at slab offset 0)
};

struct NodeAln {
  const uint8_t* p;  // rows x L, row stride L
  int R, L;
};

// ---- block-wide exclusive prefix max over MSA_T threads ---------------------------------
__device__ __forceinline__ int block_excl_prefmax(int v, int* sm /* >= 8 ints */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x = max(x, y);
  }
  int ex = __shfl_up_sync(0xffffffffu, x, 1);
  if (lane == 0) ex = MSA_NEG;
  if (lane == 31) sm[warp] = x;
  __syncthreads();
  int carry = MSA_NEG;
#pragma unroll
  for (int w = 0; w < MSA_T / 32; ++w)
    if (w < warp) carry = max(carry, sm[w]);
  return max(ex, carry);  // the caller's row-end barrier protects sm against the next row's writes
}

// first/last non-gap column of every row (one warp per row, ballots)
__device__ void row_spans(const NodeAln& a, int* first, int* last) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < a.R; r += MSA_T / 32) {
    const uint8_t* row = a.p + (size_t) r * a.L;
    int f = -1, l = -1;
    for (int j0 = 0; j0 < a.L; j0 += 32) {
      int j = j0 + lane;
      bool ng = (j < a.L) && (row[j] != '-');
      unsigned b = __ballot_sync(0xffffffffu, ng);
      if (b) {
        if (f < 0) f = j0 + __ffs(b) - 1;
        l = j0 + 31 - __clz(b);
      }
    }
    if (lane == 0) { first[r] = f; last[r] = l; }
  }
  __syncthreads();
}

// _createProfile (src/align.h:128-171): p[k][j], k = A,C,G,T,N (the '-' row is never read by _score)
// Also emits key[j] = the six integers the column's profile is made of (counts of A,C,G,T,N and the covering-row count,
// 6 bits each): two columns with equal keys have bit-identical float profiles, which is what the score table exploits.
__device__ void make_profile(const NodeAln& a, const int* first, const int* last, float* p /* 5 x L, stride L */, unsigned long long* key) {
  for (int j = threadIdx.x; j < a.L; j += MSA_T) {
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
#pragma unroll
    for (int k = 0; k < 5; ++k) p[k * a.L + j] = __fdiv_rn((float) cnt[k], (float) sum);
    unsigned long long kk = (unsigned long long) (uint32_t) sum;
#pragma unroll
    for (int k = 0; k < 5; ++k) kk = (kk << 6) | (unsigned long long) (uint32_t) cnt[k];
    key[j] = kk;
  }
  __syncthreads();
}

constexpr int MSA_UMAX = 64;  // distinct column profiles per side that the score table holds

// Map every column to the index of its distinct profile. Returns the number of distinct profiles, or -1 if > MSA_UMAX.
__device__ int dedup_profiles(const unsigned long long* key, int L, uint8_t* ids, unsigned long long* uniq, int* firstOcc /* L ints, scratch */, int* sm_ret) {
  for (int j = threadIdx.x; j < L; j += MSA_T) {
    const unsigned long long k = key[j];
    int f = j;
    for (int i = 0; i < j; ++i) if (key[i] == k) { f = i; break; }
    firstOcc[j] = f;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int cnt = 0;
    for (int j = 0; j < L; ++j) {
      if (firstOcc[j] == j) {
        if (cnt < MSA_UMAX) { ids[j] = (uint8_t) cnt; uniq[cnt] = key[j]; }
        ++cnt;
        if (cnt > MSA_UMAX) break;
      } else ids[j] = ids[firstOcc[j]];
    }
    *sm_ret = (cnt > MSA_UMAX) ? -1 : cnt;
  }
  __syncthreads();
  return *sm_ret;
}

// Exact _score (src/align.h:104-109) for every pair of distinct profiles: S[u1 * MSA_UMAX + u2].
__device__ void build_score_table(const unsigned long long* uniq1, int U1, const unsigned long long* uniq2, int U2, int match, int mismatch, int* S) {
  const float fm = (float) match, fx = (float) mismatch;
  for (int x = threadIdx.x; x < U1 * U2; x += MSA_T) {
    const int u1 = x / U2, u2 = x % U2;
    float p1[5], p2[5];
    unsigned long long k1 = uniq1[u1], k2 = uniq2[u2];
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
    S[u1 * MSA_UMAX + u2] = (int) acc;
  }
  __syncthreads();
}

// The DP of gotoh() for CPT columns per thread. Writes trace nibbles; returns nothing (score unused).
// nibble bits: 1 = bit1 of column c+1, 2 = bit2, 4 = bit3, 8 = bit4 of this cell.
template <int CPT>
__device__ void gotoh_dp(const NodeAln& a1, const NodeAln& a2, const float* __restrict__ p1, const float* __restrict__ p2,
                         const MsaArgs& A, uint32_t* tr, int rowwords, int* sm_scan, int* sm_edge,
                         const uint8_t* id1, const uint8_t* id2, const int* S /* NULL: evaluate the float sum per cell */) {
  const int m = a1.L, n = a2.L;
  const int tid = threadIdx.x, lane = tid & 31;
  const int c0 = tid * CPT;
  const bool charmode = (a1.R == 1 && a2.R == 1);
  const float fm = (float) A.match, fx = (float) A.mismatch;
  const int o = A.go + A.ge, e = A.ge;

  const bool table = (!charmode) && (S != nullptr);
  float q2[CPT][5];
  uint32_t ch2[CPT];  // char mode: the column character; table mode: the column's distinct-profile index
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    int c = c0 + j;
    ch2[j] = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) q2[j][k] = 0.f;
    if (c >= 1 && c <= n) {
      if (charmode) ch2[j] = a2.p[c - 1];
      else if (table) ch2[j] = id2[c - 1];
      else {
#pragma unroll
        for (int k = 0; k < 5; ++k) q2[j][k] = p2[k * n + (c - 1)];
      }
    }
  }
  int sprev[CPT], vprev[CPT];
#pragma unroll
  for (int j = 0; j < CPT; ++j) { sprev[j] = 0; vprev[j] = -MSA_NEGINF; }

  // row 0: bit3 for every column >= 1 (src/gotoh.h:113-117)
  {
    uint32_t bits = 0;
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      int c = c0 + j;
      if (c >= 1 && c <= n) bits |= 4u << (4 * j);
    }
    if (CPT == 8) { if (c0 <= n) tr[c0 >> 3] = bits; }
    else {
      uint32_t hi = __shfl_down_sync(0xffffffffu, bits, 1);
      if (!(lane & 1) && c0 <= n) tr[c0 >> 3] = bits | (hi << 16);
    }
  }

  if (lane == 31) sm_edge[tid >> 5] = 0;  // s[0][*] = 0
  __syncthreads();
  for (int r = 1; r <= m; ++r) {
    const int er = (r == m) ? 0 : e;
    const int orr = (r == m) ? 0 : o;
    float q1[5];
    uint32_t ch1 = 0;
    const int* Srow = nullptr;
    if (charmode) ch1 = a1.p[r - 1];
    else if (table) Srow = S + (int) id1[r - 1] * MSA_UMAX;
    else {
#pragma unroll
      for (int k = 0; k < 5; ++k) q1[k] = p1[k * m + (r - 1)];
    }
    // diagonal input of this thread's first column: s[r-1][c0-1]; warp-boundary values were published at the end of the
    // previous row (same barrier that frees the scan buffer), so a row costs two barriers, not three
    int leftS = __shfl_up_sync(0xffffffffu, sprev[CPT - 1], 1);
    if (lane == 0 && tid > 0) leftS = sm_edge[(tid >> 5) - 1];

    int Tq[CPT], vn[CPT], exF[CPT];
    int run = MSA_NEG;
    int diag = leftS;
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const int c = c0 + j;
      int t, v;
      if (c == 0) { t = 0; v = 0; }  // s[r][0] = v[r][0] = 0 (vertical end gap free at column 0)
      else if (c <= n) {
        int sc;
        if (charmode) sc = (ch1 == ch2[j]) ? A.match : A.mismatch;
        else if (table) sc = Srow[ch2[j]];
        else {
          float acc = 0.f;
#pragma unroll
          for (int k1 = 0; k1 < 5; ++k1) {
            if (q1[k1] != 0.f) {  // CTA-uniform skip of all-zero terms (adding +-0 is exact)
#pragma unroll
              for (int k2 = 0; k2 < 5; ++k2)
                acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(q1[k1], q2[j][k2]), (k1 == k2) ? fm : fx));
            }
          }
          sc = (int) acc;
        }
        const int vgo = (c == n) ? 0 : o, vge = (c == n) ? 0 : e;
        v = max(sprev[j] + vgo, vprev[j] + vge);
        t = max(diag + sc, v);
      } else { t = MSA_NEG; v = MSA_NEG; }
      diag = sprev[j];
      Tq[j] = t; vn[j] = v;
      exF[j] = run;
      int F = (c <= n) ? t - er * c : MSA_NEG;
      run = max(run, F);
    }
    const int carry = block_excl_prefmax(run, sm_scan);
    uint32_t bits = 0;
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const int c = c0 + j;
      if (c > n) continue;
      const int hx = max(exF[j], carry);            // max_{c'<c} F[c']
      const int F = Tq[j] - er * c;
      uint32_t nb = 0;
      int s;
      if (c == 0) { s = 0; nb |= 8u; }               // bit4 on column 0 (src/gotoh.h:118-123)
      else {
        const int h = orr + (c - 1) * er + hx;
        s = max(Tq[j], h);
        // F of this cell for the next column's bit1 uses the FINAL s: F = s - er*c == max(T,h) - er*c.
        if (s == h) nb |= 4u; else if (s == vn[j]) nb |= 8u;
        const int vge = (c == n) ? 0 : e;
        if (vn[j] != vprev[j] + vge) nb |= 2u;
      }
      // bit1 of column c+1: h[c+1] != h[c] + e  <=>  s[c] + o > h[c] + e  <=>  F[c] > hx[c] (o <= e), with
      // F[c] = T'[c] - er*c. (s[c] = max(T,h) but the h branch can never open strictly better than extending.)
      if (F > hx) nb |= 1u;
      bits |= nb << (4 * j);
      sprev[j] = s; vprev[j] = vn[j];
    }
    if (CPT == 8) { if (c0 <= n) tr[(size_t) r * rowwords + (c0 >> 3)] = bits; }
    else {
      uint32_t hi = __shfl_down_sync(0xffffffffu, bits, 1);
      if (!(lane & 1) && c0 <= n) tr[(size_t) r * rowwords + (c0 >> 3)] = bits | (hi << 16);
    }
    if (lane == 31) sm_edge[tid >> 5] = sprev[CPT - 1];
    __syncthreads();  // publishes sm_edge for the next row and frees the scan buffer
  }
  __syncthreads();
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

__global__ void __launch_bounds__(MSA_T, 5) msa_kernel(MsaArgs A) {
  extern __shared__ __align__(16) uint8_t dyn_smem[];
  uint32_t* tr_sm = (uint32_t*) dyn_smem;                         // MSA_TRACE_SMEM bytes
  __shared__ int8_t d[MSA_N * MSA_N];                             // guide-tree similarity matrix (values -2..100)
  __shared__ int Stab[MSA_UMAX * MSA_UMAX];                       // exact _score per pair of distinct column profiles
  __shared__ uint8_t ids1[MSA_LCAP + 1], ids2[MSA_LCAP + 1];
  __shared__ unsigned long long uniq1[MSA_UMAX], uniq2[MSA_UMAX];
  __shared__ int par[MSA_N], lc[MSA_N], rc[MSA_N];
  __shared__ int nodeR[MSA_N], nodeL[MSA_N];
  __shared__ unsigned long long nodeP[MSA_N];
  __shared__ int first1[MSA_MAXR], last1[MSA_MAXR], first2[MSA_MAXR], last2[MSA_MAXR];
  __shared__ int sm_scan[8], sm_edge[8], sm_misc[8];
  __shared__ unsigned long long sm_key[8];
  __shared__ int16_t src1[2 * MSA_LCAP + 4], src2[2 * MSA_LCAP + 4];
  __shared__ uint32_t peq[MSA_MAXR * 5 * LCS_W];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint8_t* slab = A.work + (size_t) blockIdx.x * A.work_stride;
  uint8_t* alnbuf = slab;
  float* prof1 = (float*) (slab + A.off_prof);
  float* prof2 = prof1 + 5 * (MSA_LCAP + 1);
  unsigned long long* keys1 = (unsigned long long*) (slab + A.off_keys);
  unsigned long long* keys2 = keys1 + (MSA_LCAP + 1);
  int* firstOcc = (int*) (keys2 + (MSA_LCAP + 1));
  uint32_t* tr_gl = (uint32_t*) (slab + A.off_trace);
  int* lcsrow = (int*) (slab + A.off_lcs);

  for (uint32_t cl = blockIdx.x; cl < A.nclusters; cl += gridDim.x) {
    __syncthreads();
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
        for (uint32_t j = tid; j < L; j += MSA_T)
          if (dna_code(s[j]) > 4) bad |= 1;
      }
      bad = __syncthreads_or(bad);
      if (bad & 2) status = ST_TOO_LONG; else if (bad & 1) status = ST_BAD_CHAR;
    }
    if (status != ST_OK || num == 0) {
      if (tid == 0) { A.status[cl] = status; A.cons_len[cl] = 0; A.n_rows[cl] = 0; if (A.aln_cols) A.aln_cols[cl] = 0; }
      continue;
    }
    const int N = 2 * num + 1;

    // ---- distance matrix (src/msa.h:32-44, :190-195) ---------------------------------
    for (int x = tid; x < N * N; x += MSA_T) d[x] = -1;
    for (int x = tid; x < N; x += MSA_T) { par[x] = -1; lc[x] = -1; rc[x] = -1; }
    // Peq masks of reads that fit the bit-parallel path
    for (int x = tid; x < num * 5 * LCS_W; x += MSA_T) peq[x] = 0;
    __syncthreads();
    for (int i = warp; i < num; i += MSA_T / 32) {
      const uint32_t L = A.read_len[rbeg + i];
      const uint8_t* s = A.seqs + A.read_off[rbeg + i];
      if (L <= 32 * LCS_W) {
        for (uint32_t j0 = 0; j0 < L; j0 += 32) {
          uint32_t j = j0 + lane;
          uint32_t code = (j < L) ? dna_code(s[j]) : 7u;
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            unsigned b = __ballot_sync(0xffffffffu, code == (uint32_t) k);
            if (lane == 0) peq[(i * 5 + k) * LCS_W + (j0 >> 5)] = b;
          }
        }
      }
    }
    __syncthreads();
    {
      const int npairs = num * (num - 1) / 2;
      for (int pidx = tid; pidx < npairs; pidx += MSA_T) {
        // unrank (i<j) in row-major order
        int i = 0, rem = pidx;
        while (rem >= num - 1 - i) { rem -= num - 1 - i; ++i; }
        const int j = i + 1 + rem;
        const int li = (int) A.read_len[rbeg + i], lj = (int) A.read_len[rbeg + j];
        int l;
        if (li <= 32 * LCS_W) l = lcs_bitpar(&peq[i * 5 * LCS_W], li, A.seqs + A.read_off[rbeg + j], lj);
        else if (lj <= 32 * LCS_W) l = lcs_bitpar(&peq[j * 5 * LCS_W], lj, A.seqs + A.read_off[rbeg + i], li);
        else l = -2;  // long pair: plain DP below
        if (l >= 0) d[i * N + j] = (l * 100) / min(li, lj);
        else d[i * N + j] = -2;
      }
      __syncthreads();
      // plain LCS DP for pairs of two long reads (rare): one warp-serial pass by thread 0
      if (tid == 0) {
        for (int i = 0; i < num; ++i)
          for (int j = i + 1; j < num; ++j)
            if (d[i * N + j] == -2) {
              const int li = (int) A.read_len[rbeg + i], lj = (int) A.read_len[rbeg + j];
              const uint8_t* si = A.seqs + A.read_off[rbeg + i];
              const uint8_t* sj = A.seqs + A.read_off[rbeg + j];
              for (int c = 0; c <= lj; ++c) lcsrow[c] = 0;
              for (int a = 1; a <= li; ++a) {
                int diag = 0;
                for (int c = 1; c <= lj; ++c) {
                  int up = lcsrow[c];
                  lcsrow[c] = (si[a - 1] == sj[c - 1]) ? diag + 1 : max(up, lcsrow[c - 1]);
                  diag = up;
                }
              }
              d[i * N + j] = (lcsrow[lj] * 100) / min(li, lj);
            }
      }
      __syncthreads();
    }

    // ---- UPGMA (src/msa.h:46-89) ---------------------------------------------------------
    int nn = num;
    for (; nn < N; ++nn) {
      unsigned long long key = 0;
      for (int x = tid; x < nn * nn; x += MSA_T) {
        int i = x / nn, j = x % nn;
        if (j > i) {
          int v = d[i * N + j];
          unsigned long long k2 = ((unsigned long long) (uint32_t) (v + 2) << 32) | (uint32_t) (0x7fffffff - (i * MSA_N + j));
          key = k2 > key ? k2 : key;
        }
      }
#pragma unroll
      for (int dd = 16; dd >= 1; dd >>= 1) {
        unsigned long long y = __shfl_xor_sync(0xffffffffu, key, dd);
        key = y > key ? y : key;
      }
      if (lane == 0) sm_key[warp] = key;
      __syncthreads();
      key = sm_key[0];
#pragma unroll
      for (int w = 1; w < MSA_T / 32; ++w) key = sm_key[w] > key ? sm_key[w] : key;
      __syncthreads();
      const int dMax = (int) (uint32_t) (key >> 32) - 2;
      if (key == 0 || dMax == -1) break;
      const int code = 0x7fffffff - (int) (uint32_t) (key & 0xffffffffu);
      const int dI = code / MSA_N, dJ = code % MSA_N;
      if (tid == 0) { par[dI] = nn; par[dJ] = nn; lc[nn] = dI; rc[nn] = dJ; }
      __syncthreads();
      for (int i = tid; i < nn; i += MSA_T)
        if (par[i] == -1)
          d[i * N + nn] = (((dI < i) ? d[dI * N + i] : d[i * N + dI]) + ((dJ < i) ? d[dJ * N + i] : d[i * N + dJ])) / 2;
      __syncthreads();
      for (int i = tid; i < nn + 1; i += MSA_T) {
        if (i < dI) d[i * N + dI] = -1;
        if (i > dI) d[dI * N + i] = -1;
        if (i < dJ) d[i * N + dJ] = -1;
        if (i > dJ) d[dJ * N + i] = -1;
      }
      __syncthreads();
    }
    const int root = (nn > 0) ? nn - 1 : 0;

    // ---- progressive alignment in node-creation order ---------------------------------------
    if (tid < num) {
      nodeR[tid] = 1;
      nodeL[tid] = (int) A.read_len[rbeg + tid];
      nodeP[tid] = (unsigned long long) (A.seqs + A.read_off[rbeg + tid]);
    }
    __syncthreads();
    size_t bump = 0;
    for (int node = num; node <= root && status == ST_OK; ++node) {
      NodeAln a1, a2;
      a1.p = (const uint8_t*) nodeP[lc[node]]; a1.R = nodeR[lc[node]]; a1.L = nodeL[lc[node]];
      a2.p = (const uint8_t*) nodeP[rc[node]]; a2.R = nodeR[rc[node]]; a2.L = nodeL[rc[node]];
      const int m = a1.L, n = a2.L;
      const bool charmode = (a1.R == 1 && a2.R == 1);
      if (!charmode) {
        row_spans(a1, first1, last1);
        make_profile(a1, first1, last1, prof1, keys1);
        row_spans(a2, first2, last2);
        make_profile(a2, first2, last2, prof2, keys2);
      }
      const int* Sptr = nullptr;
      if (!charmode) {
        const int U1 = dedup_profiles(keys1, m, ids1, uniq1, firstOcc, &sm_misc[1]);
        const int U2 = (U1 > 0) ? dedup_profiles(keys2, n, ids2, uniq2, firstOcc, &sm_misc[2]) : -1;
        if (U1 > 0 && U2 > 0) { build_score_table(uniq1, U1, uniq2, U2, A.match, A.mismatch, Stab); Sptr = Stab; }
      }
      const int rowwords = (n + 1 + 7) >> 3;
      const size_t trbytes = (size_t) (m + 1) * rowwords * 4;
      uint32_t* tr = (trbytes <= MSA_TRACE_SMEM) ? tr_sm : tr_gl;
      if (n + 1 <= MSA_T * 4) gotoh_dp<4>(a1, a2, prof1, prof2, A, tr, rowwords, sm_scan, sm_edge, ids1, ids2, Sptr);
      else gotoh_dp<8>(a1, a2, prof1, prof2, A, tr, rowwords, sm_scan, sm_edge, ids1, ids2, Sptr);

      // traceback (src/gotoh.h:141-167): one thread, state machine over the trace nibbles
      if (tid == 0) {
        int row = m, col = n, k = 0;
        int st = 0;  // 0 = 's', 1 = 'h', 2 = 'v'
        while (row > 0 || col > 0) {
          const uint32_t w = tr[(size_t) row * rowwords + (col >> 3)];
          const uint32_t nb = (w >> ((col & 7) * 4)) & 0xfu;
          if (st == 0) {
            if (nb & 4u) st = 1;
            else if (nb & 8u) st = 2;
            else { --row; --col; src1[k] = (int16_t) row; src2[k] = (int16_t) col; ++k; }
          } else if (st == 1) {
            // bit1 of (row,col) is stored in the nibble of column col-1
            const uint32_t w2 = tr[(size_t) row * rowwords + ((col - 1) >> 3)];
            if ((w2 >> (((col - 1) & 7) * 4)) & 1u) st = 0;
            --col; src1[k] = -1; src2[k] = (int16_t) col; ++k;
          } else {
            if (nb & 2u) st = 0;
            --row; src1[k] = (int16_t) row; src2[k] = -1; ++k;
          }
          if (k > 2 * MSA_LCAP) break;
        }
        sm_misc[0] = k;
      }
      __syncthreads();
      const int L = sm_misc[0];
      const int R = a1.R + a2.R;
      if (L > MSA_LCAP || bump + (size_t) R * L > A.aln_cap) { status = ST_TOO_LONG; break; }
      uint8_t* out = alnbuf + bump;
      bump += ((size_t) R * L + 15) & ~(size_t) 15;
      // _createAlignment (src/align.h:202-229): rows of a1, then rows of a2
      for (int x = tid; x < R * L; x += MSA_T) {
        const int i = x / L, ai = x % L;
        const int k = L - 1 - ai;
        uint8_t ch;
        if (i < a1.R) { const int s = src1[k]; ch = (s >= 0) ? a1.p[(size_t) i * a1.L + s] : (uint8_t) '-'; }
        else { const int s = src2[k]; ch = (s >= 0) ? a2.p[(size_t) (i - a1.R) * a2.L + s] : (uint8_t) '-'; }
        out[x] = ch;
      }
      if (tid == 0) { nodeR[node] = R; nodeL[node] = L; nodeP[node] = (unsigned long long) out; }
      __syncthreads();
    }
    if (status != ST_OK) {
      if (tid == 0) { A.status[cl] = status; A.cons_len[cl] = 0; A.n_rows[cl] = 0; if (A.aln_cols) A.aln_cols[cl] = 0; }
      continue;
    }

    // ---- consensus (src/msa.h:111-173) -------------------------------------------------------
    NodeAln ra;
    ra.p = (const uint8_t*) nodeP[root]; ra.R = nodeR[root]; ra.L = nodeL[root];
    row_spans(ra, first1, last1);
    uint8_t* gapped = (uint8_t*) prof1;  // reuse: L bytes
    const int thr = max(2, min(A.min_clique, ra.R));
    for (int j = tid; j < ra.L; j += MSA_T) {
      int cov = 0;
      int cnt[5] = {0, 0, 0, 0, 0};
      for (int r = 0; r < ra.R; ++r) {
        if (first1[r] >= 0 && first1[r] <= j && j <= last1[r]) {
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
    __syncthreads();
    if (tid == 0) {
      uint8_t* o = A.cons + A.cons_off[cl];
      uint32_t k = 0;
      for (int j = 0; j < ra.L; ++j) if (gapped[j] != '-') o[k++] = gapped[j];
      A.cons_len[cl] = k;
      A.n_rows[cl] = (uint32_t) ra.R;
      A.status[cl] = ST_OK;
      if (A.aln_cols) A.aln_cols[cl] = (uint32_t) ra.L;
    }
    if (A.aln) {
      uint8_t* o = A.aln + A.aln_off[cl];
      for (int x = tid; x < ra.R * ra.L; x += MSA_T) o[x] = ra.p[x];
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
  if (go > 0 || ge > 0 || go + ge > ge) return DGPU_ERR_UNSUPPORTED;  // the scan form needs open <= extend <= 0
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
  const size_t b_prof = al(2 * 5 * (size_t) (MSA_LCAP + 1) * sizeof(float));
  A.off_trace = A.off_prof + b_prof;
  const size_t b_trace = al((size_t) (MSA_LCAP + 1) * ((MSA_LCAP + 8) / 8) * 4);
  A.off_lcs = A.off_trace + b_trace;
  A.off_keys = A.off_lcs + al((MSA_LCAP + 2) * sizeof(int));
  A.work_stride = A.off_keys + al((size_t) (MSA_LCAP + 1) * (8 + 8 + 4));
  static bool attr_set = false;
  if (!attr_set) {
    DGPU_CUDA(ctx, cudaFuncSetAttribute(msa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MSA_TRACE_SMEM));
    attr_set = true;
  }
  int per_sm = 5;
  size_t grid = std::min<size_t>(nclusters, (size_t) ctx->num_sms * per_sm);
  void* p;
  int rc = dgpu_reserve(ctx, SLOT_WORK2, grid * A.work_stride, &p);
  if (rc) return rc;
  A.work = (uint8_t*) p;
  dgpu_prof_begin(ctx, st);
  msa_kernel<<<(unsigned) grid, MSA_T, MSA_TRACE_SMEM, st>>>(A);
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
  if (!ctx) return DGPU_ERR_ARG;
  if (nclusters == 0) return DGPU_OK;
  if (!seqs || !read_off || !read_len || !cluster_off || !cons || !cons_off || !cons_len || !n_rows || !status) return DGPU_ERR_ARG;
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
