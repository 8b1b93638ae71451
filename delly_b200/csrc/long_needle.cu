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
// Device design: ONE CTA PER ALIGNMENT, one DP ROW per iteration. A row is spread over the CTA's
// threads (CPT consecutive columns per thread, previous row in registers). The only intra-row
// dependency, H[c] = max(D[c], H[c-1]-1), is a max-plus prefix scan: with E[c] = D[c] + c,
// H[c] = prefmax(E)[c] - c — thread-local running max + one warp/CTA exclusive max-scan. The join of
// needle.h:104-115 is fused into the forward pass (prefix max of the mat row, suffix max of the stored
// rev row, arg-max with the reference's first-max tie rule), so mat is never materialised: only the
// rev matrix (int16, stored column-reversed so the join reads it with aligned vector loads) and 2-bit
// traceback directions go to the per-CTA workspace, which stays L2-resident for short-read shapes.
#include "common.cuh"
#include <algorithm>

namespace {

constexpr int LN_NCLS = 8;
constexpr int NEG = -(1 << 28);

struct LnArgs {
  const uint8_t* seqs;
  const uint8_t* seqs_end;
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
  uint32_t* counts;         // [0..7] class counts, [8..15] starts, [16..23] cursors
  unsigned long long* maxcells;  // [8] per class max (m+1)*rstride ; [8..15] per class max m+n
  uint8_t* work;            // per-CTA workspace slabs
  size_t work_stride;
  size_t off_dirsR, off_dirsF, off_str, off_trace;  // offsets inside a slab (rev values at 0)
};

// class by number of DP columns (n+1): 0 = trivial (m==0 || n==0), 1..7 = kernel shapes, -1 unsupported
__host__ __device__ inline int ln_class(uint32_t m, uint32_t n) {
  if (m == 0 || n == 0) return 0;
  if ((uint64_t) m + n > 32000u) return -1;  // int16 score storage
  uint32_t cols = n + 1;
  if (cols <= 512) return 1;
  if (cols <= 1024) return 2;
  if (cols <= 2048) return 3;
  if (cols <= 4096) return 4;
  if (cols <= 8192) return 5;
  if (cols <= 16384) return 6;
  if (cols <= 32768) return 7;
  return -1;
}

__host__ __device__ inline uint32_t ln_rstride(uint32_t n) { return (n + 1 + 31u) & ~31u; }  // multiple of CPT: a thread's columns never straddle rows

__global__ void ln_count_kernel(LnArgs a, int* unsupported) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  uint32_t m = a.c_len[i], n = a.r_len[i];
  int c = ln_class(m, n);
  if (c < 0) { atomicExch(unsupported, 1); return; }
  atomicAdd(&a.counts[c], 1u);
  if (c == 0) { a.ok[i] = 0; a.aln_len[i] = 0; return; }
  atomicMax(&a.maxcells[c], (unsigned long long) (m + 1) * ln_rstride(n));
  atomicMax(&a.maxcells[8 + c], (unsigned long long) (m + n));
}

__global__ void ln_offsets_kernel(uint32_t* counts) {
  uint32_t s = 0;
  for (int c = 0; c < 8; ++c) { counts[8 + c] = s; counts[16 + c] = s; s += counts[c]; }
}

__global__ void ln_scatter_kernel(LnArgs a) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  int c = ln_class(a.c_len[i], a.r_len[i]);
  if (c < 0) return;
  uint32_t p = atomicAdd(&a.counts[16 + c], 1u);
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

template <int G>
struct BlockScan {
  // Exclusive max-scan of one int per thread across the CTA, plus the CTA-wide max.
  // dirUp=true: prefix (threads 0..t-1); false: suffix (threads t+1..T-1).
  template <bool PREFIX>
  static __device__ __forceinline__ int exclusive(int v, int* sm /* G ints */, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int y = PREFIX ? __shfl_up_sync(0xffffffffu, x, d) : __shfl_down_sync(0xffffffffu, x, d);
      bool valid = PREFIX ? (lane >= d) : (lane + d < 32);
      if (valid) x = max(x, y);
    }
    // x = inclusive scan within warp
    int ex = PREFIX ? __shfl_up_sync(0xffffffffu, x, 1) : __shfl_down_sync(0xffffffffu, x, 1);
    if (PREFIX ? (lane == 0) : (lane == 31)) ex = NEG;
    if (G == 1) {
      total = __shfl_sync(0xffffffffu, x, PREFIX ? 31 : 0);
      return ex;
    }
    if (PREFIX ? (lane == 31) : (lane == 0)) sm[warp] = x;
    __syncthreads();
    int wt = (lane < G) ? sm[lane] : NEG;
    int wx = wt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int y = PREFIX ? __shfl_up_sync(0xffffffffu, wx, d) : __shfl_down_sync(0xffffffffu, wx, d);
      bool valid = PREFIX ? (lane >= d) : (lane + d < 32);
      if (valid) wx = max(wx, y);
    }
    total = __shfl_sync(0xffffffffu, wx, PREFIX ? 31 : 0);
    int src = PREFIX ? warp - 1 : warp + 1;
    int wcarry = __shfl_sync(0xffffffffu, wx, (src < 0 || src > 31) ? 0 : src);
    if (PREFIX ? (warp == 0) : (warp == G - 1)) wcarry = NEG;
    __syncthreads();  // sm reusable
    return max(ex, wcarry);
  }

  static __device__ __forceinline__ unsigned long long reduce_max(unsigned long long k, unsigned long long* sm /* G */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      unsigned long long y = __shfl_xor_sync(0xffffffffu, k, d);
      k = y > k ? y : k;
    }
    if (G == 1) return k;
    if (lane == 0) sm[warp] = k;
    __syncthreads();
    unsigned long long w = (lane < G) ? sm[lane] : 0ull;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      unsigned long long y = __shfl_xor_sync(0xffffffffu, w, d);
      w = y > w ? y : w;
    }
    __syncthreads();
    return w;
  }
};

// One DP pass over all rows. REVPASS: store values (int16, column-reversed) ; else: fused join.
// A = row sequence (length m), B = column sequence (length n).
template <int G, int CPT, bool REVPASS>
__device__ __forceinline__ void ln_pass(const uint8_t* __restrict__ A, const uint8_t* __restrict__ B, const uint32_t m, const uint32_t n,
                                        int16_t* __restrict__ revv, uint32_t* __restrict__ dirs, const uint32_t rstride,
                                        int* sm_i, unsigned long long* sm_k, int* sm_pub,
                                        int& cornerOut, int& gbest, int& consLeft, int& refLeft, int& matv) {
  constexpr int WPT = CPT / 16;  // direction words per thread per row
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const uint32_t c0 = (uint32_t) tid * CPT;
  const uint32_t wpr = rstride / 16;

  // column characters for this thread's columns (c>=1 -> B[c-1])
  uint32_t bw[CPT / 4];  // packed, 4 characters per register
#pragma unroll
  for (int j = 0; j < CPT / 4; ++j) bw[j] = 0;
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    uint32_t c = c0 + j;
    uint32_t ch = (c >= 1 && c <= n) ? (uint32_t) B[c - 1] : 0u;
    bw[j >> 2] |= ch << ((j & 3) * 8);
  }
  int prev[CPT];  // H[r-1][c]
#pragma unroll
  for (int j = 0; j < CPT; ++j) prev[j] = 0;  // row 0 is all zeros

  gbest = NEG; consLeft = 0; refLeft = 0; matv = 0;

  for (uint32_t r = 0; r <= m; ++r) {
    if (r > 0) {
      const uint8_t ach = A[r - 1];
      const int g = (r == m) ? 0 : 1;
      // diagonal input for j == 0: left neighbour's last column of the previous row
      int leftprev = __shfl_up_sync(0xffffffffu, prev[CPT - 1], 1);
      if (G > 1) {
        if (lane == 31) sm_i[32 + (tid >> 5)] = prev[CPT - 1];
        __syncthreads();
        if (lane == 0 && tid > 0) leftprev = sm_i[32 + (tid >> 5) - 1];
      }
      int cur[CPT];
      int run = NEG;
      int diag = leftprev;
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const uint32_t c = c0 + j;
        int D;
        if (c == 0) D = -(int) r;
        else if (c <= n) {
          int sub = ((uint32_t) ach == ((bw[j >> 2] >> ((j & 3) * 8)) & 0xffu)) ? 1 : -1;
          D = max(diag + sub, prev[j] - 1);
        } else D = NEG;
        diag = prev[j];
        int E = D + (g ? (int) c : 0);
        run = max(run, E);
        cur[j] = run;
      }
      int total;
      int carry = BlockScan<G>::template exclusive<true>(run, sm_i, total);
      // pass B: final values + directions
      uint32_t dw[WPT];
#pragma unroll
      for (int w = 0; w < WPT; ++w) dw[w] = 0;
      int hleft = carry - (g ? (int) c0 - 1 : 0);  // H[r][c0-1] (unused for c0 == 0)
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const uint32_t c = c0 + j;
        int Er = max(cur[j], carry);
        int H = Er - (g ? (int) c : 0);
        uint32_t code;
        if (c == 0) code = 1;                       // column 0: always vertical
        else if (H == prev[j] - 1) code = 1;        // vertical first (needle.h:160)
        else if (H == hleft - g) code = 2;          // then horizontal (:163)
        else code = 0;                              // diagonal
        dw[j >> 4] |= code << ((j & 15) * 2);
        hleft = H;
        prev[j] = (c <= n) ? H : 0;
      }
      if (c0 <= n) {
#pragma unroll
        for (int w = 0; w < WPT; ++w) dirs[(size_t) r * wpr + (c0 >> 4) + w] = dw[w];
      }
    }
    if (REVPASS) {
      // store row r column-reversed: rev[r][c] at index n-c  (so the forward pass loads index c)
      // position x = n - c ; this thread's columns map to x in (n-c0-CPT, n-c0]
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const uint32_t c = c0 + j;
        if (c <= n) revv[(size_t) r * rstride + (n - c)] = (int16_t) prev[j];
      }
    } else {
      // ---- fused join for row r (needle.h:88-115) --------------------------------------
      // bestMat: prefix max of this row
      int pm[CPT];
      int run = NEG;
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const uint32_t c = c0 + j;
        if (c <= n) run = max(run, prev[j]);
        pm[j] = run;
      }
      int tot;
      int carryM = BlockScan<G>::template exclusive<true>(run, sm_i, tot);
      // bestRev[m-r][n-c] = max over c' >= c of rev[m-r][n-c'] ; stored reversed => index c'
      int sr[CPT];
      const int16_t* rrow = revv + (size_t) (m - r) * rstride;
      {
        // aligned 16-byte L2 loads (the row was written by this CTA: bypass L1)
        const uint4* rv = (const uint4*) (rrow + c0);
#pragma unroll
        for (int q = 0; q < CPT / 8; ++q) {
          uint4 v = (c0 < rstride) ? __ldcg(rv + q) : make_uint4(0, 0, 0, 0);
          uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) sr[q * 8 + e] = (int) (int16_t) ((w4[e >> 1] >> ((e & 1) * 16)) & 0xffffu);
        }
      }
      int srun = NEG;
#pragma unroll
      for (int j = CPT - 1; j >= 0; --j) {
        const uint32_t c = c0 + j;
        if (c <= n) srun = max(srun, sr[j]);
        sr[j] = srun;
      }
      int carryR = BlockScan<G>::template exclusive<false>(srun, sm_i, tot);
      int bval = NEG, bcol = 0, bbm = 0;
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const uint32_t c = c0 + j;
        if (c <= n) {
          int bm = max(pm[j], carryM);
          int v = bm + max(sr[j], carryR);
          if (v > bval) { bval = v; bcol = (int) c; bbm = bm; }
        }
      }
      unsigned long long key = ((unsigned long long) (uint32_t) (bval + (1 << 29)) << 32) | (uint32_t) (0x7fffffff - bcol);
      unsigned long long wk = BlockScan<G>::reduce_max(key, sm_k);
      int rowbest = (int) (uint32_t) (wk >> 32) - (1 << 29);
      if (rowbest > gbest) {  // strict: earlier rows win ties (row-major first max)
        gbest = rowbest;
        consLeft = (int) r;
        refLeft = 0x7fffffff - (int) (uint32_t) (wk & 0xffffffffu);
        if (key == wk) sm_pub[0] = bbm;  // exactly one thread owns the winning (value, column)
        if (G > 1) __syncthreads(); else __syncwarp();
        matv = sm_pub[0];
        if (G > 1) __syncthreads(); else __syncwarp();
      }
    }
  }
  // corner value H[m][n]
  {
    const uint32_t owner = n / CPT;
    if ((uint32_t) tid == owner) sm_pub[1] = prev[n % CPT];
    if (G > 1) __syncthreads(); else __syncwarp();
    cornerOut = sm_pub[1];
    if (G > 1) __syncthreads(); else __syncwarp();
  }
}

// Sequential traceback over 2-bit directions from (rr,cc) to (0,0); emits alignment characters in
// traceback order. Returns the number of columns.
__device__ __forceinline__ uint32_t ln_traceback(const uint32_t* __restrict__ dirs, uint32_t wpr, const uint8_t* A, const uint8_t* B,
                                                 uint32_t rr, uint32_t cc, uint8_t* tA, uint8_t* tB, bool complement) {
  uint32_t k = 0;
  while (rr > 0 || cc > 0) {
    uint32_t code;
    if (rr == 0) code = 2;
    else {
      uint32_t w = __ldcg(dirs + (size_t) rr * wpr + (cc >> 4));
      code = (w >> ((cc & 15) * 2)) & 3u;
    }
    uint8_t a, b;
    if (code == 1) { --rr; a = A[rr]; b = '-'; }
    else if (code == 2) { --cc; a = '-'; b = B[cc]; }
    else { --rr; --cc; a = A[rr]; b = B[cc]; }
    if (complement) { a = comp_aln(a); b = comp_aln(b); }
    tA[k] = a; tB[k] = b;
    ++k;
  }
  return k;
}

template <int G, int CPT>
__global__ void __launch_bounds__(G * 32) ln_kernel(LnArgs a, int cls) {
  constexpr int T = G * 32;
  __shared__ int sm_i[64];
  __shared__ unsigned long long sm_k[32];
  __shared__ int sm_pub[8];
  const int tid = threadIdx.x;
  const uint32_t cnt = a.counts[cls], start = a.counts[8 + cls];
  uint8_t* slab = a.work + (size_t) blockIdx.x * a.work_stride;
  int16_t* revv = (int16_t*) slab;
  uint32_t* dirsR = (uint32_t*) (slab + a.off_dirsR);
  uint32_t* dirsF = (uint32_t*) (slab + a.off_dirsF);
  uint8_t* strs = slab + a.off_str;
  uint8_t* trace = slab + a.off_trace;

  for (uint32_t idx = blockIdx.x; idx < cnt; idx += gridDim.x) {
    const uint32_t job = a.perm[start + idx];
    const uint32_t m = a.c_len[job], n = a.r_len[job];
    const uint8_t* s1 = a.seqs + a.c_off[job];
    const uint8_t* s2 = a.seqs + a.r_off[job];
    const uint32_t rstride = ln_rstride(n);
    const uint32_t wpr = rstride / 16;
    uint8_t* rc1 = strs;
    uint8_t* rc2 = strs + m;
    for (uint32_t i = tid; i < m; i += T) rc1[i] = revcomp_at(s1, m, i);
    for (uint32_t i = tid; i < n; i += T) rc2[i] = revcomp_at(s2, n, i);
    __syncthreads();

    int revCorner, matCorner, gbest, consLeft, refLeft, matv, d0, d1, d2, d3;
    ln_pass<G, CPT, true>(rc1, rc2, m, n, revv, dirsR, rstride, sm_i, sm_k, sm_pub, revCorner, d0, d1, d2, d3);
    __syncthreads();  // rev values visible to the whole CTA
    ln_pass<G, CPT, false>(s1, s2, m, n, revv, dirsF, rstride, sm_i, sm_k, sm_pub, matCorner, gbest, consLeft, refLeft, matv);
    __syncthreads();

    bool ok = (matCorner == revCorner) && (gbest != matCorner);  // needle.h:83-86, :152
    if (!ok) {
      if (tid == 0) {
        a.ok[job] = 0; a.aln_len[job] = 0;
        if (a.info) { a.info[4 * job] = consLeft; a.info[4 * job + 1] = refLeft; a.info[4 * job + 2] = 0; a.info[4 * job + 3] = gbest; }
      }
      continue;
    }
    // refRight: last right in [0, n-refLeft] with matv + rev[consRight][right] == gbest (needle.h:116-123)
    const uint32_t consRight = m - (uint32_t) consLeft;
    {
      const int16_t* rrow = revv + (size_t) consRight * rstride;
      int best = 0;
      for (uint32_t right = tid; right <= n - (uint32_t) refLeft; right += T)
        if (matv + (int) __ldcg(rrow + (n - right)) == gbest) best = max(best, (int) right);
      unsigned long long k = BlockScan<G>::reduce_max((unsigned long long) (uint32_t) best, sm_k);
      if (tid == 0) sm_pub[2] = (int) (uint32_t) k;
      __syncthreads();
    }
    const uint32_t refRight = (uint32_t) sm_pub[2];
    uint8_t* tFA = trace;
    uint8_t* tFB = trace + (m + n);
    uint8_t* tRA = trace + 2 * (size_t) (m + n);
    uint8_t* tRB = trace + 3 * (size_t) (m + n);
    if (tid == 0) sm_pub[3] = (int) ln_traceback(dirsF, wpr, s1, s2, (uint32_t) consLeft, (uint32_t) refLeft, tFA, tFB, false);
    if (tid == (T > 32 ? 32 : 1)) sm_pub[4] = (int) ln_traceback(dirsR, wpr, rc1, rc2, consRight, refRight, tRA, tRB, true);
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
    __syncthreads();  // slab reuse by the next job
  }
}

struct LnShape { int G, CPT; };

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
  a.seqs = seqs; a.seqs_end = seqs + seqs_bytes;
  a.c_off = c_off; a.c_len = c_len; a.r_off = r_off; a.r_len = r_len; a.n = (uint32_t) n;
  a.aln = aln; a.aln_off = aln_off; a.aln_len = aln_len; a.ok = ok; a.info = info;
  void* p;
  int rc;
  if ((rc = dgpu_reserve(ctx, SLOT_PERM, n * sizeof(uint32_t), &p))) return rc;
  a.perm = (uint32_t*) p;
  if ((rc = dgpu_reserve(ctx, SLOT_COUNTS, 64 * sizeof(uint64_t), &p))) return rc;
  a.counts = (uint32_t*) p;                                   // 24 uint32
  a.maxcells = (unsigned long long*) ((uint8_t*) p + 128);     // 16 uint64
  int* d_unsupported = (int*) ((uint8_t*) p + 128 + 16 * 8);
  DGPU_CUDA(ctx, cudaMemsetAsync(p, 0, 64 * sizeof(uint64_t), st));
  const uint32_t cb = (uint32_t) ((n + 255) / 256);
  ln_count_kernel<<<cb, 256, 0, st>>>(a, d_unsupported);
  DGPU_LAUNCH_CHECK(ctx, "ln_count");
  ln_offsets_kernel<<<1, 1, 0, st>>>(a.counts);
  DGPU_LAUNCH_CHECK(ctx, "ln_offsets");
  ln_scatter_kernel<<<cb, 256, 0, st>>>(a);
  DGPU_LAUNCH_CHECK(ctx, "ln_scatter");
  struct { uint32_t counts[32]; unsigned long long maxc[16]; int unsupported; } h;
  DGPU_CUDA(ctx, cudaMemcpyAsync(&h, p, sizeof(h), cudaMemcpyDeviceToHost, st));
  DGPU_CUDA(ctx, cudaStreamSynchronize(st));
  if (h.unsupported) {
    ctx->last_error = "dgpu_long_needle: a job has |cons|+|ref| > 32000 or |ref| >= 32768 (int16 score storage)";
    return DGPU_ERR_UNSUPPORTED;
  }
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  dgpu_prof_begin(ctx, st);
  for (int c = 1; c < LN_NCLS; ++c) {
    if (!h.counts[c]) continue;
    const size_t cells = (size_t) h.maxc[c];
    const size_t mn = (size_t) h.maxc[8 + c];
    auto al = [](size_t x) { return (x + 255) & ~(size_t) 255; };
    const size_t b_rev = al(cells * 2);
    const size_t b_dirs = al(cells / 4 + 64);
    const size_t b_str = al(mn + 64);
    const size_t b_trace = al(4 * mn + 64);
    a.off_dirsR = b_rev;
    a.off_dirsF = b_rev + b_dirs;
    a.off_str = b_rev + 2 * b_dirs;
    a.off_trace = a.off_str + b_str;
    a.work_stride = a.off_trace + b_trace;
    static const int G_of[8] = {0, 1, 1, 2, 4, 8, 16, 32};
    const int threads = G_of[c] * 32;
    int per_sm = std::max(1, std::min(16, 1536 / threads));
    size_t grid = std::min<size_t>(h.counts[c], (size_t) ctx->num_sms * per_sm);
    // bound the workspace to ~60 % of what is free right now
    size_t budget = (size_t) ((double) free_b * 0.6) + ctx->bufs[SLOT_WORK1].cap;
    if (grid * a.work_stride > budget) grid = std::max<size_t>(1, budget / a.work_stride);
    if ((rc = dgpu_reserve(ctx, SLOT_WORK1, grid * a.work_stride, &p))) return rc;
    a.work = (uint8_t*) p;
    switch (c) {
      case 1: ln_kernel<1, 16><<<(unsigned) grid, 32, 0, st>>>(a, c); break;
      case 2: ln_kernel<1, 32><<<(unsigned) grid, 32, 0, st>>>(a, c); break;
      case 3: ln_kernel<2, 32><<<(unsigned) grid, 64, 0, st>>>(a, c); break;
      case 4: ln_kernel<4, 32><<<(unsigned) grid, 128, 0, st>>>(a, c); break;
      case 5: ln_kernel<8, 32><<<(unsigned) grid, 256, 0, st>>>(a, c); break;
      case 6: ln_kernel<16, 32><<<(unsigned) grid, 512, 0, st>>>(a, c); break;
      case 7: ln_kernel<32, 32><<<(unsigned) grid, 1024, 0, st>>>(a, c); break;
    }
    DGPU_LAUNCH_CHECK(ctx, "ln_kernel");
  }
  dgpu_prof_end(ctx, st);
  return DGPU_OK;
}

int dgpu_long_needle(dgpu_ctx* ctx, const uint8_t* seqs, uint64_t seqs_bytes,
                     const uint32_t* c_off, const uint32_t* c_len, const uint32_t* r_off, const uint32_t* r_len,
                     uint64_t n, uint8_t* aln, const uint64_t* aln_off, uint64_t aln_bytes, uint32_t* aln_len, uint8_t* ok,
                     int32_t* info) {
  if (!ctx) return DGPU_ERR_ARG;
  if (n == 0) return DGPU_OK;
  if (!seqs || !c_off || !c_len || !r_off || !r_len || !aln || !aln_off || !aln_len || !ok) return DGPU_ERR_ARG;
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
