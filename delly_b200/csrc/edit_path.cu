// edit_path.cu — batched edlibAlign(..., EDLIB_TASK_PATH): edit distance, start/end location and the
// edit-operation path (0 match, 1 insert, 2 delete, 3 mismatch), for the regime in which the reference uses
// its plain traceback (src/edlib.cpp:1189-1212: estimated alignment data < 1 MiB). Larger problems go through
// Hirschberg's recursion in the reference (src/edlib.cpp:1232-1397) and are reported as status 2 here.
//
// Reference semantics reproduced:
//   * distance and first end location as dgpu_edit_distance (src/edlib.cpp:545-702, :728-929);
//   * HW start location: end - (LAST optimal end of the reversed query in the reversed target prefix, SHW,
//     k = distance) (src/edlib.cpp:226-254); end location -1 -> start 0 and an all-insert path (:241);
//     SHW and NW start at 0;
//   * the path is the global alignment of the query against target[start..end] traced back from the bottom
//     right cell with priority up (insert) > left (delete) > diagonal (src/edlib.cpp:1021-1131); the band the
//     reference restricts itself to never changes that choice because every cell that can be chosen has a
//     value <= the optimum and is therefore exact.
// The traceback matrix is produced by the same anti-diagonal wavefront engine as longNeedle (wavefront.cuh)
// with scoring (match 0, mismatch -1, gap -1), i.e. score = -distance.
#include "common.cuh"
#include "wavefront.cuh"
#include <algorithm>

int dgpu_edit_distance_impl(dgpu_ctx* ctx, const uint8_t* seqs, uint64_t seqs_bytes, const uint32_t* q_off, const uint32_t* q_len,
                            const uint32_t* t_off, const uint32_t* t_len, const int32_t* k, int mode, uint64_t n, int32_t* dist,
                            int32_t* end_loc, void* stream, int last_pos);

namespace {

constexpr int EP_NCLS = 12;

struct EpArgs {
  const uint8_t* seqs;
  const uint32_t* q_off;
  const uint32_t* q_len;
  const uint32_t* t_off;   // already advanced to the start location
  const uint32_t* n_aln;   // aligned target length (end - start + 1)
  const uint32_t* t_len0;  // original target length (0 => edlibAlign returns no path at all)
  const int32_t* dist;
  uint32_t n;
  uint8_t* ops;
  const uint64_t* ops_off;
  uint32_t* ops_len;
  uint32_t* status;
  uint32_t* perm;
  uint32_t* counts;              // [0..15] counts, [16..31] starts, [32..47] cursors
  unsigned long long* maxcells;  // [c] max (m+1)*dstride words ; [16+c] max m
  uint8_t* work;
  size_t work_stride;
  size_t off_trace;
};

__host__ __device__ inline int ep_cols(int cls) { return cls <= 8 ? 8 * cls : 32; }
__host__ __device__ inline uint32_t ep_dstride(uint32_t n, uint32_t C) { return ((n + C - 1) / C + 1) * (C / 8); }

// 0: nothing to trace (empty / no solution), 1..11 kernel classes, 12: Hirschberg regime (unsupported)
__host__ __device__ inline int ep_class(uint32_t m, uint32_t n, int dist) {
  if (dist < 0 || m == 0) return 0;
  if (n == 0) return 0;
  const long long est = 20ll * ((m + 63) / 64) * n + 8ll * n;   // src/edlib.cpp:1189-1191
  if (est >= 1024 * 1024) return 12;
  if (n <= 2048) return (int) ((n + 255) / 256);
  if (n <= 4096) return 9;
  if (n <= 8192) return 10;
  if (n <= 16384) return 11;
  return 12;
}

__global__ void ep_mirror_kernel(const uint8_t* seqs, uint8_t* mirror, const uint32_t* off, const uint32_t* len, uint32_t n) {
  // one warp per sequence: mirror[off+i] = seqs[off+len-1-i]
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n) return;
  const uint32_t o = off[w], L = len[w];
  for (uint32_t i = lane; i < L; i += 32) mirror[o + i] = seqs[o + L - 1 - i];
}

// stage-2 job description for the HW start location: reversed query vs reversed target prefix [0..end]
__global__ void ep_hwstart_jobs_kernel(const uint32_t* t_off, const uint32_t* t_len, const int32_t* dist, const int32_t* end_loc, uint32_t n,
                                       uint32_t* t2_off, uint32_t* t2_len, int32_t* k2) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int e = end_loc[i], d = dist[i];
  if (d < 0 || e < 0) { t2_off[i] = t_off[i]; t2_len[i] = 0; k2[i] = 0; return; }
  t2_off[i] = t_off[i] + (t_len[i] - (uint32_t) e - 1);
  t2_len[i] = (uint32_t) e + 1;
  k2[i] = d;
}

// start locations + aligned target slice
__global__ void ep_slice_kernel(int mode, const uint32_t* t_off, const uint32_t* q_len, const uint32_t* t_len, const int32_t* dist, const int32_t* end_loc,
                                const int32_t* rev_pos, uint32_t n, int32_t* start_loc, uint32_t* a_off, uint32_t* a_len) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int d = dist[i], e = end_loc[i];
  int s = 0;
  if (q_len[i] == 0 || t_len[i] == 0 || d < 0) {  // edlibAlign returns before locations/path exist (src/edlib.cpp:158-177) or found nothing
    start_loc[i] = -1; a_off[i] = t_off[i]; a_len[i] = 0;
    return;
  }
  if (mode == DGPU_MODE_HW && e >= 0) s = e - rev_pos[i];
  start_loc[i] = s;
  a_off[i] = t_off[i] + (uint32_t) s;
  a_len[i] = (uint32_t) (e - s + 1);  // e == -1 -> 0 (all-insert path)
}

__global__ void ep_count_kernel(EpArgs a) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const uint32_t m = a.q_len[i], n = a.n_aln[i];
  const int d = a.dist[i];
  const int c = ep_class(m, n, d);
  if (c == 12) { a.status[i] = 2; a.ops_len[i] = 0; atomicAdd(&a.counts[12], 1u); return; }
  a.status[i] = 0;
  if (c == 0) {
    // no cells: the path is all inserts (target slice empty) or nothing at all
    uint32_t L = (d >= 0 && m > 0 && n == 0 && a.t_len0[i] > 0) ? m : 0;
    uint8_t* o = a.ops + a.ops_off[i];
    for (uint32_t k = 0; k < L; ++k) o[k] = 1;
    a.ops_len[i] = L;
    atomicAdd(&a.counts[0], 1u);
    return;
  }
  atomicAdd(&a.counts[c], 1u);
  atomicMax(&a.maxcells[c], (unsigned long long) (m + 1) * ep_dstride(n, ep_cols(c)));
  atomicMax(&a.maxcells[16 + c], (unsigned long long) m);
  atomicMax(&a.maxcells[32 + c], (unsigned long long) (m + n));
}

__global__ void ep_offsets_kernel(uint32_t* counts) {
  uint32_t s = 0;
  for (int c = 0; c < 16; ++c) { counts[16 + c] = s; counts[32 + c] = s; s += counts[c]; }
}

__global__ void ep_scatter_kernel(EpArgs a) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const int c = ep_class(a.q_len[i], a.n_aln[i], a.dist[i]);
  if (c == 0 || c == 12) return;
  a.perm[atomicAdd(&a.counts[32 + c], 1u)] = i;
}

template <int C, bool MULTI>
__global__ void __launch_bounds__(MULTI ? 512 : 32) ep_kernel(EpArgs a, int cls) {
  extern __shared__ uint8_t sm_rows[];
  __shared__ int sm_x[128];
  __shared__ int sm_pub[4];
  const int tid = threadIdx.x, T = blockDim.x;
  const uint32_t cnt = a.counts[cls], start = a.counts[16 + cls];
  uint8_t* slab = a.work + (size_t) blockIdx.x * a.work_stride;
  uint32_t* dirs = (uint32_t*) slab;
  uint8_t* trace = slab + a.off_trace;
  const wf::Scoring sc = {0, -1, false, false};
  for (uint32_t idx = blockIdx.x; idx < cnt; idx += gridDim.x) {
    const uint32_t job = a.perm[start + idx];
    const uint32_t m = a.q_len[job], n = a.n_aln[job];
    const uint8_t* q = a.seqs + a.q_off[job];
    const uint8_t* t = a.seqs + a.t_off[job];
    const uint32_t dstride = ep_dstride(n, C);
    __syncthreads();
    for (uint32_t i = tid; i < m; i += T) sm_rows[i] = q[i];
    __syncthreads();
    wf::Best dummy;
    int corner;
    wf::pass<C, MULTI, wf::PLAIN>([&](int i) { return sm_rows[i]; }, [&](int i) { return t[i]; }, (int) m, (int) n, sc, 1, dirs, dstride, nullptr, 0, 0, sm_x,
                                  dummy, corner);
    __syncthreads();
    if (tid == 0) {
      // traceback: up (insert, 1) > left (delete, 2) > diagonal (match 0 / mismatch 3); boundaries emit the forced moves
      uint32_t rr = m, cc = n, k = 0;
      while (rr > 0 || cc > 0) {
        uint32_t code;
        if (rr == 0) code = 2;
        else if (cc == 0) code = 1;
        else {
          uint32_t w = __ldcg(dirs + (size_t) rr * dstride + ((cc - 1) >> 3));
          code = (w >> (((cc - 1) & 7) * 4)) & 3u;
        }
        uint8_t op;
        if (code == 1) { --rr; op = 1; }
        else if (code == 2) { --cc; op = 2; }
        else { --rr; --cc; op = (sm_rows[rr] == t[cc]) ? 0 : 3; }
        trace[k++] = op;
      }
      sm_pub[0] = (int) k;
      sm_pub[1] = corner;
    }
    __syncthreads();
    const uint32_t L = (uint32_t) sm_pub[0];
    uint8_t* o = a.ops + a.ops_off[job];
    for (uint32_t i = tid; i < L; i += T) o[i] = trace[L - 1 - i];
    if (tid == 0) {
      a.ops_len[job] = L;
      if (-sm_pub[1] != a.dist[job]) a.status[job] = 3;  // internal consistency: the traced matrix must reproduce the distance
    }
  }
}

template <int C, bool MULTI>
int ep_launch(dgpu_ctx* ctx, EpArgs& a, int cls, unsigned grid, unsigned threads, size_t smem, cudaStream_t st) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(ep_kernel<C, MULTI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return dgpu_set_cuda_error(ctx, e, "cudaFuncSetAttribute(ep_kernel)");
  }
  ep_kernel<C, MULTI><<<grid, threads, smem, st>>>(a, cls);
  DGPU_LAUNCH_CHECK(ctx, "ep_kernel");
  return DGPU_OK;
}

}  // namespace

extern "C" {

int dgpu_edit_path_dev(dgpu_ctx* ctx, const uint8_t* seqs, uint64_t seqs_bytes,
                       const uint32_t* q_off, const uint32_t* q_len, const uint32_t* t_off, const uint32_t* t_len,
                       int mode, uint64_t n, int32_t* dist, int32_t* start_loc, int32_t* end_loc,
                       uint8_t* ops, const uint64_t* ops_off, uint32_t* ops_len, uint32_t* status, void* stream) {
  if (!ctx) return DGPU_ERR_ARG;
  if (n == 0) return DGPU_OK;
  if (n >= (1ull << 31) || seqs_bytes >= (1ull << 32)) return DGPU_ERR_ARG;
  if (mode != DGPU_MODE_NW && mode != DGPU_MODE_SHW && mode != DGPU_MODE_HW) return DGPU_ERR_ARG;
  if (!seqs || !q_off || !q_len || !t_off || !t_len || !dist || !start_loc || !end_loc || !ops || !ops_off || !ops_len || !status) return DGPU_ERR_ARG;
  DGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t) stream : ctx->stream;
  int rc;
  // 1. distance + first end location (k = -1 at every PATH call site of the reference)
  if ((rc = dgpu_edit_distance_impl(ctx, seqs, seqs_bytes, q_off, q_len, t_off, t_len, nullptr, mode, n, dist, end_loc, st, 0))) return rc;
  const uint32_t nb = (uint32_t) ((n + 255) / 256);
  void *d_aoff, *d_alen, *d_rev = nullptr;
  if ((rc = dgpu_reserve(ctx, SLOT_A5, n * 4, &d_aoff))) return rc;
  if ((rc = dgpu_reserve(ctx, SLOT_A6, n * 4, &d_alen))) return rc;
  if (mode == DGPU_MODE_HW) {
    // 2. start location: reversed query vs reversed target prefix, SHW, k = distance, LAST optimal end
    void *d_mirror, *d_t2off, *d_t2len, *d_k2;
    if ((rc = dgpu_reserve(ctx, SLOT_WORK3, seqs_bytes + 64, &d_mirror))) return rc;
    if ((rc = dgpu_reserve(ctx, SLOT_A7, n * 4, &d_t2off))) return rc;
    if ((rc = dgpu_reserve(ctx, SLOT_A8, n * 4, &d_t2len))) return rc;
    if ((rc = dgpu_reserve(ctx, SLOT_A9, n * 8, &d_k2))) return rc;
    d_rev = (uint8_t*) d_k2 + n * 4;
    const uint32_t wb = (uint32_t) ((n * 32 + 255) / 256);
    ep_mirror_kernel<<<wb, 256, 0, st>>>(seqs, (uint8_t*) d_mirror, q_off, q_len, (uint32_t) n);
    DGPU_LAUNCH_CHECK(ctx, "ep_mirror(q)");
    ep_mirror_kernel<<<wb, 256, 0, st>>>(seqs, (uint8_t*) d_mirror, t_off, t_len, (uint32_t) n);
    DGPU_LAUNCH_CHECK(ctx, "ep_mirror(t)");
    ep_hwstart_jobs_kernel<<<nb, 256, 0, st>>>(t_off, t_len, dist, end_loc, (uint32_t) n, (uint32_t*) d_t2off, (uint32_t*) d_t2len, (int32_t*) d_k2);
    DGPU_LAUNCH_CHECK(ctx, "ep_hwstart_jobs");
    // distances of this run are discarded into start_loc (used as scratch), positions land in d_rev
    if ((rc = dgpu_edit_distance_impl(ctx, (const uint8_t*) d_mirror, seqs_bytes, q_off, q_len, (const uint32_t*) d_t2off, (const uint32_t*) d_t2len,
                                      (const int32_t*) d_k2, DGPU_MODE_SHW, n, start_loc, (int32_t*) d_rev, st, 1))) return rc;
  }
  ep_slice_kernel<<<nb, 256, 0, st>>>(mode, t_off, q_len, t_len, dist, end_loc, (const int32_t*) d_rev, (uint32_t) n, start_loc, (uint32_t*) d_aoff,
                                      (uint32_t*) d_alen);
  DGPU_LAUNCH_CHECK(ctx, "ep_slice");
  // 3. traceback matrix + path per job
  EpArgs a;
  a.seqs = seqs; a.q_off = q_off; a.q_len = q_len; a.t_off = (const uint32_t*) d_aoff; a.n_aln = (const uint32_t*) d_alen; a.t_len0 = t_len; a.dist = dist;
  a.n = (uint32_t) n; a.ops = ops; a.ops_off = ops_off; a.ops_len = ops_len; a.status = status;
  void* p;
  if ((rc = dgpu_reserve(ctx, SLOT_PERM, n * sizeof(uint32_t), &p))) return rc;
  a.perm = (uint32_t*) p;
  if ((rc = dgpu_reserve(ctx, SLOT_COUNTS, 1024, &p))) return rc;
  a.counts = (uint32_t*) p;
  a.maxcells = (unsigned long long*) ((uint8_t*) p + 256);
  DGPU_CUDA(ctx, cudaMemsetAsync(p, 0, 1024, st));
  ep_count_kernel<<<nb, 256, 0, st>>>(a);
  DGPU_LAUNCH_CHECK(ctx, "ep_count");
  ep_offsets_kernel<<<1, 1, 0, st>>>(a.counts);
  DGPU_LAUNCH_CHECK(ctx, "ep_offsets");
  ep_scatter_kernel<<<nb, 256, 0, st>>>(a);
  DGPU_LAUNCH_CHECK(ctx, "ep_scatter");
  struct { uint32_t counts[64]; unsigned long long maxc[48]; } h;
  DGPU_CUDA(ctx, cudaMemcpyAsync(&h, p, sizeof(h), cudaMemcpyDeviceToHost, st));
  DGPU_CUDA(ctx, cudaStreamSynchronize(st));
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  dgpu_prof_begin(ctx, st);
  for (int c = 1; c < EP_NCLS; ++c) {
    if (!h.counts[c]) continue;
    auto al = [](size_t x) { return (x + 255) & ~(size_t) 255; };
    const size_t b_dirs = al((size_t) h.maxc[c] * 4 + 1024);
    const size_t mmax = (size_t) h.maxc[16 + c];
    a.off_trace = b_dirs;
    a.work_stride = b_dirs + al((size_t) h.maxc[32 + c] + 64);
    const unsigned threads = c <= 8 ? 32u : (c == 9 ? 128u : (c == 10 ? 256u : 512u));
    const size_t smem = (mmax + 15) & ~(size_t) 15;
    int per_sm = c <= 8 ? 16 : (c == 9 ? 4 : (c == 10 ? 2 : 1));
    size_t grid = std::min<size_t>(h.counts[c], (size_t) ctx->num_sms * per_sm);
    size_t budget = (size_t) ((double) free_b * 0.6) + ctx->bufs[SLOT_WORK1].cap;
    if (grid * a.work_stride > budget) grid = std::max<size_t>(1, budget / a.work_stride);
    if ((rc = dgpu_reserve(ctx, SLOT_WORK1, grid * a.work_stride, &p))) return rc;
    a.work = (uint8_t*) p;
    switch (c) {
      case 1: rc = ep_launch<8, false>(ctx, a, c, (unsigned) grid, threads, smem, st); break;
      case 2: rc = ep_launch<16, false>(ctx, a, c, (unsigned) grid, threads, smem, st); break;
      case 3: rc = ep_launch<24, false>(ctx, a, c, (unsigned) grid, threads, smem, st); break;
      case 4: rc = ep_launch<32, false>(ctx, a, c, (unsigned) grid, threads, smem, st); break;
      case 5: rc = ep_launch<40, false>(ctx, a, c, (unsigned) grid, threads, smem, st); break;
      case 6: rc = ep_launch<48, false>(ctx, a, c, (unsigned) grid, threads, smem, st); break;
      case 7: rc = ep_launch<56, false>(ctx, a, c, (unsigned) grid, threads, smem, st); break;
      case 8: rc = ep_launch<64, false>(ctx, a, c, (unsigned) grid, threads, smem, st); break;
      default: rc = ep_launch<32, true>(ctx, a, c, (unsigned) grid, threads, smem, st); break;
    }
    if (rc) return rc;
  }
  dgpu_prof_end(ctx, st);
  return DGPU_OK;
}

int dgpu_edit_path(dgpu_ctx* ctx, const uint8_t* seqs, uint64_t seqs_bytes,
                   const uint32_t* q_off, const uint32_t* q_len, const uint32_t* t_off, const uint32_t* t_len,
                   int mode, uint64_t n, int32_t* dist, int32_t* start_loc, int32_t* end_loc,
                   uint8_t* ops, const uint64_t* ops_off, uint64_t ops_bytes, uint32_t* ops_len, uint32_t* status) {
  if (!ctx) return DGPU_ERR_ARG;
  if (n == 0) return DGPU_OK;
  if (!seqs || !q_off || !q_len || !t_off || !t_len || !dist || !start_loc || !end_loc || !ops || !ops_off || !ops_len || !status) return DGPU_ERR_ARG;
  DGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  void *d_seqs, *d_qoff, *d_qlen, *d_toff, *d_tlen, *d_dist, *d_start, *d_end, *d_ops, *d_opsoff, *d_opslen, *d_status;
  int rc;
  if ((rc = dgpu_reserve(ctx, SLOT_SEQS, seqs_bytes + 64, &d_seqs))) return rc;
  if ((rc = dgpu_reserve(ctx, SLOT_QOFF, n * 4, &d_qoff))) return rc;
  if ((rc = dgpu_reserve(ctx, SLOT_QLEN, n * 4, &d_qlen))) return rc;
  if ((rc = dgpu_reserve(ctx, SLOT_TOFF, n * 4, &d_toff))) return rc;
  if ((rc = dgpu_reserve(ctx, SLOT_TLEN, n * 4, &d_tlen))) return rc;
  if ((rc = dgpu_reserve(ctx, SLOT_DIST, n * 4, &d_dist))) return rc;
  if ((rc = dgpu_reserve(ctx, SLOT_ENDLOC, n * 4, &d_end))) return rc;
  if ((rc = dgpu_reserve(ctx, SLOT_K, n * 4, &d_start))) return rc;
  if ((rc = dgpu_reserve(ctx, SLOT_A0, ops_bytes + 64, &d_ops))) return rc;
  if ((rc = dgpu_reserve(ctx, SLOT_A1, n * 8, &d_opsoff))) return rc;
  if ((rc = dgpu_reserve(ctx, SLOT_A2, n * 4, &d_opslen))) return rc;
  if ((rc = dgpu_reserve(ctx, SLOT_A3, n * 4, &d_status))) return rc;
  DGPU_CUDA(ctx, cudaMemcpyAsync(d_seqs, seqs, seqs_bytes, cudaMemcpyHostToDevice, st));
  DGPU_CUDA(ctx, cudaMemcpyAsync(d_qoff, q_off, n * 4, cudaMemcpyHostToDevice, st));
  DGPU_CUDA(ctx, cudaMemcpyAsync(d_qlen, q_len, n * 4, cudaMemcpyHostToDevice, st));
  DGPU_CUDA(ctx, cudaMemcpyAsync(d_toff, t_off, n * 4, cudaMemcpyHostToDevice, st));
  DGPU_CUDA(ctx, cudaMemcpyAsync(d_tlen, t_len, n * 4, cudaMemcpyHostToDevice, st));
  DGPU_CUDA(ctx, cudaMemcpyAsync(d_opsoff, ops_off, n * 8, cudaMemcpyHostToDevice, st));
  rc = dgpu_edit_path_dev(ctx, (const uint8_t*) d_seqs, seqs_bytes, (const uint32_t*) d_qoff, (const uint32_t*) d_qlen, (const uint32_t*) d_toff,
                          (const uint32_t*) d_tlen, mode, n, (int32_t*) d_dist, (int32_t*) d_start, (int32_t*) d_end, (uint8_t*) d_ops,
                          (const uint64_t*) d_opsoff, (uint32_t*) d_opslen, (uint32_t*) d_status, st);
  if (rc) return rc;
  DGPU_CUDA(ctx, cudaMemcpyAsync(dist, d_dist, n * 4, cudaMemcpyDeviceToHost, st));
  DGPU_CUDA(ctx, cudaMemcpyAsync(start_loc, d_start, n * 4, cudaMemcpyDeviceToHost, st));
  DGPU_CUDA(ctx, cudaMemcpyAsync(end_loc, d_end, n * 4, cudaMemcpyDeviceToHost, st));
  DGPU_CUDA(ctx, cudaMemcpyAsync(ops, d_ops, ops_bytes, cudaMemcpyDeviceToHost, st));
  DGPU_CUDA(ctx, cudaMemcpyAsync(ops_len, d_opslen, n * 4, cudaMemcpyDeviceToHost, st));
  DGPU_CUDA(ctx, cudaMemcpyAsync(status, d_status, n * 4, cudaMemcpyDeviceToHost, st));
  DGPU_CUDA(ctx, cudaStreamSynchronize(st));
  return DGPU_OK;
}

}  // extern "C"
