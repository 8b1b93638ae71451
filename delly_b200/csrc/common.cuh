// common.cuh — shared device/host helpers for libdelly_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/dgpu.h"

#define DGPU_NUM_SMS_B200 148

struct DevBuf {  // grow-only device scratch buffer
  void* p = nullptr;
  size_t cap = 0;
};

struct dgpu_ctx {
  int device = 0;
  int num_sms = DGPU_NUM_SMS_B200;
  cudaStream_t stream = nullptr;
  // pipelined host-pointer calls: upload stream, download stream and their events (created on first use)
  cudaStream_t copy_stream = nullptr, out_stream = nullptr;
  std::vector<cudaEvent_t> pipe_events;
  // fork/join pool for independent launches of one batch call (job classes): [0] fork event, [1..] join events
  std::vector<cudaStream_t> fork_streams;
  std::vector<cudaEvent_t> fork_events;
  std::string last_error;
  uint64_t launches = 0;
  std::vector<DevBuf> bufs;  // indexed by slot id (see SLOT_* below)
  void* comm = nullptr;      // ncclComm_t when multi-GPU gather is initialised
  int rank = 0, world = 1;
  uint32_t async_bound = 0;  // dgpu_set_async_bound: > 0 = the caller guarantees no sequence is longer; device-form calls then never synchronise
  bool no_band = false;      // development switch (DGPU_ED_NO_BAND): long NW jobs straight to the full-matrix kernel
  uint64_t unsupported = 0;  // jobs refused per item because they exceed a device limit (dgpu_unsupported_count)
  // optional CUDA-event timing of the dominant kernels of the last batch call
  bool profiling = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool ev_pending = false;
  float last_kernel_ms = 0.f;
};

// Bracket the dominant kernels of a batch call with events on the launching stream.
inline void dgpu_prof_begin(dgpu_ctx* ctx, cudaStream_t st) {
  if (ctx->profiling) cudaEventRecord(ctx->ev0, st);
}
inline void dgpu_prof_end(dgpu_ctx* ctx, cudaStream_t st) {
  if (ctx->profiling) { cudaEventRecord(ctx->ev1, st); ctx->ev_pending = true; }
}

enum {
  SLOT_SEQS = 0, SLOT_QOFF, SLOT_QLEN, SLOT_TOFF, SLOT_TLEN, SLOT_K, SLOT_DIST, SLOT_ENDLOC,
  SLOT_PERM, SLOT_COUNTS, SLOT_WORK0, SLOT_WORK1, SLOT_WORK2, SLOT_WORK3,
  SLOT_A0, SLOT_A1, SLOT_A2, SLOT_A3, SLOT_A4, SLOT_A5, SLOT_A6, SLOT_A7, SLOT_A8, SLOT_A9,
  SLOT_EQTAB,
  SLOT_EDBAND,  // job lists of the banded NW passes (edit_distance.cu)
  SLOT_COUNT
};

// DGPU_TRACE=1: one stderr line per host-form batch call (name, items, wall ms) — where a pipeline stage's time goes, call by call
struct DgpuCallTrace {
  const char* name; uint64_t n; double t0; bool on;
  static double now();
  DgpuCallTrace(const char* nm, uint64_t items);
  ~DgpuCallTrace();
};

#define DGPU_FORK_STREAMS 4
int dgpu_fork_init(dgpu_ctx* ctx);  // creates the fork/join pool on first use
int dgpu_set_cuda_error(dgpu_ctx* ctx, cudaError_t e, const char* what);
// Returns device pointer of at least `bytes` bytes in `slot` (contents NOT preserved on growth).
int dgpu_reserve(dgpu_ctx* ctx, int slot, size_t bytes, void** out);

#define DGPU_CUDA(ctx, call)                                            \
  do {                                                                  \
    cudaError_t _e = (call);                                            \
    if (_e != cudaSuccess) return dgpu_set_cuda_error((ctx), _e, #call); \
  } while (0)

#define DGPU_LAUNCH_CHECK(ctx, name)                                   \
  do {                                                                 \
    (ctx)->launches++;                                                 \
    cudaError_t _e = cudaGetLastError();                               \
    if (_e != cudaSuccess) return dgpu_set_cuda_error((ctx), _e, name); \
  } while (0)

// ---- unaligned 16-byte chunk reader over a byte arena -------------------------------
// Reads sequence bytes [16*j, 16*j+16) of the sequence starting at `p` regardless of the
// alignment of p, with two aligned LDG.128 and PRMT realignment. Loads are clamped to the
// 16-byte blocks that contain at least one arena byte (arena_end = one past last byte).
struct ChunkReader {
  const uint4* base;   // aligned-down pointer
  uint32_t shift;      // misalignment in bytes (0..15)
  const uint8_t* end;  // arena end
  uint4 cur;           // aligned block j (already loaded)
  uint32_t j;

  __device__ __forceinline__ uint4 load(uint32_t idx) const {
    const uint4* a = base + idx;
    if ((const uint8_t*) a < end) return __ldg(a);
    return make_uint4(0, 0, 0, 0);
  }
  __device__ __forceinline__ void init(const uint8_t* p, const uint8_t* arena_end) {
    uintptr_t u = (uintptr_t) p;
    shift = (uint32_t) (u & 15);
    base = (const uint4*) (u - shift);
    end = arena_end;
    j = 0;
    cur = load(0);
  }
  // returns bytes [16j,16j+16) of the sequence and advances
  __device__ __forceinline__ uint4 next() {
    uint4 nxt = load(j + 1);
    uint4 r;
    if (shift == 0) {
      r = cur;
    } else {
      uint32_t w[8] = {cur.x, cur.y, cur.z, cur.w, nxt.x, nxt.y, nxt.z, nxt.w};
      uint32_t ws = shift >> 2, bs = (shift & 3) * 8;
      uint32_t o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        // dynamic word index resolved with selects (ws in 0..3)
        uint32_t lo = ws == 0 ? w[i] : ws == 1 ? w[i + 1] : ws == 2 ? w[i + 2] : w[i + 3];
        uint32_t hi = ws == 0 ? w[i + 1] : ws == 1 ? w[i + 2] : ws == 2 ? w[i + 3] : w[i + 4];
        o[i] = __funnelshift_r(lo, hi, bs);
      }
      r = make_uint4(o[0], o[1], o[2], o[3]);
    }
    cur = nxt;
    ++j;
    return r;
  }
};

__device__ __forceinline__ uint32_t byte_of(const uint4& v, int b) {
  uint32_t w = (b < 4) ? v.x : (b < 8) ? v.y : (b < 12) ? v.z : v.w;
  return (w >> ((b & 3) * 8)) & 0xffu;
}

// DNA byte -> code: A0 C1 G2 T3 N4, anything else 5 (exact byte equality is preserved by
// routing code 5 through an exact slow path in every kernel).
__device__ __forceinline__ uint32_t dna_code(uint32_t c) {
  uint32_t x = (c >> 1) & 3u;                           // A:0 C:1 T:2 G:3
  uint32_t expect = (0x47544341u >> (x * 8)) & 0xffu;   // 'A','C','T','G'
  uint32_t code = x ^ (x >> 1);                         // ->A0 C1 T3 G2
  return (c == expect) ? code : (c == 'N' ? 4u : 5u);
}
