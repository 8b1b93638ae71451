// ctx.cu — context, error plumbing and scratch-buffer management for libdelly_b200.
#include "common.cuh"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdio>
#include <cstring>

int dgpu_set_cuda_error(dgpu_ctx* ctx, cudaError_t e, const char* what) {
  if (ctx) {
    char buf[512];
    snprintf(buf, sizeof(buf), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    ctx->last_error = buf;
  }
  cudaGetLastError();  // clear sticky-less errors
  return DGPU_ERR_CUDA;
}

int dgpu_reserve(dgpu_ctx* ctx, int slot, size_t bytes, void** out) {
  DevBuf& b = ctx->bufs[slot];
  if (bytes > b.cap) {
    if (b.p) DGPU_CUDA(ctx, cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + (bytes >> 2) + 4096;  // 25 % headroom: batches arrive at similar sizes
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      want = bytes + 256;
      e = cudaMalloc(&b.p, want);
      if (e != cudaSuccess) return dgpu_set_cuda_error(ctx, e, "cudaMalloc(scratch)");
    }
    b.cap = want;
  }
  *out = b.p;
  return DGPU_OK;
}

int dgpu_fork_init(dgpu_ctx* ctx) {
  while (ctx->fork_streams.size() < DGPU_FORK_STREAMS) {
    cudaStream_t s;
    DGPU_CUDA(ctx, cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    ctx->fork_streams.push_back(s);
  }
  while (ctx->fork_events.size() < 1 + DGPU_FORK_STREAMS) {
    cudaEvent_t e;
    DGPU_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->fork_events.push_back(e);
  }
  return DGPU_OK;
}

extern "C" {

int dgpu_version(void) { return 1000; }

const char* dgpu_strerror(int code) {
  switch (code) {
    case DGPU_OK: return "ok";
    case DGPU_ERR_CUDA: return "CUDA runtime error";
    case DGPU_ERR_ARG: return "invalid argument";
    case DGPU_ERR_NODEVICE: return "no usable CUDA device (this library has no CPU fallback)";
    case DGPU_ERR_CAPACITY: return "capacity exceeded";
    case DGPU_ERR_UNSUPPORTED: return "unsupported shape";
    case DGPU_ERR_NCCL: return "NCCL error";
    default: return "unknown error";
  }
}

int dgpu_ctx_create(int device, dgpu_ctx** out) {
  if (!out) return DGPU_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
    cudaGetLastError();
    return DGPU_ERR_NODEVICE;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return DGPU_ERR_NODEVICE;
  if (prop.major != 10) return DGPU_ERR_NODEVICE;  // kernels are compiled for sm_100a only
  if (cudaSetDevice(device) != cudaSuccess) return DGPU_ERR_NODEVICE;
  dgpu_ctx* ctx = new dgpu_ctx();
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  ctx->bufs.resize(SLOT_COUNT);
  ctx->no_band = getenv("DGPU_ED_NO_BAND") != nullptr;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return DGPU_ERR_CUDA;
  }
  *out = ctx;
  return DGPU_OK;
}

void dgpu_ctx_destroy(dgpu_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& b : ctx->bufs)
    if (b.p) cudaFree(b.p);
  if (ctx->ev0) { cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1); }
  for (auto e : ctx->pipe_events) cudaEventDestroy(e);
  for (auto e : ctx->fork_events) cudaEventDestroy(e);
  for (auto f : ctx->fork_streams) cudaStreamDestroy(f);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->out_stream) cudaStreamDestroy(ctx->out_stream);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int dgpu_ctx_sync(dgpu_ctx* ctx) {
  if (!ctx) return DGPU_ERR_ARG;
  DGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return DGPU_OK;
}

const char* dgpu_last_error(dgpu_ctx* ctx) { return ctx ? ctx->last_error.c_str() : ""; }

uint64_t dgpu_launch_count(dgpu_ctx* ctx) { return ctx ? ctx->launches : 0; }

uint64_t dgpu_unsupported_count(dgpu_ctx* ctx) { return ctx ? ctx->unsupported : 0; }

int dgpu_set_async_bound(dgpu_ctx* ctx, uint32_t max_seq_len) {
  if (!ctx) return DGPU_ERR_ARG;
  ctx->async_bound = max_seq_len;
  return DGPU_OK;
}

int dgpu_set_profiling(dgpu_ctx* ctx, int on) {
  if (!ctx) return DGPU_ERR_ARG;
  DGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  if (on && !ctx->ev0) {
    DGPU_CUDA(ctx, cudaEventCreate(&ctx->ev0));
    DGPU_CUDA(ctx, cudaEventCreate(&ctx->ev1));
  }
  ctx->profiling = on != 0;
  return DGPU_OK;
}

float dgpu_last_kernel_ms(dgpu_ctx* ctx) {
  if (!ctx) return -1.f;
  if (ctx->ev_pending) {
    cudaEventSynchronize(ctx->ev1);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) ctx->last_kernel_ms = ms;
    ctx->ev_pending = false;
  }
  return ctx->last_kernel_ms;
}

}  // extern "C"

double DgpuCallTrace::now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
DgpuCallTrace::DgpuCallTrace(const char* nm, uint64_t items) : name(nm), n(items), t0(0), on(false) {
  static const bool enabled = getenv("DGPU_TRACE") != nullptr;
  on = enabled;
  if (on) t0 = now();
}
DgpuCallTrace::~DgpuCallTrace() { if (on) fprintf(stderr, "[dgpu] %-22s %10llu items %9.2f ms\n", name, (unsigned long long) n, now() - t0); }
