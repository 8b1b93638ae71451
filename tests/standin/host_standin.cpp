// tests/standin/host_standin.cpp — TEST INFRASTRUCTURE ONLY. Never built into, linked into or loaded by the product
// libraries (delly_b200/libdelly_b200*.so); the CPU tests compile it into tests/standin/_build/.
//
// Purpose: pin the HOST logic of the batched mirrors that call nothing on the device but edit distances and edit paths
// (annotateSVBatch, genotypeLRBatch incl. the methylation windows) on a machine without a GPU. This file is the host
// mirror's own hook file (delly_b200/host/capi.cpp, included below) with TWO C-ABI entry points defined here as forwarders
// to the reference's own edlib (ref_edlib / ref_edlib_distance_batch of oracle/_ref/libdelly_ref.so, i.e. src/edlib.cpp
// compiled verbatim): the distances and paths are the reference's, everything around them is the code under test. Every
// other dgpu_* entry point still resolves to the real CUDA library and fails without a device.
// The device kernels are not exercised by this; the `-m gpu` tests do that through the real libraries.
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../delly_b200/host/capi.cpp"

extern "C" {
int ref_edlib(const char* q, int ql, const char* t, int tl, int k, int mode, int task, const char* eq, int neq, int* dist, int* end0, int* start0, int* numloc,
              unsigned char* aln, int aln_cap, int* aln_len);
void ref_edlib_distance_batch(const char* arena, const uint64_t* q_off, const uint32_t* q_len, const uint64_t* t_off, const uint32_t* t_len, const int32_t* k,
                              int mode, uint64_t n, int32_t* dist, int threads);

int dgpu_edit_distance(dgpu_ctx*, const uint8_t* seqs, uint64_t, const uint32_t* q_off, const uint32_t* q_len, const uint32_t* t_off, const uint32_t* t_len,
                       const int32_t* k, int mode, uint64_t n, int32_t* dist, int32_t* end_loc) {
  if (end_loc) return DGPU_ERR_UNSUPPORTED;
  std::vector<uint64_t> qo(q_off, q_off + n), to(t_off, t_off + n);
  std::vector<int32_t> kk(n, -1);
  if (k) kk.assign(k, k + n);
  ref_edlib_distance_batch((const char*) seqs, qo.data(), q_len, to.data(), t_len, kk.data(), mode, n, dist, 4);
  return DGPU_OK;
}

int dgpu_edit_path(dgpu_ctx*, const uint8_t* seqs, uint64_t, const uint32_t* q_off, const uint32_t* q_len, const uint32_t* t_off, const uint32_t* t_len, int mode,
                   uint64_t n, int32_t* dist, int32_t* start_loc, int32_t* end_loc, uint8_t* ops, const uint64_t* ops_off, uint64_t, uint32_t* ops_len,
                   uint32_t* status) {
  for (uint64_t i = 0; i < n; ++i) {
    int d = 0, e = 0, s = 0, nl = 0, al = 0;
    int rc = ref_edlib((const char*) seqs + q_off[i], (int) q_len[i], (const char*) seqs + t_off[i], (int) t_len[i], -1, mode, 2, nullptr, 0, &d, &e, &s, &nl,
                       ops + ops_off[i], (int) (q_len[i] + t_len[i]), &al);
    dist[i] = d; start_loc[i] = s; end_loc[i] = e; ops_len[i] = (uint32_t) al;
    status[i] = rc ? 3u : 0u;
  }
  return DGPU_OK;
}

// a non-null context token for hooks that refuse a null one (the two forwarders above ignore it)
void* standin_ctx(void) { static int token; return &token; }
}
