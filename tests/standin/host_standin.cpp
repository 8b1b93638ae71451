// tests/standin/host_standin.cpp — TEST INFRASTRUCTURE ONLY. Never built into, linked into or loaded by the product
// libraries (delly_b200/libdelly_b200*.so); the CPU tests compile it into tests/standin/_build/.
//
// Purpose: pin the HOST logic of the batched mirrors (job construction, batching, flush points, folding of the results,
// everything around the alignments) on a machine without a GPU. This file is the host mirror's own hook file
// (delly_b200/host/capi.cpp, included below) with the alignment entry points of the C ABI — dgpu_edit_distance,
// dgpu_edit_path, dgpu_edit_path_ex, dgpu_long_needle, dgpu_msa — defined HERE as forwarders to the reference's own functions
// compiled verbatim in oracle/_ref/libdelly_ref.so (edlibAlign, longNeedle, msa): the alignments are the reference's,
// everything around them is the code under test. The two clustering-edge entry points are served by the host mirror's own
// pair scan (see below). Nothing here touches a device.
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
  if (end_loc) {   // the first end location as well: one edlibAlign per job
    for (uint64_t i = 0; i < n; ++i) {
      int d = 0, e = 0, s = 0, nl = 0, al = 0;
      ref_edlib((const char*) seqs + q_off[i], (int) q_len[i], (const char*) seqs + t_off[i], (int) t_len[i], k ? k[i] : -1, mode, 0, nullptr, 0, &d, &e, &s, &nl, nullptr, 0, &al);
      dist[i] = d;
      end_loc[i] = (d >= 0 && e != -2) ? e : -1;
    }
    return DGPU_OK;
  }
  std::vector<uint64_t> qo(q_off, q_off + n), to(t_off, t_off + n);
  std::vector<int32_t> kk(n, -1);
  if (k) kk.assign(k, k + n);
  ref_edlib_distance_batch((const char*) seqs, qo.data(), q_len, to.data(), t_len, kk.data(), mode, n, dist, 4);
  return DGPU_OK;
}

int dgpu_edit_path_ex(dgpu_ctx*, const uint8_t* seqs, uint64_t, const uint32_t* q_off, const uint32_t* q_len, const uint32_t* t_off, const uint32_t* t_len, int mode,
                      const uint8_t* eq_pairs, uint32_t n_eq, uint64_t n, int32_t* dist, int32_t* start_loc, int32_t* end_loc, uint8_t* ops, const uint64_t* ops_off,
                      uint64_t, uint32_t* ops_len, uint32_t* status) {
  for (uint64_t i = 0; i < n; ++i) {
    int d = 0, e = 0, s = 0, nl = 0, al = 0;
    int rc = ref_edlib((const char*) seqs + q_off[i], (int) q_len[i], (const char*) seqs + t_off[i], (int) t_len[i], -1, mode, 2, (const char*) eq_pairs, (int) n_eq,
                       &d, &e, &s, &nl, ops + ops_off[i], (int) (q_len[i] + t_len[i]), &al);
    dist[i] = d;
    if (start_loc) start_loc[i] = s;
    if (end_loc) end_loc[i] = e;
    ops_len[i] = (uint32_t) al;
    if (status) status[i] = rc ? 3u : 0u;
  }
  return DGPU_OK;
}

int dgpu_edit_path(dgpu_ctx* ctx, const uint8_t* seqs, uint64_t bytes, const uint32_t* q_off, const uint32_t* q_len, const uint32_t* t_off, const uint32_t* t_len,
                   int mode, uint64_t n, int32_t* dist, int32_t* start_loc, int32_t* end_loc, uint8_t* ops, const uint64_t* ops_off, uint64_t ops_bytes,
                   uint32_t* ops_len, uint32_t* status) {
  return dgpu_edit_path_ex(ctx, seqs, bytes, q_off, q_len, t_off, t_len, mode, nullptr, 0, n, dist, start_loc, end_loc, ops, ops_off, ops_bytes, ops_len, status);
}

int ref_long_needle(const char* s1, int m, const char* s2, int n, char* rows, long cap, int* alilen);
int ref_msa(const char* arena, const uint32_t* off, const uint32_t* len, int nreads, int minClique, int match, int mismatch, int go, int ge, char* cons, int cons_cap,
            int* cons_len, char* aln_out, long aln_cap, int* alnL);

// longNeedle / msa of the reference (src/needle.h, src/msa.h compiled verbatim in oracle/_ref/libdelly_ref.so), in the output layout of the C ABI
int dgpu_long_needle(dgpu_ctx*, const uint8_t* seqs, uint64_t, const uint32_t* c_off, const uint32_t* c_len, const uint32_t* r_off, const uint32_t* r_len, uint64_t n,
                     uint8_t* aln, const uint64_t* aln_off, uint64_t, uint32_t* aln_len, uint8_t* ok, int32_t* info) {
  if (info) return DGPU_ERR_UNSUPPORTED;
  std::vector<char> rows;
  for (uint64_t i = 0; i < n; ++i) {
    const long cap = 2l * ((long) c_len[i] + r_len[i]) + 16;
    rows.assign((size_t) cap, 0);
    int L = 0;
    const int r = ref_long_needle((const char*) seqs + c_off[i], (int) c_len[i], (const char*) seqs + r_off[i], (int) r_len[i], rows.data(), cap, &L);
    if (r < 0) return DGPU_ERR_CAPACITY;
    ok[i] = (uint8_t) (r == 1);
    aln_len[i] = r == 1 ? (uint32_t) L : 0u;
    if (r == 1) {
      memcpy(aln + aln_off[i], rows.data(), (size_t) L);
      memcpy(aln + aln_off[i] + c_len[i] + r_len[i], rows.data() + L, (size_t) L);
    }
  }
  return DGPU_OK;
}

int dgpu_msa(dgpu_ctx*, const uint8_t* seqs, uint64_t, const uint32_t* read_off, const uint32_t* read_len, uint32_t, const uint32_t* cluster_off, uint32_t nclusters,
             int match, int mismatch, int go, int ge, int min_clique, uint8_t* cons, const uint64_t* cons_off, uint64_t, uint32_t* cons_len, uint32_t* n_rows,
             uint32_t* status, uint8_t* aln, const uint64_t*, uint64_t, uint32_t*) {
  if (aln) return DGPU_ERR_UNSUPPORTED;
  for (uint32_t i = 0; i < nclusters; ++i) {
    const uint32_t a = cluster_off[i], b = cluster_off[i + 1];
    int cap = 0;
    for (uint32_t r = a; r < b; ++r) cap += (int) read_len[r];
    std::vector<char> buf((size_t) cap + 1);
    int cl = 0;
    const int rows = ref_msa((const char*) seqs, read_off + a, read_len + a, (int) (b - a), min_clique, match, mismatch, go, ge, buf.data(), cap, &cl, nullptr, 0, nullptr);
    if (rows < 0) return DGPU_ERR_CAPACITY;
    memcpy(cons + cons_off[i], buf.data(), (size_t) cl);
    cons_len[i] = (uint32_t) cl; n_rows[i] = (uint32_t) rows; status[i] = 0;
  }
  return DGPU_OK;
}

// The pair scans of cluster() as the host mirror itself runs them when no device is used (delly_b200/host/cluster.hpp: cluster(), pinned
// against the reference's cluster() in tests/test_host_cluster.py), emitted in the CSR layout of the C ABI.
static int emit_edges(uint64_t n, std::vector<std::vector<std::pair<uint32_t, uint32_t> > > const& adj, uint32_t* edge_off, uint32_t* edge_j, uint32_t* edge_w,
                      uint64_t edge_cap, uint64_t* n_edges) {
  uint64_t total = 0;
  for (uint64_t i = 0; i < n; ++i) { edge_off[i] = (uint32_t) total; total += adj[i].size(); }
  edge_off[n] = (uint32_t) total;
  *n_edges = total;
  if (total > edge_cap) return DGPU_ERR_CAPACITY;
  uint64_t k = 0;
  for (uint64_t i = 0; i < n; ++i)
    for (auto const& e : adj[i]) { edge_j[k] = e.first; edge_w[k] = e.second; ++k; }
  return DGPU_OK;
}

int dgpu_cluster_edges_sr(dgpu_ctx*, const int32_t* chr, const int32_t* pos, const int32_t* chr2, const int32_t* pos2, const int32_t* inslen, uint64_t n, int svt,
                          uint32_t max_read_sep, uint32_t* edge_off, uint32_t* edge_j, uint32_t* edge_w, uint64_t edge_cap, uint64_t* n_edges) {
  using namespace dellyb200;
  Config c; c.maxReadSep = max_read_sep;
  std::vector<std::vector<std::pair<uint32_t, uint32_t> > > adj(n);
  for (uint64_t i = 0; i < n; ++i) {
    const uint32_t varisize = detail::srVariability(c, svt, (double) (pos2[i] - pos[i]), inslen[i]);
    for (uint64_t j = i + 1; j < n && chr[j] == chr[i]; ++j) {
      if ((uint32_t) (pos[j] - pos[i]) > varisize) break;
      if ((svt == 4) && ((uint32_t) std::abs(inslen[j] - inslen[i]) > varisize)) continue;
      if (_translocation(svt) && (chr2[j] != chr2[i])) continue;
      if ((uint32_t) std::abs(pos2[j] - pos2[i]) < varisize)
        adj[i].push_back(std::make_pair((uint32_t) j, (uint32_t) (std::abs(pos2[j] - pos2[i]) + std::abs(pos[j] - pos[i]))));
    }
  }
  return emit_edges(n, adj, edge_off, edge_j, edge_w, edge_cap, n_edges);
}

int dgpu_cluster_edges_pe(dgpu_ctx*, const int32_t* pos, const int32_t* mpos, const int32_t* mtid, const int32_t* alen, const int32_t* median,
                          const int32_t* max_normal_isize, uint64_t n, int svt, uint32_t varisize, uint32_t* edge_off, uint32_t* edge_j, uint32_t* edge_w,
                          uint64_t edge_cap, uint64_t* n_edges) {
  using namespace dellyb200;
  std::vector<std::vector<std::pair<uint32_t, uint32_t> > > adj(n);
  for (uint64_t i = 0; i < n; ++i) {
    const int32_t aMin = _minCoord(pos[i], mpos[i], svt), aMax = _maxCoord(pos[i], mpos[i], svt);
    for (uint64_t j = i + 1; j < n; ++j) {
      const int32_t bMin = _minCoord(pos[j], mpos[j], svt), bMax = _maxCoord(pos[j], mpos[j], svt);
      if (!((uint32_t) std::abs(bMin + alen[j] - aMin) <= varisize)) break;
      if (mtid[i] != mtid[j]) continue;
      if (_pairsDisagree(aMin, aMax, alen[i], max_normal_isize[i], bMin, bMax, alen[j], max_normal_isize[j], svt)) continue;
      adj[i].push_back(std::make_pair((uint32_t) j, (uint32_t) (std::log2((double) std::abs(std::abs((bMin - aMin) - (bMax - aMax)) - std::abs(median[i] - median[j])) + 1))));
    }
  }
  return emit_edges(n, adj, edge_off, edge_j, edge_w, edge_cap, n_edges);
}

// a non-null context token for hooks that refuse a null one (the forwarders above ignore it)
void* standin_ctx(void) { static int token; return &token; }
// the edlib-compatible layer (include/dgpu_edlib.h) creates its own device context; in this build it is given the token instead
void standin_edlib_compat_init(void) { dellyb200::detail::edlibCompatCtxOverride() = (dgpu_ctx*) standin_ctx(); }
}
