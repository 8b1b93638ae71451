// tests/standin/annotate_standin.cpp — TEST INFRASTRUCTURE ONLY. Never built into, linked with or loaded by the product
// libraries (delly_b200/libdelly_b200*.so): tests/test_svanno.py compiles it into tests/standin/_build/ for the CPU suite.
//
// Purpose: pin the HOST logic of annotateSVBatch (delly_b200/host/svanno.hpp: job construction, template orientation,
// query/target swap, class selection, flank-repeat round) on a machine without a GPU. The one C-ABI entry point that
// logic calls, dgpu_edit_distance, is defined HERE as a forwarder to the reference's own edlib (ref_edlib_distance_batch of
// oracle/_ref/libdelly_ref.so), so the distances are the reference's and everything else is the code under test.
// The device kernels are not exercised by this; `-m gpu` tests do that through the real library.
#include <cstring>
#include <vector>
#include "../../delly_b200/host/svanno.hpp"

extern "C" void ref_edlib_distance_batch(const char* arena, const uint64_t* q_off, const uint32_t* q_len, const uint64_t* t_off, const uint32_t* t_len,
                                         const int32_t* k, int mode, uint64_t n, int32_t* dist, int threads);

extern "C" int dgpu_edit_distance(dgpu_ctx*, const uint8_t* seqs, uint64_t, const uint32_t* q_off, const uint32_t* q_len, const uint32_t* t_off,
                                  const uint32_t* t_len, const int32_t* k, int mode, uint64_t n, int32_t* dist, int32_t* end_loc) {
  if (end_loc) return DGPU_ERR_UNSUPPORTED;
  std::vector<uint64_t> qo(q_off, q_off + n), to(t_off, t_off + n);
  std::vector<int32_t> kk(n, -1);
  if (k) kk.assign(k, k + n);
  ref_edlib_distance_batch((const char*) seqs, qo.data(), q_len, to.data(), t_len, kk.data(), mode, n, dist, 4);
  return DGPU_OK;
}

// same layout as dh_annotate_sv (delly_b200/host/capi.cpp), without a context
extern "C" int standin_annotate_sv(const char* tpl_arena, const uint32_t* tpl_off, const char* seq, int chrLen, const int32_t* sv3, int nsv, const char* alleles,
                                   const uint32_t* al_off, float meiMinFrac, float trMinFrac, int32_t* out5) {
  using namespace dellyb200;
  MeiTemplates mei;
  for (int t = 1; t <= 6; ++t) mei.seq[t].assign(tpl_arena + tpl_off[t - 1], tpl_arena + tpl_off[t]);
  mei.polyA.assign(tpl_arena + tpl_off[6], tpl_arena + tpl_off[7]);
  AnnoConfig c;
  c.meiMinFrac = meiMinFrac;
  c.trMinFrac = trMinFrac;
  std::vector<StructuralVariantRecord> svs(nsv);
  std::vector<int32_t> which(nsv);
  for (int i = 0; i < nsv; ++i) {
    svs[i].svt = sv3[3 * i]; svs[i].svStart = sv3[3 * i + 1]; svs[i].svEnd = sv3[3 * i + 2];
    svs[i].alleles.assign(alleles + al_off[i], alleles + al_off[i + 1]);
    which[i] = i;
  }
  std::vector<const char*> chrseq(1, seq);
  std::vector<uint32_t> tlen(1, (uint32_t) chrLen);
  dgpu_ctx* none = reinterpret_cast<dgpu_ctx*>(&mei);  // any non-null token: the forwarder above ignores it
  int rc = annotateSVBatch(none, c, mei, chrseq, tlen, svs, which);
  if (rc) return rc;
  for (int i = 0; i < nsv; ++i) {
    int32_t* o = out5 + 5 * i;
    o[0] = svs[i].anno.isRC ? 1 : 0; o[1] = svs[i].anno.seqType; o[2] = svs[i].anno.homLen; o[3] = svs[i].anno.trPeriod;
    std::memcpy(&o[4], &svs[i].anno.trCopies, 4);
  }
  return 0;
}
