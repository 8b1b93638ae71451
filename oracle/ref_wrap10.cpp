// oracle/_ref wrapper, part 10 (TEST INFRASTRUCTURE ONLY): the sequence-identity helpers of `delly merge`
// (_minRotation, _seqIdentity, _bestSeqIdentity: src/merge.h:187-243 — SURVEY section 8f row 4) with the reference's own edlib.
// merge.h as a whole needs Boost uuid / icl / iostreams and the htslib BCF reader, so the three functions are taken out of
// it AT BUILD TIME: oracle/Makefile cuts the lines between "_minRotation(" and the comment that opens _svMatch from
// /root/reference/src/merge.h into oracle/_ref/merge_identity.inc (a build output, git-ignored like the libraries; nothing
// of it is committed) and this file includes that extract inside namespace torali.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "edlib.h"

namespace torali {
#include "_ref/merge_identity.inc"
}  // namespace torali

extern "C" {

double ref_seq_identity(const char* a, int la, const char* b, int lb, double minId) {
  return torali::_seqIdentity(std::string(a, a + la), std::string(b, b + lb), minId);
}

double ref_best_seq_identity(const char* a, int la, const char* b, int lb, int posOff, double minId, int seqCutoff) {
  return torali::_bestSeqIdentity(std::string(a, a + la), std::string(b, b + lb), posOff, minId, seqCutoff);
}

int ref_min_rotation(const char* s, int n, char* out) {
  const std::string r = torali::_minRotation(std::string(s, s + n));
  std::memcpy(out, r.data(), r.size());
  return (int) r.size();
}

}  // extern "C"
