// oracle/_ref wrapper, part 8 (TEST INFRASTRUCTURE ONLY): the reference's reference-based SV annotation
// (src/svanno.h:38-238: detectTandemRepeat, annotateSV — breakpoint homology / target-site duplication, mobile-element /
// NUMT / LTR classification of inserted sequence through edlib HW distances, tandem-repeat classification) compiled VERBATIM
// from /root/reference/src together with the reference's own edlib.cpp. util.h is masked by its include guard (svanno.h needs
// reverseComplement and boost::to_upper_copy from it: shim/prelude.h); bam_hdr_t comes from the htslib headers, only
// target_len is read. The mobile-element template sequences (class MEI, src/svanno.h:30-36) stay in the reference: the tests
// fetch them through ref_mei_template, nothing from the reference is copied into this repository.
#define UTIL_H
#define PANGENOME_H
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <iostream>
#include <algorithm>
#include <string>
#include <vector>
#include "shim/prelude.h"

#include <htslib/faidx.h>
#include <htslib/vcf.h>
#include <htslib/sam.h>
#include "edlib.h"
#include "tags.h"
#include "svanno.h"

namespace {
struct RefConfig8 {  // the two fields annotateSV reads from TConfig (src/tegua.h:63-64, defaults :242-243)
  float meiMinFrac = 0.8f;
  float trMinFrac = 0.85f;
};
}  // namespace

extern "C" {

// which: 1 Alu, 2 LINE1, 3 SVA, 4 NUMT, 5 solo LTR, 6 HERVK (the SVAnno::seqType numbering, src/tags.h:84), 7 polyA tail.
// Copies the sequence into buf (if it fits) and returns its length.
int ref_mei_template(int which, char* buf, int cap) {
  using torali::MEI;
  const std::string* s = nullptr;
  switch (which) {
    case 1: s = &MEI::alu; break;
    case 2: s = &MEI::line1; break;
    case 3: s = &MEI::sva; break;
    case 4: s = &MEI::numt; break;
    case 5: s = &MEI::soloLTR; break;
    case 6: s = &MEI::hervk; break;
    case 7: s = &MEI::polyA; break;
    default: return -1;
  }
  if ((int) s->size() <= cap && buf) std::memcpy(buf, s->data(), s->size());
  return (int) s->size();
}

// annotateSV for nsv records of one chromosome (seq, chrLen).
//   sv3: nsv x 3 [svt, svStart, svEnd]; alleles: concatenated "REF,ALT" strings with offsets al_off[nsv+1]
//   out5: nsv x 5 [isRC, seqType, homLen, trPeriod, trCopies as float bits]
void ref_annotate_sv(const char* seq, int chrLen, const int32_t* sv3, int nsv, const char* alleles, const uint32_t* al_off, float meiMinFrac,
                     float trMinFrac, int32_t* out5) {
  RefConfig8 c;
  c.meiMinFrac = meiMinFrac;
  c.trMinFrac = trMinFrac;
  bam_hdr_t hdr;
  std::memset(&hdr, 0, sizeof(hdr));
  uint32_t tl[1] = {(uint32_t) chrLen};
  hdr.n_targets = 1;
  hdr.target_len = tl;
  for (int i = 0; i < nsv; ++i) {
    torali::StructuralVariantRecord sv;
    sv.chr = 0;
    sv.chr2 = 0;
    sv.svt = sv3[3 * i];
    sv.svStart = sv3[3 * i + 1];
    sv.svEnd = sv3[3 * i + 2];
    sv.alleles.assign(alleles + al_off[i], alleles + al_off[i + 1]);
    torali::annotateSV(c, &hdr, seq, sv);
    int32_t* o = out5 + 5 * i;
    o[0] = sv.anno.isRC ? 1 : 0;
    o[1] = sv.anno.seqType;
    o[2] = sv.anno.homLen;
    o[3] = sv.anno.trPeriod;
    std::memcpy(&o[4], &sv.anno.trCopies, 4);
  }
}

// detectTandemRepeat on its own (src/svanno.h:39-53): returns the period, copies through *copies
int ref_detect_tandem_repeat(const char* s, int n, int maxPeriod, float minFraction, float* copies) {
  auto r = torali::detectTandemRepeat(std::string(s, s + n), maxPeriod, minFraction);
  *copies = r.second;
  return r.first;
}

}  // extern "C"
