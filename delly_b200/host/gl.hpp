// gl.hpp — genotype likelihoods and the per-sample FORMAT values derived from them (GT, GL, GQ, PL, FT, RCN),
// the host mirror of
//   src/bolog.h:12-20    BoLog: phred -> probability table, 10^(-i/10) for i = 0 .. 10000
//   src/bolog.h:25-85    _computeGLs: log10 likelihoods of 0/0, 0/1, 1/1 from the REF/ALT support qualities
//   src/modvcf.h:667-715 PL from the float GLs, phasing of het calls, copy-number estimate, LowQual/PASS filter
// Everything is double arithmetic with the libm functions the reference calls (std::pow, std::log10), in the
// reference's order of operations, so the values are bit-identical on the same libm. This stays on the host on
// purpose: device log10/pow are not correctly rounded, and a one-ulp difference can flip a rounded PL or GQ.
// boost::math::round / iround are round-half-away-from-zero (std::round) for the finite values that occur here.
#pragma once
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

namespace dellyb200 {

constexpr int DELLY_SMALLEST_GL = -1000;  // src/bolog.h:9

// htslib's genotype encoding (htslib/vcf.h: bcf_gt_phased / bcf_gt_unphased / bcf_gt_missing)
inline int32_t gtUnphased(int32_t allele) { return (allele + 1) << 1; }
inline int32_t gtPhased(int32_t allele) { return ((allele + 1) << 1) | 1; }
constexpr int32_t GT_MISSING = 0;
constexpr int32_t INT32_MISSING = INT32_MIN;  // bcf_int32_missing

struct BoLog {  // src/bolog.h:12-20 with TPrecision = double (src/modvcf.h:373)
  std::vector<double> phred2prob;
  BoLog() {
    for (int i = 0; i <= (int) std::round(-10.0 * DELLY_SMALLEST_GL); ++i) phred2prob.push_back(std::pow(10.0, -((double) i / 10.0)));
  }
};

// src/bolog.h:25-85. gls[3] = {GL(1/1)... as the reference stores them: gls[2] = gl[0], gls[1] = gl[1], gls[0] = gl[2]},
// gts[2], gq. mapqRef / mapqAlt are the per-read qualities of the REF and ALT supporting reads of one sample.
inline void _computeGLs(BoLog const& bl, std::vector<uint8_t> const& mapqRef, std::vector<uint8_t> const& mapqAlt, float* gls, int32_t* gqval,
                        int32_t* gts) {
  double gl[3] = {0, 0, 0};
  const unsigned int peDepth = (unsigned int) (mapqRef.size() + mapqAlt.size());
  for (uint8_t q : mapqRef) {
    const double p = bl.phred2prob[q];
    gl[0] += std::log10(p);
    gl[1] += std::log10(p + (1.0 - p));
    gl[2] += std::log10(1.0 - p);
  }
  for (uint8_t q : mapqAlt) {
    const double p = bl.phred2prob[q];
    gl[0] += std::log10(1.0 - p);
    gl[1] += std::log10((1.0 - p) + p);
    gl[2] += std::log10(p);
  }
  gl[1] += -(double) peDepth * std::log10(2.0);
  unsigned int glBest = 0;
  double glBestVal = gl[0];
  for (unsigned int geno = 1; geno <= 2; ++geno)
    if (gl[geno] >= glBestVal) { glBestVal = gl[geno]; glBest = geno; }
  for (unsigned int geno = 0; geno <= 2; ++geno) {
    gl[geno] -= glBestVal;
    gl[geno] = (gl[geno] > DELLY_SMALLEST_GL) ? gl[geno] : DELLY_SMALLEST_GL;
  }
  uint32_t pl[3];
  for (int g = 0; g < 3; ++g) pl[g] = (uint32_t) std::round(-10 * gl[g]);
  if (peDepth && (pl[0] + pl[1] + pl[2] > 0)) {
    double likelihood = std::log10(1 - 1 / (bl.phred2prob[pl[0]] + bl.phred2prob[pl[1]] + bl.phred2prob[pl[2]]));
    likelihood = (likelihood > DELLY_SMALLEST_GL) ? likelihood : DELLY_SMALLEST_GL;
    *gqval = (int32_t) std::round(-10 * likelihood);
    if (glBest == 0) { gts[0] = gtUnphased(1); gts[1] = gtUnphased(1); }
    else if (glBest == 1) { gts[0] = gtUnphased(0); gts[1] = gtUnphased(1); }
    else { gts[0] = gtUnphased(0); gts[1] = gtUnphased(0); }
  } else {
    gts[0] = GT_MISSING; gts[1] = GT_MISSING;
    *gqval = 0;
  }
  gls[2] = (float) gl[0];
  gls[1] = (float) gl[1];
  gls[0] = (float) gl[2];
}

struct SampleFormat {  // one sample's FORMAT values of one SV record
  int32_t gt[2] = {GT_MISSING, GT_MISSING};
  float gl[3] = {0, 0, 0};
  bool glMissing = true;
  int32_t pl[3] = {INT32_MISSING, INT32_MISSING, INT32_MISSING};
  int32_t gq = 0;
  int32_t rcn = -1;  // cnest
  bool pass = false; // FT: PASS vs LowQual
};

// src/modvcf.h:667-715 for one sample. ps = phase set (-1 unphased), hp1alt / hp2alt = ALT support per haplotype,
// (rcl, rc, rcr) = read counts left / inside / right of the SV.
inline SampleFormat sampleFormat(BoLog const& bl, std::vector<uint8_t> const& ref, std::vector<uint8_t> const& alt, int32_t ps, int32_t hp1alt,
                                 int32_t hp2alt, int32_t rcl, int32_t rc, int32_t rcr) {
  SampleFormat f;
  _computeGLs(bl, ref, alt, f.gl, &f.gq, f.gt);
  if (f.gt[0] == GT_MISSING) {
    f.glMissing = true;
  } else {
    f.glMissing = false;
    for (int k = 0; k < 3; ++k) f.pl[k] = (int32_t) std::max(0.0f, std::round(-10.0f * f.gl[k]));
  }
  if (ps != -1) {
    const bool isHet = (f.gt[0] == gtUnphased(0)) && (f.gt[1] == gtUnphased(1));
    if (isHet && (hp1alt + hp2alt) > 0 && hp1alt != hp2alt) {
      if (hp1alt > hp2alt) { f.gt[0] = gtPhased(1); f.gt[1] = gtPhased(0); }
      else { f.gt[0] = gtPhased(0); f.gt[1] = gtPhased(1); }
    }
  }
  f.rcn = -1;
  if ((rcl + rcr) > 0) {
    double cn = 2.0 * (double) rc / (double) (rcl + rcr);
    if (cn < 0) cn = 0;
    if (cn > 100000) cn = 100000;
    f.rcn = (int32_t) std::lround(cn);  // boost::math::iround
  }
  f.pass = !(f.gq < 15);
  return f;
}

}  // namespace dellyb200
