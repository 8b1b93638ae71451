// Stand-in for boost::math::normal / pdf, named by _computeCNLs in src/bolog.h (delly cnv: out of scope, never called
// by the oracle). TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cmath>
namespace boost { namespace math {
struct normal { double m, s; normal(double mean, double sd) : m(mean), s(sd) {} };
inline double pdf(normal const& n, double x) { const double z = (x - n.m) / n.s; return std::exp(-0.5 * z * z) / (n.s * std::sqrt(2 * M_PI)); }
}}  // namespace boost::math
