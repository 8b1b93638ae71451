// Stand-in for boost::math::round / iround as src/bolog.h uses them: round half away from zero on finite values,
// which is std::round. TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cmath>
namespace boost { namespace math {
template <typename T> inline double round(T v) { return std::round((double) v); }
template <typename T> inline int iround(T v) { return (int) std::lround((double) v); }
}}  // namespace boost::math
