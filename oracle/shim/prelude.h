// Prelude for compiling the reference's hot-path headers VERBATIM without Boost
// (TEST INFRASTRUCTURE ONLY — never linked into the product library).
// Provides: the handful of Boost names the headers use non-dependently, and the util.h
// helpers that split.h / needle.h pick up by prior inclusion in delly.cpp. Each helper
// cites the reference lines it restates.
#ifndef ORACLE_SHIM_PRELUDE_H
#define ORACLE_SHIM_PRELUDE_H
#include <algorithm>
#include <cctype>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include <boost/multi_array.hpp>
#include <boost/dynamic_bitset.hpp>

namespace boost {
// boost::to_upper_copy (std::toupper per char, "C" locale)
inline std::string to_upper_copy(std::string s) {
  for (std::size_t i = 0; i < s.size(); ++i) s[i] = (char) std::toupper((unsigned char) s[i]);
  return s;
}
// boost::char_separator / tokenizer / lexical_cast — only named by _alignmentScore (align.h:231-245)
template <typename C> struct char_separator {
  std::string seps;
  explicit char_separator(const C* s) : seps(s) {}
};
template <typename Sep> class tokenizer {
 public:
  typedef std::vector<std::string>::iterator iterator;
  tokenizer(std::string const& s, Sep const& sep) {
    std::string cur;
    for (char ch : s) {
      if (sep.seps.find(ch) != std::string::npos) { if (!cur.empty()) toks_.push_back(cur); cur.clear(); }
      else cur.push_back(ch);
    }
    if (!cur.empty()) toks_.push_back(cur);
  }
  iterator begin() { return toks_.begin(); }
  iterator end() { return toks_.end(); }
 private:
  std::vector<std::string> toks_;
};
template <typename T, typename S> inline T lexical_cast(S const& s) {
  std::stringstream ss; ss << s; T v; ss >> v; return v;
}
}  // namespace boost

#include <htslib/sam.h>
#include "edlib.h"

#ifndef ORACLE_REAL_UTIL_H   // a wrapper that compiles the reference's util.h itself defines this: no restatements then
namespace torali {
// util.h:549-563
inline void reverseComplement(std::string& sequence) {
  std::string rev = boost::to_upper_copy(std::string(sequence.rbegin(), sequence.rend()));
  std::size_t i = 0;
  for (std::string::iterator revIt = rev.begin(); revIt != rev.end(); ++revIt, ++i) {
    switch (*revIt) {
      case 'A': sequence[i] = 'T'; break;
      case 'C': sequence[i] = 'G'; break;
      case 'G': sequence[i] = 'C'; break;
      case 'T': sequence[i] = 'A'; break;
      case 'N': sequence[i] = 'N'; break;
      default: break;
    }
  }
}
// util.h:86-94
inline uint32_t infixStart(EdlibAlignResult const& cigar) {
  int32_t tIdx = cigar.endLocations[0];
  for (int32_t i = 0; i < cigar.alignmentLength; i++)
    if (cigar.alignment[i] != EDLIB_EDOP_INSERT) tIdx--;
  if (tIdx >= 0) return tIdx + 1;
  else return 0;
}
// util.h:96-99
inline uint32_t infixEnd(EdlibAlignResult const& cigar) { return cigar.endLocations[0]; }
// util.h:250-253
inline std::string _addAlleles(std::string const& ref, std::string const& alt) { return ref + "," + alt; }
}  // namespace torali
#endif  // ORACLE_REAL_UTIL_H

#endif
