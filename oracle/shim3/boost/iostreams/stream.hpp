// Stand-in for the Boost.Iostreams names src/coverage.h mentions in annotateCoverage (the gzip dump file, :337-341).
// TEST INFRASTRUCTURE ONLY: annotateCoverage is never instantiated by the oracle; these declarations only let the
// header parse. A filtering_ostream that is never pushed to discards what is written to it.
#pragma once
#include <ios>
#include <ostream>
#include <string>
namespace boost { namespace iostreams {
struct gzip_compressor {};
struct file_sink { file_sink(std::string const&, std::ios_base::openmode) {} };
struct filtering_ostream : std::ostream {
  filtering_ostream() : std::ostream(nullptr) {}
  template <typename T> void push(T const&) {}
  void pop() {}
};
}}  // namespace boost::iostreams
