// Stand-in for the Boost.Iostreams gzip output chain of the reference's dump file (src/coverage.h:337-341, src/genotype.h:128-131):
// a std::ostream that writes gzip through zlib once a compressor and a file sink have been pushed. TEST INFRASTRUCTURE ONLY.
#pragma once
#include <ios>
#include <ostream>
#include <streambuf>
#include <string>
#include <zlib.h>
namespace boost { namespace iostreams {
struct input {};
struct gzip_compressor {};
struct gzip_decompressor {};
struct file_sink { std::string name; file_sink(std::string const& n, std::ios_base::openmode) : name(n) {} };
class gz_outbuf : public std::streambuf {
 public:
  gzFile f = nullptr;
  ~gz_outbuf() { close(); }
  void close() { if (f) { gzclose(f); f = nullptr; } }
 protected:
  int overflow(int c) override { if (c != EOF && f) { char ch = (char) c; gzwrite(f, &ch, 1); } return c; }
  std::streamsize xsputn(const char* s, std::streamsize n) override { if (f) gzwrite(f, s, (unsigned) n); return n; }
};
struct filtering_ostream : std::ostream {
  gz_outbuf buf;
  filtering_ostream() : std::ostream(nullptr) {}
  void push(gzip_compressor const&) {}
  void push(file_sink const& s) { buf.f = gzopen(s.name.c_str(), "wb"); rdbuf(&buf); }
  void pop() { flush(); buf.close(); }
};
template <typename Mode> struct filtering_streambuf : std::streambuf { template <typename T> void push(T const&, int = 0) {} };
}}  // namespace boost::iostreams
