// Stand-in for the Boost.Filesystem / Boost.Functional-hash names the reference's util.h uses (TEST INFRASTRUCTURE ONLY; used by
// oracle/ref_wrap6.cpp and oracle/ref_main11.cpp, which compile util.h itself instead of masking it). Paths are plain strings; the file-system queries
// answer from the real file system through <sys/stat.h>. boost::hash_combine follows the classic formula
// (seed ^= h(v) + 0x9e3779b9 + (seed << 6) + (seed >> 2)); the ids built from it are opaque to everything the oracle checks.
#pragma once
#include <cstdio>
#include <fstream>
#include <functional>
#include <ostream>
#include <stdexcept>
#include <string>
#include <sys/stat.h>
namespace boost {
namespace filesystem {
struct path {
  std::string s;
  path() {}
  path(std::string const& x) : s(x) {}
  path(const char* x) : s(x) {}
  std::string const& string() const { return s; }
  const char* c_str() const { return s.c_str(); }
  void clear() { s.clear(); }
  bool has_parent_path() const { return s.find('/') != std::string::npos; }
  // file name without directories and without its last extension ("a/b/sr.bam" -> "sr")
  path stem() const { std::size_t k = s.rfind('/'); std::string f = (k == std::string::npos) ? s : s.substr(k + 1); std::size_t d = f.rfind('.'); return (d == std::string::npos || d == 0) ? path(f) : path(f.substr(0, d)); }
  path parent_path() const { std::size_t k = s.rfind('/'); return k == std::string::npos ? path() : path(s.substr(0, k)); }
};
inline std::ostream& operator<<(std::ostream& o, path const& p) { return o << p.s; }
inline path current_path() { return path("."); }
enum perms { owner_write = 0200 };
struct file_status { bool ok; unsigned mode; unsigned permissions() const { return mode; } };
inline bool exists(path const& p) { struct stat st; return ::stat(p.s.c_str(), &st) == 0; }
inline bool exists(file_status const& s) { return s.ok; }
inline file_status status(path const& p) { struct stat st; const bool ok = ::stat(p.s.c_str(), &st) == 0; return file_status{ok, ok ? (unsigned) st.st_mode : 0u}; }
inline bool is_regular_file(path const& p) { struct stat st; return ::stat(p.s.c_str(), &st) == 0 && S_ISREG(st.st_mode); }
inline bool is_regular_file(file_status const& s) { return s.ok; }
inline std::size_t file_size(path const& p) { struct stat st; return ::stat(p.s.c_str(), &st) == 0 ? (std::size_t) st.st_size : 0; }
inline bool remove(path const& p) { return std::remove(p.s.c_str()) == 0; }
struct ofstream : std::ofstream { explicit ofstream(path const& p) : std::ofstream(p.s.c_str()) {} };
struct filesystem_error : std::runtime_error { filesystem_error() : std::runtime_error("filesystem") {} };
}  // namespace filesystem
template <typename T> struct hash { std::size_t operator()(T const& v) const { return std::hash<T>()(v); } };
template <typename T> inline void hash_combine(std::size_t& seed, T const& v) { seed ^= std::hash<T>()(v) + 0x9e3779b9 + (seed << 6) + (seed >> 2); }
template <typename It> inline std::size_t hash_range(It b, It e) { std::size_t seed = 0; for (; b != e; ++b) hash_combine(seed, *b); return seed; }
struct bad_lexical_cast : std::exception {};
// boost::split(container, string, is_any_of(chars))
struct AnyOf { std::string chars; };
inline AnyOf is_any_of(std::string const& c) { return AnyOf{c}; }
template <typename TCont> inline void split(TCont& out, std::string const& in, AnyOf const& sep) {
  out.clear();
  std::string cur;
  for (char ch : in) { if (sep.chars.find(ch) != std::string::npos) { out.push_back(cur); cur.clear(); } else cur += ch; }
  out.push_back(cur);
}
}  // namespace boost
