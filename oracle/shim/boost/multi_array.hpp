// Minimal stand-in for boost::multi_array<T,2> (TEST INFRASTRUCTURE ONLY).
// Boost is not installed in this image; the reference headers (align.h, gotoh.h,
// needle.h, msa.h, split.h) only use: index, shape(), resize(extents[a][b]),
// operator[][] and construction from extents. Written from the Boost public API docs.
#ifndef ORACLE_SHIM_MULTI_ARRAY_HPP
#define ORACLE_SHIM_MULTI_ARRAY_HPP
#include <cstddef>
#include <vector>
namespace boost {
namespace detail_shim {
struct extent2 { std::size_t a, b; };
struct extent1 {
  std::size_t a;
  extent2 operator[](std::size_t b) const { return extent2{a, b}; }
};
struct extent_gen {
  extent1 operator[](std::size_t a) const { return extent1{a}; }
};
// std::vector<bool> is bit-packed and has no T& operator[]; store bools as char.
template <typename T> struct store { typedef T type; };
template <> struct store<bool> { typedef char type; };
}  // namespace detail_shim
static const detail_shim::extent_gen extents = detail_shim::extent_gen();

template <typename T, std::size_t N>
class multi_array;

template <typename T>
class multi_array<T, 2> {
  typedef typename detail_shim::store<T>::type S;
 public:
  typedef std::ptrdiff_t index;
  typedef std::size_t size_type;
  typedef T element;
  struct row_ref {
    S* p;
    S& operator[](index j) const { return p[j]; }
  };
  struct const_row_ref {
    const S* p;
    const S& operator[](index j) const { return p[j]; }
  };
  multi_array() { shp_[0] = shp_[1] = 0; }
  explicit multi_array(detail_shim::extent2 e) : d_(e.a * e.b) { shp_[0] = e.a; shp_[1] = e.b; }
  void resize(detail_shim::extent2 e) {
    // boost preserves overlapping content on resize; the reference never relies on it
    // except for fresh arrays, but keep the semantics anyway.
    std::vector<S> nd(e.a * e.b);
    std::size_t ra = e.a < shp_[0] ? e.a : shp_[0];
    std::size_t rb = e.b < shp_[1] ? e.b : shp_[1];
    for (std::size_t i = 0; i < ra; ++i)
      for (std::size_t j = 0; j < rb; ++j) nd[i * e.b + j] = d_[i * shp_[1] + j];
    d_.swap(nd);
    shp_[0] = e.a; shp_[1] = e.b;
  }
  const size_type* shape() const { return shp_; }
  row_ref operator[](index i) { return row_ref{d_.data() + i * shp_[1]}; }
  const_row_ref operator[](index i) const { return const_row_ref{d_.data() + i * shp_[1]}; }
 private:
  std::vector<S> d_;
  size_type shp_[2];
};
}  // namespace boost
#endif
