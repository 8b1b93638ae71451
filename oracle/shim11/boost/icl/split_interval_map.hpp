// Stand-in for boost::icl::interval_set<uint32_t> with right-open discrete intervals (TEST INFRASTRUCTURE ONLY): insert() joins
// overlapping or touching intervals, iteration is in increasing order — what _parseExcludeIntervals (src/util.h:666-741) and the
// region loops of the reference's drivers use.
#pragma once
#include <cstdint>
#include <map>
#include <vector>
namespace boost { namespace icl {
template <typename T> struct discrete_interval {
  T lo, hi;
  T lower() const { return lo; }
  T upper() const { return hi; }
  static discrete_interval right_open(T a, T b) { return discrete_interval{a, b}; }
};
template <typename T> class interval_set {
 public:
  typedef discrete_interval<T> interval_type;
  typedef typename std::vector<interval_type>::iterator iterator;
  typedef typename std::vector<interval_type>::const_iterator const_iterator;
  void insert(interval_type iv) {
    if (!(iv.lo < iv.hi)) return;
    std::vector<interval_type> out;
    bool placed = false;
    for (auto const& x : v_) {
      if (x.hi < iv.lo) out.push_back(x);
      else if (iv.hi < x.lo) { if (!placed) { out.push_back(iv); placed = true; } out.push_back(x); }
      else { if (x.lo < iv.lo) iv.lo = x.lo; if (iv.hi < x.hi) iv.hi = x.hi; }
    }
    if (!placed) out.push_back(iv);
    v_.swap(out);
  }
  void add(interval_type iv) { insert(iv); }
  iterator begin() { return v_.begin(); }
  iterator end() { return v_.end(); }
  const_iterator begin() const { return v_.begin(); }
  const_iterator end() const { return v_.end(); }
  bool empty() const { return v_.empty(); }
  std::size_t size() const { return v_.size(); }
  void clear() { v_.clear(); }
 private:
  std::vector<interval_type> v_;
};
}}  // namespace boost::icl
