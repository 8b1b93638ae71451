// Minimal stand-in for boost::dynamic_bitset<> (TEST INFRASTRUCTURE ONLY): size ctor + operator[].
#ifndef ORACLE_SHIM_DYNAMIC_BITSET_HPP
#define ORACLE_SHIM_DYNAMIC_BITSET_HPP
#include <cstddef>
#include <vector>
namespace boost {
template <typename B = unsigned long>
class dynamic_bitset {
 public:
  dynamic_bitset() {}
  dynamic_bitset(std::size_t n, bool v) : d_(n, v) {}
  explicit dynamic_bitset(std::size_t n) : d_(n, false) {}
  std::vector<bool>::reference operator[](std::size_t i) { return d_[i]; }
  bool operator[](std::size_t i) const { return d_[i]; }
  std::size_t size() const { return d_.size(); }
 private:
  std::vector<bool> d_;
};
}  // namespace boost
#endif
