// boost::unordered_map served by std::unordered_map (TEST INFRASTRUCTURE ONLY). Pair keys (src/delly.h:134-136, src/tegua.h:112-115)
// hash like boost::hash<std::pair<>>: hash_combine of the two members. NOTE: iteration order is libstdc++'s, not Boost's — the
// product side uses the same container, so whole-file parity is pinned modulo Boost's bucket order.
#pragma once
#include <unordered_map>
#include <utility>
#include <boost/filesystem.hpp>
namespace boost {
template <typename K> struct umap_hash { std::size_t operator()(K const& k) const { return std::hash<K>()(k); } };
template <typename A, typename B> struct umap_hash<std::pair<A, B> > {
  std::size_t operator()(std::pair<A, B> const& p) const { std::size_t seed = 0; hash_combine(seed, p.first); hash_combine(seed, p.second); return seed; }
};
template <typename K, typename V> using unordered_map = std::unordered_map<K, V, umap_hash<K> >;
}
