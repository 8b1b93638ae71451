// shard.hpp — cutting a batch of independent jobs into contiguous per-rank ranges, shared by the sharded stage mirrors.
#pragma once
#include <algorithm>
#include <cstdint>
#include <functional>
#include <vector>

#include "types.hpp"

namespace dellyb200 {

// Contiguous ranges of [0, n) with (nearly) equal total cost: bounds[r] .. bounds[r+1] belongs to rank r. Prefix cut at the multiples of
// total / nranks; deterministic, identical on every rank (every rank computes it from the same list).
inline std::vector<std::size_t> partitionByCost(std::vector<uint64_t> const& cost, int nranks) {
  const std::size_t n = cost.size();
  std::vector<std::size_t> bounds((std::size_t) nranks + 1, n);
  bounds[0] = 0;
  uint64_t total = 0;
  for (uint64_t c : cost) total += std::max<uint64_t>(c, 1);
  uint64_t acc = 0;
  std::size_t i = 0;
  for (int r = 1; r < nranks; ++r) {
    const uint64_t target = (uint64_t) (((unsigned __int128) total * (unsigned) r) / (unsigned) nranks);
    while (i < n && acc + std::max<uint64_t>(cost[i], 1) / 2 < target) { acc += std::max<uint64_t>(cost[i], 1); ++i; }
    bounds[r] = i;
  }
  return bounds;
}

// The split-read assembly of one rank (assembleSplitReadsBatch / assembleLRBatch): the queued SVs are cut by cost, the rank builds the
// consensus and aligns it for its range only, `exchange` all-gathers the finished records (mine -> all, in queue order).
struct AssembleShard {
  int rank = 0, nranks = 1;
  std::function<int(std::vector<StructuralVariantRecord> const& mineWork, std::vector<uint8_t> const& mineOk, std::vector<std::size_t> const& bounds,
                    std::vector<StructuralVariantRecord>& allWork, std::vector<uint8_t>& allOk)> exchange;
};

}  // namespace dellyb200
