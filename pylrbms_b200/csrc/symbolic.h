// Host-only symbolic phase of the mu-batched reduced solve: 8x8-tile sparse Cholesky schedule.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

struct lrbms_symbolic {
  int32_t n_sub = 0;
  std::vector<int32_t> sizes, offsets;   // N_i, prefix sums (n_sub + 1)
  int32_t n_red = 0;                     // sum N_i
  int32_t n_pad = 0;                     // n_red rounded up to a multiple of 8
  int32_t ntc = 0;                       // tile columns
  // L tile pattern, compressed by tile column; rows ascending, the diagonal tile first
  std::vector<int32_t> col_ptr, row_idx;
  // update pairs per target: targets are the L tiles (slot order) followed by one rhs target per tile column.
  // pair (a, b): target -= tile[a] * tile[b]^T with a, b slots (slots >= n_tiles address the forward-solve row y_K)
  std::vector<int32_t> pair_ptr, pair_a, pair_b;
  // a_map[slot] = index of the tile in the assembled-operator tile list, or -1 for pure fill
  std::vector<int32_t> a_map;
  int32_t n_a_tiles = 0;
  // scatter list to build operator tiles from reduced blocks: for every stored block (i >= j) entry
  // (block index b, a, c) -> (a-tile index, position in tile); generated on demand by the plan
  std::vector<int32_t> block_i, block_j;
  int64_t flops = 0;
  int32_t max_targets = 0;

  int64_t n_tiles() const { return (int64_t)row_idx.size(); }
  int64_t n_pairs() const { return (int64_t)pair_a.size(); }
};

int lrbms_symbolic_build(lrbms_symbolic& S, int32_t n_sub, const int32_t* sizes, int32_t n_blocks, const int32_t* bi,
                         const int32_t* bj, std::string* err);
