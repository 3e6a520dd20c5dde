// Stage lists: the neighbour table nbr[K][n_pad] of a SUBMANIFOLD rulebook re-laid for the gather of k_conv_tc.
//
// A "stage" of the convolution kernel is (tile of 128 output rows, kernel offset k).  Building a stage from the table
// costs the gathering warp 128 index loads, ballots, prefix sums and list stores -- most of its instructions at the
// shallow levels -- and it is repeated by every convolution that shares the rulebook (8 layers x forward + dgrad per
// resolution level).  The stage lists hold that work done once: per stage the compacted live rows as ready-made
// (source row, swizzled shared-memory offset) pairs, the count, and the disable-output-lane mask.
//
// One device buffer (scn_stage_lists_bytes):
//   [0,16)                  u32 entries_used (bump allocator of the builder), u32 K, u32 n_tiles, u32 magic
//   hdr  [n_tiles][K] int2  .x = first entry of the stage's list (a multiple of 8: 64-byte aligned), .y = entries in the
//                           list = live rows rounded up to a multiple of 8 (0..128)
//   msk  [n_tiles][K] uint4 bit r of word r/32 set <=> output row r of the tile has NO neighbour through offset k
//   ent  [...]        int2  .x = source row, .y = (r << 7) + ((r & 7) << 4): byte offset of (row r, 16-byte chunk 0)
//                           in a 128-row SWIZZLE_128B tile.  A list is padded to a multiple of 8 entries by REPEATING
//                           its last entry (copying a row twice is harmless), so the gather loop runs whole passes
//                           with no per-item predicate
// Lists of one tile are contiguous (k ascending); tiles are placed in bump-allocation order.
#pragma once
#include <stdint.h>

namespace sl {
constexpr uint32_t MAGIC = 0x534C3031u;   // "SL01"
constexpr int TILE = 128;

__host__ __device__ inline size_t hdr_offset() { return 16; }
__host__ __device__ inline size_t msk_offset(int64_t n_tiles, int K) {
  return 16 + (((size_t)n_tiles * (size_t)K * 8 + 15) & ~(size_t)15);
}
__host__ __device__ inline size_t ent_offset(int64_t n_tiles, int K) {
  return msk_offset(n_tiles, K) + (size_t)n_tiles * (size_t)K * 16;
}
constexpr int PAD = 8;                     // entries per list are a multiple of this
// upper bound: every table entry live (a full list needs no padding)
__host__ __device__ inline size_t total_bytes(int64_t n_tiles, int K) {
  return ent_offset(n_tiles, K) + (size_t)n_tiles * K * TILE * 8;
}
}  // namespace sl
