// First-appearance order of the groups WITHOUT a sort.
//
// The global-table and resample paths end with G compacted (first_row, slot) pairs in table order and have to emit
// the groups ordered by first row.  Round 1 sorted the pairs (cub::DeviceRadixSort: +18-23 ms at 100 M groups).
// First rows are distinct numbers in [0, n_rows), so the position of a group is simply the number of groups whose
// first row is smaller — a rank query on a bitmap of the first rows:
//   k_bm_set     bitmap[first_row] = 1                                  (n_rows / 8 bytes: 125 MB for 1 B rows, L2 resident)
//   k_bm_count8  number of set bits per 256-bit block                   (n_rows / 256 counters)
//   k_scan_*     exclusive prefix of the block counts                   (three small kernels)
//   k_bm_rank    rank = prefix[block] + popc(words of the block before the bit) -> sorted_slot[rank] = slot
// Traffic: 2 random 4-byte accesses per group on L2-resident arrays + two streaming passes over n_rows / 8 bytes.
// Replaces the ordering step after Grouper::Consume-equivalent group discovery (dataframe.cpp:1584-1591 gets its
// order from arrow's Grouper); no library code.
#pragma once
#include "common.cuh"

namespace pa {

constexpr int BM_BLOCK_WORDS = 8;   // 32-bit words per prefix block (one 32-byte sector)

__global__ void __launch_bounds__(256) k_bm_set(const uint32_t* first_row, uint32_t G, uint32_t* bitmap) {
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i < G) {
    const uint32_t r = first_row[i];
    atomicOr(bitmap + (r >> 5), 1u << (r & 31u));
  }
}

// one thread per 8-word block
__global__ void __launch_bounds__(256) k_bm_count8(const uint32_t* bitmap, uint64_t nblocks, uint32_t* counts) {
  uint64_t b = blockIdx.x * 256ull + threadIdx.x;
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * 256ull;
  for (; b < nblocks; b += stride) {
    const uint4 lo = *reinterpret_cast<const uint4*>(bitmap + b * BM_BLOCK_WORDS);
    const uint4 hi = *reinterpret_cast<const uint4*>(bitmap + b * BM_BLOCK_WORDS + 4);
    counts[b] = __popc(lo.x) + __popc(lo.y) + __popc(lo.z) + __popc(lo.w) + __popc(hi.x) + __popc(hi.y) + __popc(hi.z) + __popc(hi.w);
  }
}

// ---- device-wide exclusive scan of uint32 (in place), three kernels: tiles of 2048 -> tile sums -> apply ----
constexpr int SC_TILE = 2048;   // 256 threads x 8 items

__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t* total) {
  __shared__ uint32_t wsum[8];
  const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
    if (lane >= static_cast<uint32_t>(d)) incl += o;
  }
  if (lane == 31) wsum[w] = incl;
  __syncthreads();
  uint32_t base = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t s = wsum[i];
    if (static_cast<uint32_t>(i) < w) base += s;
    tot += s;
  }
  __syncthreads();
  *total = tot;
  return base + incl - v;
}

__global__ void __launch_bounds__(256) k_scan_tiles(uint32_t* data, uint64_t n, uint32_t* tile_sums) {
  const uint64_t base = static_cast<uint64_t>(blockIdx.x) * SC_TILE + threadIdx.x * 8ull;
  uint32_t v[8], sum = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { v[j] = base + j < n ? data[base + j] : 0u; sum += v[j]; }
  uint32_t total;
  uint32_t excl = block_exclusive_scan_256(sum, &total);
#pragma unroll
  for (int j = 0; j < 8; ++j) { if (base + j < n) data[base + j] = excl; excl += v[j]; }
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single CTA: exclusive scan of the tile sums (any length, sequential over chunks of 256 x 8)
__global__ void __launch_bounds__(256) k_scan_sums(uint32_t* sums, uint32_t n) {
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t c0 = 0; c0 < n; c0 += SC_TILE) {
    const uint32_t base = c0 + threadIdx.x * 8u;
    uint32_t v[8], sum = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { v[j] = base + j < n ? sums[base + j] : 0u; sum += v[j]; }
    uint32_t total;
    uint32_t excl = block_exclusive_scan_256(sum, &total) + carry;
#pragma unroll
    for (int j = 0; j < 8; ++j) { if (base + j < n) sums[base + j] = excl; excl += v[j]; }
    __syncthreads();
    if (threadIdx.x == 0) carry += total;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) k_scan_apply(uint32_t* data, uint64_t n, const uint32_t* tile_sums) {
  const uint64_t base = static_cast<uint64_t>(blockIdx.x) * SC_TILE + threadIdx.x * 8ull;
  const uint32_t add = tile_sums[blockIdx.x];
#pragma unroll
  for (int j = 0; j < 8; ++j) if (base + j < n) data[base + j] += add;
}

// sorted_slot[rank of first_row[i]] = slot[i]   (slot == null: i itself)
__global__ void __launch_bounds__(256) k_bm_rank(const uint32_t* first_row, const uint32_t* slot, uint32_t G, const uint32_t* bitmap,
                                                 const uint32_t* prefix8, uint32_t* sorted_slot) {
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i >= G) return;
  const uint32_t r = first_row[i];
  const uint32_t word = r >> 5, blk = word / BM_BLOCK_WORDS;
  uint32_t rank = prefix8[blk];
  for (uint32_t w = blk * BM_BLOCK_WORDS; w < word; ++w) rank += __popc(bitmap[w]);
  rank += __popc(bitmap[word] & ((1u << (r & 31u)) - 1u));
  sorted_slot[rank] = slot ? slot[i] : i;
}

}  // namespace pa
