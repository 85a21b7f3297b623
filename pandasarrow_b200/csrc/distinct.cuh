// count_distinct per group (GROUPBY_NUMERIC_AGG(count_distinct, int64_t), /root/reference/src/dataframe.cpp:1528;
// arrow::compute "count_distinct" with default CountOptions: distinct NON-NULL values, a memo table keyed by the
// value's bits — so -0.0 / +0.0 and NaNs with different payloads count separately, which is reproduced here).
//
// Exact, sort based:  (group id, widened value bits) of every row  ->  stable radix sort by value, then by id
// (cub::DeviceRadixSort, library code)  ->  one pass counting the positions where (id, bits) changes.
#pragma once
#include "rowids.cuh"

namespace pa {

constexpr uint32_t kNoGroup = 0xFFFFFFFFu;

struct DistinctFillArgs {
  RowIdArgs ids;
  const void* vals;
  const uint8_t* vvalid;
  int64_t voff;
  int vw;
  uint64_t* out_bits;
  uint32_t* out_id;      // kNoGroup for null values: sorts behind every real group
};

template <int VC>
__global__ void __launch_bounds__(256) k_distinct_fill(DistinctFillArgs a) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (; i < a.ids.n; i += stride) {
    const bool valid = !a.vvalid || bit_at(a.vvalid, a.voff + i);
    a.out_id[i] = valid ? rowid_lookup(a.ids, i) : kNoGroup;
    a.out_bits[i] = valid ? load_wide_rt<VC>(a.vals, i, a.vw) : 0ull;
  }
}

constexpr int CD_PER_THREAD = 8;

// ids ascending (kNoGroup last), bits ascending inside an id.  Every thread walks CD_PER_THREAD consecutive
// positions; the run it ends with (for large groups: its whole chunk) is combined across the warp before the
// one global add per distinct id and warp.
__global__ void __launch_bounds__(256) k_distinct_count(const uint32_t* ids, const uint64_t* bits, int64_t n, unsigned long long* distinct) {
  const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t begin = t * CD_PER_THREAD;
  uint32_t cur = kNoGroup, cnt = 0;
  if (begin < n) {
    uint32_t prev_id = begin ? ids[begin - 1] : kNoGroup;
    uint64_t prev_bits = begin ? bits[begin - 1] : 0ull;
    const bool have_prev = begin > 0;
    for (int j = 0; j < CD_PER_THREAD; ++j) {
      const int64_t i = begin + j;
      if (i >= n) break;
      const uint32_t id = ids[i];
      const uint64_t b = bits[i];
      if (id != cur) {
        if (cur != kNoGroup && cnt) atomicAdd(distinct + cur, static_cast<unsigned long long>(cnt));
        cur = id;
        cnt = 0;
      }
      const bool is_new = !(have_prev || j > 0) || id != prev_id || b != prev_bits;
      if (id != kNoGroup && is_new) ++cnt;
      prev_id = id;
      prev_bits = b;
    }
  }
  const uint32_t peers = __match_any_sync(0xFFFFFFFFu, cur);
  const uint32_t total = __reduce_add_sync(peers, cnt);
  if (cur != kNoGroup && total && static_cast<int>(lane_id()) == __ffs(peers) - 1) atomicAdd(distinct + cur, static_cast<unsigned long long>(total));
}

__global__ void __launch_bounds__(256) k_distinct_emit(const unsigned long long* distinct, uint32_t G, int64_t* out) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < G) out[g] = static_cast<int64_t>(distinct[g]);
}

}  // namespace pa
