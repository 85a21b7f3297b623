// Group materialisation on the device: what the reference gets from
// arrow::compute::Grouper::MakeGroupings (row numbers of every group, group-contiguous, ascending inside
// a group; /root/reference/src/dataframe.cpp:1586-1588) and Grouper::ApplyGroupings (every column gathered
// into that order; dataframe.cpp:1546,1562) — its 82 % time sink.
//
//   ids      k_rowid_scan (rowids.cuh): group id of every row
//   order    a STABLE sort of (id, row) by id — cub::DeviceRadixSort over ceil(log2 G) bits (library code,
//            bookkeeping like the first-appearance ordering; counted as such in DESIGN.md)
//   offsets  k_group_offsets: offsets[id] = first position of id in the sorted ids (every group has a row)
//   take     k_take_grouped: out[i] = column[order[i]] for 1/2/4/8-byte elements, validity bits gathered
//            with one ballot per 32 outputs
#pragma once
#include "common.cuh"

namespace pa {

__global__ void __launch_bounds__(256) k_iota_u32(uint32_t* out, int64_t n) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) out[i] = static_cast<uint32_t>(i);
}

// offsets has G + 1 entries; sorted_ids ascending, every id in [0, G) present
__global__ void __launch_bounds__(256) k_group_offsets(const uint32_t* sorted_ids, int64_t n, uint32_t G, int32_t* offsets) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  if (i == 0) offsets[G] = static_cast<int32_t>(n);
  for (; i < n; i += stride) {
    const uint32_t id = sorted_ids[i];
    if (i == 0 || sorted_ids[i - 1] != id) offsets[id] = static_cast<int32_t>(i);
  }
}

struct TakeArgs {
  const void* col;          // already advanced by the column's offset
  const uint8_t* valid;     // or null
  int64_t bit_off;
  int width;                // 1, 2, 4, 8
  const uint32_t* order;    // [n]
  int64_t n;
  void* out;
  uint32_t* out_valid;      // [ceil(n / 32)] or null
};

__global__ void __launch_bounds__(256) k_take_grouped(TakeArgs a) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t n_round = (a.n + 31) / 32 * 32;     // whole warps: the validity word is a ballot
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n_round; i += stride) {
    bool v = false;
    if (i < a.n) {
      const uint32_t r = __ldg(a.order + i);
      switch (a.width) {
        case 8: static_cast<uint64_t*>(a.out)[i] = __ldg(static_cast<const uint64_t*>(a.col) + r); break;
        case 4: static_cast<uint32_t*>(a.out)[i] = __ldg(static_cast<const uint32_t*>(a.col) + r); break;
        case 2: static_cast<uint16_t*>(a.out)[i] = __ldg(static_cast<const uint16_t*>(a.col) + r); break;
        default: static_cast<uint8_t*>(a.out)[i] = __ldg(static_cast<const uint8_t*>(a.col) + r); break;
      }
      v = a.valid ? bit_at(a.valid, a.bit_off + r) : true;
    }
    if (a.out_valid) {
      const uint32_t m = __ballot_sync(0xFFFFFFFFu, v);
      if (lane_id() == 0) a.out_valid[i >> 5] = m;
    }
  }
}

}  // namespace pa
