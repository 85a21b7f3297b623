// Group materialisation on the device: what the reference gets from
// arrow::compute::Grouper::MakeGroupings (row numbers of every group, group-contiguous, ascending inside
// a group; /root/reference/src/dataframe.cpp:1586-1588) and Grouper::ApplyGroupings (every column gathered
// into that order; dataframe.cpp:1546,1562) — its 82 % time sink.
//
//   ids      k_rowid_scan (rowids.cuh): group id of every row
//   order    a STABLE counting sort of the rows by id (csort.cuh: own kernels, one pass per 10 bits of the id);
//            the last pass also writes dest[row] = position of the row in group order
//   offsets  straight from the scanned digit counts (<= 1024 groups) or k_group_offsets on the sorted ids
//   take     k_take_scatter: out[dest[i]] = column[i] — the column is READ coalesced once and written to its group
//            position (the writes of one group are consecutive over time, so L2 merges them into whole sectors);
//            round 1's gather (out[i] = column[order[i]]) fetched a 128-byte line per 8-byte element: 24.9 GB of
//            DRAM reads for 1.6 GB gathered.  Booleans (bit-packed) are gathered bit by bit with a ballot per
//            32 outputs (k_take_bits; the packed input is n / 8 bytes and stays in L2).
#pragma once
#include "common.cuh"

namespace pa {

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

struct TakeScatterArgs {
  const void* col;          // already advanced by the column's offset
  const uint8_t* valid;     // or null
  int64_t bit_off;
  int width;                // 1, 2, 4, 8
  const uint32_t* dest;     // [n] position of row i in group order
  int64_t n;
  void* out;
  uint32_t* out_valid;      // [ceil(n / 32)] zero-initialised, or null
};

__global__ void __launch_bounds__(256) k_take_scatter(TakeScatterArgs a) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < a.n; i += stride) {
    const uint32_t d = a.dest[i];
    switch (a.width) {
      case 8: static_cast<uint64_t*>(a.out)[d] = static_cast<const uint64_t*>(a.col)[i]; break;
      case 4: static_cast<uint32_t*>(a.out)[d] = static_cast<const uint32_t*>(a.col)[i]; break;
      case 2: static_cast<uint16_t*>(a.out)[d] = static_cast<const uint16_t*>(a.col)[i]; break;
      default: static_cast<uint8_t*>(a.out)[d] = static_cast<const uint8_t*>(a.col)[i]; break;
    }
    if (a.out_valid && bit_at(a.valid, a.bit_off + i)) atomicOr(a.out_valid + (d >> 5), 1u << (d & 31u));
  }
}

// bit-packed columns (Arrow booleans, or any validity bitmap): out bit i = in bit order[i]
__global__ void __launch_bounds__(256) k_take_bits(const uint8_t* bits, int64_t bit_off, const uint32_t* order, int64_t n, uint32_t* out) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t n_round = (n + 31) / 32 * 32;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n_round; i += stride) {
    const bool v = i < n ? bit_at(bits, bit_off + __ldg(order + i)) : false;
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, v);
    if (lane_id() == 0) out[i >> 5] = m;
  }
}

}  // namespace pa
