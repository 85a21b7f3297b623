// Radix-partition pre-pass for very many groups (table far larger than L2).
//
// With > ~2 M groups the global table leaves L2 and every row costs a random 32-byte DRAM read plus a
// write-back (profiles/r1_gtable_1M_ncu_full.md).  This pre-pass reorders the rows by the TOP bits of
// the table hash — the same bits that select the table region of a key (gtable_home) — so that the
// scan kernel afterwards walks the table region by region and its working set stays L2 resident:
//   k_part_hist     rows per partition (shared-memory histogram per CTA, one global add per bin)
//   k_part_offsets  exclusive prefix -> start offset (write cursor) of every partition
//   k_part_scatter  per tile of 4096 rows: shared-memory histogram gives every row its rank inside the
//                   tile's run for its partition; the tile is regrouped in shared memory and written
//                   out as contiguous runs (one global cursor add per partition and tile) together
//                   with the original row numbers, which first/last need
// The scan then reads (key, value, row) from the permuted arrays.  Extra traffic: 8 B/row read by the
// histogram, 16 B/row read + 20 B/row written by the scatter, 20 instead of 16 B/row read by the scan.
// Only for 8-byte keys and values without validity bitmaps (otherwise the direct scan is used).
#pragma once
#include "gtable.cuh"

namespace pa {

constexpr int PH_THREADS = 1024;                      // histogram kernel
constexpr int PT_THREADS = 512;
constexpr int PT_ROWS = 8;                            // rows per thread
constexpr int PT_TILE = PT_THREADS * PT_ROWS;         // 4096 rows per CTA tile
constexpr int PT_MAX_PARTS = 1024;

struct PartArgs {
  const uint64_t* keys;
  const uint64_t* vals;
  int64_t n;
  int log_parts;                 // partitions = 1 << log_parts (<= PT_MAX_PARTS)
  unsigned long long* counts;    // [parts] rows per partition, then (after k_part_offsets) the write cursors
  uint64_t* out_keys;
  uint64_t* out_vals;
  uint32_t* out_rows;
};

__device__ __forceinline__ uint32_t part_of(uint64_t key, int log_parts) {
  return static_cast<uint32_t>(gtable_mix(key) >> (64 - log_parts));
}

__global__ void __launch_bounds__(PH_THREADS) k_part_hist(PartArgs a) {
  __shared__ unsigned int s_hist[PT_MAX_PARTS];
  const int parts = 1 << a.log_parts;
  for (int i = threadIdx.x; i < parts; i += PH_THREADS) s_hist[i] = 0;
  __syncthreads();
  const int64_t n2 = a.n / 2;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(PH_THREADS) + threadIdx.x; i < n2; i += static_cast<int64_t>(gridDim.x) * PH_THREADS) {
    const ulonglong2 k = ldg_stream_u64x2(a.keys + 2 * i);
    atomicAdd(&s_hist[part_of(k.x, a.log_parts)], 1u);
    atomicAdd(&s_hist[part_of(k.y, a.log_parts)], 1u);
  }
  if ((a.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&s_hist[part_of(a.keys[a.n - 1], a.log_parts)], 1u);
  __syncthreads();
  for (int i = threadIdx.x; i < parts; i += PH_THREADS)
    if (s_hist[i]) atomicAdd(a.counts + i, static_cast<unsigned long long>(s_hist[i]));
}

// one CTA: counts -> exclusive prefix (in place)
__global__ void __launch_bounds__(PT_MAX_PARTS) k_part_offsets(unsigned long long* counts, int parts) {
  __shared__ unsigned long long s[PT_MAX_PARTS];
  const int t = threadIdx.x;
  s[t] = t < parts ? counts[t] : 0ull;
  __syncthreads();
  for (int d = 1; d < PT_MAX_PARTS; d <<= 1) {
    const unsigned long long v = t >= d ? s[t - d] : 0ull;
    __syncthreads();
    s[t] += v;
    __syncthreads();
  }
  if (t < parts) counts[t] = s[t] - (t < parts ? counts[t] : 0ull);   // inclusive - own = exclusive
}

struct PtSmem {
  static constexpr size_t OFF_KEY = 0;
  static constexpr size_t OFF_VAL = OFF_KEY + sizeof(uint64_t) * PT_TILE;
  static constexpr size_t OFF_DELTA = OFF_VAL + sizeof(uint64_t) * PT_TILE;
  static constexpr size_t OFF_ROW = OFF_DELTA + sizeof(unsigned long long) * PT_MAX_PARTS;
  static constexpr size_t OFF_HIST = OFF_ROW + sizeof(uint32_t) * PT_TILE;
  static constexpr size_t OFF_OFF = OFF_HIST + sizeof(uint32_t) * PT_MAX_PARTS;
  static constexpr size_t TOTAL = OFF_OFF + sizeof(uint32_t) * PT_MAX_PARTS;
};
static_assert(2 * (PtSmem::TOTAL + 1024) <= 228 * 1024, "two scatter CTAs per SM");

// Two CTAs per SM: the phases of a tile are separated by barriers (load -> rank -> claim -> regroup -> write),
// so a second resident CTA keeps the memory system busy while the first one regroups.
__global__ void __launch_bounds__(PT_THREADS, 2) k_part_scatter(PartArgs a) {
  extern __shared__ __align__(16) unsigned char pt_smem[];
  uint64_t* st_key = reinterpret_cast<uint64_t*>(pt_smem + PtSmem::OFF_KEY);
  uint64_t* st_val = reinterpret_cast<uint64_t*>(pt_smem + PtSmem::OFF_VAL);
  uint32_t* st_row = reinterpret_cast<uint32_t*>(pt_smem + PtSmem::OFF_ROW);
  unsigned int* s_hist = reinterpret_cast<unsigned int*>(pt_smem + PtSmem::OFF_HIST);
  unsigned int* s_off = reinterpret_cast<unsigned int*>(pt_smem + PtSmem::OFF_OFF);
  // s_delta[p] = (start of this tile's run for p in the output) - (start of the run in the regrouped tile):
  // regrouped position i of partition p goes to output position s_delta[p] + i
  unsigned long long* s_delta = reinterpret_cast<unsigned long long*>(pt_smem + PtSmem::OFF_DELTA);
  __shared__ unsigned int s_wsum[PT_THREADS / 32];
  constexpr int BINS = PT_MAX_PARTS / PT_THREADS;    // histogram bins per thread in the scan
  const int parts = 1 << a.log_parts;
  const int64_t ntiles = (a.n + PT_TILE - 1) / PT_TILE;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t tile0 = t * PT_TILE;
    const int64_t left = a.n - tile0;
    const int cnt = left < PT_TILE ? static_cast<int>(left) : PT_TILE;
    for (int i = threadIdx.x; i < parts; i += PT_THREADS) s_hist[i] = 0;
    __syncthreads();
    // rows 2 * (threadIdx.x + h * PT_THREADS) + {0, 1} of the tile: 128-bit loads
    uint64_t key[PT_ROWS], val[PT_ROWS];
    uint32_t pr[PT_ROWS];    // partition << 13 | rank inside the tile's run for that partition
#pragma unroll
    for (int h = 0; h < PT_ROWS / 2; ++h) {
      const int r = 2 * (threadIdx.x + h * PT_THREADS);
      if (r + 1 < cnt) {
        const ulonglong2 k = ldg_stream_u64x2(a.keys + tile0 + r), v = ldg_stream_u64x2(a.vals + tile0 + r);
        key[2 * h] = k.x; key[2 * h + 1] = k.y; val[2 * h] = v.x; val[2 * h + 1] = v.y;
      } else {
        key[2 * h] = key[2 * h + 1] = 0; val[2 * h] = val[2 * h + 1] = 0;
        if (r < cnt) { key[2 * h] = a.keys[tile0 + r]; val[2 * h] = a.vals[tile0 + r]; }
      }
    }
#pragma unroll
    for (int j = 0; j < PT_ROWS; ++j) {
      const int r = 2 * (threadIdx.x + (j >> 1) * PT_THREADS) + (j & 1);
      pr[j] = 0xFFFFFFFFu;
      if (r < cnt) {
        const uint32_t p = part_of(key[j], a.log_parts);
        pr[j] = (p << 13) | atomicAdd(&s_hist[p], 1u);
      }
    }
    __syncthreads();
    // exclusive scan of the tile histogram (BINS adjacent bins per thread) + claim the output ranges
    {
      unsigned int mine[BINS], tot = 0;
#pragma unroll
      for (int b = 0; b < BINS; ++b) {
        const int bin = threadIdx.x * BINS + b;
        mine[b] = bin < parts ? s_hist[bin] : 0u;
        tot += mine[b];
      }
      unsigned int incl = tot;
      const uint32_t lane = lane_id(), w = threadIdx.x >> 5;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned int v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= static_cast<uint32_t>(d)) incl += v;
      }
      if (lane == 31) s_wsum[w] = incl;
      __syncthreads();
      if (w == 0) {
        unsigned int x = lane < PT_THREADS / 32 ? s_wsum[lane] : 0u;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const unsigned int v = __shfl_up_sync(0xFFFFFFFFu, x, d);
          if (lane >= static_cast<uint32_t>(d)) x += v;
        }
        if (lane < PT_THREADS / 32) s_wsum[lane] = x;
      }
      __syncthreads();
      unsigned int excl = incl - tot + (w ? s_wsum[w - 1] : 0u);
#pragma unroll
      for (int b = 0; b < BINS; ++b) {
        const int bin = threadIdx.x * BINS + b;
        if (bin < parts) {
          s_off[bin] = excl;
          const unsigned long long base = mine[b] ? atomicAdd(a.counts + bin, static_cast<unsigned long long>(mine[b])) : 0ull;
          s_delta[bin] = base - excl;
        }
        excl += mine[b];
      }
    }
    __syncthreads();
    // regroup the tile in shared memory
#pragma unroll
    for (int j = 0; j < PT_ROWS; ++j) {
      if (pr[j] == 0xFFFFFFFFu) continue;
      const uint32_t pos = s_off[pr[j] >> 13] + (pr[j] & 0x1FFFu);
      const int r = 2 * (threadIdx.x + (j >> 1) * PT_THREADS) + (j & 1);
      st_key[pos] = key[j];
      st_val[pos] = val[j];
      st_row[pos] = static_cast<uint32_t>(tile0 + r);
    }
    __syncthreads();
    // contiguous runs out (the partition of a regrouped row is recomputed from its key)
    for (int i = threadIdx.x; i < cnt; i += PT_THREADS) {
      const uint64_t k = st_key[i];
      const unsigned long long d = s_delta[part_of(k, a.log_parts)] + static_cast<unsigned long long>(i);
      a.out_keys[d] = k;
      a.out_vals[d] = st_val[i];
      a.out_rows[d] = st_row[i];
    }
    __syncthreads();
  }
}

}  // namespace pa
