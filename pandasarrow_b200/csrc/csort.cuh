// Stable counting / LSD radix sort of (key, payload) pairs by a digit of the key — the counting sort behind
// Grouper::MakeGroupings (/root/reference/src/dataframe.cpp:1586-1588: row numbers of every group, group-contiguous,
// ascending inside a group).  Own kernels (round 1 called cub::DeviceRadixSort here):
//   k_cs_hist     every CTA owns one contiguous chunk of the rows and counts its digits (shared-memory histogram)
//   (scan)        exclusive prefix over the digit-major [digit][cta] count matrix (order.cuh k_scan_*)
//   k_cs_scatter  the CTA walks its chunk again, tile by tile IN ORDER: every warp ranks the 32 rows of a step among
//                 themselves with MATCH.ANY (rank = earlier lanes with the same digit), keeps per-warp digit counters
//                 in shared memory, the warps' counters are prefix-summed per tile on top of the CTA's running cursor
//                 -> every row knows its final position; rows of one digit keep their input order (stable)
// One pass handles up to 1024 digit values (groups <= 1024: a single pass, which also yields the group offsets
// straight from the scanned counts); more groups take ceil(log2 G / 10) passes.  The last pass also writes
// dest[row] = position, so that ApplyGroupings can SCATTER columns from coalesced reads (groupings.cuh k_take_scatter)
// instead of gathering 8 bytes out of every 128-byte line.
#pragma once
#include "common.cuh"

namespace pa {

constexpr int CS_BITS = 10;
constexpr int CS_R = 1 << CS_BITS;
constexpr int CS_WARPS = 16;
constexpr int CS_THREADS = CS_WARPS * 32;
constexpr int CS_STEPS = 8;                          // 32-row steps per warp and tile
constexpr int CS_TILE = CS_THREADS * CS_STEPS;       // 4096 rows
constexpr size_t CS_SMEM = sizeof(uint32_t) * (CS_WARPS * CS_R + CS_R);

struct CsArgs {
  const uint32_t* keys;        // [n]
  const uint32_t* payload;     // [n] or null (payload = row number)
  int64_t n;
  int shift;                   // digit = (key >> shift) & (CS_R - 1)
  int nb;                      // number of chunks (= grid of both kernels)
  int64_t chunk;               // rows per chunk, a multiple of CS_TILE
  uint32_t* counts;            // [CS_R][nb] digit-major; after the scan: first output position of (digit, chunk)
  uint32_t* out_keys;          // [n] or null
  uint32_t* out_payload;       // [n]
  uint32_t* out_dest;          // [n] or null: out_dest[payload] = output position (last pass)
};

__global__ void __launch_bounds__(CS_THREADS) k_cs_hist(CsArgs a) {
  __shared__ uint32_t hist[CS_R];
  for (int i = threadIdx.x; i < CS_R; i += CS_THREADS) hist[i] = 0;
  __syncthreads();
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * a.chunk;
  const int64_t r1 = r0 + a.chunk < a.n ? r0 + a.chunk : a.n;
  for (int64_t r = r0 + threadIdx.x; r < r1; r += CS_THREADS) atomicAdd(&hist[(a.keys[r] >> a.shift) & (CS_R - 1)], 1u);
  __syncthreads();
  for (int d = threadIdx.x; d < CS_R; d += CS_THREADS) a.counts[static_cast<size_t>(d) * a.nb + blockIdx.x] = hist[d];
}

__global__ void __launch_bounds__(CS_THREADS, 1) k_cs_scatter(CsArgs a) {
  extern __shared__ __align__(16) uint32_t cs_smem[];
  uint32_t* whist = cs_smem;                       // [CS_WARPS][CS_R]
  uint32_t* cursor = cs_smem + CS_WARPS * CS_R;    // [CS_R]
  const uint32_t lane = lane_id(), w = threadIdx.x >> 5;
  for (int d = threadIdx.x; d < CS_R; d += CS_THREADS) cursor[d] = a.counts[static_cast<size_t>(d) * a.nb + blockIdx.x];
  for (int i = threadIdx.x; i < CS_WARPS * CS_R; i += CS_THREADS) whist[i] = 0;
  __syncthreads();
  const int64_t c0 = static_cast<int64_t>(blockIdx.x) * a.chunk;
  const int64_t c1 = c0 + a.chunk < a.n ? c0 + a.chunk : a.n;
  uint32_t* mine = whist + w * CS_R;
  for (int64_t t0 = c0; t0 < c1; t0 += CS_TILE) {
    uint32_t key[CS_STEPS], local[CS_STEPS];
    // phase A: rank inside the warp's 256-row segment, in row order
#pragma unroll
    for (int s = 0; s < CS_STEPS; ++s) {
      const int64_t r = t0 + w * (32 * CS_STEPS) + s * 32 + lane;
      const bool ok = r < c1;
      key[s] = ok ? a.keys[r] : 0u;
      const uint32_t digit = ok ? ((key[s] >> a.shift) & (CS_R - 1)) : static_cast<uint32_t>(CS_R);
      const uint32_t peers = __match_any_sync(0xFFFFFFFFu, digit);
      const int leader = __ffs(peers) - 1;
      uint32_t old = 0;
      if (ok && static_cast<int>(lane) == leader) {
        old = mine[digit];
        mine[digit] = old + __popc(peers);
      }
      old = __shfl_sync(0xFFFFFFFFu, old, leader);
      local[s] = old + __popc(peers & ((1u << lane) - 1u));
      __syncwarp();
    }
    __syncthreads();
    // phase B: per digit, the warps' counts -> output positions on top of the CTA's running cursor
    for (int d = threadIdx.x; d < CS_R; d += CS_THREADS) {
      uint32_t run = cursor[d];
#pragma unroll
      for (int ww = 0; ww < CS_WARPS; ++ww) {
        const uint32_t t = whist[ww * CS_R + d];
        whist[ww * CS_R + d] = run;
        run += t;
      }
      cursor[d] = run;
    }
    __syncthreads();
    // phase C: write
#pragma unroll
    for (int s = 0; s < CS_STEPS; ++s) {
      const int64_t r = t0 + w * (32 * CS_STEPS) + s * 32 + lane;
      if (r < c1) {
        const uint32_t digit = (key[s] >> a.shift) & (CS_R - 1);
        const uint32_t pos = mine[digit] + local[s];
        const uint32_t pl = a.payload ? a.payload[r] : static_cast<uint32_t>(r);
        if (a.out_keys) a.out_keys[pos] = key[s];
        a.out_payload[pos] = pl;
        if (a.out_dest) a.out_dest[pl] = pos;
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < CS_WARPS * CS_R; i += CS_THREADS) whist[i] = 0;
    __syncthreads();
  }
}

// single pass over <= CS_R groups: offsets[g] = first position of group g = scanned counts of (digit g, chunk 0)
__global__ void __launch_bounds__(256) k_cs_offsets(const uint32_t* counts, int nb, uint32_t G, int64_t n, int32_t* offsets) {
  const uint32_t g = blockIdx.x * 256u + threadIdx.x;
  if (g < G) offsets[g] = static_cast<int32_t>(counts[static_cast<size_t>(g) * nb]);
  if (g == G) offsets[G] = static_cast<int32_t>(n);
}

}  // namespace pa
