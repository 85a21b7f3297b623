// Stable argsort of one column on the device — what Series::sort / Series::argsort / DataFrame::sort_index /
// DataFrame::sort_values get from arrow::compute "array_sort_indices" (+ Take) in the reference
// (/root/reference/src/series.cpp:864-868,978-992, dataframe.cpp:1062-1071,1188-1208).
//
//   k_sort_keys   value -> order-preserving 64-bit key (descending: complemented, so one ascending stable sort serves
//                 both orders and equal values keep their input order, as arrow's stable sort does); NaN and null rows
//                 get a class byte (1 / 2) because arrow places them after every number in BOTH orders (NaNs before
//                 nulls); a global histogram of all seven 10-bit digits tells the host which digits are constant
//   k_ls_hist / k_ls_scatter   one stable counting-sort pass over a 10-bit digit of the 64-bit key (the scheme of
//                 csort.cuh: per-CTA chunk counts -> scan -> MATCH.ANY ranks inside a warp, per-warp counters in shared
//                 memory), carrying (key, row number); passes whose digit is constant are skipped
//   class pass    only when NaNs or nulls exist: one more stable pass on the class byte
// The last pass also writes dest[row] = position, so that columns are taken into sorted order by the scatter-shaped
// k_take_scatter of groupings.cuh (coalesced reads) rather than by a gather.
#pragma once
#include "csort.cuh"
#include "group_result.cuh"

namespace pa {

constexpr int LS_DIGITS = 7;   // ceil(64 / CS_BITS)

struct SortKeyArgs {
  const void* vals;
  const uint8_t* valid;
  int64_t voff;
  int64_t n;
  int vw;
  int descending;
  uint64_t* keys;          // [n]
  uint8_t* cls;            // [n] 0 regular, 1 NaN, 2 null
  uint32_t* hist;          // [LS_DIGITS][CS_R] global digit histograms (zeroed by the host)
  uint32_t* flags;         // [0] |= 1 when a NaN or null exists
};

template <int VC>
__global__ void __launch_bounds__(512) k_sort_keys(SortKeyArgs a) {
  __shared__ uint32_t sh[LS_DIGITS * CS_R];
  for (int i = threadIdx.x; i < LS_DIGITS * CS_R; i += 512) sh[i] = 0;
  __syncthreads();
  bool special = false;
  for (int64_t r = blockIdx.x * 512ll + threadIdx.x; r < a.n; r += static_cast<int64_t>(gridDim.x) * 512) {
    const bool ok = !a.valid || bit_at(a.valid, a.voff + r);
    uint64_t key = 0;
    uint8_t c = 2;
    if (ok) {
      uint64_t vb = load_wide_rt<VC>(a.vals, r, a.vw);
      if constexpr (VC == VC_F) {
        if (vb == 0x8000000000000000ull) vb = 0;      // -0.0 == +0.0 for arrow's comparator: ties keep their input order
      }
      if (Wide<VC>::is_nan(vb)) c = 1;
      else {
        c = 0;
        key = Wide<VC>::ord(vb);
        if (a.descending) key = ~key;
      }
    }
    a.keys[r] = key;
    a.cls[r] = c;
    special |= c != 0;
#pragma unroll
    for (int d = 0; d < LS_DIGITS; ++d) atomicAdd(&sh[d * CS_R + ((key >> (d * CS_BITS)) & (CS_R - 1))], 1u);
  }
  if (special) atomicOr(a.flags, 1u);
  __syncthreads();
  for (int i = threadIdx.x; i < LS_DIGITS * CS_R; i += 512)
    if (sh[i]) atomicAdd(a.hist + i, sh[i]);
}

struct LsArgs {
  const uint64_t* keys;        // [n] keys in the current order
  const uint32_t* payload;     // [n] row numbers in the current order, or null (identity)
  const uint8_t* cls;          // class pass: class byte by ORIGINAL row number (digit = cls[payload]); else null
  int64_t n;
  int shift;                   // digit = (key >> shift) & (CS_R - 1)
  int nb;
  int64_t chunk;               // rows per chunk, a multiple of CS_TILE
  uint32_t* counts;            // [CS_R][nb]
  uint64_t* out_keys;          // [n] or null
  uint32_t* out_payload;       // [n]
  uint32_t* out_dest;          // [n] or null: out_dest[payload] = output position (last pass)
};

__device__ __forceinline__ uint32_t ls_digit(const LsArgs& a, int64_t r, uint64_t key) {
  if (a.cls) return a.cls[a.payload ? a.payload[r] : static_cast<uint32_t>(r)];
  return static_cast<uint32_t>(key >> a.shift) & (CS_R - 1);
}

__global__ void __launch_bounds__(CS_THREADS) k_ls_hist(LsArgs a) {
  __shared__ uint32_t hist[CS_R];
  for (int i = threadIdx.x; i < CS_R; i += CS_THREADS) hist[i] = 0;
  __syncthreads();
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * a.chunk;
  const int64_t r1 = r0 + a.chunk < a.n ? r0 + a.chunk : a.n;
  for (int64_t r = r0 + threadIdx.x; r < r1; r += CS_THREADS) atomicAdd(&hist[ls_digit(a, r, a.cls ? 0ull : a.keys[r])], 1u);
  __syncthreads();
  for (int d = threadIdx.x; d < CS_R; d += CS_THREADS) a.counts[static_cast<size_t>(d) * a.nb + blockIdx.x] = hist[d];
}

__global__ void __launch_bounds__(CS_THREADS, 1) k_ls_scatter(LsArgs a) {
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
    uint64_t key[CS_STEPS];
    uint32_t local[CS_STEPS], dig[CS_STEPS];
    // phase A: rank inside the warp's 256-row segment, in row order
#pragma unroll
    for (int s = 0; s < CS_STEPS; ++s) {
      const int64_t r = t0 + w * (32 * CS_STEPS) + s * 32 + lane;
      const bool ok = r < c1;
      key[s] = ok ? a.keys[r] : 0ull;
      const uint32_t digit = ok ? ls_digit(a, r, key[s]) : static_cast<uint32_t>(CS_R);
      dig[s] = digit;
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
        const uint32_t pos = mine[dig[s]] + local[s];
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

// identity order (nothing to sort: every digit constant, no NaN / null)
__global__ void __launch_bounds__(256) k_ls_identity(uint32_t* order, uint32_t* dest, int64_t n) {
  int64_t i = blockIdx.x * 256ll + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * 256;
  for (; i < n; i += stride) { order[i] = static_cast<uint32_t>(i); dest[i] = static_cast<uint32_t>(i); }
}

// uint32 row numbers -> the uint64 indices arrow's array_sort_indices returns
__global__ void __launch_bounds__(256) k_ls_widen(const uint32_t* in, uint64_t* out, int64_t n) {
  int64_t i = blockIdx.x * 256ll + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * 256;
  for (; i < n; i += stride) out[i] = in[i];
}

}  // namespace pa
