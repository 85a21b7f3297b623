// utf8 / large_utf8 key columns on the device (buffers[1] = offsets, buffers[2] = bytes).
// makeGroups groups any key type through arrow's Grouper (/root/reference/src/dataframe.cpp:1579-1584) and the
// reference's golden tests key on strings (tests/cudf_examples/dataframe_resample_test.cpp:8-69); round 1
// dictionary-encoded such keys on the HOST.  Here:
//   k_str_hash    every row's string -> a seeded 64-bit hash; the ordinary stages 1-3 then run on that 8-byte column
//   k_str_verify  exactness: every row's bytes are compared with the bytes of its group's first row (the group's
//                 representative); two different strings with one hash set a flag and the host re-hashes the
//                 column with another seed (probability ~ n_groups^2 / 2^65 per attempt)
//   k_str_lengths / k_str_gather   unique(): lengths of the representatives -> offsets (scan) -> bytes
#pragma once
#include "rowids.cuh"

namespace pa {

struct StrCol {
  const void* offsets;       // int32 or int64, already advanced by the array offset
  const uint8_t* bytes;
  int wide;                  // 1 = int64 offsets (large_utf8)
};

__device__ __forceinline__ int64_t str_off(const StrCol& c, int64_t i) {
  return c.wide ? static_cast<const int64_t*>(c.offsets)[i] : static_cast<int64_t>(static_cast<const int32_t*>(c.offsets)[i]);
}

// 64-bit hash of a byte string: 8 bytes per step through a multiply-xorshift mix (wyhash-like), length folded in.
__device__ __forceinline__ uint64_t str_hash64(const uint8_t* p, int64_t len, uint64_t seed) {
  uint64_t h = seed ^ (static_cast<uint64_t>(len) * 0x9E3779B97F4A7C15ull);
  int64_t i = 0;
  for (; i + 8 <= len; i += 8) {
    uint64_t w = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b) w |= static_cast<uint64_t>(p[i + b]) << (8 * b);
    h = (h ^ w) * 0xFF51AFD7ED558CCDull;
    h ^= h >> 32;
  }
  uint64_t w = 0;
  for (int b = 0; i < len; ++i, ++b) w |= static_cast<uint64_t>(p[i]) << (8 * b);
  h = (h ^ w) * 0xC4CEB9FE1A85EC53ull;
  h ^= h >> 29;
  h *= 0x94D049BB133111EBull;
  return h ^ (h >> 32);
}

__global__ void __launch_bounds__(256) k_str_hash(StrCol c, const uint8_t* valid, int64_t bit_off, int64_t n, uint64_t seed, uint64_t* out) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) {
    if (valid && !bit_at(valid, bit_off + i)) { out[i] = 0; continue; }
    const int64_t a = str_off(c, i), b = str_off(c, i + 1);
    out[i] = str_hash64(c.bytes + a, b - a, seed);
  }
}

// Every row against the representative (first row) of the group its HASH belongs to; `a` is the read-only
// hash -> group lookup of rowids.cuh built over the hash column.
__global__ void __launch_bounds__(256) k_str_verify(StrCol c, RowIdArgs a, const uint32_t* first_row, uint32_t* collision) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (; i < a.n; i += stride) {
    if (a.kvalid && !bit_at(a.kvalid, a.koff + i)) continue;
    const uint32_t g = rowid_lookup(a, i);
    if (g >= a.G) { atomicExch(collision, 2u); continue; }
    const int64_t r = first_row[g];
    if (r == i) continue;
    const int64_t x = str_off(c, i), lx = str_off(c, i + 1) - x;
    const int64_t y = str_off(c, r), ly = str_off(c, r + 1) - y;
    bool same = lx == ly;
    for (int64_t j = 0; same && j < lx; ++j) same = c.bytes[x + j] == c.bytes[y + j];
    if (!same) atomicExch(collision, 1u);
  }
}

__global__ void __launch_bounds__(256) k_str_lengths(StrCol c, const uint32_t* first_row, const uint8_t* key_kind, uint32_t G, uint32_t* len_out) {
  const uint32_t g = blockIdx.x * 256u + threadIdx.x;
  if (g < G) {
    const int64_t r = first_row[g];
    len_out[g] = key_kind[g] == 1 /* KK_NULL */ ? 0u : static_cast<uint32_t>(str_off(c, r + 1) - str_off(c, r));
  }
  if (g == G) len_out[G] = 0;
}

__global__ void __launch_bounds__(256) k_sum_u32(const uint32_t* v, uint32_t n, unsigned long long* out) {
  unsigned long long acc = 0;
  for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < n; i += gridDim.x * 256u) acc += v[i];
#pragma unroll
  for (int d = 16; d; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
  if (lane_id() == 0 && acc) atomicAdd(out, acc);
}

// validity words of unique(): bit g = group g's key is not null (whole words, zero padded)
__global__ void __launch_bounds__(256) k_kind_validity(const uint8_t* key_kind, uint32_t G, uint32_t* out) {
  const uint32_t g = blockIdx.x * 256u + threadIdx.x;
  const bool ok = g < G && key_kind[g] != 1;
  const uint32_t m = __ballot_sync(0xFFFFFFFFu, ok);
  if (lane_id() == 0 && g < (G + 31) / 32 * 32) out[g >> 5] = m;
}

// out_offsets = exclusive scan of the lengths (G + 1 entries); one warp per group copies the bytes
__global__ void __launch_bounds__(256) k_str_gather(StrCol c, const uint32_t* first_row, const uint8_t* key_kind, uint32_t G,
                                                    const uint32_t* out_offsets, uint8_t* out_bytes) {
  const uint32_t g = (blockIdx.x * 256u + threadIdx.x) >> 5;
  if (g >= G || key_kind[g] == 1) return;
  const int64_t r = first_row[g];
  const int64_t a = str_off(c, r);
  const uint32_t len = out_offsets[g + 1] - out_offsets[g];
  for (uint32_t j = lane_id(); j < len; j += 32) out_bytes[out_offsets[g] + j] = c.bytes[a + j];
}

}  // namespace pa
