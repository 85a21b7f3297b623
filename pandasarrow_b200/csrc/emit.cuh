// GroupResult -> Arrow output buffers (device side), key packing for composite keys, and the
// synthetic workload generator.  Result dtypes follow arrow::compute's scalar aggregates as the
// reference calls them (pd_core_macros.h:31,66,103,132; dataframe.cpp:1720,1775).
#pragma once
#include "group_result.cuh"

namespace pa {

// Writes validity as a bitmap: every warp covers 32 consecutive groups and lane 0 stores the word.
__device__ __forceinline__ void store_valid_bit(uint32_t* bitmap, uint32_t g, uint32_t G, bool valid) {
  const uint32_t m = __ballot_sync(0xFFFFFFFFu, valid && g < G);
  if (lane_id() == 0 && (g & ~31u) < G) bitmap[g >> 5] = m;
}

struct EmitArgs {
  GroupResult r;
  uint32_t G;              // number of groups — or an upper bound (grid size) when G_dev is set
  const uint32_t* G_dev;   // optional: exact number of groups, on the device
  int vc, vw;              // value class / byte width of the input value column
  const void* vals;        // input value column (first / last gather)
  const uint8_t* vvalid;
  int64_t voff;
  // outputs (null when not requested); *_valid are bitmaps of ceil(G/32) words
  void* o_sum; uint32_t* o_sum_valid;
  double* o_mean; uint32_t* o_mean_valid;
  int64_t* o_count;
  void* o_min; uint32_t* o_min_valid;
  void* o_max; uint32_t* o_max_valid;
  void* o_first; uint32_t* o_first_valid;
  void* o_last; uint32_t* o_last_valid;
  // merged (multi-GPU) results carry first/last values with them instead of row numbers
  const uint64_t* m_first_val; const uint64_t* m_last_val;
  const uint8_t* m_first_valid; const uint8_t* m_last_valid;
};

__device__ __forceinline__ void store_raw(void* out, uint32_t g, uint64_t bits, int vw) {
  switch (vw) {
    case 8: static_cast<uint64_t*>(out)[g] = bits; break;
    case 4: static_cast<uint32_t*>(out)[g] = static_cast<uint32_t>(bits); break;
    case 2: static_cast<uint16_t*>(out)[g] = static_cast<uint16_t>(bits); break;
    default: static_cast<uint8_t*>(out)[g] = static_cast<uint8_t>(bits); break;
  }
}

__device__ __forceinline__ void store_narrow(void* out, uint32_t g, uint64_t bits, int vc, int vw) {
  // `bits` is the widened 64-bit representation (double bits for VC_F)
  if (vc == VC_F) {
    if (vw == 8) static_cast<uint64_t*>(out)[g] = bits;
    else static_cast<float*>(out)[g] = static_cast<float>(__longlong_as_double(static_cast<long long>(bits)));
  } else {
    switch (vw) {
      case 8: static_cast<uint64_t*>(out)[g] = bits; break;
      case 4: static_cast<uint32_t*>(out)[g] = static_cast<uint32_t>(bits); break;
      case 2: static_cast<uint16_t*>(out)[g] = static_cast<uint16_t>(bits); break;
      default: static_cast<uint8_t*>(out)[g] = static_cast<uint8_t>(bits); break;
    }
  }
}

__device__ __forceinline__ void copy_elem(void* out, uint32_t g, const void* in, uint32_t row, int vw) {
  switch (vw) {
    case 8: static_cast<uint64_t*>(out)[g] = static_cast<const uint64_t*>(in)[row]; break;
    case 4: static_cast<uint32_t*>(out)[g] = static_cast<const uint32_t*>(in)[row]; break;
    case 2: static_cast<uint16_t*>(out)[g] = static_cast<const uint16_t*>(in)[row]; break;
    default: static_cast<uint8_t*>(out)[g] = static_cast<const uint8_t*>(in)[row]; break;
  }
}

__device__ __forceinline__ uint64_t ord_to_wide(uint64_t o, int vc) {
  if (vc == VC_F) return ord_to_f64_bits(o);
  if (vc == VC_I) return o ^ 0x8000000000000000ull;
  return o;
}

// grid: ceil(G/256) blocks of 256 (whole warps so that the ballots are full).
__global__ void __launch_bounds__(256) k_emit(EmitArgs a) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (a.G_dev) a.G = *a.G_dev;
  const bool in = g < a.G;
  const uint64_t cnt = in ? (a.r.count64 ? a.r.count64[g] : static_cast<uint64_t>(a.r.count[g])) : 0;
  const bool has = cnt > 0;
  if (a.o_sum) {
    if (in) static_cast<uint64_t*>(a.o_sum)[g] = a.r.sum[g];
    store_valid_bit(a.o_sum_valid, g, a.G, has);
  }
  if (a.o_mean) {
    if (in) {
      double s;
      if (a.vc == VC_F) s = __longlong_as_double(static_cast<long long>(a.r.sum[g]));
      else s = a.r.dsum[g];
      a.o_mean[g] = has ? s / static_cast<double>(cnt) : 0.0;
    }
    store_valid_bit(a.o_mean_valid, g, a.G, has);
  }
  if (a.o_count && in) a.o_count[g] = static_cast<int64_t>(cnt);
  if (a.o_min) {
    if (in) {
      uint64_t o = a.r.min_ord[g];
      uint64_t bits = (a.vc == VC_F && o == kMinInit) ? 0x7FF8000000000000ull /* all-NaN group */ : ord_to_wide(o, a.vc);
      if (!has) bits = 0;
      store_narrow(a.o_min, g, bits, a.vc, a.vw);
    }
    store_valid_bit(a.o_min_valid, g, a.G, has);
  }
  if (a.o_max) {
    if (in) {
      uint64_t o = a.r.max_ord[g];
      uint64_t bits = (a.vc == VC_F && o == kMaxInit) ? 0x7FF8000000000000ull : ord_to_wide(o, a.vc);
      if (!has) bits = 0;
      store_narrow(a.o_max, g, bits, a.vc, a.vw);
    }
    store_valid_bit(a.o_max_valid, g, a.G, has);
  }
  if (a.o_first && a.m_first_val) {
    if (in) store_raw(a.o_first, g, a.m_first_val[g], a.vw);
    store_valid_bit(a.o_first_valid, g, a.G, in && a.m_first_valid[g]);
  } else if (a.o_first) {
    bool v = false;
    if (in) {
      const uint32_t row = a.r.first_row[g];
      copy_elem(a.o_first, g, a.vals, row, a.vw);
      v = a.vvalid ? bit_at(a.vvalid, a.voff + row) : true;
    }
    store_valid_bit(a.o_first_valid, g, a.G, v);
  }
  if (a.o_last && a.m_last_val) {
    if (in) store_raw(a.o_last, g, a.m_last_val[g], a.vw);
    store_valid_bit(a.o_last_valid, g, a.G, in && a.m_last_valid[g]);
  } else if (a.o_last) {
    bool v = false;
    if (in) {
      const uint32_t row = a.r.last_row[g];
      copy_elem(a.o_last, g, a.vals, row, a.vw);
      v = a.vvalid ? bit_at(a.vvalid, a.voff + row) : true;
    }
    store_valid_bit(a.o_last_valid, g, a.G, v);
  }
}

// Whole-column first / last with skip_nulls (arrow's scalar `first` / `last`, NDFrame::first/last, ndframe.cpp:129,160):
// lowest / highest VALID row.  range[0] = min valid row (init kNoRow), range[1] = max valid row + 1 (init 0).
__global__ void __launch_bounds__(256) k_valid_row_range(const uint8_t* valid, int64_t bit_off, int64_t n, uint32_t* range) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  uint32_t lo = kNoRow, hi = 0;
  for (; i < n; i += stride) {
    if (bit_at(valid, bit_off + i)) {
      const uint32_t r = static_cast<uint32_t>(i);
      lo = r < lo ? r : lo;
      hi = r + 1 > hi ? r + 1 : hi;
    }
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    const uint32_t ol = __shfl_xor_sync(0xFFFFFFFFu, lo, d), oh = __shfl_xor_sync(0xFFFFFFFFu, hi, d);
    lo = ol < lo ? ol : lo;
    hi = oh > hi ? oh : hi;
  }
  if (lane_id() == 0) {
    if (lo != kNoRow) atomicMin(range, lo);
    if (hi) atomicMax(range + 1, hi);
  }
}
// Points the single group's first / last row at the valid range (no valid row: keeps the positional rows, whose
// values are null anyway).
__global__ void k_apply_row_range(const uint32_t* range, uint32_t* first_row, uint32_t* last_row) {
  if (range[0] != kNoRow) { first_row[0] = range[0]; last_row[0] = range[1] - 1u; }
}

// Arrow boolean values (bit-packed, `bit_off` < 8 bits into the first byte) -> one byte per value.
__global__ void __launch_bounds__(256) k_unpack_bool(const uint8_t* bits, int64_t bit_off, int64_t n, uint8_t* out) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) out[i] = bit_at(bits, bit_off + i) ? 1 : 0;
}

// all / any per group from the min / max of the unpacked bytes; bit-packed values + validity (whole warps).
__global__ void __launch_bounds__(256) k_emit_bool(GroupResult r, uint32_t G, uint32_t* o_all, uint32_t* o_all_valid,
                                                   uint32_t* o_any, uint32_t* o_any_valid) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in = g < G;
  const uint64_t cnt = in ? (r.count64 ? r.count64[g] : static_cast<uint64_t>(r.count[g])) : 0;
  const bool has = cnt > 0;
  const uint32_t mv = __ballot_sync(0xFFFFFFFFu, has);
  const uint32_t ma = __ballot_sync(0xFFFFFFFFu, has && o_all && r.min_ord[g] != 0);
  const uint32_t my = __ballot_sync(0xFFFFFFFFu, has && o_any && r.max_ord[g] != 0);
  if (lane_id() == 0 && (g & ~31u) < G) {
    if (o_all) { o_all[g >> 5] = ma; o_all_valid[g >> 5] = mv; }
    if (o_any) { o_any[g >> 5] = my; o_any_valid[g >> 5] = mv; }
  }
}

// ---- composite keys: field j occupies bits [shift_j, shift_j + bits_j) of the packed key; a
// nullable field has one extra top bit that is set (value bits zero) for null. ----
constexpr int kMaxKeyCols = 4;
struct KeyPackArgs {
  const void* col[kMaxKeyCols];
  const uint8_t* valid[kMaxKeyCols];
  int64_t off[kMaxKeyCols];
  int width[kMaxKeyCols];     // bytes: 4 or 8
  int bits[kMaxKeyCols];      // value bits (without the null bit)
  int shift[kMaxKeyCols];
  int nullable[kMaxKeyCols];
  int n_cols;
  int64_t n;
  uint64_t* out;
};

__global__ void __launch_bounds__(256) k_pack_keys(KeyPackArgs a) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (; i < a.n; i += stride) {
    uint64_t k = 0;
    for (int j = 0; j < a.n_cols; ++j) {
      uint64_t v = a.width[j] == 8 ? static_cast<const uint64_t*>(a.col[j])[i]
                                   : static_cast<uint64_t>(static_cast<const uint32_t*>(a.col[j])[i]);
      const uint64_t mask = a.bits[j] >= 64 ? ~0ull : ((1ull << a.bits[j]) - 1ull);
      v &= mask;
      if (a.nullable[j] && a.valid[j] && !bit_at(a.valid[j], a.off[j] + i)) v = 1ull << a.bits[j];
      k |= v << a.shift[j];
    }
    a.out[i] = k;
  }
}

// unique-key emit: unpack column j of the packed key
struct KeyEmitArgs {
  const uint64_t* key;
  const uint8_t* key_kind;
  uint32_t G;
  int width, bits, shift, nullable, packed;   // packed == 0: single key stored verbatim
  void* out;
  uint32_t* out_valid;
};

__global__ void __launch_bounds__(256) k_emit_key(KeyEmitArgs a) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in = g < a.G;
  bool valid = in;
  if (in) {
    uint64_t k = a.key[g];
    if (a.packed) {
      k >>= a.shift;
      if (a.nullable && ((k >> a.bits) & 1ull)) valid = false;
      k &= a.bits >= 64 ? ~0ull : ((1ull << a.bits) - 1ull);
    } else if (a.key_kind[g] == KK_NULL) {
      valid = false;
      k = 0;
    }
    if (a.width == 8) static_cast<uint64_t*>(a.out)[g] = k;
    else static_cast<uint32_t*>(a.out)[g] = static_cast<uint32_t>(k);
  }
  store_valid_bit(a.out_valid, g, a.G, valid);
}

// ---- synthetic generator: counter-based, identical on host and device ----
__global__ void __launch_bounds__(256) k_synth_keys(int64_t* out, int64_t n, int64_t first_row, uint64_t G, uint64_t seed) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) out[i] = static_cast<int64_t>(splitmix64(static_cast<uint64_t>(i + first_row) ^ seed) % G);
}
__global__ void __launch_bounds__(256) k_synth_vals(double* out, int64_t n, int64_t first_row, uint64_t seed) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride)
    out[i] = static_cast<double>(splitmix64(static_cast<uint64_t>(i + first_row) + seed) >> 11) * 0x1.0p-53;
}
// one thread per output byte; bit = 1 (valid) unless splitmix64(i + seed) % null_every == 0
__global__ void __launch_bounds__(256) k_synth_validity(uint8_t* out, int64_t n, int64_t first_row, uint64_t seed,
                                                        uint32_t null_every) {
  const int64_t nbytes = (n + 7) / 8;
  int64_t b = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (; b < nbytes; b += stride) {
    uint8_t v = 0;
    for (int j = 0; j < 8; ++j) {
      const int64_t i = b * 8 + j;
      if (i < n && splitmix64(static_cast<uint64_t>(i + first_row) + seed) % null_every != 0) v |= 1u << j;
    }
    out[b] = v;
  }
}
__global__ void __launch_bounds__(256) k_synth_ts(int64_t* out, int64_t n, int64_t first_row, int64_t t0, int64_t step,
                                                  uint64_t seed) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) {
    const int64_t r = i + first_row;
    out[i] = t0 + r * step + static_cast<int64_t>(splitmix64(static_cast<uint64_t>(r) + seed) % static_cast<uint64_t>(step));
  }
}

}  // namespace pa
