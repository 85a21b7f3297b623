// Shared device helpers for the group-by kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pa {

// ---- aggregate mask bits (mirror include/pa_b200.h) ----
constexpr uint32_t AGG_SUM = 1u, AGG_MEAN = 2u, AGG_COUNT = 4u, AGG_MIN = 8u, AGG_MAX = 16u,
                   AGG_FIRST = 32u, AGG_LAST = 64u,
                   AGG_PRODUCT = 128u, AGG_VARIANCE = 256u, AGG_STDDEV = 512u;   // second-stage aggregates (stage2.cuh)

// Value classes: how a value column is widened for accumulation.
//   VC_F: float/double -> double sum, min/max on double
//   VC_I: signed ints  -> int64 wrapping sum (+ double sum for mean), min/max on int64
//   VC_U: unsigned     -> uint64 wrapping sum (+ double sum for mean), min/max on uint64
enum ValClass : int { VC_F = 0, VC_I = 1, VC_U = 2 };

// Table sentinel.  A real key equal to it is routed to a dedicated group (SPECIAL_EMPTYKEY) so
// that every 64-bit key value is representable; null keys go to SPECIAL_NULL.
constexpr uint64_t kEmptyKey = 0x9E3779B97F4A7C15ull;
constexpr uint32_t kNoRow = 0xFFFFFFFFu;
constexpr uint64_t kMinInit = 0xFFFFFFFFFFFFFFFFull;  // order-mapped accumulators
constexpr uint64_t kMaxInit = 0ull;

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// Table hash: two 32-bit multiplies, top bits taken by the caller.
__device__ __forceinline__ uint32_t hash_key(uint64_t k) {
  uint32_t lo = static_cast<uint32_t>(k), hi = static_cast<uint32_t>(k >> 32);
  uint32_t h = (lo * 0x9E3779B1u) ^ (hi * 0x85EBCA77u);
  h ^= h >> 15;
  return h * 0x2C1B3C6Du;
}
__device__ __forceinline__ uint64_t hash_key64(uint64_t k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return k;
}

// ---- order-preserving maps to uint64 so that min/max become unsigned integer atomics ----
__device__ __forceinline__ uint64_t f64_to_ord(double d) {
  uint64_t b = static_cast<uint64_t>(__double_as_longlong(d));
  uint64_t m = (static_cast<uint64_t>(static_cast<int64_t>(b) >> 63)) | 0x8000000000000000ull;
  return b ^ m;
}
__host__ __device__ __forceinline__ uint64_t ord_to_f64_bits(uint64_t u) {
  uint64_t m = ((u >> 63) - 1ull) | 0x8000000000000000ull;
  return u ^ m;
}
__device__ __forceinline__ uint64_t i64_to_ord(int64_t v) { return static_cast<uint64_t>(v) ^ 0x8000000000000000ull; }

// Widened value: 64 raw bits + what to add to the double sum.
template <int VC>
struct Wide;
template <>
struct Wide<VC_F> {
  static __device__ __forceinline__ uint64_t ord(uint64_t bits) { return f64_to_ord(__longlong_as_double(bits)); }
  static __device__ __forceinline__ bool is_nan(uint64_t bits) {
    double d = __longlong_as_double(bits);
    return d != d;
  }
  static __device__ __forceinline__ double as_double(uint64_t bits) { return __longlong_as_double(bits); }
};
template <>
struct Wide<VC_I> {
  static __device__ __forceinline__ uint64_t ord(uint64_t bits) { return bits ^ 0x8000000000000000ull; }
  static __device__ __forceinline__ bool is_nan(uint64_t) { return false; }
  static __device__ __forceinline__ double as_double(uint64_t bits) { return static_cast<double>(static_cast<int64_t>(bits)); }
};
template <>
struct Wide<VC_U> {
  static __device__ __forceinline__ uint64_t ord(uint64_t bits) { return bits; }
  static __device__ __forceinline__ bool is_nan(uint64_t) { return false; }
  static __device__ __forceinline__ double as_double(uint64_t bits) { return static_cast<double>(bits); }
};

// Load element i of a value column of byte width W and class VC, widened to 64 bits.
template <int VC, int W>
__device__ __forceinline__ uint64_t load_wide(const void* p, int64_t i) {
  if constexpr (VC == VC_F) {
    if constexpr (W == 8) return static_cast<const uint64_t*>(p)[i];
    else return static_cast<uint64_t>(__double_as_longlong(static_cast<double>(static_cast<const float*>(p)[i])));
  } else if constexpr (VC == VC_I) {
    if constexpr (W == 8) return static_cast<const uint64_t*>(p)[i];
    else if constexpr (W == 4) return static_cast<uint64_t>(static_cast<int64_t>(static_cast<const int32_t*>(p)[i]));
    else if constexpr (W == 2) return static_cast<uint64_t>(static_cast<int64_t>(static_cast<const int16_t*>(p)[i]));
    else return static_cast<uint64_t>(static_cast<int64_t>(static_cast<const int8_t*>(p)[i]));
  } else {
    if constexpr (W == 8) return static_cast<const uint64_t*>(p)[i];
    else if constexpr (W == 4) return static_cast<const uint32_t*>(p)[i];
    else if constexpr (W == 2) return static_cast<const uint16_t*>(p)[i];
    else return static_cast<const uint8_t*>(p)[i];
  }
}

template <int W>
__device__ __forceinline__ uint64_t load_key(const void* p, int64_t i) {
  if constexpr (W == 8) return static_cast<const uint64_t*>(p)[i];
  else return static_cast<uint64_t>(static_cast<const uint32_t*>(p)[i]);  // int32/uint32: zero-extended bit pattern
}

__device__ __forceinline__ bool bit_at(const uint8_t* bitmap, int64_t i) {
  return (bitmap[i >> 3] >> (i & 7)) & 1;
}

// ---- mbarrier + bulk async copy (TMA engine, 1-D) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// 128-bit / 8-bit shared-memory accesses on 32-bit shared addresses (nvcc otherwise splits a uint4
// store whose components are not already in an aligned register quad into STS.64 + 2 x STS.32).
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void sts8(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t lds64(uint32_t addr) {
  uint64_t v;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, uint64_t v) {
  asm volatile("st.shared.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}

// streaming 128-bit global load that does not allocate in L1
__device__ __forceinline__ ulonglong2 ldg_stream_u64x2(const void* p) {
  ulonglong2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
  return r;
}

// streaming 256-bit global load, no L1 allocation, evict-first in L2 (LDG.E.NA.EFL2.256): input rows are
// read exactly once and must not push the group table out of L2.  32-byte aligned address.
struct u64x4 { unsigned long long a, b, c, d; };
__device__ __forceinline__ u64x4 ldg_stream_u64x4(const void* p) {
  u64x4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.u64 {%0, %1, %2, %3}, [%4];"
               : "=l"(r.a), "=l"(r.b), "=l"(r.c), "=l"(r.d) : "l"(p));
  return r;
}

}  // namespace pa
