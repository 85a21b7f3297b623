// Second-stage aggregates: product, variance, stddev.
//
// Reference: GROUPBY_AGG(product) and GROUPBY_NUMERIC_AGG(variance | stddev, double)
// (/root/reference/src/dataframe.cpp:1516-1536, pd_core_macros.h:5-147): one arrow::compute call per
// group with default options — Product: wrapping (u)int64 for integers, double for floats, nulls
// skipped, min_count 1; Variance/Stddev: ddof 0, TWO PASSES per group (mean first, then the sum of the
// squared deviations from it), nulls skipped.
//
// The same two passes here: the first pass is the ordinary fused scan (sum / mean / count per group,
// whichever path fits the key cardinality); this file is the second pass.  It looks every row's key up
// in a read-only key -> group table built from the finished GroupResult (rowids.cuh), subtracts the
// group mean and accumulates (x - mean)^2 — and the running product — per group:
//   G <= S2_SMEM_SLOTS             CTA-shared accumulators (shared-memory atomics; the product is a CAS loop),
//                                  replicated R = 2^rlog times (slot = id * R + lane % R, R <= 32 and
//                                  G * R <= S2_REPL_SLOTS) so that the lanes of a warp hitting one
//                                  group do not serialise on one address; the replicas are folded in slot
//                                  order and flushed with one global update per group and CTA
//   more                           L2 atomics straight into the per-group arrays
// m2 is a sum of non-negative terms (condition number 1), so its value is insensitive to the order of
// summation to ~n eps; the product of floats is order dependent in the last bits like every
// floating-point product and is not reproducible run to run; integer products are exact (wrapping).
#pragma once
#include "rowids.cuh"

namespace pa {

constexpr int S2_THREADS = 512;
constexpr uint32_t S2_SMEM_SLOTS = 4096;     // x 16 bytes of shared memory
constexpr uint32_t S2_REPL_SLOTS = 1024;     // replicate the accumulators only up to this many slots

struct Stage2Args {
  RowIdArgs ids;              // lookup table + key column (`out` unused)
  const void* vals;
  const uint8_t* vvalid;
  int64_t voff;
  int vw;
  const double* mean;         // [G]
  double* m2;                 // [G] zero-initialised, or null
  unsigned long long* prod;   // [G] initialised to one, or null
  uint32_t use_smem;
  int rlog;                   // shared-memory accumulators: log2 of the replication
};

template <int VC>
__device__ __forceinline__ unsigned long long s2_one() {
  if constexpr (VC == VC_F) return static_cast<unsigned long long>(__double_as_longlong(1.0));
  else return 1ull;
}

template <int VC>
__device__ __forceinline__ unsigned long long s2_mul(unsigned long long a, unsigned long long b) {
  if constexpr (VC == VC_F) return static_cast<unsigned long long>(__double_as_longlong(__longlong_as_double(static_cast<long long>(a)) * __longlong_as_double(static_cast<long long>(b))));
  else return a * b;   // wraps, as arrow's MultiplyTraits does through unsigned arithmetic
}

template <int VC>
__device__ __forceinline__ void s2_atomic_mul(unsigned long long* p, unsigned long long v) {
  unsigned long long old = *reinterpret_cast<volatile unsigned long long*>(p);
  for (;;) {
    const unsigned long long seen = atomicCAS(p, old, s2_mul<VC>(old, v));
    if (seen == old) return;
    old = seen;
  }
}

// mean[g] = sum / count as the reference's second pass sees it (double); accumulators reset
__global__ void __launch_bounds__(256) k_stage2_init(GroupResult r, uint32_t G, int vc, double* mean, double* m2, unsigned long long* prod) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const uint64_t cnt = r.count64 ? r.count64[g] : static_cast<uint64_t>(r.count[g]);
  double s = 0.0;
  if (cnt) s = (vc == VC_F || !r.dsum) ? __longlong_as_double(static_cast<long long>(r.sum[g])) : r.dsum[g];
  mean[g] = cnt ? s / static_cast<double>(cnt) : 0.0;
  if (m2) m2[g] = 0.0;
  if (prod) prod[g] = vc == VC_F ? static_cast<unsigned long long>(__double_as_longlong(1.0)) : 1ull;
}

template <int VC>
__global__ void __launch_bounds__(S2_THREADS) k_stage2(Stage2Args a) {
  extern __shared__ __align__(16) unsigned char s2_smem[];
  const uint32_t G = a.ids.G;
  const uint32_t nslots = G << a.rlog;
  double* s_m2 = reinterpret_cast<double*>(s2_smem);
  unsigned long long* s_prod = reinterpret_cast<unsigned long long*>(s2_smem + sizeof(double) * nslots);
  const unsigned long long one = s2_one<VC>();
  if (a.use_smem) {
    for (uint32_t i = threadIdx.x; i < nslots; i += S2_THREADS) { s_m2[i] = 0.0; s_prod[i] = one; }
    __syncthreads();
  }
  const uint32_t rep = lane_id() & ((1u << a.rlog) - 1u);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * S2_THREADS;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(S2_THREADS) + threadIdx.x; i < a.ids.n; i += stride) {
    const uint32_t id = rowid_lookup(a.ids, i);
    if (id == 0xFFFFFFFFu || (a.vvalid && !bit_at(a.vvalid, a.voff + i))) continue;
    const uint64_t bits = load_wide_rt<VC>(a.vals, i, a.vw);
    const double d = Wide<VC>::as_double(bits) - __ldg(a.mean + id);
    if (a.use_smem) {
      const uint32_t s = (id << a.rlog) | rep;
      if (a.m2) atomicAdd(s_m2 + s, d * d);
      if (a.prod) s2_atomic_mul<VC>(s_prod + s, bits);
    } else {
      if (a.m2) atomicAdd(a.m2 + id, d * d);
      if (a.prod) s2_atomic_mul<VC>(a.prod + id, bits);
    }
  }
  if (a.use_smem) {
    __syncthreads();
    const uint32_t R = 1u << a.rlog;
    for (uint32_t g = threadIdx.x; g < G; g += S2_THREADS) {
      double m = 0.0;
      unsigned long long p = one;
      for (uint32_t r = 0; r < R; ++r) {
        m += s_m2[(g << a.rlog) | r];
        p = s2_mul<VC>(p, s_prod[(g << a.rlog) | r]);
      }
      if (a.m2 && m != 0.0) atomicAdd(a.m2 + g, m);
      if (a.prod && p != one) s2_atomic_mul<VC>(a.prod + g, p);
    }
  }
}

struct Stage2Emit {
  GroupResult r;
  uint32_t G;
  int vc;
  const double* m2;
  const unsigned long long* prod;
  unsigned long long* o_prod; uint32_t* o_prod_valid;
  double* o_var; uint32_t* o_var_valid;
  double* o_std; uint32_t* o_std_valid;
};

// grid: ceil(G/256) blocks of 256 (whole warps: the validity words are ballots)
__global__ void __launch_bounds__(256) k_stage2_emit(Stage2Emit a) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in = g < a.G;
  const uint64_t cnt = in ? (a.r.count64 ? a.r.count64[g] : static_cast<uint64_t>(a.r.count[g])) : 0;
  const bool has = cnt > 0;
  const uint32_t m = __ballot_sync(0xFFFFFFFFu, has);
  const bool writer = lane_id() == 0 && (g & ~31u) < a.G;
  const double var = has && a.m2 ? a.m2[g] / static_cast<double>(cnt) : 0.0;
  if (a.o_prod) {
    if (in) a.o_prod[g] = has ? a.prod[g] : 0ull;
    if (writer) a.o_prod_valid[g >> 5] = m;
  }
  if (a.o_var) {
    if (in) a.o_var[g] = var;
    if (writer) a.o_var_valid[g >> 5] = m;
  }
  if (a.o_std) {
    if (in) a.o_std[g] = sqrt(var);
    if (writer) a.o_std_valid[g >> 5] = m;
  }
}

}  // namespace pa
