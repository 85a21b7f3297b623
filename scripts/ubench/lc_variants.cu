// Micro-benchmark behind the round-2 redesign of the shared-memory group-by scan (lowcard.cuh):
// the streaming loader + warp-private accumulators of the real kernel, with the pieces that cost
// shared-memory wavefronts made switchable.
//   DET 0: claim tag (STS.8 + LDS.32 read-back)      -- the round-1 kernel
//   DET 1: MATCH.ANY on the slot id, leaders do the read-modify-write
//   DET 2: no duplicate detection (WRONG sums; lower bound of the accumulate traffic)
//   DET 3: loads only (HBM bound)
//   LOOK 0: id = key - base (dense)
//   LOOK 1: bucket of two 8-byte keys (LDS.128) + LDS.U16 id      -- the round-1 hash mode
//   LOOK 2: one packed 8-byte entry {52 key bits | 2 displacement | 10 id} (LDS.64)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o lc_variants lc_variants.cu
// run:   ./lc_variants [rows_millions] [groups]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__host__ __device__ inline uint64_t mix_bij(uint64_t k) { return (k ^ (k >> 32)) * 0x9E3779B97F4A7C15ull; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint32_t lds16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint64_t lds64(uint32_t a) { uint64_t v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts64(uint32_t a, uint64_t v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ ulonglong2 lds128(uint32_t a) {
  ulonglong2 v; asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"(a) : "memory"); return v;
}
__device__ __forceinline__ uint64_t ldg_stream(const void* p) {
  uint64_t r; asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(r) : "l"(p)); return r;
}

constexpr int NB = 8;
constexpr int GP = 1026;
constexpr uint32_t FULL = 0xFFFFFFFFu;

struct Args {
  const uint64_t* keys; const double* vals; int64_t n; uint64_t base;
  const uint64_t* tab1_keys; const uint16_t* tab1_ids;   // LOOK 1: 2048 buckets x 2
  const uint64_t* tab2;                                  // LOOK 2: 4096 packed entries
  double* out_sum; unsigned long long* out_cnt;          // [GP] accumulated with global atomics at the end
};

template <int DET, int LOOK, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k(Args a) {
  extern __shared__ __align__(16) unsigned char smem[];
  // layout: [table][per-warp: sum 8 x GP | cw 4 x GP]
  constexpr uint32_t TAB_BYTES = LOOK == 1 ? (4096 * 8 + 4096 * 2) : (LOOK == 2 ? 4096 * 8 : 0);
  constexpr uint32_t PER_WARP = GP * 12 + 8;
  const int warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (LOOK == 1) {
    uint64_t* tk = reinterpret_cast<uint64_t*>(smem);
    uint16_t* ti = reinterpret_cast<uint16_t*>(smem + 4096 * 8);
    for (int i = threadIdx.x; i < 4096; i += WARPS * 32) { tk[i] = a.tab1_keys[i]; ti[i] = a.tab1_ids[i]; }
  }
  if (LOOK == 2) {
    uint64_t* t = reinterpret_cast<uint64_t*>(smem);
    for (int i = threadIdx.x; i < 4096; i += WARPS * 32) t[i] = a.tab2[i];
  }
  unsigned char* mine = smem + TAB_BYTES + PER_WARP * warp;
  for (int i = lane; i < GP; i += 32) {
    reinterpret_cast<uint64_t*>(mine)[i] = 0;
    reinterpret_cast<uint32_t*>(mine + GP * 8)[i] = 0;
  }
  __syncthreads();
  const uint32_t s_sum = smem_u32(mine), s_cw = s_sum + GP * 8, s_tab = smem_u32(smem);

  const int64_t n_grp = a.n / 256;
  const int64_t gw = (int64_t)blockIdx.x * WARPS + warp, nw = (int64_t)gridDim.x * WARPS;
  uint64_t ck[NB], cv[NB], nk[NB], nv[NB];
  int64_t g = gw;
  if (g < n_grp) {
#pragma unroll
    for (int e = 0; e < NB; ++e) ck[e] = ldg_stream(a.keys + g * 256 + e * 32 + lane);
#pragma unroll
    for (int e = 0; e < NB; ++e) cv[e] = ldg_stream(a.vals + g * 256 + e * 32 + lane);
  }
  uint64_t sink = 0;
  while (g < n_grp) {
    const int64_t gn = g + nw;
    if (gn < n_grp) {
#pragma unroll
      for (int e = 0; e < NB; ++e) nk[e] = ldg_stream(a.keys + gn * 256 + e * 32 + lane);
#pragma unroll
      for (int e = 0; e < NB; ++e) nv[e] = ldg_stream(a.vals + gn * 256 + e * 32 + lane);
    }
    uint32_t id[NB];
    if (DET == 3) {
#pragma unroll
      for (int e = 0; e < NB; ++e) sink += ck[e] ^ cv[e];
    } else {
      // ---- lookup ----
#pragma unroll
      for (int e = 0; e < NB; ++e) {
        if (LOOK == 0) {
          id[e] = static_cast<uint32_t>(ck[e] - a.base);
        } else if (LOOK == 1) {
          const uint32_t lo = (uint32_t)ck[e], hi = (uint32_t)(ck[e] >> 32);
          uint32_t b = ((lo * 0x9E3779B1u) ^ (hi * 0x85EBCA77u) ^ (lo >> 15)) * 0x2C1B3C6Du >> 21;
          id[e] = 0xFFFF;
          for (int p = 0; p < 8 && id[e] == 0xFFFF; ++p) {
            const ulonglong2 kk = lds128(s_tab + b * 16);
            const bool h1 = kk.y == ck[e];
            if (kk.x == ck[e] || h1) id[e] = lds16(s_tab + 4096 * 8 + (2 * b + (h1 ? 1 : 0)) * 2);
            b = (b + 1) & 2047;
          }
        } else {
          const uint64_t m = mix_bij(ck[e]);
          const uint32_t home = (uint32_t)(m >> 52);
          const uint64_t want = m << 12;                 // low 52 bits in the top of the word
          id[e] = 0xFFFF;
          for (uint32_t d = 0; d < 4 && id[e] == 0xFFFF; ++d) {
            const uint64_t en = lds64(s_tab + ((home + d) & 4095) * 8);
            if ((en & ~0xFFFull) == want && ((en >> 10) & 3) == d) id[e] = (uint32_t)en & 1023u;
          }
        }
      }
      // ---- accumulate ----
#pragma unroll
      for (int e = 0; e < NB; ++e) {
        const uint32_t sa = s_sum + id[e] * 8, ca = s_cw + id[e] * 4;
        const double v = __longlong_as_double((long long)cv[e]);
        if (DET == 0) {
          sts8(ca + 3, lane);
          __syncwarp();
          const uint32_t cw = lds32(ca);
          uint64_t s = lds64(sa);
          const uint32_t tag = cw >> 24;
          const bool win = tag == lane;
          const uint32_t losers = __ballot_sync(FULL, !win);
          double add = v; uint32_t cnt = 1;
          if (losers) {
            // ordered fold through the winners (simplified general path: every loser in turn)
            uint32_t rem = losers;
            while (rem) {
              const int L = __ffs(rem) - 1; rem &= rem - 1;
              const uint32_t tw = __shfl_sync(FULL, tag, L);
              const double ov = __shfl_sync(FULL, v, L);
              if (lane == tw) { add += ov; cnt += 1; }
            }
          }
          if (win) {
            sts64(sa, (uint64_t)__double_as_longlong(__longlong_as_double((long long)s) + add));
            sts32(ca, (cw & 0xFFFFFFu) + cnt);
          }
          __syncwarp();
        } else if (DET == 1) {
          const uint32_t peers = __match_any_sync(FULL, id[e]);
          const uint32_t lanebit = 1u << lane;
          const bool leader = (peers & (lanebit - 1)) == 0;
          double add = v; uint32_t cnt = 1;
          if (__any_sync(FULL, peers != lanebit)) {
            uint32_t rem = leader ? (peers & ~lanebit) : 0u;
            while (__any_sync(FULL, rem != 0)) {
              const int src = rem ? (__ffs(rem) - 1) : (int)lane;
              const double ov = __shfl_sync(FULL, v, src);
              if (rem) { add += ov; cnt += 1; rem &= rem - 1; }
            }
          }
          if (leader) {
            const uint32_t cw = lds32(ca);
            const uint64_t s = lds64(sa);
            sts64(sa, (uint64_t)__double_as_longlong(__longlong_as_double((long long)s) + add));
            sts32(ca, cw + cnt);
          }
          __syncwarp();
        } else if (DET == 2) {
          const uint32_t cw = lds32(ca);
          const uint64_t s = lds64(sa);
          sts64(sa, (uint64_t)__double_as_longlong(__longlong_as_double((long long)s) + v));
          sts32(ca, cw + 1);
          __syncwarp();
        }
      }
    }
#pragma unroll
    for (int e = 0; e < NB; ++e) { ck[e] = nk[e]; cv[e] = nv[e]; }
    g = gn;
  }
  __syncthreads();
  if (DET == 3) { if (sink == 0x1234567) a.out_cnt[0] = 1; return; }
  for (int i = threadIdx.x; i < GP; i += WARPS * 32) {
    double s = 0; unsigned long long c = 0;
    for (int w = 0; w < WARPS; ++w) {
      const unsigned char* wa = smem + TAB_BYTES + PER_WARP * w;
      s += reinterpret_cast<const double*>(wa)[i];
      c += reinterpret_cast<const uint32_t*>(wa + GP * 8)[i] & 0xFFFFFFu;
    }
    if (c) { atomicAdd(a.out_sum + i, s); atomicAdd(a.out_cnt + i, c); }
  }
}

template <int DET, int LOOK, int WARPS>
void run(const char* name, Args a, int G, double want_sum) {
  constexpr uint32_t TAB_BYTES = LOOK == 1 ? (4096 * 8 + 4096 * 2) : (LOOK == 2 ? 4096 * 8 : 0);
  const size_t smem = TAB_BYTES + (size_t)(GP * 12 + 8) * WARPS;
  if (smem > 232448) { printf("%-34s smem %zu too large\n", name, smem); return; }
  auto kern = k<DET, LOOK, WARPS>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int it = 0; it < 4; ++it) {
    CK(cudaMemset(a.out_sum, 0, GP * 8)); CK(cudaMemset(a.out_cnt, 0, GP * 8));
    cudaEventRecord(e0);
    kern<<<148, WARPS * 32, smem>>>(a);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  std::vector<double> hs(GP); std::vector<unsigned long long> hc(GP);
  CK(cudaMemcpy(hs.data(), a.out_sum, GP * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hc.data(), a.out_cnt, GP * 8, cudaMemcpyDeviceToHost));
  double ts = 0; unsigned long long tc = 0;
  for (int i = 0; i < GP; ++i) { ts += hs[i]; tc += hc[i]; }
  const int64_t nproc = a.n / 256 * 256;
  printf("%-34s %7.3f ms  %6.1f Grows/s  %5.0f GB/s  count %s  sum rel err %.2e\n", name, best, nproc / best / 1e6, nproc * 16.0 / best / 1e6,
         (int64_t)tc == nproc ? "ok" : "WRONG", fabs(ts - want_sum) / want_sum);
}

__global__ void gen(uint64_t* keys, double* vals, int64_t n, uint64_t G, int scattered) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t g = splitmix64((uint64_t)i ^ 42) % G;
    keys[i] = scattered ? splitmix64(g * 7919 + 17) : g;
    vals[i] = (double)(splitmix64((uint64_t)i + 1337) >> 11) * 0x1.0p-53;
  }
}

int main(int argc, char** argv) {
  const int64_t n = (argc > 1 ? atoll(argv[1]) : 512) * 1000000ll;
  const int G = argc > 2 ? atoi(argv[2]) : 1000;
  uint64_t *keys; double* vals;
  CK(cudaMalloc(&keys, n * 8)); CK(cudaMalloc(&vals, n * 8));
  Args a{};
  CK(cudaMalloc(&a.out_sum, GP * 8)); CK(cudaMalloc(&a.out_cnt, GP * 8));
  // tables for the scattered key set
  std::vector<uint64_t> t1k(4096, ~0ull), t2(4096, ~0ull);
  std::vector<uint16_t> t1i(4096, 0xFFFF);
  int maxd = 0, over = 0;
  for (int g = 0; g < G; ++g) {
    const uint64_t key = splitmix64((uint64_t)g * 7919 + 17);
    const uint32_t lo = (uint32_t)key, hi = (uint32_t)(key >> 32);
    uint32_t b = ((lo * 0x9E3779B1u) ^ (hi * 0x85EBCA77u) ^ (lo >> 15)) * 0x2C1B3C6Du >> 21;
    for (;; b = (b + 1) & 2047) {
      if (t1k[2 * b] == ~0ull) { t1k[2 * b] = key; t1i[2 * b] = g; break; }
      if (t1k[2 * b + 1] == ~0ull) { t1k[2 * b + 1] = key; t1i[2 * b + 1] = g; break; }
    }
    const uint64_t m = mix_bij(key);
    const uint32_t home = (uint32_t)(m >> 52);
    int d = 0;
    while (d < 4 && t2[(home + d) & 4095] != ~0ull) ++d;
    if (d < 4) t2[(home + d) & 4095] = (m << 12) | ((uint64_t)d << 10) | (uint64_t)g; else ++over;
    if (d > maxd) maxd = d;
  }
  printf("rows %lld groups %d; packed table: max displacement %d, %d keys did not fit\n", (long long)n, G, maxd, over);
  uint64_t *d1k, *d2; uint16_t* d1i;
  CK(cudaMalloc(&d1k, 4096 * 8)); CK(cudaMalloc(&d2, 4096 * 8)); CK(cudaMalloc(&d1i, 4096 * 2));
  CK(cudaMemcpy(d1k, t1k.data(), 4096 * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d2, t2.data(), 4096 * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d1i, t1i.data(), 4096 * 2, cudaMemcpyHostToDevice));
  a.keys = keys; a.vals = vals; a.n = n; a.base = 0; a.tab1_keys = d1k; a.tab1_ids = d1i; a.tab2 = d2;
  const double want = (n / 256 * 256) * 0.5;

  gen<<<148 * 8, 256>>>(keys, vals, n, G, 0);
  CK(cudaDeviceSynchronize());
  printf("--- dense keys ---\n");
  run<3, 0, 16>("loads only, 16 warps", a, G, want);
  run<3, 0, 12>("loads only, 12 warps", a, G, want);
  run<2, 0, 16>("no detection (bound), 16 warps", a, G, want);
  run<0, 0, 16>("claim tag, 16 warps", a, G, want);
  run<1, 0, 16>("match.any, 16 warps", a, G, want);
  run<0, 0, 12>("claim tag, 12 warps", a, G, want);
  run<1, 0, 12>("match.any, 12 warps", a, G, want);
  gen<<<148 * 8, 256>>>(keys, vals, n, G, 1);
  CK(cudaDeviceSynchronize());
  printf("--- scattered keys ---\n");
  run<0, 1, 12>("bucket LDS.128+U16, tag, 12 warps", a, G, want);
  run<1, 1, 12>("bucket LDS.128+U16, match, 12 w", a, G, want);
  run<0, 2, 16>("packed LDS.64, tag, 16 warps", a, G, want);
  run<1, 2, 16>("packed LDS.64, match, 16 warps", a, G, want);
  run<0, 2, 14>("packed LDS.64, tag, 14 warps", a, G, want);
  run<1, 2, 14>("packed LDS.64, match, 14 warps", a, G, want);
  run<2, 2, 16>("packed LDS.64, no detection, 16 w", a, G, want);
  return 0;
}
