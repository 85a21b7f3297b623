// Micro-benchmark: shared-memory atomic accumulation on random slots (CTA-shared accumulators).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_atomics smem_atomics.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__device__ __forceinline__ bool cas128(uint32_t addr, ulonglong2 expected, ulonglong2 desired, ulonglong2* old) {
  unsigned long long o0, o1;
  asm volatile(
      "{\n\t.reg .b128 e, d, o;\n\t"
      "mov.b128 e, {%2, %3};\n\t"
      "mov.b128 d, {%4, %5};\n\t"
      "atom.shared.cas.b128 o, [%6], e, d;\n\t"
      "mov.b128 {%0, %1}, o;\n\t}"
      : "=l"(o0), "=l"(o1)
      : "l"(expected.x), "l"(expected.y), "l"(desired.x), "l"(desired.y), "r"(addr)
      : "memory");
  old->x = o0; old->y = o1;
  return o0 == expected.x && o1 == expected.y;
}

// MODE 0: atomicAdd u32 | 1: f64 CAS loop + atomicAdd u32 | 2: 128-bit CAS loop {f64 sum, u32 count, u32 pad}
template <int MODE>
__global__ void k(int S, int64_t n, double* out) {
  extern __shared__ __align__(16) unsigned char sm[];
  ulonglong2* t = reinterpret_cast<ulonglong2*>(sm);
  for (int i = threadIdx.x; i < S; i += blockDim.x) t[i] = make_ulonglong2(0ull, 0ull);
  __syncthreads();
  const uint32_t base = static_cast<uint32_t>(__cvta_generic_to_shared(sm));
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const uint64_t h = splitmix64((uint64_t)i ^ 42);
    const uint32_t s = (uint32_t)(h % (uint64_t)S);
    const double v = (double)(h >> 11) * 0x1.0p-53;
    if (MODE == 0) atomicAdd(reinterpret_cast<unsigned int*>(&t[s].y), 1u);
    if (MODE == 1) {
      atomicAdd(reinterpret_cast<double*>(&t[s].x), v);
      atomicAdd(reinterpret_cast<unsigned int*>(&t[s].y), 1u);
    }
    if (MODE == 2) {
      ulonglong2 cur = t[s];
      for (;;) {
        ulonglong2 want;
        want.x = (unsigned long long)__double_as_longlong(__longlong_as_double((long long)cur.x) + v);
        want.y = cur.y + 1ull;
        ulonglong2 old;
        if (cas128(base + s * 16u, cur, want, &old)) break;
        cur = old;
      }
    }
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = __longlong_as_double((long long)t[0].x) + (double)t[0].y;
}

template <int MODE>
float run(int S, int64_t n, int threads, double* out) {
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, S * 16);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<148, threads, S * 16>>>(S, n / 8, out);
  cudaEventRecord(a);
  k<MODE><<<148, threads, S * 16>>>(S, n, out);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
  return ms;
}

int main() {
  const int64_t n = 1ll << 28;
  double* out; cudaMalloc(&out, 8);
  printf("%8s %8s %10s %10s %10s  (Grows/s)\n", "slots", "threads", "add_u32", "f64+u32", "cas128");
  for (int S : {1024, 4096, 12288})
    for (int th : {512, 1024}) {
      float a = run<0>(S, n, th, out), b = run<1>(S, n, th, out), c = run<2>(S, n, th, out);
      printf("%8d %8d %10.2f %10.2f %10.2f\n", S, th, n / a / 1e6, n / b / 1e6, n / c / 1e6);
      fflush(stdout);
    }
  return 0;
}
