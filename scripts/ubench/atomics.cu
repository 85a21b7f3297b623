// Micro-benchmark: L2 atomic (RED/ATOM) throughput on random slots, as a function of table size.
// Decides the design of the global-table (high-cardinality) path.  Not product code.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atomics atomics.cu
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

struct Slot { unsigned long long key; unsigned first, last; double sum; unsigned count, pad; };

// MODE: 0 red f64 | 1 red f64 + red u32 (same sector) | 2 ld key + red f64 + red u32 | 3 atom f64 (returning)
//       4 red u32 | 5 red u64 | 6 ld key + ld first + red f64 + red u32 (what a lean scan does)
//       7 red f64 on a dense double[] (8 B stride)  | 8 plain ld of the slot only (no atomics)
template <int MODE>
__global__ void __launch_bounds__(256) k(Slot* t, double* dense, uint64_t G, int64_t n, unsigned long long* sink) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned long long acc = 0;
  for (; i < n; i += stride) {
    const uint64_t h = splitmix64((uint64_t)i ^ 42);
    const uint64_t s = h % G;
    const double v = (double)(h >> 11) * 0x1.0p-53;
    Slot* p = t + s;
    if (MODE == 0) atomicAdd(&p->sum, v);
    if (MODE == 1) { atomicAdd(&p->sum, v); atomicAdd(&p->count, 1u); }
    if (MODE == 2) { acc += __ldcg(&p->key); atomicAdd(&p->sum, v); atomicAdd(&p->count, 1u); }
    if (MODE == 3) acc += (unsigned long long)atomicAdd(&p->sum, v);
    if (MODE == 4) atomicAdd(&p->count, 1u);
    if (MODE == 5) atomicAdd((unsigned long long*)&p->sum, (unsigned long long)h);
    if (MODE == 6) {
      const ulonglong2 kf = __ldcg(reinterpret_cast<const ulonglong2*>(p));
      acc += kf.x;
      if ((unsigned)i < (unsigned)kf.y) atomicMin(&p->first, (unsigned)i);
      atomicAdd(&p->sum, v); atomicAdd(&p->count, 1u);
    }
    if (MODE == 7) atomicAdd(dense + s, v);
    if (MODE == 8) { const ulonglong2 kf = __ldcg(reinterpret_cast<const ulonglong2*>(p)); acc += kf.x + kf.y; }
  }
  if (acc == 0x1234567ull) *sink = acc;
}

template <int MODE>
float run(Slot* t, double* dense, uint64_t G, int64_t n, unsigned long long* sink, int grid) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<grid, 256>>>(t, dense, G, n / 8, sink);  // warm
  cudaEventRecord(a);
  k<MODE><<<grid, 256>>>(t, dense, G, n, sink);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
  return ms;
}

int main(int argc, char** argv) {
  const int64_t n = argc > 1 ? atoll(argv[1]) : (1ll << 28);
  const uint64_t Gs[] = {1000, 4096, 65536, 1u << 20, 1u << 24, 100000000ull};
  Slot* t; double* dense; unsigned long long* sink;
  cudaMalloc(&t, sizeof(Slot) * 100000000ull);
  cudaMalloc(&dense, 8 * 100000000ull);
  cudaMalloc(&sink, 8);
  cudaMemset(t, 0xFF, sizeof(Slot) * 100000000ull);
  cudaMemset(dense, 0, 8 * 100000000ull);
  const int grid = 148 * 8;
  printf("n=%lld rows; Grows/s per mode\n", (long long)n);
  printf("%10s %9s %9s %9s %9s %9s %9s %9s %9s %9s\n", "G", "redf64", "f64+u32", "ld+2red", "atomf64", "redu32", "redu64", "lean", "dense8B", "ldonly");
  for (uint64_t G : Gs) {
    float ms[9];
    ms[0] = run<0>(t, dense, G, n, sink, grid);
    ms[1] = run<1>(t, dense, G, n, sink, grid);
    ms[2] = run<2>(t, dense, G, n, sink, grid);
    ms[3] = run<3>(t, dense, G, n, sink, grid);
    ms[4] = run<4>(t, dense, G, n, sink, grid);
    ms[5] = run<5>(t, dense, G, n, sink, grid);
    ms[6] = run<6>(t, dense, G, n, sink, grid);
    ms[7] = run<7>(t, dense, G, n, sink, grid);
    ms[8] = run<8>(t, dense, G, n, sink, grid);
    printf("%10llu", (unsigned long long)G);
    for (int m = 0; m < 9; ++m) printf(" %9.2f", n / ms[m] / 1e6);
    printf("\n");
    fflush(stdout);
  }
  return 0;
}
