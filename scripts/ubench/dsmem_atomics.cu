// Micro-benchmark for the next round: can a group table distributed over the shared memory
// of a thread-block cluster (DSMEM) beat L2 atomics for 64 K - 128 K groups?  Every CTA of a cluster owns
// SLOTS slots {double sum, u32 count}; a row's owner CTA and slot come from its hashed key; the update goes to
// the owner's shared memory through cluster.map_shared_rank (remote shared atomics).  Not product code.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_atomics dsmem_atomics.cu
// run:   ./dsmem_atomics [rows]          prints G updates/s for cluster sizes 1 (local only), 2, 4, 8, 16
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
namespace cg = cooperative_groups;

__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

constexpr int SLOTS = 8192;          // per CTA: 8192 x 12 B = 96 KB
constexpr int THREADS = 1024;

// MODE 0: count only (native u32 shared atomic)   MODE 1: count + f64 sum (the compiler's CAS loop for the sum)
template <int MODE>
__global__ void __launch_bounds__(THREADS, 1) k(int64_t n, unsigned long long* sink) {
  extern __shared__ __align__(16) unsigned char smem[];
  double* sum = reinterpret_cast<double*>(smem);
  unsigned* cnt = reinterpret_cast<unsigned*>(smem + sizeof(double) * SLOTS);
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned csize = cluster.num_blocks();
  for (int i = threadIdx.x; i < SLOTS; i += THREADS) { sum[i] = 0.0; cnt[i] = 0u; }
  cluster.sync();
  const int64_t per_cluster = n / (gridDim.x / csize);
  const int64_t c0 = (blockIdx.x / csize) * per_cluster;
  for (int64_t i = c0 + cluster.block_rank() * THREADS + threadIdx.x; i < c0 + per_cluster; i += (int64_t)csize * THREADS) {
    const uint64_t h = splitmix64((uint64_t)i ^ 42);
    const unsigned owner = (unsigned)(h >> 40) % csize;
    const unsigned slot = (unsigned)h % SLOTS;
    unsigned* rc = cluster.map_shared_rank(cnt, owner);
    atomicAdd(rc + slot, 1u);
    if (MODE == 1) {
      double* rs = cluster.map_shared_rank(sum, owner);
      atomicAdd(rs + slot, (double)(h >> 11) * 0x1.0p-53);
    }
  }
  cluster.sync();
  unsigned long long acc = 0;
  for (int i = threadIdx.x; i < SLOTS; i += THREADS) acc += cnt[i];
  if (acc == 0x123456789ull) *sink = acc;   // keep the table alive
}

template <int MODE>
static double run(int csize, int64_t n, unsigned long long* sink) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t smem = (sizeof(double) + sizeof(unsigned)) * SLOTS;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (csize > 8) cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((sms / csize) * csize);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int it = 0; it < 4; ++it) {
    cudaEventRecord(e0);
    cudaError_t err = cudaLaunchKernelEx(&cfg, k<MODE>, n, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    if (err != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("  cluster %d: launch failed (%s)\n", csize, cudaGetErrorString(err)); return 0; }
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    if (it && ms < best) best = ms;
  }
  return (double)n / best / 1e6;   // G rows/s
}

int main(int argc, char** argv) {
  const int64_t n = argc > 1 ? atoll(argv[1]) : 1000000000ll;
  unsigned long long* sink; cudaMalloc(&sink, 8);
  printf("rows %lld, %d slots per CTA (groups held by a cluster = cluster size x %d)\n", (long long)n, SLOTS, SLOTS);
  for (int cs : {1, 2, 4, 8, 16}) {
    const double a = run<0>(cs, n, sink), b = run<1>(cs, n, sink);
    printf("cluster %2d (%6d groups): count only %7.1f G rows/s   count + f64 sum %7.1f G rows/s\n", cs, cs * SLOTS, a, b);
  }
  return 0;
}
