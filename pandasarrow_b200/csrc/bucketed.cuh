// Medium / high cardinality (more groups than one shared-memory table holds): radix-partition the rows into
// buckets whose distinct keys fit a CTA's shared-memory table, then aggregate bucket by bucket in shared memory.
// No global-memory atomics per row, no global hash table in the steady state: every key lives in exactly one bucket.
//
//   m = rp_mix(key)               bijective 64-bit mix; its TOP bits pick the bucket, the next bits the table slot
//   k_rp_hist<true>               rows per (chunk of 65 536 rows, one of 1024 fine buckets) + a HyperLogLog sketch of the
//                                 keys (4096 registers, fed by the 1/8 of the key space whose mix ends in 000) -> estimated
//                                 group count, from which the host picks B = log2(#buckets) (buckets of ~CAP/2 keys).
//                                 Keys only: computed ONCE per handle (a handle's keys never change) and reused by every
//                                 later aggregate on it, like the reference reuses its constructor's groupings.
//   k_rp_fold + scan + k_rp_ends  fine counts -> "flat" layout [bucket][chunk] -> ONE exclusive scan = the output position
//                                 of every chunk's first row of every bucket (buckets contiguous, chunks in row order)
//   k_rp_scatter                  per chunk: cursors loaded into shared memory and advanced privately; per tile of 4096
//                                 rows a shared-memory histogram ranks the rows, the tile is regrouped in shared memory
//                                 and written out as contiguous runs together with the ORIGINAL ROW NUMBERS (first / last
//                                 need them).  No global atomics, stable at tile granularity, same output every pass.
//                                 Cost grows with the fan-out (runs of 4096 / fan rows): 7.7 / 8.4 / 10.5 / 15 / 25 / 43 ms
//                                 per 1 B rows at 32 / 64 / 128 / 256 / 512 / 1024 ways (DRAM sees 128-byte pieces)
//   B <= 7 : one level.   B > 7 : two levels, ceil(B/2) bits then the rest inside every level-1 bucket
//   k_rp_cprefix, k_rp_hist<false>, scan, k_rp_ends, k_rp_scatter again, reading the level-1 output (retaken every pass:
//                                 inside a tile's run the level-1 scatter orders rows by atomic arrival)
//   k_bucket_agg                  CTA-shared open-addressing table in shared memory (8192 slots narrow / 4096 wide),
//                                 ATOMS.CAS.64 to claim a key, shared-memory atomics for sum / count / first / last / min /
//                                 max exactly like k_smemtab_scan.  Two modes:
//                                   whole buckets (>= 4 x #SM buckets): one CTA per bucket at a time (atomic work counter);
//                                     a finished bucket's groups are appended to the unordered result as sector-sized
//                                     records and the slots are reset in the same sweep;
//                                   ranged (fewer buckets): every CTA takes an equal, contiguous share of the partitioned
//                                     rows and flushes its table into the global table (gtable.cuh) at every bucket
//                                     boundary — a few L2 atomics per group and CTA, none per row — after which the
//                                     ordinary compaction / ordering of the global path runs.
//   k_bm_* + k_bm_rank + k_bk_gather  first-appearance order by bitmap rank (order.cuh): a 4-byte permutation is scattered,
//                                 then every output position fetches its group's record (one or two adjacent sectors)
//
// DRAM traffic per row (8-byte key + 8-byte value): scatter 16 + 20 B, aggregate 20 B = 56 B for one level (+ 8 B for
// the histogram on a handle's first pass); + 8 + 20 + 20 B for the second level = 104 B, against 16 B algorithmic: the
// floor of this design is 0.29 (one level) / 0.15 (two levels) of the 16 B/row roofline.  Only for 8-byte keys without
// a validity bitmap and 8-byte values (a nullable value column below 2^31 rows: the validity bit travels as RP_NULL_BIT
// of the row numbers); anything else, and any bucket that turns out to hold more keys than its table (ST_OVERFLOW), goes
// to the global-table path (gtable.cuh).
// Replaces Grouper::Consume + per-group CallFunction (/root/reference/src/dataframe.cpp:1582-1584,
// pd_core_macros.h:114-147) for the cardinalities where the reference's per-group loop takes seconds.
#pragma once
#include "gtable.cuh"
#include "merge.cuh"
#include "order.cuh"

namespace pa {

constexpr int RP_THREADS = 512;
constexpr int RP_ROWS = 8;                            // rows per thread
constexpr int RP_TILE = RP_THREADS * RP_ROWS;         // 4096 rows per tile
constexpr int RP_CHUNK_TILES = 16;
constexpr int RP_CHUNK_ROWS = RP_TILE * RP_CHUNK_TILES;   // rows per chunk (the unit of the offset bookkeeping)
constexpr int RP_MAX_FAN = 1024;
constexpr int RP_L1_LOG = 10;                         // level 1 always splits 1024 ways
constexpr int RP_MAX_BITS = 20;                       // at most 2^20 buckets
constexpr int RP_SINGLE_MAX_LOG = 7;                  // one level up to 128 buckets: the scatter slows down with its fan-out (runs of
                                                      // 4096 / fan rows: 7.7 / 8.4 / 10.5 / 15 / 25 / 43 ms per 1 B rows at 32 ... 1024)
constexpr int RH_THREADS = 1024;
constexpr int HLL_LOG2 = 12;
constexpr int HLL_M = 1 << HLL_LOG2;
constexpr int HLL_SAMPLE_LOG2 = 3;                    // one key value in 8 feeds the sketch
constexpr int BK_THREADS = 768;
constexpr int BK_MAX_PROBE = 64;
constexpr uint32_t RP_NULL_BIT = 0x80000000u;         // row-number word of the partitioned rows: the row's value is null

__device__ __forceinline__ uint64_t rp_mix(uint64_t key) { return (key ^ (key >> 32)) * 0x9E3779B97F4A7C15ull; }

__device__ __forceinline__ uint64_t ldg_stream_u64_na(const void* p) {
  uint64_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ldg_stream_u32_na(const void* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}

// ---------------------------------------------------------------------------------------------
// partition: chunks, histograms, offsets, scatter — no global atomics, stable, deterministic
// ---------------------------------------------------------------------------------------------
// A level splits every PARENT range of rows (level 1: the whole input; level 2: every level-1 bucket) `fan` ways.
// Parents are cut into CHUNKS of RP_CHUNK_ROWS consecutive rows (the last one of a parent ragged).  The histogram
// kernel counts rows per (chunk, bucket); the counts are laid out "flat" as
//     flat[cbase_p * fan + bucket * chunks_p + chunk_in_parent]        (cbase_p = chunks of the parents before p)
// so that ONE exclusive scan of the flat array yields, for every chunk, the absolute output position of its first row
// of every bucket: buckets of a parent follow each other, chunks inside a bucket follow each other in row order.
// The scatter kernel loads a chunk's `fan` cursors into shared memory and advances them privately tile by tile.
struct RpArgs {
  const uint64_t* keys;
  const uint64_t* vals;          // may be null (keys-only pass)
  const uint32_t* rows;          // original row numbers of the source rows; null = the source IS the original order
  const uint8_t* vvalid;         // level 1 only: validity bitmap of the value column (or null); a null value travels as
  int64_t voff;                  //   bit 31 of the row number (RP_NULL_BIT; the host admits < 2^31 rows then)
  int64_t n;
  int shift;                     // bucket inside the parent = (rp_mix(key) >> shift) & (fan - 1)
  int log_fan;
  int n_parents;                 // 1 (level 1) or the number of level-1 buckets
  const unsigned int* parent_end;   // [n_parents] end of every parent in the source order; null for level 1
  const unsigned int* cprefix;      // [n_parents + 1] chunks before parent p; null for level 1
  unsigned int n_chunks1;        // level 1: number of chunks
  unsigned int* counts;          // histogram out — level 1: fine counts [chunk][1024]; level 2: the flat layout
  const unsigned int* offsets;   // scatter in: the scanned flat layout
  unsigned int* hll;             // level 1 histogram: sketch registers
  uint64_t* out_keys;
  uint64_t* out_vals;
  uint32_t* out_rows;
};

struct RpChunk { uint32_t cl, chunks_p, cbase; int64_t row0; int cnt; };

__device__ __forceinline__ uint32_t rp_num_chunks(const RpArgs& a) {
  return a.cprefix ? a.cprefix[a.n_parents] : a.n_chunks1;
}

__device__ __forceinline__ RpChunk rp_chunk(const RpArgs& a, uint32_t c) {
  RpChunk r;
  if (!a.cprefix) {
    r.cl = c; r.chunks_p = a.n_chunks1; r.cbase = 0;
    r.row0 = static_cast<int64_t>(c) * RP_CHUNK_ROWS;
    const int64_t left = a.n - r.row0;
    r.cnt = left < RP_CHUNK_ROWS ? static_cast<int>(left) : RP_CHUNK_ROWS;
    return r;
  }
  uint32_t lo = 0, hi = static_cast<uint32_t>(a.n_parents);      // largest p with cprefix[p] <= c
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(a.cprefix + mid) <= c) lo = mid; else hi = mid;
  }
  const uint32_t start = lo ? __ldg(a.parent_end + lo - 1) : 0u, end = __ldg(a.parent_end + lo);
  r.cbase = __ldg(a.cprefix + lo);
  r.chunks_p = __ldg(a.cprefix + lo + 1) - r.cbase;
  r.cl = c - r.cbase;
  r.row0 = static_cast<int64_t>(start) + static_cast<int64_t>(r.cl) * RP_CHUNK_ROWS;
  const int64_t left = static_cast<int64_t>(end) - r.row0;
  r.cnt = left < RP_CHUNK_ROWS ? static_cast<int>(left) : RP_CHUNK_ROWS;
  return r;
}

__device__ __forceinline__ void rp_hll_add(uint64_t m, unsigned int* s_hll) {
  uint64_t h = (m ^ (m >> 29)) * 0xBF58476D1CE4E5B9ull;
  h ^= h >> 32;
  if ((h & ((1u << HLL_SAMPLE_LOG2) - 1u)) == 0) {
    const uint32_t idx = static_cast<uint32_t>(h >> HLL_SAMPLE_LOG2) & (HLL_M - 1);
    const uint64_t rest = h >> (HLL_SAMPLE_LOG2 + HLL_LOG2);                 // 49 bits
    const uint32_t rank = rest ? static_cast<uint32_t>(__clzll(static_cast<long long>(rest << (HLL_SAMPLE_LOG2 + HLL_LOG2)))) + 1u : 50u;
    atomicMax(&s_hll[idx], rank);
  }
}

// Rows per (chunk, bucket).  L1: 1024 fine buckets (the top 10 bits of the mix; the host folds them to the fan-out it
// picks once the sketch has told it how many groups there are) + the HyperLogLog sketch.
template <bool L1>
__global__ void __launch_bounds__(RH_THREADS) k_rp_hist(RpArgs a) {
  __shared__ unsigned int s_hist[RP_MAX_FAN];
  __shared__ unsigned int s_hll[L1 ? HLL_M : 1];
  const int fan = L1 ? RP_MAX_FAN : (1 << a.log_fan);
  const uint32_t nchunks = rp_num_chunks(a);
  if constexpr (L1) {
    for (int i = threadIdx.x; i < HLL_M; i += RH_THREADS) s_hll[i] = 0;
  }
  for (uint32_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
    for (int i = threadIdx.x; i < fan; i += RH_THREADS) s_hist[i] = 0;
    __syncthreads();
    const RpChunk ch = rp_chunk(a, c);
    const uint64_t* kp = a.keys + ch.row0;
#pragma unroll 4
    for (int r = threadIdx.x; r < ch.cnt; r += RH_THREADS) {
      const uint64_t m = rp_mix(ldg_stream_u64_na(kp + r));
      if constexpr (L1) {
        atomicAdd(&s_hist[m >> (64 - RP_L1_LOG)], 1u);
        rp_hll_add(m, s_hll);
      } else {
        atomicAdd(&s_hist[static_cast<uint32_t>(m >> a.shift) & (fan - 1)], 1u);
      }
    }
    __syncthreads();
    if constexpr (L1) {
      for (int i = threadIdx.x; i < fan; i += RH_THREADS) a.counts[static_cast<size_t>(c) * RP_MAX_FAN + i] = s_hist[i];
    } else {
      unsigned int* dst = a.counts + static_cast<size_t>(ch.cbase) * fan + ch.cl;
      for (int i = threadIdx.x; i < fan; i += RH_THREADS) dst[static_cast<size_t>(i) * ch.chunks_p] = s_hist[i];
    }
    __syncthreads();
  }
  if constexpr (L1) {
    for (int i = threadIdx.x; i < HLL_M; i += RH_THREADS)
      if (s_hll[i]) atomicMax(a.hll + i, s_hll[i]);
  }
}

// level 1: fine counts [chunk][1024] -> flat[bucket * n_chunks + chunk] with 2^b1 buckets (bucket = fine >> (10 - b1))
__global__ void __launch_bounds__(256) k_rp_fold(const unsigned int* fine, uint32_t n_chunks, int b1, unsigned int* flat) {
  const uint64_t i = blockIdx.x * 256ull + threadIdx.x;
  const uint32_t c = static_cast<uint32_t>(i >> b1), B = static_cast<uint32_t>(i) & ((1u << b1) - 1u);
  if (c >= n_chunks) return;
  const int s = RP_L1_LOG - b1;
  const unsigned int* src = fine + static_cast<size_t>(c) * RP_MAX_FAN + (static_cast<size_t>(B) << s);
  unsigned int sum = 0;
  for (int j = 0; j < (1 << s); ++j) sum += src[j];
  flat[static_cast<size_t>(B) * n_chunks + c] = sum;
}

// ends[parent << log_fan | bucket] = end of that bucket in the output order, read off the scanned flat layout (which
// carries one extra entry holding n)
__global__ void __launch_bounds__(256) k_rp_ends(RpArgs a, unsigned int* ends) {
  const uint32_t e = blockIdx.x * 256u + threadIdx.x;
  if (e >= (static_cast<uint32_t>(a.n_parents) << a.log_fan)) return;
  const uint32_t p = e >> a.log_fan, sb = e & ((1u << a.log_fan) - 1u);
  const uint32_t cbase = a.cprefix ? a.cprefix[p] : 0u;
  const uint32_t chunks_p = a.cprefix ? a.cprefix[p + 1] - cbase : a.n_chunks1;
  ends[e] = a.offsets[(static_cast<size_t>(cbase) << a.log_fan) + static_cast<size_t>(sb + 1) * chunks_p];
}

// chunks per parent -> exclusive prefix [n_parents + 1] (n_parents <= 1024)
__global__ void __launch_bounds__(1024) k_rp_cprefix(const unsigned int* ends, int n_parents, unsigned int* cprefix) {
  __shared__ unsigned int s[1024];
  const int t = threadIdx.x;
  unsigned int own = 0;
  if (t < n_parents) {
    const unsigned int start = t ? ends[t - 1] : 0u;
    own = (ends[t] - start + RP_CHUNK_ROWS - 1) / RP_CHUNK_ROWS;
  }
  s[t] = own;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    const unsigned int v = t >= d ? s[t - d] : 0u;
    __syncthreads();
    s[t] += v;
    __syncthreads();
  }
  if (t < n_parents) cprefix[t] = s[t] - own;
  if (t == n_parents - 1) cprefix[n_parents] = s[t];
}

template <int THREADS>
struct RpSmemT {
  static constexpr int TILE = THREADS * RP_ROWS;
  static constexpr size_t OFF_KEY = 0;
  static constexpr size_t OFF_VAL = OFF_KEY + sizeof(uint64_t) * TILE;
  static constexpr size_t OFF_ROW = OFF_VAL + sizeof(uint64_t) * TILE;
  static constexpr size_t OFF_HIST = OFF_ROW + sizeof(uint32_t) * TILE;
  static constexpr size_t OFF_OFF = OFF_HIST + sizeof(uint32_t) * RP_MAX_FAN;
  static constexpr size_t OFF_DELTA = OFF_OFF + sizeof(uint32_t) * RP_MAX_FAN;
  static constexpr size_t OFF_CUR = OFF_DELTA + sizeof(uint32_t) * RP_MAX_FAN;
  static constexpr size_t TOTAL = OFF_CUR + sizeof(uint32_t) * RP_MAX_FAN;
};
using RpSmem = RpSmemT<RP_THREADS>;
constexpr int RP_THREADS_BIG = 1024;          // 8192-row tiles, one CTA per SM: runs twice as long at the same fan-out
using RpSmemBig = RpSmemT<RP_THREADS_BIG>;
static_assert(2 * (RpSmem::TOTAL + 1024) <= 227 * 1024, "two scatter CTAs per SM");
static_assert(RpSmemBig::TOTAL + 1024 <= 227 * 1024, "one big-tile scatter CTA per SM");
static_assert(RP_CHUNK_ROWS % RpSmemBig::TILE == 0 && RpSmemBig::TILE <= (1 << 13), "tile ranks are 13 bits");

// Two CTAs per SM: the phases of a tile are separated by barriers (load -> rank -> offsets -> regroup -> write), so a
// second resident CTA keeps the memory system busy while the first one regroups.
// THREADS = 1024 (8192-row tiles, one CTA per SM) for fan-outs whose 4096-row runs would be 128 bytes or shorter
// (launch_rp_scatter in capi.cu picks; measured there).
// NULLABLE is a template parameter of the scatter and of k_bucket_agg: as a run-time test inside their row loops the
// handling of null values cost the plain path 5-8 % (64 K / 1 M / 100 M groups 15.0 / 25.1 / 53.9 -> 15.9 / 27.4 / 57.1 ms
// per 1 B rows, scripts/r2_run51.sh).
template <int THREADS, bool NULLABLE>
__global__ void __launch_bounds__(THREADS, THREADS == 512 ? 2 : 1) k_rp_scatter_t(RpArgs a) {
  using S = RpSmemT<THREADS>;
  constexpr int RP_THREADS = THREADS;                // (shadows the namespace constants inside this kernel)
  constexpr int RP_TILE = S::TILE;
  extern __shared__ __align__(16) unsigned char rp_smem[];
  uint64_t* st_key = reinterpret_cast<uint64_t*>(rp_smem + S::OFF_KEY);
  uint64_t* st_val = reinterpret_cast<uint64_t*>(rp_smem + S::OFF_VAL);
  uint32_t* st_row = reinterpret_cast<uint32_t*>(rp_smem + S::OFF_ROW);
  unsigned int* s_hist = reinterpret_cast<unsigned int*>(rp_smem + S::OFF_HIST);
  unsigned int* s_off = reinterpret_cast<unsigned int*>(rp_smem + S::OFF_OFF);
  // s_delta[b] = (start of this tile's run for b in the output) - (start of the run in the regrouped tile)
  unsigned int* s_delta = reinterpret_cast<unsigned int*>(rp_smem + S::OFF_DELTA);
  unsigned int* s_cur = reinterpret_cast<unsigned int*>(rp_smem + S::OFF_CUR);   // this chunk's write cursors
  __shared__ unsigned int s_wsum[RP_THREADS / 32];
  constexpr int BINS = RP_MAX_FAN / RP_THREADS;      // histogram bins per thread in the scan
  const int fan = 1 << a.log_fan;
  const uint32_t nchunks = rp_num_chunks(a);
  for (uint32_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const RpChunk ch = rp_chunk(a, c);
    {
      const unsigned int* src = a.offsets + (static_cast<size_t>(ch.cbase) << a.log_fan) + ch.cl;
      for (int i = threadIdx.x; i < fan; i += RP_THREADS) s_cur[i] = src[static_cast<size_t>(i) * ch.chunks_p];
    }
    for (int t0 = 0; t0 < ch.cnt; t0 += RP_TILE) {
      const int64_t row0 = ch.row0 + t0;
      const int cnt = ch.cnt - t0 < RP_TILE ? ch.cnt - t0 : RP_TILE;
      for (int i = threadIdx.x; i < fan; i += RP_THREADS) s_hist[i] = 0;
      __syncthreads();
      uint64_t key[RP_ROWS], val[RP_ROWS];
      uint32_t pr[RP_ROWS];    // bucket << 13 | rank inside the tile's run for that bucket
#pragma unroll
      for (int j = 0; j < RP_ROWS; ++j) {
        const int r = threadIdx.x + j * RP_THREADS;
        key[j] = 0; val[j] = 0;
        if (r < cnt) {
          key[j] = ldg_stream_u64_na(a.keys + row0 + r);
          if (a.vals) val[j] = ldg_stream_u64_na(a.vals + row0 + r);
        }
      }
#pragma unroll
      for (int j = 0; j < RP_ROWS; ++j) {
        const int r = threadIdx.x + j * RP_THREADS;
        pr[j] = 0xFFFFFFFFu;
        if (r < cnt) {
          const uint32_t b = static_cast<uint32_t>(rp_mix(key[j]) >> a.shift) & (fan - 1);
          pr[j] = (b << 13) | atomicAdd(&s_hist[b], 1u);
        }
      }
      __syncthreads();
      // exclusive scan of the tile histogram (BINS adjacent bins per thread); the output ranges come from the chunk's
      // private cursors (every bin is owned by one thread)
      {
        unsigned int mine[BINS], tot = 0;
#pragma unroll
        for (int b = 0; b < BINS; ++b) {
          const int bin = threadIdx.x * BINS + b;
          mine[b] = bin < fan ? s_hist[bin] : 0u;
          tot += mine[b];
        }
        unsigned int incl = tot;
        const uint32_t lane = lane_id(), w = threadIdx.x >> 5;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const unsigned int v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
          if (lane >= static_cast<uint32_t>(d)) incl += v;
        }
        if (lane == 31) s_wsum[w] = incl;
        __syncthreads();
        if (w == 0) {
          unsigned int x = lane < RP_THREADS / 32 ? s_wsum[lane] : 0u;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const unsigned int v = __shfl_up_sync(0xFFFFFFFFu, x, d);
            if (lane >= static_cast<uint32_t>(d)) x += v;
          }
          if (lane < RP_THREADS / 32) s_wsum[lane] = x;
        }
        __syncthreads();
        unsigned int excl = incl - tot + (w ? s_wsum[w - 1] : 0u);
#pragma unroll
        for (int b = 0; b < BINS; ++b) {
          const int bin = threadIdx.x * BINS + b;
          if (bin < fan) {
            s_off[bin] = excl;
            const unsigned int base = s_cur[bin];
            s_delta[bin] = base - excl;
            s_cur[bin] = base + mine[b];
          }
          excl += mine[b];
        }
      }
      __syncthreads();
      // regroup the tile in shared memory (the original row numbers are fetched only now: 64 registers per thread)
      uint32_t row[RP_ROWS];
#pragma unroll
      for (int j = 0; j < RP_ROWS; ++j) {
        const int r = threadIdx.x + j * RP_THREADS;
        row[j] = static_cast<uint32_t>(row0 + r);
        if (a.rows && pr[j] != 0xFFFFFFFFu) row[j] = ldg_stream_u32_na(a.rows + row0 + r);
        if constexpr (NULLABLE) {
          if (!a.rows && a.vvalid && pr[j] != 0xFFFFFFFFu && !bit_at(a.vvalid, a.voff + row0 + r)) row[j] |= RP_NULL_BIT;
        }
      }
#pragma unroll
      for (int j = 0; j < RP_ROWS; ++j) {
        if (pr[j] == 0xFFFFFFFFu) continue;
        const uint32_t pos = s_off[pr[j] >> 13] + (pr[j] & 0x1FFFu);
        st_key[pos] = key[j];
        st_val[pos] = val[j];
        st_row[pos] = row[j];
      }
      __syncthreads();
      // contiguous runs out (the bucket of a regrouped row is recomputed from its key)
      for (int i = threadIdx.x; i < cnt; i += RP_THREADS) {
        const uint64_t k = st_key[i];
        const unsigned int d = s_delta[static_cast<uint32_t>(rp_mix(k) >> a.shift) & (fan - 1)] + static_cast<unsigned int>(i);
        a.out_keys[d] = k;
        if (a.vals) a.out_vals[d] = st_val[i];
        a.out_rows[d] = st_row[i];
      }
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// bucket aggregation
// ---------------------------------------------------------------------------------------------
struct __align__(32) BkRec32 { uint64_t key, sum; uint32_t count, first; uint64_t pad; };
struct __align__(64) BkRec64 { uint64_t key, sum; uint32_t count, first, last, pad; uint64_t mn, mx; double dsum; uint64_t pad2; };
static_assert(sizeof(BkRec32) == 32 && sizeof(BkRec64) == 64, "group records are whole sectors");
template <bool WIDE> struct BkRecOf { using type = BkRec32; };
template <> struct BkRecOf<true> { using type = BkRec64; };

struct BkArgs {
  const uint64_t* keys;          // partitioned rows (arrays padded by 4 elements: 256-bit loads may read past a bucket)
  const uint64_t* vals;          // may be null
  const uint32_t* rows;
  const unsigned int* bucket_end;   // [n_buckets]
  uint32_t n_buckets;
  int part_bits;                 // bits of rp_mix consumed by the partitioning
  unsigned int* next_bucket;     // work counter
  uint32_t agg_mask;
  uint32_t max_keys;             // keys a bucket's table admits
  int nullable;                  // the row-number words carry RP_NULL_BIT for rows whose value is null
  // unordered groups: one record per group (BkRec: whole sectors, so the ordering pass reads a group with one or two
  // adjacent sector fetches) + their first rows as a plain array (bitmap / rank passes)
  void* u_rec;
  uint32_t* u_first;
  uint32_t u_cap;
  uint32_t* status;              // ST_OVERFLOW, ST_COUNTER (groups appended)
  // ranged mode (fewer buckets than would keep every SM busy): every CTA takes an equal, contiguous share of the
  // partitioned rows and flushes its table into the global table (gtable.cuh) whenever its rows move on to the next
  // bucket; a bucket shared by several CTAs is merged there by the L2 atomics (a few per group and CTA, not per row)
  int ranged;
  int64_t n;
  void* table;
  uint64_t cap_mask;
  int shift;
};

template <int VC, bool WIDE, bool NULLABLE>
__global__ void __launch_bounds__(BK_THREADS, 1) k_bucket_agg(BkArgs a) {
  using T = SmTab<VC, WIDE>;
  extern __shared__ __align__(16) unsigned char st_smem[];
  unsigned long long* s_key = reinterpret_cast<unsigned long long*>(st_smem + T::OFF_KEY);
  unsigned long long* s_sum = reinterpret_cast<unsigned long long*>(st_smem + T::OFF_SUM);
  unsigned long long* s_mn = reinterpret_cast<unsigned long long*>(st_smem + T::OFF_MN);
  unsigned long long* s_mx = reinterpret_cast<unsigned long long*>(st_smem + T::OFF_MX);
  double* s_dsum = reinterpret_cast<double*>(st_smem + T::OFF_DSUM);
  uint32_t* s_cnt = reinterpret_cast<uint32_t*>(st_smem + T::OFF_CNT);
  uint32_t* s_first = reinterpret_cast<uint32_t*>(st_smem + T::OFF_FIRST);
  uint32_t* s_last = reinterpret_cast<uint32_t*>(st_smem + T::OFF_LAST);
  uint32_t* s_misc = reinterpret_cast<uint32_t*>(st_smem + T::OFF_MISC);   // [0] keys in the table, [1] bucket, [2] output base, [3] emit cursor
  for (int i = threadIdx.x; i < T::NSLOT; i += BK_THREADS) {
    s_key[i] = kEmptyKey;
    s_sum[i] = 0ull;
    s_cnt[i] = 0u;
    s_first[i] = kNoRow;
    if constexpr (WIDE) { s_mn[i] = kMinInit; s_mx[i] = kMaxInit; s_last[i] = 0u; }
    if constexpr (T::DSUM) s_dsum[i] = 0.0;
  }
  if (threadIdx.x < 4) s_misc[threadIdx.x] = 0u;
  __syncthreads();
  using SlotT = typename SlotOf<WIDE>::type;
  constexpr int SLOT_LOG2 = WIDE ? 6 : 5;
  int64_t lo_c = 0, hi_c = 0;
  uint32_t bcur = 0;
  if (a.ranged) {
    const int64_t per = ((a.n + gridDim.x - 1) / gridDim.x + 3) & ~3ll;
    lo_c = static_cast<int64_t>(blockIdx.x) * per;
    hi_c = lo_c + per < a.n ? lo_c + per : a.n;
    uint32_t l = 0, h = a.n_buckets;                       // first bucket that ends after lo_c
    while (l < h) {
      const uint32_t mid = (l + h) >> 1;
      if (static_cast<int64_t>(a.bucket_end[mid]) > lo_c) h = mid; else l = mid + 1;
    }
    bcur = l;
  }
  for (;;) {
    uint32_t b, start, end;
    if (!a.ranged) {
      if (threadIdx.x == 0) {
        s_misc[1] = atomicAdd(a.next_bucket, 1u);
        s_misc[3] = 0u;
      }
      __syncthreads();
      b = s_misc[1];
      if (b >= a.n_buckets) break;
      if (*reinterpret_cast<volatile uint32_t*>(a.status + ST_OVERFLOW)) break;
      start = b ? a.bucket_end[b - 1] : 0u;
      end = a.bucket_end[b];
    } else {
      b = bcur++;
      if (b >= a.n_buckets || lo_c >= hi_c) break;
      if (*reinterpret_cast<volatile uint32_t*>(a.status + ST_OVERFLOW)) break;
      const int64_t be = a.bucket_end[b];
      start = static_cast<uint32_t>(lo_c);
      end = static_cast<uint32_t>(be < hi_c ? be : hi_c);
      lo_c = end;
      if (start >= end) continue;
    }
    bool failed = false;
    for (uint64_t base = (start & ~3u) + 4u * threadIdx.x; base < end; base += 4u * BK_THREADS) {
      const u64x4 k4 = ldg_stream_u64x4(a.keys + base);
      u64x4 v4{0ull, 0ull, 0ull, 0ull};
      if (a.vals) v4 = ldg_stream_u64x4(a.vals + base);
      const uint4 r4 = *reinterpret_cast<const uint4*>(a.rows + base);
      const uint64_t keyv[4] = {k4.a, k4.b, k4.c, k4.d};
      const uint64_t valv[4] = {v4.a, v4.b, v4.c, v4.d};
      const uint32_t rowv[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint64_t idx = base + j;
        if (idx < start || idx >= end) continue;
        const uint64_t key = keyv[j];
        uint32_t row = rowv[j];
        bool vvalid = true;
        if constexpr (NULLABLE) { vvalid = (row & RP_NULL_BIT) == 0u; row &= ~RP_NULL_BIT; }   // a null value: the row still makes
        uint32_t s;                                                                    // the group and counts for first / last
        bool found = false;
        if (key == kEmptyKey) { s = T::CAP + 1; found = true; }
        else {
          s = static_cast<uint32_t>((rp_mix(key) << a.part_bits) >> (64 - T::CAP_LOG2));
          for (int probe = 0; probe < BK_MAX_PROBE; ++probe) {
            const uint64_t k = *reinterpret_cast<volatile unsigned long long*>(s_key + s);
            if (k == key) { found = true; break; }
            if (k == kEmptyKey) {
              if (*reinterpret_cast<volatile uint32_t*>(s_misc) >= a.max_keys) break;
              const uint64_t old = atomicCAS(s_key + s, static_cast<unsigned long long>(kEmptyKey), static_cast<unsigned long long>(key));
              if (old == kEmptyKey) { atomicAdd(s_misc, 1u); found = true; break; }
              if (old == key) { found = true; break; }
            }
            s = (s + 1) & (T::CAP - 1);
          }
        }
        if (!found) { failed = true; continue; }
        if (row < *reinterpret_cast<volatile uint32_t*>(s_first + s)) atomicMin(s_first + s, row);
        if constexpr (WIDE) {
          if ((a.agg_mask & AGG_LAST) && row > *reinterpret_cast<volatile uint32_t*>(s_last + s)) atomicMax(s_last + s, row);
        }
        if (!a.vals || !vvalid) continue;
        atomicAdd(s_cnt + s, 1u);
        if constexpr (VC == VC_F) atomicAdd(reinterpret_cast<double*>(s_sum + s), __longlong_as_double(static_cast<long long>(valv[j])));
        else atomicAdd(s_sum + s, static_cast<unsigned long long>(valv[j]));
        if constexpr (WIDE) {
          if constexpr (T::DSUM) {
            if (a.agg_mask & AGG_MEAN) atomicAdd(s_dsum + s, Wide<VC>::as_double(valv[j]));
          }
          if ((a.agg_mask & (AGG_MIN | AGG_MAX)) && !Wide<VC>::is_nan(valv[j])) {
            const uint64_t o = Wide<VC>::ord(valv[j]);
            if (o < *reinterpret_cast<volatile unsigned long long*>(s_mn + s)) atomicMin(s_mn + s, static_cast<unsigned long long>(o));
            if (o > *reinterpret_cast<volatile unsigned long long*>(s_mx + s)) atomicMax(s_mx + s, static_cast<unsigned long long>(o));
          }
        }
      }
    }
    if (failed) atomicExch(a.status + ST_OVERFLOW, 1u);
    __syncthreads();
    if (*reinterpret_cast<volatile uint32_t*>(a.status + ST_OVERFLOW)) break;
    if (a.ranged) {
      // ---- flush into the global table, resetting the slots in the same sweep ----
      SlotT* table = static_cast<SlotT*>(a.table);
      const uint64_t cap = a.cap_mask + 1;
      for (int i = threadIdx.x; i < T::NSLOT; i += BK_THREADS) {
        const uint32_t first = s_first[i];
        if (first == kNoRow) continue;
        const uint64_t fkey = s_key[i];
        uint64_t sl = i == T::CAP + 1 ? cap + 1 : gtable_home(fkey, a.shift);
        const Sector kf = gtable_peek<WIDE>(table, sl, SLOT_LOG2);
        GProbe pr;
        if (sl >= cap) pr = GProbe{sl, static_cast<uint32_t>(kf.fl), static_cast<uint32_t>(kf.fl >> 32), kf.mn, kf.mx};
        else pr = gtable_find_or_insert<WIDE>(table, a.cap_mask, SLOT_LOG2, fkey, sl, kf);
        if (pr.slot == ~0ull) { atomicExch(a.status + ST_OVERFLOW, 1u); break; }
        SlotT* g = table + pr.slot;
        if (first < pr.first) atomicMin(&g->first_row, first);
        const uint32_t c = s_cnt[i];
        if (c) {
          atomicAdd(&g->count, c);
          if constexpr (VC == VC_F) atomicAdd(reinterpret_cast<double*>(&g->sum), __longlong_as_double(static_cast<long long>(s_sum[i])));
          else atomicAdd(reinterpret_cast<unsigned long long*>(&g->sum), s_sum[i]);
        }
        if constexpr (WIDE) {
          if ((a.agg_mask & AGG_LAST) && s_last[i] > pr.last) atomicMax(&g->last_row, s_last[i]);
          if (s_mn[i] < pr.mn) atomicMin(reinterpret_cast<unsigned long long*>(&g->min_ord), s_mn[i]);
          if (s_mx[i] > pr.mx) atomicMax(reinterpret_cast<unsigned long long*>(&g->max_ord), s_mx[i]);
          if constexpr (T::DSUM) { if (c) atomicAdd(&g->dsum, s_dsum[i]); }
        }
        s_key[i] = kEmptyKey;
        s_sum[i] = 0ull;
        s_cnt[i] = 0u;
        s_first[i] = kNoRow;
        if constexpr (WIDE) { s_mn[i] = kMinInit; s_mx[i] = kMaxInit; s_last[i] = 0u; }
        if constexpr (T::DSUM) s_dsum[i] = 0.0;
      }
      __syncthreads();
      if (threadIdx.x == 0) s_misc[0] = 0u;
      __syncthreads();
      continue;
    }
    // ---- append this bucket's groups, resetting the slots in the same sweep ----
    if (threadIdx.x == 0) {
      const uint32_t total = s_misc[0] + (s_first[T::CAP + 1] != kNoRow ? 1u : 0u);
      const uint32_t gbase = atomicAdd(a.status + ST_COUNTER, total);
      s_misc[2] = gbase;
      if (static_cast<uint64_t>(gbase) + total > a.u_cap) atomicExch(a.status + ST_OVERFLOW, 2u);
    }
    __syncthreads();
    const uint32_t gbase = s_misc[2];
    const bool fits = static_cast<uint64_t>(gbase) + s_misc[0] + 1u <= a.u_cap;
    for (int i0 = 0; i0 < T::NSLOT; i0 += BK_THREADS) {
      const int i = i0 + threadIdx.x;
      const uint32_t first = i < T::NSLOT ? s_first[i] : kNoRow;
      const bool occ = first != kNoRow;
      const uint32_t m = __ballot_sync(0xFFFFFFFFu, occ);
      if (m == 0) continue;
      uint32_t wbase = 0;
      if (lane_id() == 0) wbase = atomicAdd(s_misc + 3, __popc(m));
      wbase = __shfl_sync(0xFFFFFFFFu, wbase, 0);
      if (occ) {
        const uint32_t pos = gbase + wbase + __popc(m & ((1u << lane_id()) - 1u));
        if (fits) {
          using Rec = typename BkRecOf<WIDE>::type;
          Rec r{};
          r.key = i == T::CAP + 1 ? kEmptyKey : s_key[i];
          r.sum = s_sum[i];
          r.count = s_cnt[i];
          r.first = first;
          if constexpr (WIDE) {
            r.last = s_last[i];
            r.mn = s_mn[i];
            r.mx = s_mx[i];
            if constexpr (T::DSUM) r.dsum = s_dsum[i];
          }
          uint4* dst = reinterpret_cast<uint4*>(static_cast<Rec*>(a.u_rec) + pos);
          const uint4* src = reinterpret_cast<const uint4*>(&r);
#pragma unroll
          for (int q = 0; q < static_cast<int>(sizeof(Rec) / 16); ++q) dst[q] = src[q];
          a.u_first[pos] = first;
        }
        s_key[i] = kEmptyKey;
        s_sum[i] = 0ull;
        s_cnt[i] = 0u;
        s_first[i] = kNoRow;
        if constexpr (WIDE) { s_mn[i] = kMinInit; s_mx[i] = kMaxInit; s_last[i] = 0u; }
        if constexpr (T::DSUM) s_dsum[i] = 0.0;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) s_misc[0] = 0u;
    __syncthreads();
  }
}

// GroupResult[r] = unordered group perm[r] (perm = k_bm_rank of the first rows, order.cuh): one random record fetch
// per group, everything else streams
template <bool WIDE>
__global__ void __launch_bounds__(256) k_bk_gather(const void* recs, const uint32_t* perm, uint32_t G, GroupResult out) {
  using Rec = typename BkRecOf<WIDE>::type;
  const uint32_t r = blockIdx.x * 256u + threadIdx.x;
  if (r >= G) return;
  const uint4* src = reinterpret_cast<const uint4*>(static_cast<const Rec*>(recs) + perm[r]);
  Rec rec;
  uint4* dst = reinterpret_cast<uint4*>(&rec);
#pragma unroll
  for (int q = 0; q < static_cast<int>(sizeof(Rec) / 16); ++q) dst[q] = __ldg(src + q);
  out.key[r] = rec.key;
  out.key_kind[r] = KK_REGULAR;
  out.sum[r] = rec.sum;
  out.count[r] = rec.count;
  out.first_row[r] = rec.first;
  if constexpr (WIDE) {
    out.last_row[r] = rec.last;
    out.min_ord[r] = rec.mn;
    out.max_ord[r] = rec.mx;
    if (out.dsum) out.dsum[r] = rec.dsum;
  } else {
    out.last_row[r] = 0u;
  }
}

// ---------------------------------------------------------------------------------------------
// Sharded step with compact partial records (merge.cuh): the unordered group records of the bucket aggregation go
// straight to their owners — no ranking, no gather into the GroupResult, no second read of it by the export.
// Same owner function, same per-CTA cursor claim as k_partials_count / k_partials_scatter.  nparts <= 64.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bkrec_count(const BkRec32* recs, uint32_t G, uint32_t nparts, unsigned long long* counts) {
  __shared__ unsigned int s_cnt[64];
  if (threadIdx.x < 64) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < G) atomicAdd(&s_cnt[owner_of(recs[g].key, KK_REGULAR, nparts)], 1u);
  __syncthreads();
  if (threadIdx.x < nparts && s_cnt[threadIdx.x]) atomicAdd(counts + threadIdx.x, static_cast<unsigned long long>(s_cnt[threadIdx.x]));
}

__global__ void __launch_bounds__(256) k_bkrec_scatter(const BkRec32* recs, uint32_t G, uint32_t nparts, int64_t row_base,
                                                       unsigned long long* cursor, uint64_t* records) {
  __shared__ unsigned int s_cnt[64];
  __shared__ unsigned long long s_base[64];
  if (threadIdx.x < 64) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t o = 0, local = 0;
  uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
  if (g < G) {
    const uint4* src = reinterpret_cast<const uint4*>(recs + g);
    lo = __ldg(src);          // key, sum
    hi = __ldg(src + 1);      // count, first row, pad
    o = owner_of(static_cast<uint64_t>(lo.x) | (static_cast<uint64_t>(lo.y) << 32), KK_REGULAR, nparts);
    local = atomicAdd(&s_cnt[o], 1u);
  }
  __syncthreads();
  if (threadIdx.x < nparts && s_cnt[threadIdx.x])
    s_base[threadIdx.x] = atomicAdd(cursor + threadIdx.x, static_cast<unsigned long long>(s_cnt[threadIdx.x]));
  __syncthreads();
  if (g < G) {
    ulonglong2 a, b;
    a.x = static_cast<uint64_t>(lo.x) | (static_cast<uint64_t>(lo.y) << 32);
    a.y = static_cast<uint64_t>(lo.z) | (static_cast<uint64_t>(lo.w) << 32);
    b.x = static_cast<uint64_t>(hi.x);                                        // count (a bucket record never has a null key)
    b.y = static_cast<uint64_t>(row_base + static_cast<int64_t>(hi.y));       // global first row
    ulonglong2* dst = reinterpret_cast<ulonglong2*>(records + (s_base[o] + local) * REC_WORDS_COMPACT);
    dst[0] = a;
    dst[1] = b;
  }
}

}  // namespace pa
