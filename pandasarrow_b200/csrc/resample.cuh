// Time-bucket specialisation (pd::resample with a fixed-width rule): the index is sorted, so every
// bucket is one contiguous run of rows and no hash table is needed.
//
//   bucket(ts) = floor((ts' - first) / freq),  ts' = ts (closed left) or ts - 1 (closed right)
//
// Each warp owns a contiguous chunk of rows and walks it 32 rows at a time.  While all 32 rows
// fall into the warp's current bucket (the common case: ~1000 ticks per bucket) the batch is
// reduced with a fixed shuffle tree and folded into register accumulators; the 64-bit division is
// only executed when a bucket boundary is crossed.  Runs that start and end inside the chunk are
// complete and are written straight to the bucket table (no atomics); the first and last run of a
// chunk may continue in the neighbouring chunks, so they go to a boundary list that a second
// kernel folds in row order.  All floating-point additions therefore happen in a fixed order.
//
// Replaces generate_bins_dt64 + GroupInfo::downsample + the hash group-by on the label column
// (/root/reference/src/resample.cpp:11-83, resample.h:19-43,91-122, group_by.h:255-299).
#pragma once
#include "gtable.cuh"

namespace pa {

constexpr int RS_CHUNK = 1024;     // rows per warp-chunk
constexpr int RS_THREADS = 256;

struct ResampleSpec {
  int64_t first;        // left edge of bucket 0 (ns)
  int64_t freq;         // bucket width (ns)
  int64_t nbins;
  int64_t label_off;    // 0 or freq (label_right)
  int closed_right;
};

struct __align__(16) RsPartial {
  int64_t bucket;       // -1: empty
  uint64_t sum;
  double dsum;
  uint64_t mn, mx;
  uint32_t cnt, first_row, last_row, pad;
};

struct RsArgs {
  const int64_t* ts;
  const void* vals;
  const uint8_t* vvalid;
  int64_t voff;
  int vw;
  int64_t n;
  ResampleSpec spec;
  void* table;          // nbins (+2) slots
  RsPartial* bnd;       // [nchunks][2]
  int64_t nchunks;
  uint32_t* status;
  uint32_t agg_mask;
};

template <int VC, bool WIDE>
struct RsAcc {
  uint64_t sum = 0;     // double bits or wrapping int
  double dsum = 0.0;
  uint64_t mn = kMinInit, mx = kMaxInit;
  uint32_t cnt = 0, first_row = kNoRow, last_row = 0;
  __device__ __forceinline__ void reset() { sum = 0; dsum = 0.0; mn = kMinInit; mx = kMaxInit; cnt = 0; first_row = kNoRow; last_row = 0; }
  __device__ __forceinline__ void add(uint64_t x_sum, double x_dsum, uint64_t x_mn, uint64_t x_mx, uint32_t x_cnt) {
    if constexpr (VC == VC_F) sum = static_cast<uint64_t>(__double_as_longlong(__longlong_as_double(static_cast<long long>(sum)) +
                                                                             __longlong_as_double(static_cast<long long>(x_sum))));
    else sum += x_sum;
    cnt += x_cnt;
    if constexpr (WIDE) {
      dsum += x_dsum;
      mn = x_mn < mn ? x_mn : mn;
      mx = x_mx > mx ? x_mx : mx;
    }
  }
};

template <int VC, bool WIDE>
__device__ __forceinline__ void rs_store_slot(void* table, int64_t b, const ResampleSpec& sp, const RsAcc<VC, WIDE>& a) {
  using SlotT = typename SlotOf<WIDE>::type;
  SlotT* s = static_cast<SlotT*>(table) + b;
  Slot32* q = reinterpret_cast<Slot32*>(s);
  q->key = static_cast<uint64_t>(sp.first + b * sp.freq + sp.label_off);
  q->first_row = a.first_row;
  q->last_row = a.last_row;
  q->sum = a.sum;
  q->count = a.cnt;
  if constexpr (WIDE) {
    Slot64* w = reinterpret_cast<Slot64*>(s);
    w->min_ord = a.mn;
    w->max_ord = a.mx;
    w->dsum = a.dsum;
  }
}

// fixed shuffle tree over the lanes in `mask` (others contribute the identity)
template <int VC, bool WIDE>
__device__ __forceinline__ void rs_warp_reduce(uint32_t mask, uint32_t lane, uint64_t vb, bool valid, uint64_t* o_sum,
                                               double* o_dsum, uint64_t* o_mn, uint64_t* o_mx, uint32_t* o_cnt) {
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  const bool in = ((mask >> lane) & 1u) && valid;
  *o_cnt = __popc(__ballot_sync(FULL, in));
  uint64_t mn = kMinInit, mx = kMaxInit;
  double ds = 0.0;
  if constexpr (VC == VC_F) {
    double s = in ? __longlong_as_double(static_cast<long long>(vb)) : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    *o_sum = static_cast<uint64_t>(__double_as_longlong(s));
  } else {
    uint64_t s = in ? vb : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    *o_sum = s;
    if constexpr (WIDE) {
      ds = in ? Wide<VC>::as_double(vb) : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ds += __shfl_xor_sync(FULL, ds, o);
    }
  }
  if constexpr (WIDE) {
    if (in && !Wide<VC>::is_nan(vb)) { mn = Wide<VC>::ord(vb); mx = mn; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const uint64_t a = __shfl_xor_sync(FULL, mn, o), b = __shfl_xor_sync(FULL, mx, o);
      mn = a < mn ? a : mn;
      mx = b > mx ? b : mx;
    }
  }
  *o_dsum = ds;
  *o_mn = mn;
  *o_mx = mx;
}

template <int VC, bool WIDE>
__global__ void __launch_bounds__(RS_THREADS) k_resample_scan(RsArgs a) {
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  const uint32_t lane = lane_id();
  const int64_t gw = (static_cast<int64_t>(blockIdx.x) * RS_THREADS + threadIdx.x) >> 5;
  const int64_t nw = (static_cast<int64_t>(gridDim.x) * RS_THREADS) >> 5;
  const ResampleSpec sp = a.spec;
  for (int64_t c = gw; c < a.nchunks; c += nw) {
    const int64_t row0 = c * RS_CHUNK;
    const int64_t row_end = row0 + RS_CHUNK < a.n ? row0 + RS_CHUNK : a.n;
    RsAcc<VC, WIDE> acc;
    int64_t cur_b = -1, cur_lo = 0, cur_hi = 0;   // current bucket and its [lo, hi) range on ts'
    bool first_seg = true;                         // the open run started at the chunk's first row
    bool wrote_first = false;
    int64_t prev_t = row0 > 0 ? a.ts[row0 - 1] : INT64_MIN;   // sortedness across chunk borders
    bool unsorted = false;
    auto flush = [&](bool chunk_end) {
      // called by all lanes with identical (uniform) state; lane 0 stores
      if (cur_b < 0) return;
      if (lane == 0) {
        if (first_seg || chunk_end) {
          RsPartial p;
          p.bucket = cur_b; p.sum = acc.sum; p.dsum = acc.dsum; p.mn = acc.mn; p.mx = acc.mx;
          p.cnt = acc.cnt; p.first_row = acc.first_row; p.last_row = acc.last_row; p.pad = 0;
          a.bnd[c * 2 + (first_seg ? 0 : 1)] = p;
        } else {
          rs_store_slot<VC, WIDE>(a.table, cur_b, sp, acc);
        }
      }
      if (first_seg) wrote_first = true;
      first_seg = false;
    };
    for (int64_t r0 = row0; r0 < row_end; r0 += 32) {
      const int64_t row = r0 + lane;
      const bool active = row < row_end;
      int64_t t = INT64_MAX;
      uint64_t vb = 0;
      bool valid = false;
      if (active) {
        t = a.ts[row];
        if (a.vals) {
          vb = load_wide_rt<VC>(a.vals, row, a.vw);
          valid = a.vvalid ? bit_at(a.vvalid, a.voff + row) : true;
        }
      }
      // sortedness: every row must be >= its predecessor
      int64_t before = __shfl_up_sync(FULL, t, 1);
      if (lane == 0) before = prev_t;
      if (active && t < before) unsorted = true;
      const int last_active = __popc(__ballot_sync(FULL, active)) - 1;
      prev_t = __shfl_sync(FULL, t, last_active);
      const int64_t tp = sp.closed_right ? t - 1 : t;
      uint32_t todo = __ballot_sync(FULL, active);
      while (todo) {
        const uint32_t in_cur = __ballot_sync(FULL, active && tp >= cur_lo && tp < cur_hi && cur_b >= 0) & todo;
        if (in_cur) {
          uint64_t s, mn, mx;
          double ds;
          uint32_t cnt;
          rs_warp_reduce<VC, WIDE>(in_cur, lane, vb, valid, &s, &ds, &mn, &mx, &cnt);
          acc.add(s, ds, mn, mx, cnt);
          const uint32_t lo_row = static_cast<uint32_t>(r0) + (__ffs(in_cur) - 1);
          const uint32_t hi_row = static_cast<uint32_t>(r0) + (31 - __clz(in_cur));
          if (acc.first_row == kNoRow) acc.first_row = lo_row;
          acc.last_row = hi_row;
          todo &= ~in_cur;
        }
        if (todo) {
          // the lowest remaining row opens a new run: close the current one, locate the next bucket
          flush(false);
          acc.reset();
          const int src = __ffs(todo) - 1;
          const int64_t t0 = __shfl_sync(FULL, tp, src);
          int64_t b = (t0 - sp.first) / sp.freq;
          if (t0 < sp.first || b >= sp.nbins) {   // outside the anchored range: input was not sorted
            unsorted = true;
            cur_b = -1;
            todo &= ~(1u << src);
            continue;
          }
          cur_b = b;
          cur_lo = sp.first + b * sp.freq;
          cur_hi = cur_lo + sp.freq;
        }
      }
    }
    // close the last run of the chunk
    if (cur_b >= 0) flush(true);
    if (lane == 0) {
      // mark the slots this chunk did not use as empty
      if (!wrote_first) a.bnd[c * 2].bucket = -1;
      // slot 1 is used only when the chunk's last run is not also its first run
      // (flush(true) with first_seg == true wrote slot 0)
    }
    if (__any_sync(FULL, unsorted) && lane == 0) atomicExch(a.status + ST_UNSORTED, 1u);
  }
}

// Fold the boundary partials (row order) into the bucket table.  Thread per entry; the first
// entry of every run of equal buckets folds the run sequentially.
template <int VC, bool WIDE>
__global__ void __launch_bounds__(256) k_resample_fixup(RsArgs a) {
  const int64_t ne = a.nchunks * 2;
  const int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (e >= ne) return;
  const int64_t b = a.bnd[e].bucket;
  if (b < 0) return;
  // previous non-empty entry
  for (int64_t p = e - 1; p >= 0; --p) {
    const int64_t pb = a.bnd[p].bucket;
    if (pb < 0) continue;
    if (pb == b) return;   // not the head of its run
    break;
  }
  RsAcc<VC, WIDE> acc;
  for (int64_t q = e; q < ne; ++q) {
    const RsPartial p = a.bnd[q];
    if (p.bucket < 0) continue;
    if (p.bucket != b) break;
    acc.add(p.sum, p.dsum, p.mn, p.mx, p.cnt);
    if (acc.first_row == kNoRow) acc.first_row = p.first_row;
    acc.last_row = p.last_row;
  }
  rs_store_slot<VC, WIDE>(a.table, b, a.spec, acc);
}

__global__ void k_rs_init_bnd(RsPartial* bnd, int64_t n) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) bnd[i].bucket = -1;
}

}  // namespace pa
