// Time-bucket specialisation (pd::resample with a fixed-width rule): the index is sorted, so every
// bucket is one contiguous run of rows and no hash table is needed.
//
//   bucket(ts) = floor((ts' - first) / freq),  ts' = ts (closed left) or ts - 1 (closed right)
//
// Each warp owns a contiguous chunk of rows and walks it 128 rows (4 coalesced loads per column) at
// a time.  While all rows of a 32-row batch fall into the warp's current bucket (the common case:
// ~1000 ticks per bucket) every lane folds its own row into lane-private register accumulators — no
// shuffles; the lanes are combined with a fixed xor-shuffle tree only when a run ends, and the
// 64-bit division is only executed when a bucket boundary is crossed.  Runs that start and end inside the chunk are
// complete and are written straight to the bucket table (no atomics); the first and last run of a
// chunk may continue in the neighbouring chunks, so they go to a boundary list that a second
// kernel folds in row order.  All floating-point additions therefore happen in a fixed order.
//
// Replaces generate_bins_dt64 + GroupInfo::downsample + the hash group-by on the label column
// (/root/reference/src/resample.cpp:11-83, resample.h:19-43,91-122, group_by.h:255-299).
#pragma once
#include "gtable.cuh"

namespace pa {

constexpr int RS_CHUNK = 1024;     // rows per warp-chunk (small inputs: enough chunks for every warp of the grid)
constexpr int RS_CHUNK_BIG = 4096; // from RS_BIG_ROWS rows on: a chunk's first iteration never finds an open run and every chunk
                                   // leaves two boundary partials to the fixup kernel — a quarter of both
constexpr int64_t RS_BIG_ROWS = 32 << 20;
inline int64_t rs_chunk_rows(int64_t n) { return n >= RS_BIG_ROWS ? RS_CHUNK_BIG : RS_CHUNK; }
constexpr int RS_THREADS = 256;

struct ResampleSpec {
  int64_t first;        // left edge of bucket 0 (ns)
  int64_t freq;         // bucket width (ns)
  int64_t nbins;
  int64_t label_off;    // 0 or freq (label_right)
  int closed_right;
  // calendar (DateOffset) rules: bucket b = [edges[b], edges[b + 1]) on ts' with the label labels[b]; both arrays are
  // computed on the host (O(#buckets), makeGroupInfo's DateOffset branch) and live on the device.  null = fixed width.
  const int64_t* edges;   // [nbins + 1], ascending
  const int64_t* labels;  // [nbins]
};

// Bucket of ts' (= ts, or ts - 1 when the buckets are closed on the right); -1 = outside the anchored range.
// Only evaluated where a new run of rows begins, never per row.
__device__ __forceinline__ int64_t rs_locate(const ResampleSpec& sp, int64_t tp, int64_t* lo, uint64_t* width) {
  if (!sp.edges) {
    if (tp < sp.first) return -1;
    const int64_t b = (tp - sp.first) / sp.freq;
    if (b >= sp.nbins) return -1;
    *lo = sp.first + b * sp.freq;
    *width = static_cast<uint64_t>(sp.freq);
    return b;
  }
  if (tp < __ldg(sp.edges) || tp >= __ldg(sp.edges + sp.nbins)) return -1;
  int64_t l = 0, h = sp.nbins;                    // largest b with edges[b] <= tp
  while (h - l > 1) {
    const int64_t mid = (l + h) >> 1;
    if (__ldg(sp.edges + mid) <= tp) l = mid; else h = mid;
  }
  const int64_t e0 = __ldg(sp.edges + l);
  *lo = e0;
  *width = static_cast<uint64_t>(__ldg(sp.edges + l + 1) - e0);
  return l;
}
__device__ __forceinline__ int64_t rs_label(const ResampleSpec& sp, int64_t b) {
  return sp.labels ? __ldg(sp.labels + b) : sp.first + b * sp.freq + sp.label_off;
}

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
  int64_t chunk;        // rows per warp-chunk (a multiple of 32 * RS_U)
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
  SlotT* q = static_cast<SlotT*>(table) + b;
  q->key = static_cast<uint64_t>(rs_label(sp, b));
  q->first_row = a.first_row;
  q->last_row = a.last_row;
  q->sum = a.sum;
  q->count = a.cnt;
  if constexpr (WIDE) {
    q->min_ord = a.mn;
    q->max_ord = a.mx;
    q->dsum = a.dsum;
  }
}

// Lane-private partial of the open run: every lane folds its own rows (row order), the lanes are
// combined with a fixed xor-shuffle tree only when the run is closed.
// fp64 values keep min / max as DOUBLES folded with one DSETP + two SEL each (an ordered compare is false for a NaN, which
// is exactly "NaN is skipped"), converted to order-mapped words when the run closes: the order map + NaN test + two 64-bit
// integer compare-and-selects cost ~15 instructions per row and made OHLC issue bound (sm_100a has no DMNMX; fmin / fmax
// compile to DSETP + FSEL + LOP3 + moves, no cheaper).  -0.0 and +0.0 compare equal here, so which zero a lane keeps is
// the first it met (the lanes are then combined in the order map, -0.0 < +0.0) — the sign of a zero minimum / maximum is
// the documented tie difference against Arrow's fmin / fmax.  +inf / -inf together mean "no number seen" (a lane that saw
// +inf only has mx = +inf).
template <int VC, bool WIDE>
struct RsLane {
  static constexpr uint64_t MN0 = VC == VC_F ? 0x7FF0000000000000ull : kMinInit;   // fp64: +inf bits
  static constexpr uint64_t MX0 = VC == VC_F ? 0xFFF0000000000000ull : kMaxInit;   // fp64: -inf bits
  uint64_t sum = 0;     // double bits or wrapping int
  double dsum = 0.0;
  uint64_t mn = MN0, mx = MX0;   // fp64 values: double bits; integers: order-mapped
  uint32_t cnt = 0;
  __device__ __forceinline__ void reset() { sum = 0; dsum = 0.0; mn = MN0; mx = MX0; cnt = 0; }
  __device__ __forceinline__ void fold_mm(uint64_t vb) {
    if constexpr (VC == VC_F) {
      const double v = __longlong_as_double(static_cast<long long>(vb));
      mn = v < __longlong_as_double(static_cast<long long>(mn)) ? vb : mn;   // false for a NaN
      mx = v > __longlong_as_double(static_cast<long long>(mx)) ? vb : mx;
    } else {
      const uint64_t o = Wide<VC>::ord(vb);
      mn = o < mn ? o : mn;
      mx = o > mx ? o : mx;
    }
  }
  __device__ __forceinline__ void add_row(uint64_t vb, uint32_t agg_mask) {
    if constexpr (VC == VC_F) sum = static_cast<uint64_t>(__double_as_longlong(__longlong_as_double(static_cast<long long>(sum)) +
                                                                             __longlong_as_double(static_cast<long long>(vb))));
    else sum += vb;
    ++cnt;
    if constexpr (WIDE) {
      if constexpr (VC != VC_F) dsum += Wide<VC>::as_double(vb);
      if (agg_mask & (AGG_MIN | AGG_MAX)) fold_mm(vb);
    }
  }
  // the same without a branch (the steady state of the scan: every row of the batch is valid and in the open run);
  // `want_mm` is warp-uniform
  __device__ __forceinline__ void add_row_nb(uint64_t vb, bool want_mm) {
    if constexpr (VC == VC_F) sum = static_cast<uint64_t>(__double_as_longlong(__longlong_as_double(static_cast<long long>(sum)) +
                                                                             __longlong_as_double(static_cast<long long>(vb))));
    else sum += vb;
    ++cnt;
    if constexpr (WIDE) {
      if constexpr (VC != VC_F) dsum += Wide<VC>::as_double(vb);
      if (want_mm) fold_mm(vb);
    }
  }
  // all lanes end up with the warp total
  __device__ __forceinline__ void warp_total(RsAcc<VC, WIDE>& out) const {
    constexpr uint32_t FULL = 0xFFFFFFFFu;
    uint32_t c = cnt;
    uint64_t n = mn, x = mx;
    if constexpr (VC == VC_F && WIDE) {
      const bool none = mn == MN0 && mx == MX0;
      n = none ? kMinInit : f64_to_ord(__longlong_as_double(static_cast<long long>(mn)));
      x = none ? kMaxInit : f64_to_ord(__longlong_as_double(static_cast<long long>(mx)));
    }
    double ds = dsum;
    if constexpr (VC == VC_F) {
      double v = __longlong_as_double(static_cast<long long>(sum));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
      out.sum = static_cast<uint64_t>(__double_as_longlong(v));
    } else {
      uint64_t v = sum;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
      out.sum = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
    out.cnt = c;
    if constexpr (WIDE) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const uint64_t a = __shfl_xor_sync(FULL, n, o), b = __shfl_xor_sync(FULL, x, o);
        n = a < n ? a : n;
        x = b > x ? b : x;
        if constexpr (VC != VC_F) ds += __shfl_xor_sync(FULL, ds, o);
      }
      out.mn = n;
      out.mx = x;
      out.dsum = ds;
    }
  }
};

constexpr int RS_U = 4;   // 32-row batches loaded per iteration (loads issued before any dependent work)

// Sortedness: the algorithm needs the BUCKET sequence to be non-decreasing (every bucket is one run
// of rows); the order of timestamps inside a bucket is irrelevant — as it is for the reference's
// two-pointer scan (resample.cpp:43-80).  Inside a chunk every new run must open a later bucket
// than the one before; across chunks the fixup kernel compares the boundary runs.
template <int VC, bool WIDE>
__global__ void __launch_bounds__(RS_THREADS) k_resample_scan(RsArgs a) {
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  const uint32_t lane = lane_id();
  const int64_t gw = (static_cast<int64_t>(blockIdx.x) * RS_THREADS + threadIdx.x) >> 5;
  const int64_t nw = (static_cast<int64_t>(gridDim.x) * RS_THREADS) >> 5;
  const ResampleSpec sp = a.spec;
  const bool fast_vals = a.vals != nullptr && a.vw == 8 && a.vvalid == nullptr;
  const int64_t shift = sp.closed_right ? 1 : 0;
  for (int64_t c = gw; c < a.nchunks; c += nw) {
    const int64_t row0 = c * a.chunk;
    const int64_t row_end = row0 + a.chunk < a.n ? row0 + a.chunk : a.n;
    RsLane<VC, WIDE> part;                         // this lane's share of the open run
    uint32_t run_first = kNoRow, run_last = 0;     // first / last row of the open run (uniform)
    int64_t cur_b = -1;                            // bucket of the open run
    int64_t cur_lo = INT64_MIN;                    // its left edge on ts'
    uint64_t freq = 0;                             // its width (fixed rules: the same for every bucket)
    bool have_run = false;
    bool first_seg = true;                         // the open run started at the chunk's first row
    bool wrote_first = false;
    bool unsorted = false;
    auto flush = [&](bool chunk_end) {
      // called by all lanes with identical (uniform) run state
      if (!have_run) return;
      RsAcc<VC, WIDE> acc;
      part.warp_total(acc);
      acc.first_row = run_first;
      acc.last_row = run_last;
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
    for (int64_t g0 = row0; g0 < row_end; g0 += 32 * RS_U) {
      int64_t tt[RS_U];
      uint64_t vv[RS_U];
      bool ok[RS_U];
      const bool full = g0 + 32 * RS_U <= row_end;
      if (full && fast_vals) {
        // Steady state (a run is ~1000 rows in config 4, an iteration 128): unguarded loads, ONE vote for the whole
        // iteration, straight-line branch-free folds.  (The per-batch code below costs ~95 warp instructions per 32 rows
        // — votes, reconvergence points, guarded loads — and made the kernel issue bound at 0.60 of the HBM peak.)
        const int64_t* tp0 = a.ts + g0 + lane;
        const uint64_t* vp0 = static_cast<const uint64_t*>(a.vals) + g0 + lane;
#pragma unroll
        for (int u = 0; u < RS_U; ++u) tt[u] = tp0[u * 32];
#pragma unroll
        for (int u = 0; u < RS_U; ++u) { vv[u] = vp0[u * 32]; ok[u] = true; }
        bool in_all = have_run;
#pragma unroll
        for (int u = 0; u < RS_U; ++u) in_all = in_all && static_cast<uint64_t>(tt[u] - shift - cur_lo) < freq;
        if (__all_sync(FULL, in_all)) {
          const bool want_mm = (a.agg_mask & (AGG_MIN | AGG_MAX)) != 0;
#pragma unroll
          for (int u = 0; u < RS_U; ++u) part.add_row_nb(vv[u], want_mm);
          run_last = static_cast<uint32_t>(g0) + 32u * RS_U - 1u;
          continue;
        }
      } else {
#pragma unroll
      for (int u = 0; u < RS_U; ++u) {
        const int64_t row = g0 + u * 32 + lane;
        tt[u] = INT64_MAX;
        vv[u] = 0;
        ok[u] = false;
        if (full || row < row_end) {
          tt[u] = a.ts[row];
          if (fast_vals) {
            vv[u] = static_cast<const uint64_t*>(a.vals)[row];
            ok[u] = true;
          } else if (a.vals) {
            vv[u] = load_wide_rt<VC>(a.vals, row, a.vw);
            ok[u] = a.vvalid ? bit_at(a.vvalid, a.voff + row) : true;
          }
        }
      }
      }
#pragma unroll
      for (int u = 0; u < RS_U; ++u) {
        const int64_t r0 = g0 + u * 32;
        if (!full && r0 >= row_end) break;
        const int64_t tp = tt[u] - shift;
        // one unsigned compare: ts' in [cur_lo, cur_lo + freq)
        const bool in0 = static_cast<uint64_t>(tp - cur_lo) < freq && have_run;
        if (full && __all_sync(FULL, in0)) {
          // common case: the whole batch continues the open run; no shuffles, no division
          if (fast_vals || ok[u]) part.add_row(vv[u], a.agg_mask);
          run_last = static_cast<uint32_t>(r0) + 31u;
          continue;
        }
        const bool active = full || (r0 + lane < row_end);
        uint32_t todo = __ballot_sync(FULL, active);
        while (todo) {
          const uint32_t in_cur = __ballot_sync(FULL, active && have_run && static_cast<uint64_t>(tp - cur_lo) < freq) & todo;
          if (in_cur) {
            if (((in_cur >> lane) & 1u) && ok[u]) part.add_row(vv[u], a.agg_mask);
            if (run_first == kNoRow) run_first = static_cast<uint32_t>(r0) + (__ffs(in_cur) - 1);
            run_last = static_cast<uint32_t>(r0) + (31 - __clz(in_cur));
            todo &= ~in_cur;
          }
          if (todo) {
            // the lowest remaining row opens a new run: close the current one, locate the next bucket
            flush(false);
            part.reset();
            run_first = kNoRow;
            run_last = 0;
            const int src = __ffs(todo) - 1;
            const int64_t t0 = __shfl_sync(FULL, tp, src);
            int64_t lo_b = 0;
            uint64_t w_b = 0;
            const int64_t b = rs_locate(sp, t0, &lo_b, &w_b);
            if (b < 0 || b <= cur_b) {
              // outside the anchored range, or an earlier bucket again: the input was not sorted
              unsorted = true;
              have_run = false;
              todo &= ~(1u << src);
              continue;
            }
            cur_b = b;
            cur_lo = lo_b;
            freq = w_b;
            have_run = true;
          }
        }
      }
    }
    // close the last run of the chunk
    flush(true);
    if (lane == 0) {
      // mark the slots this chunk did not use as empty
      if (!wrote_first) a.bnd[c * 2].bucket = -1;
      // slot 1 is used only when the chunk's last run is not also its first run
      // (flush(true) with first_seg == true wrote slot 0)
      if (unsorted) atomicExch(a.status + ST_UNSORTED, 1u);
    }
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
    if (pb > b) atomicExch(a.status + ST_UNSORTED, 1u);   // bucket sequence decreases across a chunk border
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
