// High-cardinality path: open-addressing hash table in global memory, accumulated with L2
// atomics (stage 2 "spill" target of the north-star design), then compacted, ordered by first
// appearance and gathered into a GroupResult.
//
// Slot layout is array-of-structs so that one row touches one 32-byte DRAM sector (narrow
// aggregates: sum/mean/count/first/last) or two (wide: + min/max/dsum):
//   Slot32 { key, first_row, last_row, sum, count }            32 B
//   Slot64 { Slot32, min_ord, max_ord, dsum }                  64 B
// Replaces arrow::compute::Grouper::Consume + per-group CallFunction
// (/root/reference/src/dataframe.cpp:1582-1584, pd_core_macros.h:114-147).
#pragma once
#include "group_result.cuh"

namespace pa {

struct __align__(32) Slot32 {
  uint64_t key;
  uint32_t first_row;
  uint32_t last_row;
  uint64_t sum;
  uint32_t count;
  uint32_t pad;
};
struct __align__(64) Slot64 {
  Slot32 b;
  uint64_t min_ord;
  uint64_t max_ord;
  double dsum;
  uint64_t pad;
};

template <bool WIDE>
struct SlotOf { using type = Slot32; };
template <>
struct SlotOf<true> { using type = Slot64; };

struct GScanArgs {
  const void* keys;        // key column (already offset), width kw bytes
  const void* vals;        // value column (already offset), width vw bytes; may be null (keys only)
  const uint8_t* kvalid;   // key validity bitmap or null
  const uint8_t* vvalid;   // value validity bitmap or null
  int64_t koff, voff;      // bit offsets into the bitmaps
  int64_t n;
  int kw, vw;
  void* table;             // cap + 2 slots; [cap] = null-key group, [cap+1] = key == kEmptyKey group
  uint64_t cap_mask;       // cap - 1 (cap is a power of two)
  int shift;               // 64 - log2(cap)
  uint32_t* status;
  uint32_t agg_mask;
};

template <bool WIDE>
__global__ void __launch_bounds__(256) k_gtable_init(typename SlotOf<WIDE>::type* table, uint64_t nslots) {
  uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (; i < nslots; i += stride) {
    // two (four) 16-byte stores per slot
    ulonglong2* p = reinterpret_cast<ulonglong2*>(table + i);
    p[0] = make_ulonglong2(kEmptyKey, (static_cast<uint64_t>(0u) << 32) | kNoRow);  // key ; first_row=NoRow,last_row=0
    p[1] = make_ulonglong2(0ull, 0ull);                                             // sum ; count,pad
    if constexpr (WIDE) {
      p[2] = make_ulonglong2(kMinInit, kMaxInit);
      p[3] = make_ulonglong2(0ull, 0ull);
    }
  }
}

__device__ __forceinline__ uint64_t load_key_rt(const void* p, int64_t i, int kw) {
  return kw == 8 ? static_cast<const uint64_t*>(p)[i] : static_cast<uint64_t>(static_cast<const uint32_t*>(p)[i]);
}

template <int VC>
__device__ __forceinline__ uint64_t load_wide_rt(const void* p, int64_t i, int vw) {
  switch (vw) {
    case 8: return load_wide<VC, 8>(p, i);
    case 4: return load_wide<VC, 4>(p, i);
    case 2: if constexpr (VC != VC_F) return load_wide<VC, 2>(p, i); else return 0;
    default: if constexpr (VC != VC_F) return load_wide<VC, 1>(p, i); else return 0;
  }
}

// Slot of a key: Fibonacci hashing (one 64-bit multiply, top bits).
__device__ __forceinline__ uint64_t gtable_home(uint64_t key, int shift) { return (key * 0x9E3779B97F4A7C15ull) >> shift; }

struct GProbe {
  uint64_t slot;    // ~0 = overflow
  uint32_t first;   // first_row / last_row of the slot as seen by the probe (stale reads are safe:
  uint32_t last;    // first_row only decreases, last_row only increases)
};

// One probe step: a single 16-byte L2 load returns the slot's key and its {first_row, last_row}.
__device__ __forceinline__ ulonglong2 gtable_peek(const void* table, uint64_t slot, int slot_log2) {
  return __ldcg(reinterpret_cast<const ulonglong2*>(static_cast<const char*>(table) + (slot << slot_log2)));
}

// Continue a probe sequence whose first load `kf` (slot `s`) is already in registers.
__device__ __forceinline__ GProbe gtable_find_or_insert(void* table, uint64_t cap_mask, int slot_log2, uint64_t key,
                                                        uint64_t s, ulonglong2 kf) {
  const uint32_t max_probe = cap_mask + 1 < 4096 ? static_cast<uint32_t>(cap_mask + 1) : 4096u;
  for (uint32_t probe = 0; probe < max_probe; ++probe) {
    if (kf.x == key) return GProbe{s, static_cast<uint32_t>(kf.y), static_cast<uint32_t>(kf.y >> 32)};
    if (kf.x == kEmptyKey) {
      unsigned long long* kp = reinterpret_cast<unsigned long long*>(static_cast<char*>(table) + (s << slot_log2));
      const uint64_t old = atomicCAS(kp, static_cast<unsigned long long>(kEmptyKey), static_cast<unsigned long long>(key));
      if (old == kEmptyKey || old == key) return GProbe{s, kNoRow, 0u};
    }
    s = (s + 1) & cap_mask;
    kf = gtable_peek(table, s, slot_log2);
  }
  return GProbe{~0ull, kNoRow, 0u};
}

template <int VC, bool WIDE>
__device__ __forceinline__ void gtable_accumulate(typename SlotOf<WIDE>::type* slot, const GProbe& pr, uint32_t row,
                                                  bool vvalid, uint64_t vbits, uint32_t agg_mask) {
  Slot32* b = reinterpret_cast<Slot32*>(slot);
  if (row < pr.first) atomicMin(&b->first_row, row);
  if ((agg_mask & AGG_LAST) && row > pr.last) atomicMax(&b->last_row, row);
  if (!vvalid) return;
  atomicAdd(&b->count, 1u);
  if constexpr (VC == VC_F) {
    atomicAdd(reinterpret_cast<double*>(&b->sum), __longlong_as_double(static_cast<long long>(vbits)));
  } else {
    atomicAdd(reinterpret_cast<unsigned long long*>(&b->sum), static_cast<unsigned long long>(vbits));
  }
  if constexpr (WIDE) {
    Slot64* w = reinterpret_cast<Slot64*>(slot);
    if constexpr (VC != VC_F) {
      if (agg_mask & AGG_MEAN) atomicAdd(&w->dsum, Wide<VC>::as_double(vbits));
    }
    if ((agg_mask & (AGG_MIN | AGG_MAX)) && !Wide<VC>::is_nan(vbits)) {
      const uint64_t o = Wide<VC>::ord(vbits);
      const ulonglong2 mm = __ldcg(reinterpret_cast<const ulonglong2*>(&w->min_ord));
      if (o < mm.x) atomicMin(reinterpret_cast<unsigned long long*>(&w->min_ord), static_cast<unsigned long long>(o));
      if (o > mm.y) atomicMax(reinterpret_cast<unsigned long long*>(&w->max_ord), static_cast<unsigned long long>(o));
    }
  }
}

// One thread per row, R rows per thread per tile: all key / value loads of the tile are issued
// first, then the R first-probe loads (one 16-byte L2 transaction each), and only then the
// dependent work, so R independent L2 round trips are in flight per thread.  Per row in steady
// state: one L2 load + two L2 reductions (RED.ADD.F64 sum, RED.ADD.U32 count); first/last/min/max
// are pre-checked against the loaded slot and only rarely issue an atomic.
// FAST: 8-byte keys and values, no validity bitmaps; thread t of a tile takes the adjacent rows
// 2t, 2t+1 (128-bit loads) — row order is irrelevant here, first/last rows are atomicMin/Max.
template <int VC, bool WIDE, bool FAST>
__global__ void __launch_bounds__(256, 4) k_gtable_scan(GScanArgs a) {
  using SlotT = typename SlotOf<WIDE>::type;
  constexpr int SLOT_LOG2 = WIDE ? 6 : 5;
  SlotT* table = static_cast<SlotT*>(a.table);
  const uint64_t cap = a.cap_mask + 1;
  constexpr int R = 4;
  const int64_t tile_rows = static_cast<int64_t>(blockDim.x) * R;
  const int64_t ntiles = (a.n + tile_rows - 1) / tile_rows;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    if (*reinterpret_cast<volatile uint32_t*>(a.status + ST_OVERFLOW)) return;
    const int64_t tile0 = t * tile_rows;
    uint64_t key[R], vb[R];
    int64_t row[R];
    bool act[R], kv[R], vv[R];
    if (FAST && tile0 + tile_rows <= a.n) {
#pragma unroll
      for (int h = 0; h < R / 2; ++h) {
        const int64_t r0 = tile0 + static_cast<int64_t>(h) * blockDim.x * 2 + 2 * threadIdx.x;
        const ulonglong2 k2 = ldg_stream_u64x2(static_cast<const uint64_t*>(a.keys) + r0);
        ulonglong2 v2 = make_ulonglong2(0ull, 0ull);
        if (a.vals) v2 = ldg_stream_u64x2(static_cast<const uint64_t*>(a.vals) + r0);
        key[2 * h] = k2.x; key[2 * h + 1] = k2.y;
        vb[2 * h] = v2.x; vb[2 * h + 1] = v2.y;
        row[2 * h] = r0; row[2 * h + 1] = r0 + 1;
        act[2 * h] = act[2 * h + 1] = true;
        kv[2 * h] = kv[2 * h + 1] = true;
        vv[2 * h] = vv[2 * h + 1] = a.vals != nullptr;
      }
    } else {
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int64_t i = tile0 + threadIdx.x + static_cast<int64_t>(j) * blockDim.x;
        row[j] = i;
        act[j] = i < a.n;
        key[j] = 0; vb[j] = 0; kv[j] = true; vv[j] = false;
        if (act[j]) {
          key[j] = load_key_rt(a.keys, i, a.kw);
          if (a.kvalid) kv[j] = bit_at(a.kvalid, a.koff + i);
          if (a.vals) {
            vb[j] = load_wide_rt<VC>(a.vals, i, a.vw);
            vv[j] = a.vvalid ? bit_at(a.vvalid, a.voff + i) : true;
          }
        }
      }
    }
    uint64_t s[R];
    ulonglong2 kf[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {
      s[j] = gtable_home(key[j], a.shift);
      if (!kv[j]) s[j] = cap;
      else if (key[j] == kEmptyKey) s[j] = cap + 1;
      kf[j] = make_ulonglong2(0ull, 0ull);
      if (act[j]) kf[j] = gtable_peek(table, s[j], SLOT_LOG2);
    }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      if (!act[j]) continue;
      GProbe pr;
      if (s[j] >= cap) pr = GProbe{s[j], static_cast<uint32_t>(kf[j].y), static_cast<uint32_t>(kf[j].y >> 32)};
      else pr = gtable_find_or_insert(table, a.cap_mask, SLOT_LOG2, key[j], s[j], kf[j]);
      if (pr.slot == ~0ull) { atomicExch(a.status + ST_OVERFLOW, 1u); return; }
      gtable_accumulate<VC, WIDE>(table + pr.slot, pr, static_cast<uint32_t>(row[j]), vv[j], vb[j], a.agg_mask);
    }
  }
}

// Append (first_row, slot) of every occupied slot; order is irrelevant (sorted afterwards).
template <bool WIDE>
__global__ void __launch_bounds__(256) k_gtable_compact(const typename SlotOf<WIDE>::type* table, uint64_t nslots,
                                                        uint32_t* out_first, uint32_t* out_slot, uint32_t* status) {
  uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  const uint64_t nround = (nslots + stride - 1) / stride * stride;
  for (; i < nround; i += stride) {
    uint32_t f = kNoRow;
    if (i < nslots) f = reinterpret_cast<const Slot32*>(table + i)->first_row;
    const bool occ = f != kNoRow;
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, occ);
    if (m) {
      uint32_t basepos = 0;
      if (lane_id() == 0) basepos = atomicAdd(status + ST_COUNTER, __popc(m));
      basepos = __shfl_sync(0xFFFFFFFFu, basepos, 0);
      if (occ) {
        const uint32_t pos = basepos + __popc(m & ((1u << lane_id()) - 1u));
        out_first[pos] = f;
        out_slot[pos] = static_cast<uint32_t>(i);
      }
    }
  }
}

// order[r] = slot index of the r-th group in first-appearance order -> GroupResult
template <bool WIDE>
__global__ void __launch_bounds__(256) k_gtable_gather(const typename SlotOf<WIDE>::type* table, uint64_t cap,
                                                       const uint32_t* order, uint32_t G, GroupResult r) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const uint32_t s = order[g];
  const auto* slot = table + s;
  const Slot32* b = reinterpret_cast<const Slot32*>(slot);
  r.key[g] = b->key;
  r.key_kind[g] = (s == cap) ? KK_NULL : KK_REGULAR;
  if (s == cap + 1) r.key[g] = kEmptyKey;
  r.sum[g] = b->sum;
  r.count[g] = b->count;
  r.first_row[g] = b->first_row;
  r.last_row[g] = b->last_row;
  if constexpr (WIDE) {
    const Slot64* w = reinterpret_cast<const Slot64*>(slot);
    r.min_ord[g] = w->min_ord;
    r.max_ord[g] = w->max_ord;
    if (r.dsum) r.dsum[g] = w->dsum;
  }
}

}  // namespace pa
