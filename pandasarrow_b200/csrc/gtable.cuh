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

// Find-or-insert `key`; returns the slot index or ~0 on overflow (probe budget exhausted).
template <typename SlotT>
__device__ __forceinline__ uint64_t gtable_find_or_insert(SlotT* table, uint64_t cap_mask, uint64_t key) {
  uint64_t s = hash_key64(key) & cap_mask;
  const uint32_t max_probe = cap_mask + 1 < 4096 ? static_cast<uint32_t>(cap_mask + 1) : 4096u;
  for (uint32_t probe = 0; probe < max_probe; ++probe) {
    unsigned long long* kp = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(table + s));
    uint64_t k = __ldcg(kp);
    if (k == key) return s;
    if (k == kEmptyKey) {
      uint64_t old = atomicCAS(kp, static_cast<unsigned long long>(kEmptyKey), static_cast<unsigned long long>(key));
      if (old == kEmptyKey || old == key) return s;
    }
    s = (s + 1) & cap_mask;
  }
  return ~0ull;
}

template <int VC, bool WIDE>
__device__ __forceinline__ void gtable_accumulate(typename SlotOf<WIDE>::type* slot, uint32_t row, bool vvalid,
                                                  uint64_t vbits, uint32_t agg_mask) {
  Slot32* b = reinterpret_cast<Slot32*>(slot);
  // first_row / last_row share one 8-byte word with a cheap pre-check (stale reads are safe:
  // first_row only decreases, last_row only increases).
  const uint64_t fl = __ldcg(reinterpret_cast<const unsigned long long*>(&b->first_row));
  const uint32_t f = static_cast<uint32_t>(fl), l = static_cast<uint32_t>(fl >> 32);
  if (row < f) atomicMin(&b->first_row, row);
  if ((agg_mask & AGG_LAST) && row > l) atomicMax(&b->last_row, row);
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

// One thread per row, 4 rows per thread per tile so that four independent key/value loads are in
// flight before the dependent probe chain starts.  Grid: multiple of the SM count (host side).
template <int VC, bool WIDE>
__global__ void __launch_bounds__(256) k_gtable_scan(GScanArgs a) {
  using SlotT = typename SlotOf<WIDE>::type;
  SlotT* table = static_cast<SlotT*>(a.table);
  const uint64_t cap = a.cap_mask + 1;
  constexpr int R = 4;
  const int64_t tile_rows = static_cast<int64_t>(blockDim.x) * R;
  const int64_t ntiles = (a.n + tile_rows - 1) / tile_rows;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    if (*reinterpret_cast<volatile uint32_t*>(a.status + ST_OVERFLOW)) return;
    const int64_t base = t * tile_rows + threadIdx.x;
    uint64_t key[R], vb[R];
    bool act[R], kv[R], vv[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int64_t i = base + static_cast<int64_t>(j) * blockDim.x;
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
#pragma unroll
    for (int j = 0; j < R; ++j) {
      if (!act[j]) continue;
      const int64_t i = base + static_cast<int64_t>(j) * blockDim.x;
      uint64_t s;
      if (!kv[j]) s = cap;
      else if (key[j] == kEmptyKey) s = cap + 1;
      else {
        s = gtable_find_or_insert(table, a.cap_mask, key[j]);
        if (s == ~0ull) { atomicExch(a.status + ST_OVERFLOW, 1u); return; }
      }
      gtable_accumulate<VC, WIDE>(table + s, static_cast<uint32_t>(i), vv[j], vb[j], a.agg_mask);
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
