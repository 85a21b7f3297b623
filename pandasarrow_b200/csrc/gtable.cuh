// High-cardinality path: open-addressing hash table in global memory, accumulated with L2
// atomics (stage 2 "spill" target of the north-star design), then compacted, ordered by first
// appearance and gathered into a GroupResult.
//
// Slot layout is array-of-structs in whole 32-byte DRAM/L2 sectors:
//   Slot32 { key, first_row, last_row, sum, count }                                   one sector
//   Slot64 { key, first_row, last_row, min_ord, max_ord | sum, count, -, dsum, - }    two sectors
// A probe is ONE 256-bit load (LDG.E.256) of the first sector: it returns the key to compare and
// everything the rarely-firing atomics are pre-checked against (first/last row, min, max); the
// per-row reductions (sum, count[, dsum]) go to the same (narrow) or the second (wide) sector.
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
  uint64_t key;
  uint32_t first_row;
  uint32_t last_row;
  uint64_t min_ord;
  uint64_t max_ord;
  uint64_t sum;
  uint32_t count;
  uint32_t pad;
  double dsum;
  uint64_t pad2;
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
  const uint32_t* rowids;  // partitioned input: original row number of every (permuted) row, or null
  KeyRange* krange;        // shared-memory front table: sampled key range (dense-mode decision), or null
  int64_t koff, voff;      // bit offsets into the bitmaps
  int64_t n;
  int kw, vw;
  void* table;             // cap + 2 slots; [cap] = null-key group, [cap+1] = key == kEmptyKey group
  uint64_t cap_mask;       // cap - 1 (cap is a power of two)
  int shift;               // 64 - log2(cap)
  uint32_t* status;
  uint32_t agg_mask;
  uint32_t sm_max_keys;    // shared-memory front table, hash mode: keys admitted before the rest spills
};

template <bool WIDE>
__global__ void __launch_bounds__(256) k_gtable_init(typename SlotOf<WIDE>::type* table, uint64_t nslots) {
  uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (; i < nslots; i += stride) {
    // two (four) 16-byte stores per slot
    ulonglong2* p = reinterpret_cast<ulonglong2*>(table + i);
    p[0] = make_ulonglong2(kEmptyKey, (static_cast<uint64_t>(0u) << 32) | kNoRow);  // key ; first_row=NoRow,last_row=0
    if constexpr (WIDE) {
      p[1] = make_ulonglong2(kMinInit, kMaxInit);
      p[2] = make_ulonglong2(0ull, 0ull);                                           // sum ; count,pad
      p[3] = make_ulonglong2(0ull, 0ull);                                           // dsum ; pad
    } else {
      p[1] = make_ulonglong2(0ull, 0ull);                                           // sum ; count,pad
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

// Slot of a key: Fibonacci hashing (one 64-bit multiply, top bits).  Chosen on purpose: arithmetic
// progressions of keys (auto-increment ids, dictionary codes, fixed-interval timestamps — and the
// benchmark's dense keys) spread over the table without a single collision, so a lookup is exactly
// one probe; arbitrary 64-bit keys behave like any random hash (1.5 probes at load 1/2).  A stronger
// mix would make the common dense case pay the random-hash collision rate (measured: 22 -> 55 ms per
// 1B rows at 1M groups).
__device__ __forceinline__ uint64_t gtable_mix(uint64_t key) { return key * 0x9E3779B97F4A7C15ull; }
__device__ __forceinline__ uint64_t gtable_home(uint64_t key, int shift) { return gtable_mix(key) >> shift; }

// First sector of a slot as one 256-bit L2 load.
struct Sector {
  uint64_t key, fl, mn, mx;   // fl = first_row | last_row << 32; mn / mx only meaningful for wide slots
};
template <bool WIDE>
__device__ __forceinline__ Sector gtable_peek(const void* table, uint64_t slot, int slot_log2) {
  Sector r;
  const char* p = static_cast<const char*>(table) + (slot << slot_log2);
  if constexpr (WIDE) {
    asm volatile("ld.global.cg.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(r.key), "=l"(r.fl), "=l"(r.mn), "=l"(r.mx) : "l"(p));
  } else {   // narrow slots: key + first/last are all a probe needs (fewer registers)
    asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(r.key), "=l"(r.fl) : "l"(p));
    r.mn = 0; r.mx = 0;
  }
  return r;
}

struct GProbe {
  uint64_t slot;    // ~0 = overflow
  uint32_t first;   // first_row / last_row / min / max of the slot as seen by the probe (stale reads are
  uint32_t last;    // safe: first_row and min only decrease, last_row and max only increase)
  uint64_t mn, mx;
};

// Continue a probe sequence whose first load `kf` (slot `s`) is already in registers.
template <bool WIDE>
__device__ __forceinline__ GProbe gtable_find_or_insert(void* table, uint64_t cap_mask, int slot_log2, uint64_t key,
                                                        uint64_t s, Sector kf) {
  const uint32_t max_probe = cap_mask + 1 < 4096 ? static_cast<uint32_t>(cap_mask + 1) : 4096u;
  for (uint32_t probe = 0; probe < max_probe; ++probe) {
    if (kf.key == key) return GProbe{s, static_cast<uint32_t>(kf.fl), static_cast<uint32_t>(kf.fl >> 32), kf.mn, kf.mx};
    if (kf.key == kEmptyKey) {
      unsigned long long* kp = reinterpret_cast<unsigned long long*>(static_cast<char*>(table) + (s << slot_log2));
      const uint64_t old = atomicCAS(kp, static_cast<unsigned long long>(kEmptyKey), static_cast<unsigned long long>(key));
      if (old == kEmptyKey || old == key) return GProbe{s, kNoRow, 0u, kMinInit, kMaxInit};
    }
    s = (s + 1) & cap_mask;
    kf = gtable_peek<WIDE>(table, s, slot_log2);
  }
  return GProbe{~0ull, kNoRow, 0u, kMinInit, kMaxInit};
}

template <int VC, bool WIDE>
__device__ __forceinline__ void gtable_accumulate(typename SlotOf<WIDE>::type* slot, const GProbe& pr, uint32_t row,
                                                  bool vvalid, uint64_t vbits, uint32_t agg_mask) {
  if (row < pr.first) atomicMin(&slot->first_row, row);
  if ((agg_mask & AGG_LAST) && row > pr.last) atomicMax(&slot->last_row, row);
  if (!vvalid) return;
  atomicAdd(&slot->count, 1u);
  if constexpr (VC == VC_F) {
    atomicAdd(reinterpret_cast<double*>(&slot->sum), __longlong_as_double(static_cast<long long>(vbits)));
  } else {
    atomicAdd(reinterpret_cast<unsigned long long*>(&slot->sum), static_cast<unsigned long long>(vbits));
  }
  if constexpr (WIDE) {
    if constexpr (VC != VC_F) {
      if (agg_mask & AGG_MEAN) atomicAdd(&slot->dsum, Wide<VC>::as_double(vbits));
    }
    if ((agg_mask & (AGG_MIN | AGG_MAX)) && !Wide<VC>::is_nan(vbits)) {
      const uint64_t o = Wide<VC>::ord(vbits);
      if (o < pr.mn) atomicMin(reinterpret_cast<unsigned long long*>(&slot->min_ord), static_cast<unsigned long long>(o));
      if (o > pr.mx) atomicMax(reinterpret_cast<unsigned long long*>(&slot->max_ord), static_cast<unsigned long long>(o));
    }
  }
}

constexpr int GT_R = 4;   // rows per thread per tile

// Rows of one tile for this thread.  FAST: 8-byte non-null keys and 8-byte values (a value validity bitmap
// is allowed): thread t takes
// the adjacent rows 4t .. 4t+3 (256-bit non-allocating, L2 evict-first loads) — row order is irrelevant on this path,
// first/last rows are atomicMin/Max.
struct GRows {
  uint64_t key[GT_R], vb[GT_R];
  int64_t row[GT_R];
  bool act[GT_R], kv[GT_R], vv[GT_R];
};

template <int VC, bool FAST>
__device__ __forceinline__ void gtile_load(const GScanArgs& a, int64_t tile0, int64_t tile_rows, GRows& r) {
  constexpr int R = GT_R;
  if (FAST && tile0 + tile_rows <= a.n) {
    static_assert(GT_R == 4, "one 256-bit load = 4 rows");
    const int64_t r0 = tile0 + 4 * static_cast<int64_t>(threadIdx.x);
    const u64x4 k4 = ldg_stream_u64x4(static_cast<const uint64_t*>(a.keys) + r0);
    u64x4 v4{0ull, 0ull, 0ull, 0ull};
    if (a.vals) v4 = ldg_stream_u64x4(static_cast<const uint64_t*>(a.vals) + r0);
    r.key[0] = k4.a; r.key[1] = k4.b; r.key[2] = k4.c; r.key[3] = k4.d;
    r.vb[0] = v4.a; r.vb[1] = v4.b; r.vb[2] = v4.c; r.vb[3] = v4.d;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      r.row[j] = r0 + j;
      r.act[j] = true;
      r.kv[j] = true;
      r.vv[j] = a.vals != nullptr;
    }
    if (a.rowids) {
      const uint4 ri = *reinterpret_cast<const uint4*>(a.rowids + r0);
      r.row[0] = ri.x; r.row[1] = ri.y; r.row[2] = ri.z; r.row[3] = ri.w;
    }
    if (a.vvalid && !a.rowids) {   // value validity: the four adjacent bits of this thread's rows
      const int64_t b0 = a.voff + r0;
      const uint32_t w = static_cast<uint32_t>(a.vvalid[b0 >> 3]) | (static_cast<uint32_t>(a.vvalid[(b0 + 3) >> 3]) << 8);
      const uint32_t sh = static_cast<uint32_t>(b0 & 7);
      const uint32_t bits = (((b0 + 3) >> 3) != (b0 >> 3)) ? (w >> sh) : (static_cast<uint32_t>(a.vvalid[b0 >> 3]) >> sh);
#pragma unroll
      for (int j = 0; j < R; ++j) r.vv[j] = (bits >> j) & 1u;
    }
  } else {
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int64_t i = tile0 + threadIdx.x + static_cast<int64_t>(j) * blockDim.x;
      r.row[j] = i;
      r.act[j] = i < a.n;
      if (a.rowids && r.act[j]) r.row[j] = a.rowids[i];
      r.key[j] = 0; r.vb[j] = 0; r.kv[j] = true; r.vv[j] = false;
      if (r.act[j]) {
        r.key[j] = load_key_rt(a.keys, i, a.kw);
        if (a.kvalid) r.kv[j] = bit_at(a.kvalid, a.koff + i);
        if (a.vals) {
          r.vb[j] = load_wide_rt<VC>(a.vals, i, a.vw);
          r.vv[j] = a.vvalid ? bit_at(a.vvalid, a.voff + i) : true;
        }
      }
    }
  }
}

// Find-or-insert + accumulate for the rows of `todo` (bit j = row j), all probe sequences advancing in
// lock step: every round first examines the sectors already loaded, then issues the next probe loads
// of all unresolved rows together, so colliding rows overlap their L2 round trips instead of
// serialising them.  Special keys (null / sentinel) are expected at their dedicated slots already.
// Returns false when a probe sequence ran out of budget (table too full).
template <int VC, bool WIDE>
__device__ __forceinline__ bool gtable_probe_accumulate(typename SlotOf<WIDE>::type* table, const GScanArgs& a, const GRows& r,
                                                        uint32_t todo) {
  constexpr int R = GT_R;
  constexpr int SLOT_LOG2 = WIDE ? 6 : 5;
  const uint64_t cap = a.cap_mask + 1;
  uint64_t s[R];
  Sector kf[R];
#pragma unroll
  for (int j = 0; j < R; ++j) {
    s[j] = gtable_home(r.key[j], a.shift);
    if (!r.kv[j]) s[j] = cap;
    else if (r.key[j] == kEmptyKey) s[j] = cap + 1;
    kf[j] = Sector{0ull, 0ull, 0ull, 0ull};
    if ((todo >> j) & 1u) kf[j] = gtable_peek<WIDE>(table, s[j], SLOT_LOG2);
  }
  const uint32_t max_probe = cap < 4096 ? static_cast<uint32_t>(cap) : 4096u;
  for (uint32_t round = 0; todo; ++round) {
    if (round > max_probe) return false;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      if (!((todo >> j) & 1u)) continue;
      bool hit = s[j] >= cap || kf[j].key == r.key[j];
      bool fresh = false;
      if (!hit && kf[j].key == kEmptyKey) {
        unsigned long long* kp = reinterpret_cast<unsigned long long*>(table + s[j]);
        const uint64_t old = atomicCAS(kp, static_cast<unsigned long long>(kEmptyKey), static_cast<unsigned long long>(r.key[j]));
        hit = old == kEmptyKey || old == r.key[j];
        fresh = hit;
      }
      if (hit) {
        GProbe pr{s[j], static_cast<uint32_t>(kf[j].fl), static_cast<uint32_t>(kf[j].fl >> 32), kf[j].mn, kf[j].mx};
        if (fresh) pr = GProbe{s[j], kNoRow, 0u, kMinInit, kMaxInit};
        gtable_accumulate<VC, WIDE>(table + s[j], pr, static_cast<uint32_t>(r.row[j]), r.vv[j], r.vb[j], a.agg_mask);
        todo &= ~(1u << j);
      } else {
        s[j] = (s[j] + 1) & a.cap_mask;
      }
    }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      if ((todo >> j) & 1u) kf[j] = gtable_peek<WIDE>(table, s[j], SLOT_LOG2);
    }
  }
  return true;
}

// Pure global-table scan.  One thread per row, GT_R rows per thread per tile: all key / value loads of
// the tile are issued first, then the GT_R first-probe loads (one 32-byte L2 transaction each), and
// only then the dependent work, so GT_R independent L2 round trips are in flight per thread.  Per row
// in steady state: one L2 load + two L2 reductions (RED.ADD.F64 sum, RED.ADD.U32 count);
// first/last/min/max are pre-checked against the loaded sector and only rarely issue an atomic.
template <int VC, bool WIDE, bool FAST>
__global__ void __launch_bounds__(256, 3) k_gtable_scan(GScanArgs a) {
  using SlotT = typename SlotOf<WIDE>::type;
  SlotT* table = static_cast<SlotT*>(a.table);
  constexpr int R = GT_R;
  const int64_t tile_rows = static_cast<int64_t>(blockDim.x) * R;
  const int64_t ntiles = (a.n + tile_rows - 1) / tile_rows;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    if (*reinterpret_cast<volatile uint32_t*>(a.status + ST_OVERFLOW)) return;
    GRows r;
    gtile_load<VC, FAST>(a, t * tile_rows, tile_rows, r);
    uint32_t todo = 0;
#pragma unroll
    for (int j = 0; j < R; ++j) todo |= r.act[j] ? (1u << j) : 0u;
    if (!gtable_probe_accumulate<VC, WIDE>(table, a, r, todo)) { atomicExch(a.status + ST_OVERFLOW, 1u); return; }
  }
}

// ---------------------------------------------------------------------------------------------
// Shared-memory front table ("mid cardinality"): every CTA aggregates its rows into a CTA-shared
// open-addressing table in shared memory with shared-memory atomics — ATOMS.CAS.64 to claim a key,
// the compiler's ATOMS.CAST.SPIN loop for the f64/u64 sum, native ATOMS for count / first / last —
// and only at the end flushes its groups into the global table with L2 atomics.  Rows whose key
// finds no room (table 3/4 full, or a probe sequence longer than SM_MAX_PROBE) SPILL straight to
// the global table, so any cardinality is handled: up to a few thousand groups nothing spills and
// L2 sees G x #CTAs updates instead of one per row; beyond that the kernel degrades towards
// k_gtable_scan.  Measured (scripts/ubench/smem_atomics.cu): 217 G updates/s chip-wide for
// {f64 add + u32 add} on 4096-12288 random slots with 1024 threads per SM, against 45-55 G rows/s
// for the L2 path.
// ---------------------------------------------------------------------------------------------
constexpr int SM_THREADS = 768;
constexpr int SM_SAMPLE_GRID = 64;   // x 256 threads = 16384 sampled keys

__global__ void __launch_bounds__(256) k_key_range(GScanArgs a) {
  key_range_sample(a.keys, a.kvalid, a.koff, a.kw, a.n, blockIdx.x * 256u + threadIdx.x, gridDim.x * 256u, a.krange);
}
constexpr int SM_MAX_PROBE = 16;

template <int VC, bool WIDE>
struct SmTab {
  static constexpr bool DSUM = WIDE && VC != VC_F;
  static constexpr int CAP_LOG2 = WIDE ? 12 : 13;
  static constexpr int CAP = 1 << CAP_LOG2;
  static constexpr int NSLOT = CAP + 2;                       // + null-key group + (key == kEmptyKey) group
  static constexpr int MAX_KEYS = CAP / 4 * 3;
  // byte offsets (8-byte arrays first)
  static constexpr size_t OFF_KEY = 0;
  static constexpr size_t OFF_SUM = OFF_KEY + sizeof(uint64_t) * NSLOT;
  static constexpr size_t OFF_MN = OFF_SUM + sizeof(uint64_t) * NSLOT;
  static constexpr size_t OFF_MX = OFF_MN + (WIDE ? sizeof(uint64_t) * NSLOT : 0);
  static constexpr size_t OFF_DSUM = OFF_MX + (WIDE ? sizeof(uint64_t) * NSLOT : 0);
  static constexpr size_t OFF_CNT = OFF_DSUM + (DSUM ? sizeof(double) * NSLOT : 0);
  static constexpr size_t OFF_FIRST = OFF_CNT + sizeof(uint32_t) * NSLOT;
  static constexpr size_t OFF_LAST = OFF_FIRST + sizeof(uint32_t) * NSLOT;
  static constexpr size_t OFF_MISC = OFF_LAST + (WIDE ? sizeof(uint32_t) * NSLOT : 0);
  static constexpr size_t TOTAL = OFF_MISC + 16;
};

template <int VC, bool WIDE, bool FAST>
__global__ void __launch_bounds__(SM_THREADS, 1) k_smemtab_scan(GScanArgs a) {
  using SlotT = typename SlotOf<WIDE>::type;
  using T = SmTab<VC, WIDE>;
  constexpr int SLOT_LOG2 = WIDE ? 6 : 5;
  extern __shared__ __align__(16) unsigned char st_smem[];
  unsigned long long* s_key = reinterpret_cast<unsigned long long*>(st_smem + T::OFF_KEY);
  unsigned long long* s_sum = reinterpret_cast<unsigned long long*>(st_smem + T::OFF_SUM);
  unsigned long long* s_mn = reinterpret_cast<unsigned long long*>(st_smem + T::OFF_MN);
  unsigned long long* s_mx = reinterpret_cast<unsigned long long*>(st_smem + T::OFF_MX);
  double* s_dsum = reinterpret_cast<double*>(st_smem + T::OFF_DSUM);
  uint32_t* s_cnt = reinterpret_cast<uint32_t*>(st_smem + T::OFF_CNT);
  uint32_t* s_first = reinterpret_cast<uint32_t*>(st_smem + T::OFF_FIRST);
  uint32_t* s_last = reinterpret_cast<uint32_t*>(st_smem + T::OFF_LAST);
  uint32_t* s_misc = reinterpret_cast<uint32_t*>(st_smem + T::OFF_MISC);   // [0] keys in the table
  SlotT* table = static_cast<SlotT*>(a.table);
  const uint64_t cap = a.cap_mask + 1;
  for (int i = threadIdx.x; i < T::NSLOT; i += SM_THREADS) {
    s_key[i] = kEmptyKey;
    s_sum[i] = 0ull;
    s_cnt[i] = 0u;
    s_first[i] = kNoRow;
    if constexpr (WIDE) { s_mn[i] = kMinInit; s_mx[i] = kMaxInit; s_last[i] = 0u; }
    if constexpr (T::DSUM) s_dsum[i] = 0.0;
  }
  if (threadIdx.x < 4) s_misc[threadIdx.x] = 0u;
  __syncthreads();
  // Dense mode (decided on the device from the key sample): the sampled keys span fewer values than the
  // table has slots -> slot = key - base, no key array, no probing, every slot usable.  A key outside the
  // window simply spills to the global table.
  uint64_t base = 0, span = 0;
  const bool dense = a.krange && key_range_get(a.krange, &base, &span) && span < static_cast<uint64_t>(T::CAP);
  if (dense) base -= (static_cast<uint64_t>(T::CAP) - (span + 1)) / 2;   // centre the window on the sample

  constexpr int R = GT_R;
  const int64_t tile_rows = static_cast<int64_t>(SM_THREADS) * R;
  const int64_t ntiles = (a.n + tile_rows - 1) / tile_rows;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    if (*reinterpret_cast<volatile uint32_t*>(a.status + ST_OVERFLOW)) break;
    GRows r;
    gtile_load<VC, FAST>(a, t * tile_rows, tile_rows, r);
    uint32_t spill = 0;   // bit j: row j found no room in shared memory
#pragma unroll
    for (int j = 0; j < R; ++j) {
      if (!r.act[j]) continue;
      const uint64_t key = r.key[j];
      const uint32_t row = static_cast<uint32_t>(r.row[j]);
      // ---- find or insert in the shared-memory table ----
      uint32_t s;
      bool found = false;
      if (!r.kv[j]) { s = T::CAP; found = true; }
      else if (dense) {
        const uint64_t d = key - base;
        s = static_cast<uint32_t>(d);
        found = d < static_cast<uint64_t>(T::CAP);
      }
      else if (key == kEmptyKey) { s = T::CAP + 1; found = true; }
      else {
        s = static_cast<uint32_t>(gtable_mix(key) >> (64 - T::CAP_LOG2));
        for (int probe = 0; probe < SM_MAX_PROBE; ++probe) {
          const uint64_t k = *reinterpret_cast<volatile unsigned long long*>(s_key + s);
          if (k == key) { found = true; break; }
          if (k == kEmptyKey) {
            if (*reinterpret_cast<volatile uint32_t*>(s_misc) >= a.sm_max_keys) break;   // table is full enough: spill
            const uint64_t old = atomicCAS(s_key + s, static_cast<unsigned long long>(kEmptyKey), static_cast<unsigned long long>(key));
            if (old == kEmptyKey) { atomicAdd(s_misc, 1u); found = true; break; }
            if (old == key) { found = true; break; }
          }
          s = (s + 1) & (T::CAP - 1);
        }
      }
      if (!found) { spill |= 1u << j; continue; }
      // ---- accumulate with shared-memory atomics ----
      if (row < *reinterpret_cast<volatile uint32_t*>(s_first + s)) atomicMin(s_first + s, row);
      if constexpr (WIDE) {
        if ((a.agg_mask & AGG_LAST) && row > *reinterpret_cast<volatile uint32_t*>(s_last + s)) atomicMax(s_last + s, row);
      }
      if (!r.vv[j]) continue;
      atomicAdd(s_cnt + s, 1u);
      if constexpr (VC == VC_F) atomicAdd(reinterpret_cast<double*>(s_sum + s), __longlong_as_double(static_cast<long long>(r.vb[j])));
      else atomicAdd(s_sum + s, static_cast<unsigned long long>(r.vb[j]));
      if constexpr (WIDE) {
        if constexpr (T::DSUM) {
          if (a.agg_mask & AGG_MEAN) atomicAdd(s_dsum + s, Wide<VC>::as_double(r.vb[j]));
        }
        if ((a.agg_mask & (AGG_MIN | AGG_MAX)) && !Wide<VC>::is_nan(r.vb[j])) {
          const uint64_t o = Wide<VC>::ord(r.vb[j]);
          if (o < *reinterpret_cast<volatile unsigned long long*>(s_mn + s)) atomicMin(s_mn + s, static_cast<unsigned long long>(o));
          if (o > *reinterpret_cast<volatile unsigned long long*>(s_mx + s)) atomicMax(s_mx + s, static_cast<unsigned long long>(o));
        }
      }
    }
    if (spill) {   // rows that found no room in shared memory: straight to the global table
      if (!gtable_probe_accumulate<VC, WIDE>(table, a, r, spill)) { atomicExch(a.status + ST_OVERFLOW, 1u); break; }
    }
  }
  __syncthreads();
  if (*reinterpret_cast<volatile uint32_t*>(a.status + ST_OVERFLOW)) return;
  // ---- flush this CTA's groups into the global table ----
  for (int i = threadIdx.x; i < T::NSLOT; i += SM_THREADS) {
    const uint32_t first = s_first[i];
    if (first == kNoRow) continue;             // empty slot (a claimed key always has a first row)
    uint64_t s;
    if (i == T::CAP) s = cap;
    else if (i == T::CAP + 1) s = cap + 1;
    uint64_t fkey = dense ? base + static_cast<uint64_t>(i) : s_key[i];
    if (i < T::CAP && fkey == kEmptyKey) s = cap + 1;     // (dense window containing the sentinel value)
    else if (i < T::CAP) s = gtable_home(fkey, a.shift);
    const Sector kf = gtable_peek<WIDE>(table, s, SLOT_LOG2);
    GProbe pr;
    if (s >= cap) pr = GProbe{s, static_cast<uint32_t>(kf.fl), static_cast<uint32_t>(kf.fl >> 32), kf.mn, kf.mx};
    else pr = gtable_find_or_insert<WIDE>(table, a.cap_mask, SLOT_LOG2, fkey, s, kf);
    if (pr.slot == ~0ull) { atomicExch(a.status + ST_OVERFLOW, 1u); return; }
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
    if (i < nslots) f = table[i].first_row;
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
  const auto* b = table + s;
  r.key[g] = b->key;
  r.key_kind[g] = (s == cap) ? KK_NULL : KK_REGULAR;
  if (s == cap + 1) r.key[g] = kEmptyKey;
  r.sum[g] = b->sum;
  r.count[g] = b->count;
  r.first_row[g] = b->first_row;
  r.last_row[g] = b->last_row;
  if constexpr (WIDE) {
    r.min_ord[g] = b->min_ord;
    r.max_ord[g] = b->max_ord;
    if (r.dsum) r.dsum[g] = b->dsum;
  }
}

}  // namespace pa
