// Low-cardinality path (stages 1-3 fused): one persistent CTA per SM, accumulators privatised per
// warp in shared memory.  Two kernels: the dense-mode kernel (16 warps, no key table) and the
// hash-mode kernel (12 warps + a 40 KB CTA-shared key table).
//
//  * Input: every warp streams "row groups" of 256 rows (8 batches of 32 rows) with coalesced
//    non-allocating loads (256 contiguous bytes per warp per instruction); the next group is
//    loaded into a second register buffer before the current one is processed, so 4 KB per warp /
//    32-48 KB per SM are in flight at all times without relying on occupancy.  (The first version
//    staged the rows through shared memory with bulk async copies; ncu showed the kernel bound by
//    shared-memory wavefronts and instruction latency, so the staging traffic was removed —
//    profiles/r1_lowcard_v3_*.)
//  * Stage 1-2, key -> dense group id.  The dense kernel is launched first and decides ON THE DEVICE,
//    from a key-range sample taken by k_lowcard_prep, whether it applies:
//      dense : all sampled keys lie in a window of <= GMAX values -> id = key - base, no table at all.
//              Keys not dense (declined before the first row) or a key outside the window met later
//              (ST_DENSE_MISS): the host launches the hash-mode kernel.
//      hash  : CTA-shared open-addressing table in shared memory: 2048 buckets of two PACKED 8-byte entries
//              {53 key bits | 1-bit bucket displacement | 10-bit id}: the key is first mixed by a BIJECTION of 64 bits
//              whose top 11 bits are the home bucket, so home + displacement + the other 53 bits identify the key
//              exactly and ONE LDS.128 per probe answers "which id" for both entries of a bucket (round 1: LDS.128 of
//              a two-key bucket + LDS.U16 of the id; round 2's first table: single entries, LDS.64, three predicated
//              rounds for the 12 % of displaced keys — the kernel is bound by instruction issue, 133 warp instructions
//              per 32 rows, so the rounds were what had to go).  Read-only in steady state; 3 % of a 1000-key set sit in
//              the next bucket (second probe, those lanes only), ~0.2 % in a 16-entry overflow list.  New keys take an
//              out-of-line slow path (ATOMS.CAS.64) that also obtains a GLOBAL id for the key from a small directory
//              in global memory, so that ids mean the same group in every CTA and the merge needs no join.  Null keys
//              and the key equal to the directory sentinel have dedicated ids.
//  * Stage 3: accumulators are WARP-PRIVATE arrays in shared memory (sum 8 B + count word 4 B per
//    id; the wide variant adds last row, {min, max} and a double sum), updated with plain loads and
//    stores — no atomics (sm_100a has no native 64-bit shared-memory atomics: f64/u64 adds compile
//    to ATOMS.CAST.SPIN loops).  Lanes of a warp that hit the same group in a 32-row batch are
//    found with a claim tag: each lane stores its lane number into the top byte of the group's
//    count word and reads the word back; one lane per group reads its own number (the winner).
//    With one losing lane in the batch (the common case) its group has two members and the winner
//    adds the loser's value to its own (commutative, so it does not matter which lane won); with
//    more, an out-of-line routine folds every group in ascending row order.  The order of
//    floating-point additions is therefore a fixed function of (n_rows, grid size): results are
//    run-to-run deterministic.
//  * At the end each CTA folds its warps in warp order into a partial table; k_lowcard_merge folds
//    the partial tables (one warp per group id, fixed order), k_lowcard_rank orders the groups by
//    first row and writes the GroupResult.
// More than GMAX distinct keys -> ST_OVERFLOW, the host reruns on the global-table path.
//
// Replaces Grouper::Consume + MakeGroupings + ApplyGroupings + per-group CallFunction
// (/root/reference/src/dataframe.cpp:1571-1600, pd_core_macros.h:5-147).
#pragma once
#include "group_result.cuh"

namespace pa {

constexpr int LC_HB_LOG2 = 11;
constexpr int LC_HB = 1 << LC_HB_LOG2;       // buckets of the hash-mode table: two packed 8-byte entries each (one LDS.128)
constexpr int LC_HT = 2 * LC_HB;             // packed entries
constexpr int LC_HT_MAXD = 1;                // largest bucket displacement an entry can record (1 bit)
constexpr int LC_OVF = 16;                   // overflow list: keys displaced further than that
constexpr uint64_t LC_HE_EMPTY = ~0ull;      // (id field 0x3FF is never assigned)
constexpr uint32_t LC_HE_PENDING = 0x3FEu;   // id field while the inserting lane fetches the global id
constexpr uint32_t LC_HE_OVF = 0x3FDu;       // id field of a key that did not get an id (more than GMAX keys)
constexpr int LC_NB = 8;                     // 32-row batches per row group
constexpr int LC_GROUP_ROWS = LC_NB * 32;    // rows per warp per row group
constexpr uint32_t LC_ID_UNSET = 0xFFFFu;
constexpr uint32_t LC_ID_OVF = 0xFFFEu;
constexpr uint32_t LC_NOID = 0xFFFFFFFFu;
constexpr uint32_t LC_CNT_MASK = 0x00FFFFFFu;   // count field of a count word; byte 3 = claim tag
                                                // (count == LC_CNT_MASK: the warp has not met this id yet)
constexpr int LC_GMAX_NARROW = 1024;   // sum / mean(float) / count / first : 12 B per id per warp
constexpr int LC_GMAX_WIDE_F = 1024;   // + min / max / last               : + 20 B per id, CTA shared (order independent)
constexpr int LC_GMAX_WIDE_I = 640;    // + double sum (mean of integers)  : 20 B per id per warp
constexpr int LC_GMAX_MAX = LC_GMAX_NARROW;
constexpr int LC_GMAX_HASH = 1000;     // hash-mode kernel: ids fit 10 bits (< LC_HE_OVF) and 16 warps' accumulators fit next to the table

// global key -> id directory (hash mode)
constexpr int LC_GT_LOG2 = 13;
constexpr int LC_GT_CAP = 1 << LC_GT_LOG2;
constexpr uint32_t LC_GID_UNSET = 0xFFFFFFFFu;
constexpr uint32_t LC_GID_OVF = 0xFFFFFFFEu;
constexpr int LC_PREP_GRID = 64;       // x 256 threads = 16384 sampled keys

template <int VC, bool WIDE>
struct LcCfg {
  static constexpr bool DSUM = WIDE && VC != VC_F;
  static constexpr int WARPS = 16;                // dense-mode kernel (no key table)
  static constexpr int GMAX = WIDE ? (DSUM ? LC_GMAX_WIDE_I : LC_GMAX_WIDE_F) : LC_GMAX_NARROW;
  static constexpr int GP = GMAX + 2;             // + null-key group + (key == kEmptyKey) group
  static constexpr int ID_NULL = GMAX;
  static constexpr int ID_EMPTYKEY = GMAX + 1;
  // hash-mode kernel: its own (smaller) id space so that the table and 16 warps' accumulators share 227 KB
  static constexpr int GMAX_H = GMAX < LC_GMAX_HASH ? GMAX : LC_GMAX_HASH;
  static constexpr int GP_H = GMAX_H + 2;
};
inline int lc_gmax(int vc, bool wide) { return wide ? (vc != VC_F ? LC_GMAX_WIDE_I : LC_GMAX_WIDE_F) : LC_GMAX_NARROW; }
inline int lc_gmax_hash(int vc, bool wide) { const int g = lc_gmax(vc, wide); return g < LC_GMAX_HASH ? g : LC_GMAX_HASH; }

// Written by k_lowcard_prep (zero-initialised by the host), read by scan / rank.
struct LcPrep {
  unsigned long long nmin_ord;   // max over sampled valid keys of ~ord(key)   (ord = key ^ 2^63: signed order)
  unsigned long long max_ord;    // max over sampled valid keys of  ord(key)
  unsigned int next_id;          // hash mode: next global id
  unsigned int pad[3];
};

struct LcDir {                   // global-memory directory (hash mode ids) + key sample
  LcPrep* prep;
  unsigned long long* gt_keys;   // [LC_GT_CAP], kEmptyKey = free
  unsigned int* gt_ids;          // [LC_GT_CAP]
  unsigned long long* key_by_id; // [LC_GMAX_MAX]
};

// Dense-mode decision; identical in every kernel that needs it.  `window` = number of key values
// the id space covers, `rlog` = log2 of the accumulator replication: when the sampled key range is
// much smaller than GMAX, every id gets 2^rlog accumulator slots per warp and lane L uses replica
// L mod 2^rlog, so lanes of a batch collide 2^rlog times less often (with 32 replicas every lane
// owns its slots and accesses are bank-conflict free).
__device__ __forceinline__ bool lc_dense_mode(const LcPrep* p, int force_hash, uint32_t gmax, uint64_t* base,
                                              uint32_t* window, uint32_t* rlog) {
  const uint64_t a = p->nmin_ord, b = p->max_ord;
  *base = 0;
  *window = gmax;
  *rlog = 0;
  if (force_hash || (a == 0 && b == 0)) return false;
  const uint64_t smin = (~a) ^ 0x8000000000000000ull, smax = b ^ 0x8000000000000000ull;   // two's complement bit patterns
  const uint64_t span = smax - smin;           // smax >= smin in signed order, so this does not wrap
  if (span >= gmax) return false;
  uint32_t gpow = 1;                           // largest power of two <= gmax
  while (gpow * 2 <= gmax) gpow *= 2;
  uint32_t w = 1;                              // power of two >= 1.25 x the sampled range
  while (w < (span + 1) + (span + 1) / 4) w *= 2;
  if (w * 2 <= gpow) {
    *window = w;
    uint32_t r = gpow / w, l = 0;
    if (r > 32) r = 32;
    while ((1u << (l + 1)) <= r) ++l;
    *rlog = l;
  }
  *base = smin - (*window - (span + 1)) / 2;   // centre the window on the sample (wrapping arithmetic)
  return true;
}

struct LcArgs {
  const void* keys;
  const void* vals;        // may be null (keys-only pass)
  const uint8_t* kvalid;
  const uint8_t* vvalid;
  int64_t koff, voff;
  int64_t n;
  int kw, vw;              // element widths in bytes (generic loader)
  int force_hash;
  int hash_rlog;           // hash-mode kernel: log2 of the accumulator replicas per id (host: from the known group count)
  uint32_t agg_mask;
  LcDir dir;
  // per-CTA partial tables, [grid][GP]
  uint64_t* p_sum;
  double* p_dsum;
  uint32_t* p_count;
  uint32_t* p_first;
  uint32_t* p_last;
  uint64_t* p_min;
  uint64_t* p_max;
  uint32_t* status;
};

// Shared-memory layout.  CTA shared: first rows, (wide) {min, max} 16 B + last row 4 B per id — order
// independent, updated by all warps with shared-memory atomics after a plain pre-check — and, in the hash-mode
// kernel (HASHK), the packed key table + overflow list.  Per warp: [sum 8 B x GP] [dsum 8 B x GP (mean of
// integers)] [count word 4 B x GP].  The hash-mode kernel uses its own, slightly smaller id space (GP_H).
template <int VC, bool WIDE, bool HASHK>
struct LcSmem {
  using Cfg = LcCfg<VC, WIDE>;
  static constexpr size_t GP = HASHK ? Cfg::GP_H : Cfg::GP;
  static constexpr size_t OFF_MISC = 0;                                          // 4 x u32
  static constexpr size_t OFF_MM = 16;                                           // GP x {min, max} (wide)
  static constexpr size_t OFF_FIRST = OFF_MM + (WIDE ? GP * 16 : 0);             // GP u32
  static constexpr size_t OFF_LAST = OFF_FIRST + GP * 4;                         // GP u32 (wide)
  static constexpr bool BOUNDS = WIDE && !HASHK && VC == VC_F;                   // fp32 {min, max} pre-check words (lc_wide_update)
  static constexpr size_t OFF_BND = ((OFF_LAST + (WIDE ? GP * 4 : 0) + 7) / 8) * 8;      // GP x {float, float}
  static constexpr size_t OFF_TAB = ((OFF_BND + (BOUNDS ? GP * 8 : 0) + 15) / 16) * 16;  // LC_HT packed entries (hash)
  static constexpr size_t OFF_OVFK = OFF_TAB + (HASHK ? LC_HT * 8 : 0);          // LC_OVF keys (hash)
  static constexpr size_t OFF_OVFI = OFF_OVFK + (HASHK ? LC_OVF * 8 : 0);        // LC_OVF ids (hash)
  static constexpr size_t OFF_ACC = ((OFF_OVFI + (HASHK ? LC_OVF * 4 : 0) + 15) / 16) * 16;
  static constexpr size_t W_SUM = 0;
  static constexpr size_t W_DSUM = W_SUM + GP * 8;
  static constexpr size_t W_CW = W_DSUM + (Cfg::DSUM ? GP * 8 : 0);
  static constexpr size_t ACC_PER_WARP = ((W_CW + GP * 4 + 15) / 16) * 16;
  static constexpr size_t BUDGET = 227 * 1024;
  static constexpr int WARPS_FIT = static_cast<int>((BUDGET - OFF_ACC) / ACC_PER_WARP);
#ifndef LC_HASH_WARPS_MAX
#define LC_HASH_WARPS_MAX 16
#endif
#ifndef LC_HASH_EARLY_LOADS
#define LC_HASH_EARLY_LOADS 0
#endif
  static constexpr int WARPS = HASHK ? (WARPS_FIT < LC_HASH_WARPS_MAX ? WARPS_FIT : LC_HASH_WARPS_MAX) : Cfg::WARPS;
  static constexpr size_t TOTAL = OFF_ACC + ACC_PER_WARP * WARPS;
  static_assert(TOTAL <= BUDGET, "shared-memory layout exceeds 227 KB");
  static_assert(WARPS >= 8, "too few warps to keep HBM busy");
};

// per-thread view of the CTA's shared state (32-bit shared-memory addresses for the hot arrays)
struct LcCtx {
  uint32_t sum, cw, dsum;             // this warp's accumulator arrays
  uint32_t mm;                        // CTA-shared {min, max} array (wide), shared-memory address
  unsigned long long* mm_p;           // the same, generic pointer (atomics)
  uint32_t* last_p;                   // CTA-shared last-row array (wide)
  uint32_t bnd;                       // CTA-shared fp32 {min rounded up, max rounded down} pre-check words (wide fp64, dense mode)
  uint32_t* bnd_p;                    // the same, generic pointer (atomics)
  uint32_t tab;                       // hash mode: packed table, shared-memory address
  unsigned long long* tab_p;          // the same, generic pointer (atomics)
  unsigned long long* ovf_keys;       // hash mode: overflow list
  uint32_t* ovf_ids;
  uint32_t* cta_first;
  uint32_t* misc;                     // [1] abort seen by this CTA, [2] entries of the overflow list
  int nwarps;                         // warps of the CTA that scan rows
  uint64_t base;                      // dense mode: id = key - base, must be < window
  uint32_t window, rlog, rmask;       // dense mode: accumulator replication (slot = id << rlog | lane & rmask)
};

// Hash mode: bijective mix of the key; the top LC_HB_LOG2 bits are the home bucket, the rest is what an entry stores.
__device__ __forceinline__ uint64_t lc_mix(uint64_t key) { return (key ^ (key >> 32)) * 0x9E3779B97F4A7C15ull; }

// One probe of the packed table at bucket displacement `d`: ONE LDS.128 fetches both entries of the bucket.  Returns the
// id, or a value >= LC_HE_OVF when neither entry holds the key (or its id is not published yet).
__device__ __forceinline__ uint32_t lc_probe(uint64_t m, uint32_t d, uint32_t tab) {
  const uint32_t b = (static_cast<uint32_t>(m >> (64 - LC_HB_LOG2)) + d) & (LC_HB - 1);
  const uint4 e = lds128(tab + b * 16u);
  const uint64_t want = (m << LC_HB_LOG2) | (static_cast<uint64_t>(d) << 10);
  const uint32_t wl = static_cast<uint32_t>(want), wh = static_cast<uint32_t>(want >> 32);
  const bool hit0 = (((e.x ^ wl) & ~0x3FFu) | (e.y ^ wh)) == 0u;
  const bool hit1 = (((e.z ^ wl) & ~0x3FFu) | (e.w ^ wh)) == 0u;
  const uint32_t id = (hit0 ? e.x : e.z) & 0x3FFu;
  return (hit0 || hit1) ? id : static_cast<uint32_t>(LC_ID_UNSET);
}

// Global id of a key that this CTA sees for the first time.  LC_NOID when more than gmax keys exist.
__device__ __noinline__ uint32_t lc_global_id(uint64_t key, LcDir d, uint32_t gmax) {
  uint32_t s = static_cast<uint32_t>(hash_key64(key)) & (LC_GT_CAP - 1);
  for (int probe = 0; probe < LC_GT_CAP; ++probe) {
    if (*reinterpret_cast<volatile unsigned int*>(&d.prep->next_id) > gmax) return LC_NOID;   // already overflowed: do not crawl a full directory
    uint64_t k = *reinterpret_cast<volatile unsigned long long*>(d.gt_keys + s);
    if (k == kEmptyKey) {
      k = atomicCAS(d.gt_keys + s, static_cast<unsigned long long>(kEmptyKey), static_cast<unsigned long long>(key));
      if (k == kEmptyKey) {
        const uint32_t gid = atomicAdd(&d.prep->next_id, 1u);
        if (gid >= gmax) {
          *reinterpret_cast<volatile unsigned int*>(d.gt_ids + s) = LC_GID_OVF;
          return LC_NOID;
        }
        d.key_by_id[gid] = key;
        __threadfence();
        *reinterpret_cast<volatile unsigned int*>(d.gt_ids + s) = gid;
        return gid;
      }
    }
    if (k == key) {
      uint32_t v;
      do { v = *reinterpret_cast<volatile unsigned int*>(d.gt_ids + s); } while (v == LC_GID_UNSET);
      return v == LC_GID_OVF ? LC_NOID : v;
    }
    s = (s + 1) & (LC_GT_CAP - 1);
  }
  return LC_NOID;
}

__device__ __forceinline__ void lc_give_up(uint32_t* misc, uint32_t* status) {
  misc[1] = 1u;
  atomicExch(status + ST_OVERFLOW, 1u);
  atomicExch(status + ST_ABORT, 1u);
}

// Hash-mode miss path (out of line, per-lane divergent code): walks the key's probe sequence, inserts on first
// sight (entry first, then the global id), falls back to the overflow list past LC_HT_MAXD.  LC_NOID on overflow.
__device__ __noinline__ uint32_t lc_miss_resolve(uint64_t key, const LcCtx& c, LcDir dir, uint32_t gmax, uint32_t* status) {
  if (*reinterpret_cast<volatile uint32_t*>(c.misc + 1)) return LC_NOID;   // this CTA already gave up
  const uint64_t m = lc_mix(key);
  const uint32_t home = static_cast<uint32_t>(m >> (64 - LC_HB_LOG2));
  for (uint32_t q = 0; q < 2u * (LC_HT_MAXD + 1); ++q) {      // home bucket entry 0, 1, next bucket entry 0, 1
    const uint32_t d = q >> 1;
    volatile unsigned long long* slot = c.tab_p + (((home + d) & (LC_HB - 1)) * 2u + (q & 1u));
    const uint64_t mine = (m << LC_HB_LOG2) | (static_cast<uint64_t>(d) << 10);
    uint64_t en = *slot;
    if (en == LC_HE_EMPTY) {
      en = atomicCAS(const_cast<unsigned long long*>(slot), static_cast<unsigned long long>(LC_HE_EMPTY),
                     static_cast<unsigned long long>(mine | LC_HE_PENDING));
      if (en == LC_HE_EMPTY) {   // this lane inserted the key into the CTA's table: fetch its global id
        const uint32_t gid = lc_global_id(key, dir, gmax);
        *slot = mine | (gid == LC_NOID ? LC_HE_OVF : gid);
        if (gid == LC_NOID) lc_give_up(c.misc, status);
        return gid;
      }
    }
    if (((en ^ mine) >> 10) == 0) {   // the key's own entry (possibly still pending)
      uint32_t id = static_cast<uint32_t>(en) & 0x3FFu;
      while (id == LC_HE_PENDING) id = static_cast<uint32_t>(*slot) & 0x3FFu;
      return id == LC_HE_OVF ? LC_NOID : id;
    }
  }
  // every slot of the probe window holds another key: overflow list (full keys, linear)
  for (int i = 0; i < LC_OVF; ++i) {
    volatile unsigned long long* ok = c.ovf_keys + i;
    uint64_t k = *ok;
    if (k == kEmptyKey) {
      k = atomicCAS(const_cast<unsigned long long*>(ok), static_cast<unsigned long long>(kEmptyKey), static_cast<unsigned long long>(key));
      if (k == kEmptyKey) {
        const uint32_t gid = lc_global_id(key, dir, gmax);
        *reinterpret_cast<volatile uint32_t*>(c.ovf_ids + i) = gid == LC_NOID ? LC_ID_OVF : gid;
        atomicMax(c.misc + 2, static_cast<uint32_t>(i + 1));
        if (gid == LC_NOID) lc_give_up(c.misc, status);
        return gid;
      }
    }
    if (k == key) {
      uint32_t id;
      do { id = *reinterpret_cast<volatile uint32_t*>(c.ovf_ids + i); } while (id == LC_ID_UNSET);
      return id == LC_ID_OVF ? LC_NOID : id;
    }
  }
  lc_give_up(c.misc, status);   // pathological key set (everything hashes together): the global-table path takes over
  return LC_NOID;
}

// What one lane (later: one group) adds to the warp-private accumulators.
template <bool DSUM>
struct LcContrib {
  uint64_t sum;      // double bits (VC_F) or wrapping integer
  uint32_t cnt;
  uint32_t lo;       // lowest lane of the group
  uint32_t doer;     // this lane performs the read-modify-write
};
template <>
struct LcContrib<true> {
  uint64_t sum;
  uint32_t cnt;
  uint32_t lo;
  uint32_t doer;
  double dsum;       // sum of the values converted to double (mean of integers)
};

template <int VC, bool DSUM>
__device__ __forceinline__ void lc_add(LcContrib<DSUM>& t, const LcContrib<DSUM>& x) {
  if constexpr (VC == VC_F) {
    t.sum = static_cast<uint64_t>(__double_as_longlong(__longlong_as_double(static_cast<long long>(t.sum)) +
                                                       __longlong_as_double(static_cast<long long>(x.sum))));
  } else {
    t.sum += x.sum;
  }
  t.cnt += x.cnt;
  if constexpr (DSUM) t.dsum += x.dsum;
}

template <int VC, bool DSUM>
__device__ __forceinline__ LcContrib<DSUM> lc_shfl(const LcContrib<DSUM>& k, int src) {
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  LcContrib<DSUM> o = k;
  o.sum = __shfl_sync(FULL, k.sum, src);
  o.cnt = __shfl_sync(FULL, k.cnt, src);
  if constexpr (DSUM) o.dsum = __shfl_sync(FULL, k.dsum, src);
  return o;
}

// Two or more losing lanes in the batch: fold every group in ascending lane (= row) order, so that
// the result does not depend on which lanes the hardware let win.  Out of line (rare for ~1000
// groups; for a handful of groups nearly every batch comes here and MATCH.ANY is cheap).
template <int VC, bool DSUM>
__device__ __noinline__ LcContrib<DSUM> lc_fold_general(LcContrib<DSUM> k, uint32_t id, uint32_t tag, uint32_t live,
                                                        uint32_t losers) {
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  const uint32_t lane = lane_id();
  const bool winner = live && tag == lane;
  if (__popc(losers) <= 4) {
    // few duplicates: one warp-uniform iteration per loser lane, ascending; the winner inserts its
    // own value at its own lane position
    LcContrib<DSUM> f = k;
    f.sum = 0; f.cnt = 0;
    if constexpr (DSUM) f.dsum = 0.0;
    bool own_added = false;
    uint32_t rem = losers;
    while (rem) {
      const int L = __ffs(rem) - 1;
      rem &= rem - 1;
      const uint32_t tw = __shfl_sync(FULL, tag, L);
      const LcContrib<DSUM> o = lc_shfl<VC, DSUM>(k, L);
      if (winner && tw == lane) {
        if (!own_added && static_cast<uint32_t>(L) > lane) {
          lc_add<VC, DSUM>(f, k);
          own_added = true;
        }
        lc_add<VC, DSUM>(f, o);
        f.lo = static_cast<uint32_t>(L) < f.lo ? static_cast<uint32_t>(L) : f.lo;
      }
    }
    if (winner) {
      if (!own_added) lc_add<VC, DSUM>(f, k);
      k = f;
    }
    k.doer = winner;
  } else {
    // many duplicates = few distinct ids: the lowest lane of every group pulls its peers in
    // ascending lane order and does the read-modify-write
    const uint32_t peers = __match_any_sync(FULL, live ? id : LC_NOID);
    const uint32_t lanebit = 1u << lane;
    const bool leader = (peers & (lanebit - 1u)) == 0;
    uint32_t rem = leader ? (peers & ~lanebit) : 0u;
    while (__any_sync(FULL, rem != 0)) {
      const int src = rem ? (__ffs(rem) - 1) : static_cast<int>(lane);
      const LcContrib<DSUM> o = lc_shfl<VC, DSUM>(k, src);
      if (rem) {
        lc_add<VC, DSUM>(k, o);
        rem &= rem - 1;
      }
    }
    k.doer = live && leader;
    k.lo = lane;
  }
  return k;
}

// min / max / last row of one row into the CTA-shared arrays (order independent: plain pre-check, rare atomic).  BOUNDS (fp64 values, dense-mode kernel): the pre-check reads
// 8 bytes instead of 16 — {the group's minimum rounded UP to fp32, its maximum rounded DOWN to fp32} (NaN = no number seen
// yet) against the row's value rounded down / up: rd(v) >= ru(min) implies v >= min, ru(v) <= rd(max) implies v <= max, so
// a row that passes both cannot change either and is done after one LDS.64 (~5.8 shared-memory wavefronts per 32-row batch
// on random slots against ~10.5 for the LDS.128 of the exact pair: profiles/r2_lowcard_wide_ncu_full.md).  Everything
// else — a new extreme, a value within fp32 rounding of one, the first number of a group — takes the exact path below and
// then tightens the fp32 words (CAS loops on a rare path).  The words only ever move towards the exact pair and are
// written AFTER it, so a stale word is merely conservative.
template <int VC, bool BOUNDS>
__device__ __forceinline__ void lc_wide_update(uint32_t gid, uint64_t vbits, bool vv, uint32_t row, uint32_t agg_mask, const LcCtx& c) {
  if (agg_mask & AGG_LAST) atomicMax(c.last_p + gid, row);
  if (vv && (agg_mask & (AGG_MIN | AGG_MAX)) && !Wide<VC>::is_nan(vbits)) {
    float lo = 0.f, hi = 0.f;
    if constexpr (BOUNDS) {
      const double v = __longlong_as_double(static_cast<long long>(vbits));
      lo = __double2float_rd(v);
      hi = __double2float_ru(v);
      const uint64_t b = lds64(c.bnd + gid * 8u);
      const uint32_t bmn = static_cast<uint32_t>(b), bmx = static_cast<uint32_t>(b >> 32);
      // equal counts only bit for bit: -0.0 against a +0.0 word goes on to the exact pair, whose order map tells the two
      // zeros apart — the result must not depend on which zero a CTA met first.  All four tests fail on a NaN word.
      const bool in_lo = lo > __uint_as_float(bmn) || __float_as_uint(lo) == bmn;
      const bool in_hi = hi < __uint_as_float(bmx) || __float_as_uint(hi) == bmx;
      if (in_lo && in_hi) return;
    }
    const uint64_t o = Wide<VC>::ord(vbits);
    const uint4 m4 = lds128(c.mm + gid * 16u);
    const uint64_t mn = static_cast<uint64_t>(m4.x) | (static_cast<uint64_t>(m4.y) << 32);
    const uint64_t mx = static_cast<uint64_t>(m4.z) | (static_cast<uint64_t>(m4.w) << 32);
    if (o < mn) atomicMin(c.mm_p + 2 * gid, static_cast<unsigned long long>(o));
    if (o > mx) atomicMax(c.mm_p + 2 * gid + 1, static_cast<unsigned long long>(o));
    if constexpr (BOUNDS) {
      // fp32 words: min word = the smallest ru(v) seen, max word = the largest rd(v) seen (>= / <= the exact pair)
      __threadfence_block();                            // the exact pair first, then the words that vouch for it
      volatile uint32_t* bp = c.bnd_p + 2 * gid;
      uint32_t cur = bp[0];
      // smaller wins; of two equal zeros the negative one (the smaller in the exact pair's order)
      while (!(__uint_as_float(cur) < hi || cur == __float_as_uint(hi) || (__uint_as_float(cur) == hi && (cur >> 31)))) {
        const uint32_t seen = atomicCAS(c.bnd_p + 2 * gid, cur, __float_as_uint(hi));   // word is NaN (unset) or above ru(v)
        if (seen == cur) break;
        cur = seen;
      }
      cur = bp[1];
      while (!(__uint_as_float(cur) > lo || cur == __float_as_uint(lo) || (__uint_as_float(cur) == lo && !(cur >> 31)))) {
        const uint32_t seen = atomicCAS(c.bnd_p + 2 * gid + 1, cur, __float_as_uint(lo));
        if (seen == cur) break;
        cur = seen;
      }
    }
  }
}

// Accumulate one 32-row batch whose ids are known.  Warp-synchronous; lane L holds row `row`
// (= batch row base + L).  CLEAN: every lane holds a row with a resolved id and a valid value.
// `id` is the lane's accumulator SLOT (dense mode: group id << rlog | replica), `gid` the group id.
// Lanes that hit the same slot in one batch find each other through a claim tag: every lane stores its lane number
// into byte 3 of the slot's count word (STS.8) and reads the word back; one lane per slot reads its own number.
// (Measured alternative, scripts/ubench/lc_variants.cu: MATCH.ANY on the slot number needs no shared-memory traffic
// but issues at ~1 per 40 cycles per SM on sm_100a — 3.64 ms against 1.70 ms per 512 M rows — so the tags stay.)
template <int VC, bool WIDE, bool CLEAN, bool DENSE>
__device__ __forceinline__ void lc_accumulate(uint32_t id, uint64_t vbits, bool vvalid, uint32_t row, uint32_t agg_mask,
                                              const LcCtx& c) {
  using Cfg = LcCfg<VC, WIDE>;
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  const uint32_t lane = lane_id();
  const bool live = CLEAN ? true : id != LC_NOID;
  const uint32_t ida = live ? id : 0u;
  const uint32_t cw_addr = c.cw + ida * 4u, sum_addr = c.sum + ida * 8u;
  const bool vv = CLEAN ? true : vvalid;
  const uint32_t gid = ida < static_cast<uint32_t>(DENSE ? Cfg::GMAX : Cfg::GMAX_H) ? (ida >> c.rlog) : ida;
  LcContrib<Cfg::DSUM> k;
  k.sum = vv ? vbits : 0ull;
  k.cnt = vv ? 1u : 0u;
  k.lo = lane;
  if constexpr (Cfg::DSUM) k.dsum = vv ? Wide<VC>::as_double(vbits) : 0.0;
  if (live) sts8(cw_addr + 3u, lane);
  __syncwarp();
  const uint32_t cw = lds32(cw_addr);
  uint64_t s = lds64(sum_addr);
  if constexpr (WIDE) {
    if (live) lc_wide_update<VC, DENSE && VC == VC_F>(gid, vbits, vv, row, agg_mask, c);
  }
  const uint32_t tag = cw >> 24;
  const bool winner = live && (tag == lane);
  const uint32_t losers = __ballot_sync(FULL, live && !winner);
  k.doer = winner;
  if (losers) {
    if ((losers & (losers - 1u)) == 0u) {
      // exactly one losing lane: its group has two members, a + b is commutative
      const int L = __ffs(losers) - 1;
      const uint32_t tw = __shfl_sync(FULL, tag, L);
      const LcContrib<Cfg::DSUM> o = lc_shfl<VC, Cfg::DSUM>(k, L);
      if (lane == tw) {
        lc_add<VC, Cfg::DSUM>(k, o);
        k.lo = static_cast<uint32_t>(L) < lane ? static_cast<uint32_t>(L) : lane;
      }
    } else {
      // two or three losing lanes whose slots differ (their winners' tags are pairwise distinct): as many
      // independent two-member groups, each folded commutatively like the single case.  With ~1000 slots
      // 8 % of the batches have more than one loser, with 512 slots 25 %; nearly all of them are of this kind.
      bool independent = false;
      int l0 = 0, l1 = 0, l2 = -1;
      uint32_t t0 = 0, t1 = 0, t2 = 0xFFFFFFFFu;
      if (__popc(losers) <= 3) {
        l0 = __ffs(losers) - 1;
        const uint32_t rest = losers & (losers - 1u);
        l1 = __ffs(rest) - 1;
        const uint32_t rest2 = rest & (rest - 1u);
        t0 = __shfl_sync(FULL, tag, l0);
        t1 = __shfl_sync(FULL, tag, l1);
        independent = t0 != t1;
        if (rest2) {
          l2 = __ffs(rest2) - 1;
          t2 = __shfl_sync(FULL, tag, l2);
          independent = independent && t2 != t0 && t2 != t1;
        }
      }
      if (independent) {
        const LcContrib<Cfg::DSUM> o0 = lc_shfl<VC, Cfg::DSUM>(k, l0);
        const LcContrib<Cfg::DSUM> o1 = lc_shfl<VC, Cfg::DSUM>(k, l1);
        if (lane == t0) { lc_add<VC, Cfg::DSUM>(k, o0); k.lo = static_cast<uint32_t>(l0) < lane ? static_cast<uint32_t>(l0) : lane; }
        if (lane == t1) { lc_add<VC, Cfg::DSUM>(k, o1); k.lo = static_cast<uint32_t>(l1) < lane ? static_cast<uint32_t>(l1) : lane; }
        if (l2 >= 0) {
          const LcContrib<Cfg::DSUM> o2 = lc_shfl<VC, Cfg::DSUM>(k, l2);
          if (lane == t2) { lc_add<VC, Cfg::DSUM>(k, o2); k.lo = static_cast<uint32_t>(l2) < lane ? static_cast<uint32_t>(l2) : lane; }
        }
      } else {
        k = lc_fold_general<VC, Cfg::DSUM>(k, id, tag, live, losers);
      }
    }
  }
  // one non-atomic read-modify-write per distinct slot
  if (k.doer) {
    const uint32_t old = cw & LC_CNT_MASK;
    const bool first_seen = old == LC_CNT_MASK;   // first time this warp meets the slot
    if constexpr (VC == VC_F) {
      s = static_cast<uint64_t>(__double_as_longlong(__longlong_as_double(static_cast<long long>(s)) +
                                                     __longlong_as_double(static_cast<long long>(k.sum))));
    } else {
      s += k.sum;
    }
    sts64(sum_addr, s);
    sts32(cw_addr, (first_seen ? 0u : old) + k.cnt);   // also clears the claim tag
    if constexpr (Cfg::DSUM) {
      const uint32_t da = c.dsum + id * 8u;
      sts64(da, static_cast<uint64_t>(__double_as_longlong(__longlong_as_double(static_cast<long long>(lds64(da))) + k.dsum)));
    }
    if (first_seen) atomicMin(c.cta_first + gid, row + k.lo - lane);   // candidate for the CTA's first row
  }
  __syncwarp();
}

// One row group in registers: LC_NB batches of 32 rows, entry e = row g0 + 32 e + lane.
struct LcBuf {
  uint64_t key[LC_NB];
  uint64_t val[LC_NB];
  uint32_t kv, vv, act;    // bit e: key valid / value valid / row exists
};

__device__ __forceinline__ uint64_t ldg_stream_u64(const void* p) {
  uint64_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(r) : "l"(p));
  return r;
}

// FAST loader: 8-byte keys and values, every row exists, no bitmaps.  Each load instruction reads
// 256 contiguous bytes per warp.
__device__ __forceinline__ void lc_load_fast(LcBuf& b, const LcArgs& a, int64_t g0, uint32_t lane) {
  const char* kp = static_cast<const char*>(a.keys) + (g0 + lane) * 8;
  const char* vp = static_cast<const char*>(a.vals) + (g0 + lane) * 8;
#pragma unroll
  for (int e = 0; e < LC_NB; ++e) b.key[e] = ldg_stream_u64(kp + e * 256);
#pragma unroll
  for (int e = 0; e < LC_NB; ++e) b.val[e] = ldg_stream_u64(vp + e * 256);
  b.kv = 0xFFFFFFFFu;
  b.vv = 0xFFFFFFFFu;
  b.act = 0xFFFFFFFFu;
}

template <int VC>
__device__ __forceinline__ void lc_load_generic(LcBuf& b, const LcArgs& a, int64_t g0, uint32_t lane) {
  b.kv = 0; b.vv = 0; b.act = 0;
#pragma unroll
  for (int e = 0; e < LC_NB; ++e) {
    const int64_t r = g0 + e * 32 + lane;
    b.key[e] = 0;
    b.val[e] = 0;
    if (r < a.n) {
      b.act |= 1u << e;
      b.key[e] = a.kw == 8 ? static_cast<const uint64_t*>(a.keys)[r]
                           : static_cast<uint64_t>(static_cast<const uint32_t*>(a.keys)[r]);
      if (!a.kvalid || bit_at(a.kvalid, a.koff + r)) b.kv |= 1u << e;
      if (a.vals) {
        switch (a.vw) {
          case 8: b.val[e] = load_wide<VC, 8>(a.vals, r); break;
          case 4: b.val[e] = load_wide<VC, 4>(a.vals, r); break;
          case 2: if constexpr (VC != VC_F) b.val[e] = load_wide<VC, 2>(a.vals, r); break;
          default: if constexpr (VC != VC_F) b.val[e] = load_wide<VC, 1>(a.vals, r); break;
        }
        if (!a.vvalid || bit_at(a.vvalid, a.voff + r)) b.vv |= 1u << e;
      }
    }
  }
}

// One row group: resolve the ids of all LC_NB batches first (independent lookups in flight, one vote for the rare
// paths; returns false on abort), then accumulate batch by batch.  CLEAN: every row exists, keys and values are valid.
template <int VC, bool WIDE, bool DENSE, bool CLEAN>
__device__ __forceinline__ bool lc_resolve_group(const LcBuf& b, uint32_t (&id)[LC_NB], uint32_t lane, const LcCtx& c, const LcArgs& a) {
  using Cfg = LcCfg<VC, WIDE>;
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  if constexpr (DENSE) {
    bool bad = false;
#pragma unroll
    for (int e = 0; e < LC_NB; ++e) {
      const bool act = CLEAN ? true : ((b.act >> e) & 1u) != 0;
      const bool kvalid = CLEAN ? true : ((b.kv >> e) & 1u) != 0;
      const uint64_t d = b.key[e] - c.base;
      id[e] = (static_cast<uint32_t>(d) << c.rlog) | (lane & c.rmask);   // this lane's replica of the group's slot
      if (!kvalid) id[e] = Cfg::ID_NULL;
      if (!act) id[e] = LC_NOID;
      bad |= act && kvalid && d >= static_cast<uint64_t>(c.window);
    }
    if (__any_sync(FULL, bad)) {
      if (lane == 0) {
        c.misc[1] = 1u;
        atomicExch(a.status + ST_DENSE_MISS, 1u);
        atomicExch(a.status + ST_ABORT, 1u);
      }
      return false;
    }
  } else {
    // probe 0 of every batch (8 independent LDS.128 in flight), then — lanes that missed only — the displaced keys:
    // of a 1000-key set 3 % sit in the next bucket, 0.2 % in the overflow list.  Both
    // rounds are straight-line predicated code (nearly every 256-row group has a lane in every round; a loop with
    // votes per batch measured 8.8 ms per 1 B rows against 6.4 ms for round 1's bucket table); only keys that
    // are not in the table yet leave it for the out-of-line insert.
    uint32_t missmask = 0;
#pragma unroll
    for (int e = 0; e < LC_NB; ++e) {
      const bool act = CLEAN ? true : ((b.act >> e) & 1u) != 0;
      const bool kvalid = CLEAN ? true : ((b.kv >> e) & 1u) != 0;
      // (the sentinel key value is never inserted, so it misses every probe and is given its id with the rare cases below)
      id[e] = lc_probe(lc_mix(b.key[e]), 0u, c.tab);
      if (!kvalid) id[e] = Cfg::GMAX_H;
      if (!act) id[e] = LC_NOID;
      if (act && kvalid && id[e] >= LC_HE_OVF) missmask |= 1u << e;
    }
#pragma unroll
    for (uint32_t d = 1; d <= LC_HT_MAXD; ++d) {
#pragma unroll
      for (int e = 0; e < LC_NB; ++e) {
        if ((missmask >> e) & 1u) {
          const uint32_t id2 = lc_probe(lc_mix(b.key[e]), d, c.tab);   // (recomputed: eight live 64-bit mixes cost 16 registers)
          if (id2 < LC_HE_OVF) { id[e] = id2; missmask &= ~(1u << e); }
        }
      }
    }
    if (__any_sync(FULL, missmask != 0)) {
      // the sentinel key value, the overflow list (published entries), then the insert path
      const uint32_t novf = *reinterpret_cast<volatile uint32_t*>(c.misc + 2);
#pragma unroll
      for (int e = 0; e < LC_NB; ++e) {
        if ((missmask >> e) & 1u) {
          if (b.key[e] == kEmptyKey) { id[e] = Cfg::GMAX_H + 1; missmask &= ~(1u << e); continue; }
          for (uint32_t i = 0; i < novf; ++i) {
            if (lds64(smem_u32(c.ovf_keys + i)) == b.key[e]) {
              const uint32_t oid = *reinterpret_cast<volatile uint32_t*>(c.ovf_ids + i);
              if (oid < LC_ID_OVF) { id[e] = oid; missmask &= ~(1u << e); }
              break;
            }
          }
        }
      }
      if (__any_sync(FULL, missmask != 0)) {
#pragma unroll
        for (int e = 0; e < LC_NB; ++e) {
          if ((missmask >> e) & 1u) id[e] = lc_miss_resolve(b.key[e], c, a.dir, c.window, a.status);
        }
        __syncwarp();
        // a key that got no id (more than GMAX keys, overflow list full): the pass is abandoned before anything of this
        // row group is accumulated — lc_miss_resolve has raised the abort flags, the host reruns on the global path
        bool lost = false;
#pragma unroll
        for (int e = 0; e < LC_NB; ++e) lost |= id[e] == LC_NOID && (CLEAN || ((b.act >> e) & 1u) != 0);
        if (__any_sync(FULL, lost)) return false;
      }
    }
    if (c.rlog) {   // group id -> this lane's replica of the group's slot (the two special ids have no replicas)
#pragma unroll
      for (int e = 0; e < LC_NB; ++e) {
        if (id[e] < static_cast<uint32_t>(Cfg::GMAX_H)) id[e] = (id[e] << c.rlog) | (lane & c.rmask);
      }
    }
  }
  return true;
}

template <int VC, bool WIDE, bool DENSE, bool CLEAN>
__device__ __forceinline__ void lc_accumulate_group(const LcBuf& b, const uint32_t (&id)[LC_NB], int64_t g0, uint32_t lane,
                                                    const LcCtx& c, const LcArgs& a) {
#pragma unroll
  for (int e = 0; e < LC_NB; ++e) {
    const uint32_t row = static_cast<uint32_t>(g0) + e * 32 + lane;
    constexpr bool ACLEAN = CLEAN;            // (hash mode: a row group with an unresolved key never gets here)
    const bool vvalid = CLEAN ? true : ((b.vv >> e) & 1u) != 0;
    lc_accumulate<VC, WIDE, ACLEAN, DENSE>(id[e], b.val[e], vvalid && (ACLEAN || id[e] != LC_NOID), row, a.agg_mask, c);
  }
}

template <int VC, bool WIDE, bool DENSE, bool CLEAN>
__device__ __forceinline__ bool lc_process_group(const LcBuf& b, int64_t g0, uint32_t lane, const LcCtx& c, const LcArgs& a) {
  uint32_t id[LC_NB];
  if (!lc_resolve_group<VC, WIDE, DENSE, CLEAN>(b, id, lane, c, a)) return false;
  lc_accumulate_group<VC, WIDE, DENSE, CLEAN>(b, id, g0, lane, c, a);
  return true;
}

template <int VC, bool WIDE, bool FAST, bool DENSE>
__device__ __forceinline__ void lc_scan_rows(const LcArgs& a, const LcCtx& c, int warp, uint32_t lane) {
  if (warp >= c.nwarps) return;
  const int64_t gw = static_cast<int64_t>(blockIdx.x) * c.nwarps + warp;
  const int64_t nw = static_cast<int64_t>(gridDim.x) * c.nwarps;
  const int64_t n_full = a.n / LC_GROUP_ROWS;                        // full row groups
  const int64_t n_groups = (a.n + LC_GROUP_ROWS - 1) / LC_GROUP_ROWS;
  volatile uint32_t* abort_local = c.misc + 1;
  volatile uint32_t* abort_global = a.status + ST_ABORT;
  if constexpr (FAST) {
    LcBuf cur, nxt;
    int64_t g = gw;
    if (g < n_full) lc_load_fast(cur, a, g * LC_GROUP_ROWS, lane);
    while (g < n_full) {
      const int64_t gn = g + nw;
      const uint32_t gabort = *abort_global;
      if constexpr (DENSE || LC_HASH_EARLY_LOADS) {
        if (gn < n_full) lc_load_fast(nxt, a, gn * LC_GROUP_ROWS, lane);
        if (!lc_process_group<VC, WIDE, DENSE, true>(cur, g * LC_GROUP_ROWS, lane, c, a)) return;
      } else {
        // hash mode: the keys stay live through the probe rounds, so the next row group is requested after them
        // (its 4 KB per warp then travel during the 8 accumulate batches) — at 128 registers per thread the
        // prefetch buffer would otherwise spill to local memory, which shares the LSU with shared memory
        uint32_t id[LC_NB];
        if (!lc_resolve_group<VC, WIDE, DENSE, true>(cur, id, lane, c, a)) return;
        if (gn < n_full) lc_load_fast(nxt, a, gn * LC_GROUP_ROWS, lane);
        lc_accumulate_group<VC, WIDE, DENSE, true>(cur, id, g * LC_GROUP_ROWS, lane, c, a);
      }
      if (__any_sync(0xFFFFFFFFu, (*abort_local | gabort) != 0)) return;
      cur = nxt;
      g = gn;
    }
    // the last, partial row group (if any) goes through the generic loader
    if (n_groups > n_full && gw == (n_full % nw)) {
      LcBuf t;
      lc_load_generic<VC>(t, a, n_full * LC_GROUP_ROWS, lane);
      lc_process_group<VC, WIDE, DENSE, false>(t, n_full * LC_GROUP_ROWS, lane, c, a);
    }
  } else {
    LcBuf cur, nxt;
    int64_t g = gw;
    if (g < n_groups) lc_load_generic<VC>(cur, a, g * LC_GROUP_ROWS, lane);
    while (g < n_groups) {
      const int64_t gn = g + nw;
      if (gn < n_groups) lc_load_generic<VC>(nxt, a, gn * LC_GROUP_ROWS, lane);
      const uint32_t gabort = *abort_global;
      if (!lc_process_group<VC, WIDE, DENSE, false>(cur, g * LC_GROUP_ROWS, lane, c, a)) return;
      if (__any_sync(0xFFFFFFFFu, (*abort_local | gabort) != 0)) return;
      cur = nxt;
      g = gn;
    }
  }
}

// FAST: int64/uint64 keys and 8-byte values, no validity bitmaps, 8-byte aligned columns.
// DENSEK: the dense-mode kernel (16 warps, no key table; when the key sample says the keys are not dense it
// returns at once with ST_DENSE_MISS = 2 and the host launches the hash-mode kernel: packed key table + as many
// warps as fit beside it, 16 for the narrow aggregate set).
template <int VC, bool WIDE, bool FAST, bool DENSEK>
__global__ void __launch_bounds__(LcSmem<VC, WIDE, !DENSEK>::WARPS * 32, 1) k_lowcard_scan(LcArgs a) {
  using Cfg = LcCfg<VC, WIDE>;
  using L = LcSmem<VC, WIDE, !DENSEK>;
  constexpr int THREADS = L::WARPS * 32;
  constexpr int GPK = static_cast<int>(L::GP);          // ids of this kernel's accumulator arrays
  constexpr int GMAXK = GPK - 2;
  constexpr uint32_t CNT_UNSEEN = LC_CNT_MASK;
  extern __shared__ __align__(16) unsigned char smem[];
  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  LcCtx c;
  c.misc = reinterpret_cast<uint32_t*>(smem + L::OFF_MISC);
  c.tab_p = reinterpret_cast<unsigned long long*>(smem + L::OFF_TAB);
  c.tab = smem_u32(c.tab_p);
  c.ovf_keys = reinterpret_cast<unsigned long long*>(smem + L::OFF_OVFK);
  c.ovf_ids = reinterpret_cast<uint32_t*>(smem + L::OFF_OVFI);
  c.cta_first = reinterpret_cast<uint32_t*>(smem + L::OFF_FIRST);
  const bool dense = lc_dense_mode(a.dir.prep, DENSEK ? 0 : 1, Cfg::GMAX, &c.base, &c.window, &c.rlog);
  if (DENSEK && !dense) {   // not a dense key set: hand over to the hash-mode kernel
    if (threadIdx.x == 0) { atomicExch(a.status + ST_DENSE_MISS, 2u); atomicExch(a.status + ST_ABORT, 1u); }
    return;
  }
  if constexpr (!DENSEK) {
    // Few SCATTERED keys (hashed utf8 keys with a handful of values): without replicas most lanes of a batch hit the
    // same slot and every batch takes the out-of-line ordered fold (2 keys: 21 ms per 1 B rows).  When the host knows
    // the group count it asks for 2^rlog replicas per id, exactly like the dense kernel's small windows; ids beyond
    // GMAX >> rlog overflow to the global path.
    c.rlog = static_cast<uint32_t>(a.hash_rlog);
    c.window = static_cast<uint32_t>(GMAXK) >> c.rlog;
  }
  // lane-private replicas (32 per id) of the narrow layout are conflict free and 12 warps already saturate HBM
  // (2.99 against 3.05 ms at 16 groups); everything else wants all 16 warps (wide set: 3.6 against 4.2 ms at 16
  // groups, 4.4 against 5.3 ms at 64; narrow 64 groups: 3.3 against 3.7 ms)
  const int nwarps = DENSEK ? ((!WIDE && c.rlog >= 5) ? 12 : L::WARPS) : L::WARPS;
  c.nwarps = nwarps;
  unsigned char* my_acc = smem + L::OFF_ACC + L::ACC_PER_WARP * (warp < nwarps ? warp : 0);
  const uint32_t acc_s = smem_u32(my_acc);
  c.sum = acc_s + static_cast<uint32_t>(L::W_SUM);
  c.dsum = acc_s + static_cast<uint32_t>(L::W_DSUM);
  c.cw = acc_s + static_cast<uint32_t>(L::W_CW);
  c.mm_p = reinterpret_cast<unsigned long long*>(smem + L::OFF_MM);
  c.mm = smem_u32(c.mm_p);
  c.last_p = reinterpret_cast<uint32_t*>(smem + L::OFF_LAST);
  c.bnd_p = reinterpret_cast<uint32_t*>(smem + L::OFF_BND);
  c.bnd = smem_u32(c.bnd_p);
  c.rmask = (1u << c.rlog) - 1u;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.status[ST_MODE] = dense ? 1u : 2u;
    a.status[ST_RLOG] = c.rlog;
  }

  // ---- init shared state ----
  if constexpr (!DENSEK) {
    for (int i = threadIdx.x; i < LC_HT; i += THREADS) c.tab_p[i] = LC_HE_EMPTY;
    if (threadIdx.x < LC_OVF) { c.ovf_keys[threadIdx.x] = kEmptyKey; c.ovf_ids[threadIdx.x] = LC_ID_UNSET; }
  }
  for (int i = threadIdx.x; i < GPK; i += THREADS) {
    c.cta_first[i] = kNoRow;
    if constexpr (WIDE) {
      c.mm_p[2 * i] = kMinInit;
      c.mm_p[2 * i + 1] = kMaxInit;
      c.last_p[i] = 0u;
      if constexpr (L::BOUNDS) { c.bnd_p[2 * i] = 0x7FC00000u; c.bnd_p[2 * i + 1] = 0x7FC00000u; }   // NaN: no number yet
    }
  }
  if (threadIdx.x < 4) c.misc[threadIdx.x] = 0;
  if (warp < nwarps) {
    for (int i = lane; i < GPK; i += 32) {
      reinterpret_cast<uint64_t*>(my_acc + L::W_SUM)[i] = 0ull;
      reinterpret_cast<uint32_t*>(my_acc + L::W_CW)[i] = CNT_UNSEEN;
      if constexpr (Cfg::DSUM) reinterpret_cast<double*>(my_acc + L::W_DSUM)[i] = 0.0;
    }
  }
  __syncthreads();

  lc_scan_rows<VC, WIDE, FAST, DENSEK>(a, c, warp, lane);
  __syncthreads();
  if (c.misc[1]) {
    if (threadIdx.x == 0) atomicExch(a.status + ST_ABORT, 1u);
    return;
  }

  // ---- fold the warps (and, in dense mode, the replicas) in a fixed order and write this CTA's partial table ----
  // (partial tables are always Cfg::GP ids wide; the hash-mode kernel's two special ids move to the end)
  const size_t pbase = static_cast<size_t>(blockIdx.x) * Cfg::GP;
  const uint32_t nrep = 1u << c.rlog;
  for (int id = threadIdx.x; id < Cfg::GP; id += THREADS) {
    const bool regular = id < Cfg::GMAX;
    const int kid = regular ? id : id - Cfg::GMAX + GMAXK;      // this kernel's slot number of partial-table id `id`
    const bool present = regular ? id < GMAXK : true;
    uint64_t sum = 0;
    double fsum = 0.0, dsum = 0.0;
    uint32_t cnt = 0;
    const uint32_t reps = regular ? nrep : 1u;
    if (present && (!regular || static_cast<uint32_t>(id) < c.window)) {
      for (int w = 0; w < nwarps; ++w) {
        const unsigned char* wa = smem + L::OFF_ACC + L::ACC_PER_WARP * w;
        for (uint32_t r = 0; r < reps; ++r) {
          const uint32_t slot = regular ? ((static_cast<uint32_t>(kid) << c.rlog) | r) : static_cast<uint32_t>(kid);
          uint32_t cw = reinterpret_cast<const uint32_t*>(wa + L::W_CW)[slot];
          cw &= LC_CNT_MASK;
          if (cw == CNT_UNSEEN) continue;
          cnt += cw;
          const uint64_t s = reinterpret_cast<const uint64_t*>(wa + L::W_SUM)[slot];
          if constexpr (VC == VC_F) fsum += __longlong_as_double(static_cast<long long>(s));
          else sum += s;
          if constexpr (Cfg::DSUM) dsum += reinterpret_cast<const double*>(wa + L::W_DSUM)[slot];
        }
      }
    }
    if constexpr (VC == VC_F) sum = static_cast<uint64_t>(__double_as_longlong(fsum));
    a.p_sum[pbase + id] = sum;
    a.p_count[pbase + id] = cnt;
    a.p_first[pbase + id] = present ? c.cta_first[kid] : kNoRow;
    if constexpr (WIDE) {
      a.p_last[pbase + id] = present ? c.last_p[kid] : 0u;
      a.p_min[pbase + id] = present ? c.mm_p[2 * kid] : kMinInit;
      a.p_max[pbase + id] = present ? c.mm_p[2 * kid + 1] : kMaxInit;
      if constexpr (Cfg::DSUM) a.p_dsum[pbase + id] = dsum;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Prep: sample the key range (dense-mode decision) and reset the global id directory.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_lowcard_prep(LcArgs a) {
  const uint32_t t = blockIdx.x * 256u + threadIdx.x;
  const uint32_t nt = gridDim.x * 256u;
  for (uint32_t i = t; i < LC_GT_CAP; i += nt) {
    a.dir.gt_keys[i] = kEmptyKey;
    a.dir.gt_ids[i] = LC_GID_UNSET;
  }
  key_range_sample(a.keys, a.kvalid, a.koff, a.kw, a.n, t, nt, reinterpret_cast<KeyRange*>(a.dir.prep));
}

// ---------------------------------------------------------------------------------------------
// Merge: one warp per group id folds the per-CTA partials (lane-strided, then a fixed shuffle tree).
// ---------------------------------------------------------------------------------------------
struct LmArgs {
  LcArgs part;          // partial tables written by the scan
  int grid;             // number of scan CTAs
  int gp;               // ids per partial table
  // merged, indexed by id
  uint64_t* m_sum;
  double* m_dsum;
  uint32_t* m_count;
  uint32_t* m_first;
  uint32_t* m_last;
  uint64_t* m_min;
  uint64_t* m_max;
  GroupResult out;
  uint32_t* status;
};

template <int VC, bool WIDE>
__global__ void __launch_bounds__(256) k_lowcard_merge(LmArgs a) {
  using Cfg = LcCfg<VC, WIDE>;
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  const int id = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (id >= Cfg::GP) return;
  if (*reinterpret_cast<volatile uint32_t*>(a.status + ST_ABORT)) return;
  const uint32_t lane = lane_id();
  uint64_t sum = 0, mn = kMinInit, mx = kMaxInit;
  double fsum = 0.0, dsum = 0.0;
  uint32_t cnt = 0, first = kNoRow, last = 0;
  for (int b = lane; b < a.grid; b += 32) {
    const size_t p = static_cast<size_t>(b) * Cfg::GP + id;
    const uint32_t f = a.part.p_first[p];
    if (f == kNoRow) continue;
    first = f < first ? f : first;
    cnt += a.part.p_count[p];
    if constexpr (VC == VC_F) fsum += __longlong_as_double(static_cast<long long>(a.part.p_sum[p]));
    else sum += a.part.p_sum[p];
    if constexpr (WIDE) {
      const uint32_t l = a.part.p_last[p];
      last = l > last ? l : last;
      const uint64_t pmn = a.part.p_min[p], pmx = a.part.p_max[p];
      mn = pmn < mn ? pmn : mn;
      mx = pmx > mx ? pmx : mx;
      if constexpr (Cfg::DSUM) dsum += a.part.p_dsum[p];
    }
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    const uint32_t of = __shfl_xor_sync(FULL, first, d);
    first = of < first ? of : first;
    cnt += __shfl_xor_sync(FULL, cnt, d);
    if constexpr (VC == VC_F) fsum += __shfl_xor_sync(FULL, fsum, d);
    else sum += __shfl_xor_sync(FULL, sum, d);
    if constexpr (WIDE) {
      const uint32_t ol = __shfl_xor_sync(FULL, last, d);
      last = ol > last ? ol : last;
      const uint64_t on = __shfl_xor_sync(FULL, mn, d), ox = __shfl_xor_sync(FULL, mx, d);
      mn = on < mn ? on : mn;
      mx = ox > mx ? ox : mx;
      if constexpr (Cfg::DSUM) dsum += __shfl_xor_sync(FULL, dsum, d);
    }
  }
  if (lane == 0) {
    if constexpr (VC == VC_F) sum = static_cast<uint64_t>(__double_as_longlong(fsum));
    a.m_sum[id] = sum;
    a.m_count[id] = cnt;
    a.m_first[id] = first;
    if constexpr (WIDE) {
      a.m_last[id] = last;
      a.m_min[id] = mn;
      a.m_max[id] = mx;
      if constexpr (Cfg::DSUM) a.m_dsum[id] = dsum;
    }
  }
}

// Rank: orders the merged groups by first row (first rows are distinct: a row belongs to exactly one
// group), writes the GroupResult and status[ST_NGROUPS] (zeroed by the host before the pass).  One warp
// per group id: the lanes split the comparisons against all first rows (staged in shared memory).
constexpr int LR_THREADS = 256;
template <int VC, bool WIDE>
__global__ void __launch_bounds__(LR_THREADS) k_lowcard_rank(LmArgs a) {
  using Cfg = LcCfg<VC, WIDE>;
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  __shared__ uint32_t sfirst[Cfg::GP];
  if (*reinterpret_cast<volatile uint32_t*>(a.status + ST_ABORT)) return;
  for (int i = threadIdx.x; i < Cfg::GP; i += LR_THREADS) sfirst[i] = a.m_first[i];
  __syncthreads();
  const int id = blockIdx.x * (LR_THREADS / 32) + (threadIdx.x >> 5);
  if (id >= Cfg::GP) return;
  const uint32_t f = sfirst[id];
  if (f == kNoRow) return;
  const uint32_t lane = lane_id();
  uint32_t rank = 0;
  for (int j = lane; j < Cfg::GP; j += 32) rank += sfirst[j] < f;
#pragma unroll
  for (int d = 16; d; d >>= 1) rank += __shfl_xor_sync(FULL, rank, d);
  if (lane != 0) return;
  uint64_t base;
  uint32_t window, rlog;
  const bool dense = lc_dense_mode(a.part.dir.prep, a.part.force_hash, Cfg::GMAX, &base, &window, &rlog);
  uint64_t key;
  if (id == Cfg::ID_NULL) key = 0;
  else if (id == Cfg::ID_EMPTYKEY) key = kEmptyKey;
  else key = dense ? base + static_cast<uint64_t>(id) : a.part.dir.key_by_id[id];
  a.out.key[rank] = key;
  a.out.key_kind[rank] = (id == Cfg::ID_NULL) ? KK_NULL : KK_REGULAR;
  a.out.sum[rank] = a.m_sum[id];
  a.out.count[rank] = a.m_count[id];
  a.out.first_row[rank] = f;
  if constexpr (WIDE) {
    a.out.last_row[rank] = a.m_last[id];
    a.out.min_ord[rank] = a.m_min[id];
    a.out.max_ord[rank] = a.m_max[id];
    if constexpr (Cfg::DSUM) { if (a.out.dsum) a.out.dsum[rank] = a.m_dsum[id]; }
  }
  atomicAdd(a.status + ST_NGROUPS, 1u);
}

}  // namespace pa
