// Low-cardinality path (stages 1-3 fused): one persistent CTA per SM.
//
//  * Every warp streams its own row chunks HBM -> shared memory with 1-D bulk async copies
//    (cp.async.bulk, the TMA engine) through a private 3-stage mbarrier ring, so the memory
//    pipeline depth does not depend on occupancy (8 warps / SM).
//  * Keys are resolved to dense CTA-local group ids through a CTA-shared hash table in shared
//    memory: 2048 buckets of two 8-byte keys, one LDS.128 per lookup; the steady state is
//    read-only (plain loads, no atomics), a rare out-of-line slow path inserts new keys with
//    ATOMS.CAS.64 and handles null keys / the sentinel-valued key / bucket overflow.
//  * Accumulators are WARP-PRIVATE arrays in shared memory, updated without atomics (sm_100a has
//    no native 64-bit shared-memory atomics: f64/u64 adds compile to CAS spin loops).  Lanes of a
//    warp that hit the same group (MATCH.ANY) are first combined with a shuffle-based segmented
//    reduction in ascending lane (= row) order, then the lowest lane does one read-modify-write.
//    The order of floating-point additions is therefore a fixed function of (n_rows, grid size):
//    results are run-to-run deterministic.
//  * Two 32-row batches are processed interleaved per step so that two independent lookup chains
//    are in flight per warp.
//  * At the end each CTA folds its warps in warp order and writes a partial table; a single-CTA
//    merge kernel joins the partial tables by key, folds them in CTA order, ranks the groups by
//    first row and writes the GroupResult.
// If a CTA sees more than GMAX distinct keys the pass aborts (status[ST_OVERFLOW]) and the host
// reruns the global-table path.
//
// Replaces Grouper::Consume + MakeGroupings + ApplyGroupings + per-group CallFunction
// (/root/reference/src/dataframe.cpp:1571-1600, pd_core_macros.h:5-147).
#pragma once
#include "group_result.cuh"

namespace pa {

constexpr int LC_WARPS = 8;
constexpr int LC_THREADS = LC_WARPS * 32;
constexpr int LC_NB_LOG2 = 11;
constexpr int LC_NBUCKET = 1 << LC_NB_LOG2;  // buckets of two keys
constexpr int LC_TCAP = LC_NBUCKET * 2;      // key slots
constexpr int LC_STAGES = 3;
constexpr int LC_CHUNK = 192;                // rows per warp per stage (3 steps of 2 x 32 rows)
constexpr uint32_t LC_UNSEEN = 0xFFFFFFFFu;  // count sentinel: warp has not met this id yet
constexpr uint16_t LC_ID_UNSET = 0xFFFFu;
constexpr uint16_t LC_ID_OVF = 0xFFFEu;
constexpr uint32_t LC_NOID = 0xFFFFFFFFu;

constexpr int LC_GMAX_NARROW = 1024;   // sum / mean / count / first  (12 B per id per warp)
constexpr int LC_GMAX_WIDE = 352;      // + min / max / last / dsum    (40 B per id per warp)

template <bool WIDE>
struct LcCfg {
  static constexpr int GMAX = WIDE ? LC_GMAX_WIDE : LC_GMAX_NARROW;
  static constexpr int GP = GMAX + 2;             // + null-key group + (key == kEmptyKey) group
  static constexpr int ID_NULL = GMAX;
  static constexpr int ID_EMPTYKEY = GMAX + 1;
};

struct LcArgs {
  const void* keys;
  const void* vals;        // may be null (keys-only pass)
  const uint8_t* kvalid;
  const uint8_t* vvalid;
  int64_t koff, voff;
  int64_t n;
  int64_t n_bulk;          // rows [0, n_bulk) are streamed with bulk copies (multiple of LC_CHUNK; 0 disables)
  uint32_t agg_mask;
  // per-CTA partial tables, [grid][GP]
  uint64_t* p_key;
  uint64_t* p_sum;
  double* p_dsum;
  uint32_t* p_count;
  uint32_t* p_first;
  uint32_t* p_last;
  uint64_t* p_min;
  uint64_t* p_max;
  uint32_t* p_nids;        // [grid]
  uint32_t* status;
};

template <bool WIDE>
struct LcSmem {
  using Cfg = LcCfg<WIDE>;
  // byte offsets inside the dynamic shared memory block
  static constexpr size_t stage_bytes(int kw, int vw) { return static_cast<size_t>(LC_CHUNK) * (kw + vw); }
  static constexpr size_t OFF_BAR = 0;                                           // LC_WARPS*LC_STAGES u64
  static constexpr size_t OFF_TKEYS = 256;                                       // LC_TCAP u64
  static constexpr size_t OFF_TIDS = OFF_TKEYS + LC_TCAP * 8;                    // LC_TCAP u16
  static constexpr size_t OFF_FIRST = OFF_TIDS + LC_TCAP * 2;                    // GP u32 (CTA shared)
  static constexpr size_t OFF_MISC = OFF_FIRST + ((Cfg::GP * 4 + 15) / 16) * 16; // next_id, ovf
  static constexpr size_t OFF_ACC = OFF_MISC + 16;
  static constexpr size_t ACC_PER_ID = WIDE ? 40 : 12;
  static constexpr size_t ACC_PER_WARP = ((Cfg::GP * ACC_PER_ID + 15) / 16) * 16;
  static constexpr size_t OFF_STAGE = OFF_ACC + ACC_PER_WARP * LC_WARPS;
  static constexpr size_t total(int kw, int vw) { return OFF_STAGE + stage_bytes(kw, vw) * LC_STAGES * LC_WARPS; }
};

// Warp-private accumulator views (struct-of-arrays inside the warp's block).
template <bool WIDE>
struct LcAcc {
  using Cfg = LcCfg<WIDE>;
  uint64_t* sum;   // GP
  uint32_t* cnt;   // GP
  uint32_t* last;  // GP (WIDE)
  uint64_t* mn;    // GP (WIDE)
  uint64_t* mx;    // GP (WIDE)
  double* dsum;    // GP (WIDE)
  __device__ __forceinline__ explicit LcAcc(unsigned char* base) {
    sum = reinterpret_cast<uint64_t*>(base);
    if constexpr (WIDE) {
      mn = sum + Cfg::GP;
      mx = mn + Cfg::GP;
      dsum = reinterpret_cast<double*>(mx + Cfg::GP);
      cnt = reinterpret_cast<uint32_t*>(dsum + Cfg::GP);
      last = cnt + Cfg::GP;
    } else {
      cnt = reinterpret_cast<uint32_t*>(sum + Cfg::GP);
      last = nullptr; mn = nullptr; mx = nullptr; dsum = nullptr;
    }
  }
};

__device__ __forceinline__ uint32_t lc_bucket(uint64_t key) {
  const uint32_t lo = static_cast<uint32_t>(key), hi = static_cast<uint32_t>(key >> 32);
  return ((lo * 0x9E3779B1u) ^ (hi * 0x85EBCA77u) ^ (lo >> 15)) * 0x2C1B3C6Du >> (32 - LC_NB_LOG2);
}

// Steady-state lookup of bucket `b`: one LDS.128 (both keys of the bucket) + one LDS.U16.  Returns a
// value >= LC_ID_OVF when the key is not in this bucket, its id is not published yet, or the row's
// key equals the empty sentinel.
__device__ __forceinline__ uint32_t lc_lookup(uint64_t key, uint32_t b, const unsigned long long* tkeys,
                                              const uint16_t* tids) {
  const ulonglong2 kk = *reinterpret_cast<const ulonglong2*>(tkeys + 2 * b);
  const bool h1 = kk.y == key;
  const bool hit = (kk.x == key) | h1;
  const uint32_t id = tids[2 * b + (h1 ? 1u : 0u)];
  return hit ? id : static_cast<uint32_t>(LC_ID_UNSET);
}

// Slow path: null / sentinel-valued keys, insertion on first sight, keys displaced from their home
// bucket.  Per-lane (divergent) code, kept out of line.  LC_NOID on overflow.
template <bool WIDE>
__device__ __noinline__ uint32_t lc_slow_resolve(uint64_t key, bool kvalid, unsigned long long* tkeys,
                                                 volatile uint16_t* tids, uint32_t* misc, uint32_t* status) {
  using Cfg = LcCfg<WIDE>;
  if (!kvalid) return Cfg::ID_NULL;
  if (key == kEmptyKey) return Cfg::ID_EMPTYKEY;
  uint32_t b = lc_bucket(key);
  for (int probe = 0; probe < 4 * LC_NBUCKET; ++probe) {
    const uint64_t k0 = *reinterpret_cast<volatile unsigned long long*>(tkeys + 2 * b);
    const uint64_t k1 = *reinterpret_cast<volatile unsigned long long*>(tkeys + 2 * b + 1);
    uint32_t slot;
    if (k0 == key) slot = 2 * b;
    else if (k1 == key) slot = 2 * b + 1;
    else if (k0 == kEmptyKey || k1 == kEmptyKey) {
      const uint32_t s = (k0 == kEmptyKey) ? 2 * b : 2 * b + 1;   // lowest empty slot of the bucket
      const uint64_t old = atomicCAS(tkeys + s, static_cast<unsigned long long>(kEmptyKey),
                                     static_cast<unsigned long long>(key));
      if (old == kEmptyKey) {  // this lane inserted the key: hand out the next dense id
        const uint32_t nid = atomicAdd(misc, 1u);
        if (nid >= static_cast<uint32_t>(Cfg::GMAX)) {
          tids[s] = LC_ID_OVF;
          misc[1] = 1u;
          atomicExch(status + ST_OVERFLOW, 1u);
          return LC_NOID;
        }
        tids[s] = static_cast<uint16_t>(nid);
        return nid;
      }
      if (old != key) continue;   // lost the slot to another key: look at the same bucket again
      slot = s;
    } else {
      b = (b + 1) & (LC_NBUCKET - 1);
      continue;
    }
    uint16_t id;
    do { id = tids[slot]; } while (id == LC_ID_UNSET);   // inserter publishes the id right after its CAS
    return id == LC_ID_OVF ? LC_NOID : id;
  }
  misc[1] = 1u;
  atomicExch(status + ST_OVERFLOW, 1u);
  return LC_NOID;
}

constexpr uint32_t LC_CNT_MASK = 0x00FFFFFFu;   // count field of an accumulator count word; byte 3 = claim tag

// Accumulate one 32-row batch whose ids are known.  Warp-synchronous.
//
// Duplicate ids inside the batch are found without MATCH.ANY (whose latency grows with the number
// of distinct values, ~360 cycles at 32): every lane stores its lane number into the top byte of
// its group's count word, then loads the word back — one lane per group reads its own number (the
// "winner"), the others learn who won.  Winners then fold the losers' contributions in ascending
// lane order (their own value at its own lane position, so the result does not depend on which
// lane the hardware let win), and do one non-atomic read-modify-write.
template <int VC, bool WIDE>
__device__ __forceinline__ void lc_accumulate(uint32_t id, uint64_t vbits, bool vvalid, uint32_t row,
                                              uint32_t* cta_first, const LcAcc<WIDE>& acc) {
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  const uint32_t lane = lane_id();
  const bool live = id != LC_NOID;
  uint32_t* cw = acc.cnt + (live ? id : 0u);
  if (live) reinterpret_cast<volatile uint8_t*>(cw)[3] = static_cast<uint8_t>(lane);
  __syncwarp();
  const uint32_t word = live ? *reinterpret_cast<volatile uint32_t*>(cw) : 0u;
  const uint32_t w = word >> 24;
  const bool winner = live && (w == lane);
  const uint32_t losers = __ballot_sync(FULL, live && !winner);
  uint64_t c_sum = 0;           // double bits (VC_F, +0.0) or wrapping integer
  uint32_t c_cnt = 0;
  double c_dsum = 0.0;
  uint64_t c_min = kMinInit, c_max = kMaxInit;
  if (vvalid) {
    c_sum = vbits;
    c_cnt = 1;
    if constexpr (WIDE) {
      if constexpr (VC != VC_F) c_dsum = Wide<VC>::as_double(vbits);
      if (!Wide<VC>::is_nan(vbits)) { c_min = Wide<VC>::ord(vbits); c_max = c_min; }
    }
  }
  uint32_t lo_lane = lane, hi_lane = lane;   // lowest / highest lane of my group (meaningful for the doer)
  bool doer = winner;                        // the lane that performs the read-modify-write
  if (losers) {
    auto add_to = [&](uint64_t& t_sum, uint32_t& t_cnt, double& t_dsum, uint64_t& t_min, uint64_t& t_max,
                      uint64_t x_sum, uint32_t x_cnt, double x_dsum, uint64_t x_min, uint64_t x_max) {
      if constexpr (VC == VC_F) {
        t_sum = static_cast<uint64_t>(__double_as_longlong(__longlong_as_double(static_cast<long long>(t_sum)) +
                                                           __longlong_as_double(static_cast<long long>(x_sum))));
      } else {
        t_sum += x_sum;
      }
      t_cnt += x_cnt;
      if constexpr (WIDE) {
        t_dsum += x_dsum;
        t_min = x_min < t_min ? x_min : t_min;
        t_max = x_max > t_max ? x_max : t_max;
      }
    };
    if (__popc(losers) <= 4) {
      // few duplicates: one warp-uniform iteration per loser lane, ascending
      uint64_t f_sum = 0;
      uint32_t f_cnt = 0;
      double f_dsum = 0.0;
      uint64_t f_min = kMinInit, f_max = kMaxInit;
      bool own_added = false;
      uint32_t rem = losers;
      while (rem) {
        const int L = __ffs(rem) - 1;
        rem &= rem - 1;
        const uint32_t tw = __shfl_sync(FULL, w, L);
        const uint64_t o_sum = __shfl_sync(FULL, c_sum, L);
        const uint32_t o_cnt = __shfl_sync(FULL, c_cnt, L);
        double o_dsum = 0.0;
        uint64_t o_min = kMinInit, o_max = kMaxInit;
        if constexpr (WIDE) {
          if constexpr (VC != VC_F) o_dsum = __shfl_sync(FULL, c_dsum, L);
          o_min = __shfl_sync(FULL, c_min, L);
          o_max = __shfl_sync(FULL, c_max, L);
        }
        if (winner && tw == lane) {
          if (!own_added && static_cast<uint32_t>(L) > lane) {
            add_to(f_sum, f_cnt, f_dsum, f_min, f_max, c_sum, c_cnt, c_dsum, c_min, c_max);
            own_added = true;
          }
          add_to(f_sum, f_cnt, f_dsum, f_min, f_max, o_sum, o_cnt, o_dsum, o_min, o_max);
          lo_lane = static_cast<uint32_t>(L) < lo_lane ? static_cast<uint32_t>(L) : lo_lane;
          hi_lane = static_cast<uint32_t>(L) > hi_lane ? static_cast<uint32_t>(L) : hi_lane;
        }
      }
      if (winner) {
        if (!own_added) add_to(f_sum, f_cnt, f_dsum, f_min, f_max, c_sum, c_cnt, c_dsum, c_min, c_max);
        c_sum = f_sum; c_cnt = f_cnt; c_dsum = f_dsum; c_min = f_min; c_max = f_max;
      }
    } else {
      // many duplicates = few distinct ids, where MATCH.ANY is cheap: the lowest lane of every group
      // pulls its peers in ascending lane order and does the read-modify-write
      const uint32_t peers = __match_any_sync(FULL, id);
      const uint32_t lanebit = 1u << lane;
      const bool leader = (peers & (lanebit - 1u)) == 0;
      uint32_t rem = leader ? (peers & ~lanebit) : 0u;
      while (__any_sync(FULL, rem != 0)) {
        const int src = rem ? (__ffs(rem) - 1) : static_cast<int>(lane);
        const uint64_t o_sum = __shfl_sync(FULL, c_sum, src);
        const uint32_t o_cnt = __shfl_sync(FULL, c_cnt, src);
        double o_dsum = 0.0;
        uint64_t o_min = kMinInit, o_max = kMaxInit;
        if constexpr (WIDE) {
          if constexpr (VC != VC_F) o_dsum = __shfl_sync(FULL, c_dsum, src);
          o_min = __shfl_sync(FULL, c_min, src);
          o_max = __shfl_sync(FULL, c_max, src);
        }
        if (rem) {
          add_to(c_sum, c_cnt, c_dsum, c_min, c_max, o_sum, o_cnt, o_dsum, o_min, o_max);
          rem &= rem - 1;
        }
      }
      doer = live && leader;
      lo_lane = lane;
      hi_lane = 31 - __clz(peers);
    }
  }
  // one non-atomic read-modify-write per distinct id
  if (doer) {
    uint32_t old = word & LC_CNT_MASK;
    if (old == LC_CNT_MASK) {  // first time this warp meets the id: candidate for the CTA's first row
      old = 0;
      atomicMin(cta_first + id, row - lane + lo_lane);
    }
    *cw = old + c_cnt;   // also clears the claim tag
    if constexpr (VC == VC_F) {
      double* s = reinterpret_cast<double*>(acc.sum + id);
      *s = *s + __longlong_as_double(static_cast<long long>(c_sum));
    } else {
      acc.sum[id] += c_sum;
    }
    if constexpr (WIDE) {
      acc.last[id] = row - lane + hi_lane;
      if constexpr (VC != VC_F) acc.dsum[id] += c_dsum;
      if (c_min < acc.mn[id]) acc.mn[id] = c_min;
      if (c_max > acc.mx[id]) acc.mx[id] = c_max;
    }
  }
}

// Two 32-row batches (A: rows rowA + lane, B: rows rowA + 32 + lane), interleaved.
template <int VC, bool WIDE>
__device__ __forceinline__ void lc_process_pair(bool actA, bool actB, uint64_t keyA, uint64_t keyB, bool kvA, bool kvB,
                                                uint64_t vbA, uint64_t vbB, bool vvA, bool vvB, uint32_t rowA,
                                                unsigned long long* tkeys, uint16_t* tids, uint32_t* cta_first,
                                                uint32_t* misc, uint32_t* status, const LcAcc<WIDE>& acc) {
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  const uint32_t bA = lc_bucket(keyA), bB = lc_bucket(keyB);
  uint32_t idA = lc_lookup(keyA, bA, tkeys, tids);
  uint32_t idB = lc_lookup(keyB, bB, tkeys, tids);
  bool missA = actA && idA >= LC_ID_OVF, missB = actB && idB >= LC_ID_OVF;
  if (__any_sync(FULL, missA || missB)) {   // keys displaced by one bucket: second inline probe
    if (missA) idA = lc_lookup(keyA, (bA + 1) & (LC_NBUCKET - 1), tkeys, tids);
    if (missB) idB = lc_lookup(keyB, (bB + 1) & (LC_NBUCKET - 1), tkeys, tids);
    missA = actA && idA >= LC_ID_OVF;
    missB = actB && idB >= LC_ID_OVF;
  }
  const bool slowA = missA || (actA && !kvA), slowB = missB || (actB && !kvB);
  if (__any_sync(FULL, slowA || slowB)) {
    if (slowA) idA = lc_slow_resolve<WIDE>(keyA, kvA, tkeys, tids, misc, status);
    __syncwarp();
    if (slowB) idB = lc_slow_resolve<WIDE>(keyB, kvB, tkeys, tids, misc, status);
    __syncwarp();
  }
  if (!actA) idA = LC_NOID;
  if (!actB) idB = LC_NOID;
  lc_accumulate<VC, WIDE>(idA, vbA, vvA && actA && idA != LC_NOID, rowA, cta_first, acc);
  __syncwarp();
  lc_accumulate<VC, WIDE>(idB, vbB, vvB && actB && idB != LC_NOID, rowA + 32u, cta_first, acc);
  __syncwarp();
}

template <int VC, int VW, int KW, bool WIDE, bool NULLS>
__global__ void __launch_bounds__(LC_THREADS, 1) k_lowcard_scan(LcArgs a) {
  using Cfg = LcCfg<WIDE>;
  using L = LcSmem<WIDE>;
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  unsigned long long* tkeys = reinterpret_cast<unsigned long long*>(smem + L::OFF_TKEYS);
  uint16_t* tids = reinterpret_cast<uint16_t*>(smem + L::OFF_TIDS);
  uint32_t* cta_first = reinterpret_cast<uint32_t*>(smem + L::OFF_FIRST);
  uint32_t* misc = reinterpret_cast<uint32_t*>(smem + L::OFF_MISC);   // [0] next id, [1] overflow seen
  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  LcAcc<WIDE> acc(smem + L::OFF_ACC + L::ACC_PER_WARP * warp);
  constexpr size_t STAGE_BYTES = static_cast<size_t>(LC_CHUNK) * (KW + VW);
  unsigned char* my_stage = smem + L::OFF_STAGE + STAGE_BYTES * LC_STAGES * warp;
  uint64_t* my_bar = bars + warp * LC_STAGES;

  // ---- init shared state ----
  for (int i = threadIdx.x; i < LC_TCAP; i += LC_THREADS) {
    tkeys[i] = kEmptyKey;
    tids[i] = LC_ID_UNSET;
  }
  for (int i = threadIdx.x; i < Cfg::GP; i += LC_THREADS) cta_first[i] = kNoRow;
  if (threadIdx.x < 4) misc[threadIdx.x] = 0;
  for (int i = lane; i < Cfg::GP; i += 32) {
    acc.sum[i] = 0;
    acc.cnt[i] = LC_UNSEEN;
    if constexpr (WIDE) { acc.last[i] = 0; acc.mn[i] = kMinInit; acc.mx[i] = kMaxInit; acc.dsum[i] = 0.0; }
  }
  if (lane == 0) {
    for (int s = 0; s < LC_STAGES; ++s) mbar_init(my_bar + s, 1);
    mbar_fence_init();
  }
  __syncthreads();

  const bool have_vals = a.vals != nullptr;
  const bool have_kvalid = NULLS && a.kvalid != nullptr, have_vvalid = NULLS && a.vvalid != nullptr;
  const int64_t gw = static_cast<int64_t>(blockIdx.x) * LC_WARPS + warp;
  const int64_t nw = static_cast<int64_t>(gridDim.x) * LC_WARPS;
  const int64_t nchunks = a.n_bulk / LC_CHUNK;
  const char* kbase = static_cast<const char*>(a.keys);
  const char* vbase = static_cast<const char*>(a.vals);

  auto issue = [&](int64_t chunk, int s) {  // lane 0 only
    unsigned char* dst = my_stage + STAGE_BYTES * s;
    const uint32_t kbytes = LC_CHUNK * KW, vbytes = have_vals ? LC_CHUNK * VW : 0;
    mbar_expect_tx(my_bar + s, kbytes + vbytes);
    bulk_g2s(dst, kbase + chunk * (LC_CHUNK * KW), kbytes, my_bar + s);
    if (have_vals) bulk_g2s(dst + LC_CHUNK * KW, vbase + chunk * (LC_CHUNK * VW), vbytes, my_bar + s);
  };

  // ---- streamed part: chunks gw, gw + nw, ... (fixed chunk -> warp map) ----
  // Every issued bulk copy is waited for before the CTA may exit (also on overflow), so no copy
  // can land in shared memory that already belongs to another CTA.
  if (gw < nchunks) {
    int64_t last_issued = gw;
    for (int s = 0; s < LC_STAGES; ++s) {
      const int64_t c = gw + static_cast<int64_t>(s) * nw;
      if (c < nchunks) {
        if (lane == 0) issue(c, s);
        last_issued = c;
      }
    }
    int s = 0;
    uint32_t phase = 0;
    for (int64_t c = gw; c <= last_issued; c += nw) {
      mbar_wait(my_bar + s, phase);
      // (only the CTA-local flag is polled: a global poll costs one system-scope load per chunk, and
      //  every CTA meets the same key population, so each one notices an overflow by itself)
      const bool ovf = __any_sync(0xFFFFFFFFu, *reinterpret_cast<volatile uint32_t*>(misc + 1) != 0);
      if (!ovf) {
        const unsigned char* st = my_stage + STAGE_BYTES * s;
        const int64_t row0 = c * LC_CHUNK;
#pragma unroll 1
        for (int step = 0; step < LC_CHUNK / 64; ++step) {
          const int rA = step * 64 + lane, rB = rA + 32;
          uint64_t keyA, keyB;
          if constexpr (KW == 8) {
            keyA = reinterpret_cast<const uint64_t*>(st)[rA];
            keyB = reinterpret_cast<const uint64_t*>(st)[rB];
          } else {
            keyA = reinterpret_cast<const uint32_t*>(st)[rA];
            keyB = reinterpret_cast<const uint32_t*>(st)[rB];
          }
          uint64_t vbA = 0, vbB = 0;
          bool vvA = false, vvB = false, kvA = true, kvB = true;
          const int64_t rowA = row0 + rA;
          if (have_vals) {
            vbA = load_wide<VC, VW>(st + LC_CHUNK * KW, rA);
            vbB = load_wide<VC, VW>(st + LC_CHUNK * KW, rB);
            vvA = vvB = true;
            if (have_vvalid) { vvA = bit_at(a.vvalid, a.voff + rowA); vvB = bit_at(a.vvalid, a.voff + rowA + 32); }
          }
          if (have_kvalid) { kvA = bit_at(a.kvalid, a.koff + rowA); kvB = bit_at(a.kvalid, a.koff + rowA + 32); }
          lc_process_pair<VC, WIDE>(true, true, keyA, keyB, kvA, kvB, vbA, vbB, vvA, vvB, static_cast<uint32_t>(rowA),
                                    tkeys, tids, cta_first, misc, a.status, acc);
        }
      }
      __syncwarp();
      const int64_t cn = c + static_cast<int64_t>(LC_STAGES) * nw;
      if (!ovf && cn < nchunks) {
        if (lane == 0) {
          fence_proxy_async();
          issue(cn, s);
        }
        last_issued = cn;
      }
      if (++s == LC_STAGES) { s = 0; phase ^= 1; }
    }
  }

  // ---- remainder rows [n_bulk, n): direct loads, warps of the whole grid take 64-row steps ----
  for (int64_t r0 = a.n_bulk + gw * 64; r0 < a.n; r0 += nw * 64) {
    if (__any_sync(0xFFFFFFFFu, *reinterpret_cast<volatile uint32_t*>(misc + 1) != 0)) break;
    const int64_t rowA = r0 + lane, rowB = rowA + 32;
    const bool actA = rowA < a.n, actB = rowB < a.n;
    uint64_t keyA = 0, keyB = 0, vbA = 0, vbB = 0;
    bool kvA = true, kvB = true, vvA = false, vvB = false;
    if (actA) {
      keyA = load_key<KW>(a.keys, rowA);
      if (have_kvalid) kvA = bit_at(a.kvalid, a.koff + rowA);
      if (have_vals) { vbA = load_wide<VC, VW>(a.vals, rowA); vvA = have_vvalid ? bit_at(a.vvalid, a.voff + rowA) : true; }
    }
    if (actB) {
      keyB = load_key<KW>(a.keys, rowB);
      if (have_kvalid) kvB = bit_at(a.kvalid, a.koff + rowB);
      if (have_vals) { vbB = load_wide<VC, VW>(a.vals, rowB); vvB = have_vvalid ? bit_at(a.vvalid, a.voff + rowB) : true; }
    }
    lc_process_pair<VC, WIDE>(actA, actB, keyA, keyB, kvA, kvB, vbA, vbB, vvA, vvB, static_cast<uint32_t>(rowA), tkeys,
                              tids, cta_first, misc, a.status, acc);
  }
  __syncthreads();
  if (misc[1]) return;

  // ---- fold the warps in warp order and write this CTA's partial table ----
  const size_t pbase = static_cast<size_t>(blockIdx.x) * Cfg::GP;
  for (int id = threadIdx.x; id < Cfg::GP; id += LC_THREADS) {
    uint64_t sum = 0;
    double fsum = 0.0, dsum = 0.0;
    uint32_t cnt = 0, last = 0;
    uint64_t mn = kMinInit, mx = kMaxInit;
    for (int w = 0; w < LC_WARPS; ++w) {
      LcAcc<WIDE> o(smem + L::OFF_ACC + L::ACC_PER_WARP * w);
      const uint32_t c = o.cnt[id] & LC_CNT_MASK;
      if (c == LC_CNT_MASK) continue;
      cnt += c;
      if constexpr (VC == VC_F) fsum += __longlong_as_double(static_cast<long long>(o.sum[id]));
      else sum += o.sum[id];
      if constexpr (WIDE) {
        dsum += o.dsum[id];
        last = o.last[id] > last ? o.last[id] : last;
        mn = o.mn[id] < mn ? o.mn[id] : mn;
        mx = o.mx[id] > mx ? o.mx[id] : mx;
      }
    }
    if constexpr (VC == VC_F) sum = static_cast<uint64_t>(__double_as_longlong(fsum));
    a.p_sum[pbase + id] = sum;
    a.p_count[pbase + id] = cnt;
    a.p_first[pbase + id] = cta_first[id];
    if constexpr (WIDE) {
      a.p_dsum[pbase + id] = dsum;
      a.p_last[pbase + id] = last;
      a.p_min[pbase + id] = mn;
      a.p_max[pbase + id] = mx;
    }
  }
  for (int slot = threadIdx.x; slot < LC_TCAP; slot += LC_THREADS) {
    const uint16_t id = tids[slot];
    if (id < Cfg::GMAX) a.p_key[pbase + id] = tkeys[slot];
  }
  if (threadIdx.x == 0) a.p_nids[blockIdx.x] = misc[0] < static_cast<uint32_t>(Cfg::GMAX) ? misc[0] : Cfg::GMAX;
}

// ---------------------------------------------------------------------------------------------
// Merge: one CTA.  Joins the per-CTA partial tables by key, folds them in CTA order, ranks the
// merged groups by first row, writes the GroupResult and status[ST_NGROUPS].
// ---------------------------------------------------------------------------------------------
constexpr int LM_THREADS = 1024;
constexpr int LM_TCAP_LOG2 = 12;
constexpr int LM_TCAP = 1 << LM_TCAP_LOG2;

struct LmArgs {
  LcArgs part;          // partial tables written by the scan
  int grid;             // number of scan CTAs
  uint16_t* inv;        // [GP][grid] scratch, pre-filled with 0xFFFF: merged id, CTA -> CTA-local id
  GroupResult out;
  uint32_t* status;
};

template <int VC, bool WIDE>
__global__ void __launch_bounds__(LM_THREADS, 1) k_lowcard_merge(LmArgs a) {
  using Cfg = LcCfg<WIDE>;
  __shared__ unsigned long long mkeys[LM_TCAP];
  __shared__ uint16_t mids[LM_TCAP];
  __shared__ uint32_t sfirst[Cfg::GP];
  __shared__ uint32_t m_next, m_ovf;
  if (*reinterpret_cast<volatile uint32_t*>(a.status + ST_OVERFLOW)) return;
  for (int i = threadIdx.x; i < LM_TCAP; i += LM_THREADS) { mkeys[i] = kEmptyKey; mids[i] = LC_ID_UNSET; }
  for (int i = threadIdx.x; i < Cfg::GP; i += LM_THREADS) sfirst[i] = kNoRow;
  if (threadIdx.x == 0) { m_next = 0; m_ovf = 0; }
  __syncthreads();
  const int grid = a.grid;
  // phase 1: join by key
  for (int idx = threadIdx.x; idx < grid * Cfg::GMAX; idx += LM_THREADS) {
    const int b = idx / Cfg::GMAX, id = idx - b * Cfg::GMAX;
    if (static_cast<uint32_t>(id) >= a.part.p_nids[b]) continue;
    const uint64_t key = a.part.p_key[static_cast<size_t>(b) * Cfg::GP + id];
    uint32_t slot = hash_key(key) >> (32 - LM_TCAP_LOG2);
    uint32_t mid = LC_NOID;
    for (int probe = 0; probe < LM_TCAP; ++probe) {
      const uint64_t k = *reinterpret_cast<volatile unsigned long long*>(mkeys + slot);
      bool found = (k == key);
      if (!found && k == kEmptyKey) {
        const uint64_t old = atomicCAS(mkeys + slot, static_cast<unsigned long long>(kEmptyKey),
                                       static_cast<unsigned long long>(key));
        if (old == kEmptyKey) {
          const uint32_t nid = atomicAdd(&m_next, 1u);
          if (nid >= static_cast<uint32_t>(Cfg::GMAX)) { m_ovf = 1; *reinterpret_cast<volatile uint16_t*>(mids + slot) = LC_ID_OVF; break; }
          *reinterpret_cast<volatile uint16_t*>(mids + slot) = static_cast<uint16_t>(nid);
          mid = nid;
          break;
        }
        found = (old == key);
      }
      if (found) {
        uint16_t v;
        do { v = *reinterpret_cast<volatile uint16_t*>(mids + slot); } while (v == LC_ID_UNSET);
        mid = (v == LC_ID_OVF) ? LC_NOID : v;
        break;
      }
      slot = (slot + 1) & (LM_TCAP - 1);
    }
    if (mid != LC_NOID) a.inv[static_cast<size_t>(mid) * grid + b] = static_cast<uint16_t>(id);
  }
  __syncthreads();
  if (m_ovf) {
    if (threadIdx.x == 0) atomicExch(a.status + ST_OVERFLOW, 1u);
    return;
  }
  const int M = m_next;
  __threadfence_block();
  __syncthreads();
  // phase 2: fold partials in CTA order (merged groups M and M+1 are the two special groups);
  // up to two merged groups per thread (GMAX + 2 may exceed the block size)
  constexpr int U = (Cfg::GP + LM_THREADS - 1) / LM_THREADS;
  uint64_t r_sum[U], r_key[U], r_min[U], r_max[U];
  double r_dsum[U];
  uint32_t r_cnt[U], r_first[U], r_last[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int m = threadIdx.x + u * LM_THREADS;
    r_sum[u] = 0; r_key[u] = 0; r_min[u] = kMinInit; r_max[u] = kMaxInit; r_dsum[u] = 0.0;
    r_cnt[u] = 0; r_first[u] = kNoRow; r_last[u] = 0;
    if (m >= M + 2) continue;
    double fsum = 0.0;
    for (int b = 0; b < grid; ++b) {
      int id;
      if (m < M) {
        const uint16_t v = a.inv[static_cast<size_t>(m) * grid + b];
        if (v == LC_ID_UNSET) continue;
        id = v;
      } else {
        id = (m == M) ? Cfg::ID_NULL : Cfg::ID_EMPTYKEY;
      }
      const size_t p = static_cast<size_t>(b) * Cfg::GP + id;
      const uint32_t f = a.part.p_first[p];
      if (f == kNoRow) continue;
      if (m < M) r_key[u] = a.part.p_key[p];
      r_first[u] = f < r_first[u] ? f : r_first[u];
      r_cnt[u] += a.part.p_count[p];
      if constexpr (VC == VC_F) fsum += __longlong_as_double(static_cast<long long>(a.part.p_sum[p]));
      else r_sum[u] += a.part.p_sum[p];
      if constexpr (WIDE) {
        r_dsum[u] += a.part.p_dsum[p];
        const uint32_t l = a.part.p_last[p];
        r_last[u] = l > r_last[u] ? l : r_last[u];
        const uint64_t mn = a.part.p_min[p], mx = a.part.p_max[p];
        r_min[u] = mn < r_min[u] ? mn : r_min[u];
        r_max[u] = mx > r_max[u] ? mx : r_max[u];
      }
    }
    if constexpr (VC == VC_F) r_sum[u] = static_cast<uint64_t>(__double_as_longlong(fsum));
    if (m == M + 1) r_key[u] = kEmptyKey;
    sfirst[m] = r_first[u];
  }
  __syncthreads();
  // phase 3: rank by first row (first rows are distinct: a row belongs to exactly one group)
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int m = threadIdx.x + u * LM_THREADS;
    if (m >= M + 2 || r_first[u] == kNoRow) continue;
    uint32_t rank = 0;
    for (int j = 0; j < M + 2; ++j) rank += sfirst[j] < r_first[u];
    a.out.key[rank] = r_key[u];
    a.out.key_kind[rank] = (m == M) ? KK_NULL : KK_REGULAR;
    a.out.sum[rank] = r_sum[u];
    a.out.count[rank] = r_cnt[u];
    a.out.first_row[rank] = r_first[u];
    if constexpr (WIDE) {
      a.out.last_row[rank] = r_last[u];
      a.out.min_ord[rank] = r_min[u];
      a.out.max_ord[rank] = r_max[u];
      if (a.out.dsum) a.out.dsum[rank] = r_dsum[u];
    }
  }
  if (threadIdx.x == 0) {
    uint32_t G = M;
    G += sfirst[M] != kNoRow;
    G += sfirst[M + 1] != kNoRow;
    a.status[ST_NGROUPS] = G;
  }
}

}  // namespace pa
