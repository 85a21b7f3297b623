// Multi-GPU exchange + merge.  Rows are sharded by contiguous range; every GPU
// aggregates its shard (stages 1-3), then
//   1. k_partials_count / k_partials_scatter bucket the local groups by owner = hash(key) % P into
//      fixed-size records, grouped by destination rank (the send buffer of the all-to-all);
//   2. the host exchanges counts and records with an all-to-all over NVLink (NCCL);
//   3. k_merge_insert / k_merge_fold join the received records by key and fold them IN SOURCE-RANK
//      ORDER (at most one record per source and key), so fp64 sums do not depend on arrival order;
//   4. the merged groups are ordered by their global first row.
// The reference has no multi-device path at all; this is new structure.
#pragma once
#include "group_result.cuh"

namespace pa {

// One partial aggregate = PA_REC_WORDS 64-bit words.
constexpr int REC_KEY = 0, REC_FLAGS = 1, REC_SUM = 2, REC_DSUM = 3, REC_COUNT = 4, REC_FIRST_ROW = 5,
              REC_LAST_ROW = 6, REC_MIN = 7, REC_MAX = 8, REC_FIRST_VAL = 9, REC_LAST_VAL = 10, REC_WORDS = 11;
constexpr uint64_t RF_KEY_NULL = 1, RF_FIRST_VALID = 2, RF_LAST_VALID = 4;

__device__ __forceinline__ uint32_t owner_of(uint64_t key, uint8_t kind, uint32_t nparts) {
  if (kind == KK_NULL) return 0u;
  return static_cast<uint32_t>(hash_key64(key) % nparts);
}

struct PartialsArgs {
  GroupResult r;
  uint32_t G;
  uint32_t nparts;
  int64_t row_base;            // global row number of local row 0
  // optional first/last values as emitted by k_emit (typed, width vw) + validity bitmaps
  const void* first_vals; const uint32_t* first_valid;
  const void* last_vals; const uint32_t* last_valid;
  int vw;
  bool wide;
  unsigned long long* counts;  // [nparts] (device)
  unsigned long long* cursor;  // [nparts] running write offsets (device), pre-set to the exclusive prefix of counts
  uint64_t* records;
};

__global__ void __launch_bounds__(256) k_partials_count(PartialsArgs a) {
  __shared__ unsigned int s_cnt[64];
  if (threadIdx.x < 64) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < a.G) atomicAdd(&s_cnt[owner_of(a.r.key[g], a.r.key_kind[g], a.nparts)], 1u);
  __syncthreads();
  if (threadIdx.x < a.nparts && s_cnt[threadIdx.x]) atomicAdd(a.counts + threadIdx.x, static_cast<unsigned long long>(s_cnt[threadIdx.x]));
}

__device__ __forceinline__ uint64_t load_raw(const void* p, uint32_t i, int vw) {
  switch (vw) {
    case 8: return static_cast<const uint64_t*>(p)[i];
    case 4: return static_cast<const uint32_t*>(p)[i];
    case 2: return static_cast<const uint16_t*>(p)[i];
    default: return static_cast<const uint8_t*>(p)[i];
  }
}

__global__ void __launch_bounds__(256) k_partials_scatter(PartialsArgs a) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= a.G) return;
  const uint8_t kind = a.r.key_kind[g];
  const uint32_t o = owner_of(a.r.key[g], kind, a.nparts);
  const unsigned long long pos = atomicAdd(a.cursor + o, 1ull);
  uint64_t* rec = a.records + pos * REC_WORDS;
  uint64_t flags = kind == KK_NULL ? RF_KEY_NULL : 0;
  rec[REC_KEY] = a.r.key[g];
  rec[REC_SUM] = a.r.sum[g];
  rec[REC_DSUM] = (a.wide && a.r.dsum) ? static_cast<uint64_t>(__double_as_longlong(a.r.dsum[g])) : 0ull;
  rec[REC_COUNT] = a.r.count[g];
  rec[REC_FIRST_ROW] = static_cast<uint64_t>(a.row_base + a.r.first_row[g]);
  rec[REC_LAST_ROW] = a.wide ? static_cast<uint64_t>(a.row_base + a.r.last_row[g]) : 0ull;
  rec[REC_MIN] = a.wide ? a.r.min_ord[g] : kMinInit;
  rec[REC_MAX] = a.wide ? a.r.max_ord[g] : kMaxInit;
  uint64_t fv = 0, lv = 0;
  if (a.first_vals) {
    fv = load_raw(a.first_vals, g, a.vw);
    if (!a.first_valid || ((a.first_valid[g >> 5] >> (g & 31)) & 1u)) flags |= RF_FIRST_VALID;
  }
  if (a.last_vals) {
    lv = load_raw(a.last_vals, g, a.vw);
    if (!a.last_valid || ((a.last_valid[g >> 5] >> (g & 31)) & 1u)) flags |= RF_LAST_VALID;
  }
  rec[REC_FIRST_VAL] = fv;
  rec[REC_LAST_VAL] = lv;
  rec[REC_FLAGS] = flags;
}

struct MergeArgs {
  const uint64_t* records;       // all received records, grouped by source rank
  const uint64_t* src_offset;    // [nsrc + 1] exclusive prefix of the per-source record counts (device)
  uint32_t nsrc;
  uint64_t nrec;
  unsigned long long* tkeys;     // cap + 2 keys
  uint64_t cap_mask;
  uint32_t* idx;                 // [(cap + 2) * nsrc] record index per (slot, source) or 0xFFFFFFFF
  uint32_t* status;
  // compacted, unordered merged groups
  uint64_t* m_first_row;         // sort key
  uint32_t* m_slot;
  // final, ordered
  GroupResult out;
  uint64_t* o_first_val; uint64_t* o_last_val; uint8_t* o_first_valid; uint8_t* o_last_valid;
  uint64_t* o_first_row_g;       // global first row of every merged group
  const uint32_t* order;         // slot of the r-th merged group (after the sort)
  uint32_t G;
  int vc;
};

__global__ void __launch_bounds__(256) k_merge_insert(MergeArgs a) {
  const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
  if (i >= a.nrec) return;
  // source rank of record i (nsrc is small: linear scan of the prefix array)
  uint32_t s = 0;
  while (s + 1 < a.nsrc && i >= a.src_offset[s + 1]) ++s;
  const uint64_t* rec = a.records + i * REC_WORDS;
  const uint64_t key = rec[REC_KEY];
  const uint64_t cap = a.cap_mask + 1;
  uint64_t slot;
  if (rec[REC_FLAGS] & RF_KEY_NULL) slot = cap;
  else if (key == kEmptyKey) slot = cap + 1;
  else {
    slot = hash_key64(key ^ 0xA5A5A5A5A5A5A5A5ull) & a.cap_mask;   // different mix than owner_of: owners share hash residues
    bool done = false;
    for (uint64_t probe = 0; probe <= a.cap_mask; ++probe) {
      const uint64_t k = __ldcg(a.tkeys + slot);
      if (k == key) { done = true; break; }
      if (k == kEmptyKey) {
        const uint64_t old = atomicCAS(a.tkeys + slot, static_cast<unsigned long long>(kEmptyKey), static_cast<unsigned long long>(key));
        if (old == kEmptyKey || old == key) { done = true; break; }
      }
      slot = (slot + 1) & a.cap_mask;
    }
    if (!done) { atomicExch(a.status + ST_OVERFLOW, 1u); return; }
  }
  a.idx[slot * a.nsrc + s] = static_cast<uint32_t>(i);
}

// One thread per slot: if any source contributed, append (global first row, slot) to the compact list.
__global__ void __launch_bounds__(256) k_merge_compact(MergeArgs a) {
  const uint64_t nslots = a.cap_mask + 3;
  const uint64_t slot = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
  uint64_t first = ~0ull;
  if (slot < nslots) {
    for (uint32_t s = 0; s < a.nsrc; ++s) {
      const uint32_t i = a.idx[slot * a.nsrc + s];
      if (i == 0xFFFFFFFFu) continue;
      const uint64_t f = a.records[static_cast<uint64_t>(i) * REC_WORDS + REC_FIRST_ROW];
      first = f < first ? f : first;
    }
  }
  const bool occ = first != ~0ull;
  const uint32_t m = __ballot_sync(0xFFFFFFFFu, occ);
  if (m) {
    uint32_t base = 0;
    if (lane_id() == 0) base = atomicAdd(a.status + ST_COUNTER, __popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (occ) {
      const uint32_t pos = base + __popc(m & ((1u << lane_id()) - 1u));
      a.m_first_row[pos] = first;
      a.m_slot[pos] = static_cast<uint32_t>(slot);
    }
  }
}

// One thread per merged group (final order): fold the sources in rank order.
__global__ void __launch_bounds__(256) k_merge_fold(MergeArgs a) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= a.G) return;
  const uint64_t slot = a.order[g];
  uint64_t key = 0, sum = 0, mn = kMinInit, mx = kMaxInit, cnt = 0, first = ~0ull, last = 0, fv = 0, lv = 0;
  double fsum = 0.0, dsum = 0.0;
  bool knull = false, fvalid = false, lvalid = false, any_last = false;
  for (uint32_t s = 0; s < a.nsrc; ++s) {
    const uint32_t i = a.idx[slot * a.nsrc + s];
    if (i == 0xFFFFFFFFu) continue;
    const uint64_t* rec = a.records + static_cast<uint64_t>(i) * REC_WORDS;
    key = rec[REC_KEY];
    const uint64_t flags = rec[REC_FLAGS];
    knull = flags & RF_KEY_NULL;
    if (a.vc == VC_F) fsum += __longlong_as_double(static_cast<long long>(rec[REC_SUM]));
    else sum += rec[REC_SUM];
    dsum += __longlong_as_double(static_cast<long long>(rec[REC_DSUM]));
    cnt += rec[REC_COUNT];
    mn = rec[REC_MIN] < mn ? rec[REC_MIN] : mn;
    mx = rec[REC_MAX] > mx ? rec[REC_MAX] : mx;
    if (rec[REC_FIRST_ROW] < first) { first = rec[REC_FIRST_ROW]; fv = rec[REC_FIRST_VAL]; fvalid = flags & RF_FIRST_VALID; }
    if (!any_last || rec[REC_LAST_ROW] >= last) { last = rec[REC_LAST_ROW]; lv = rec[REC_LAST_VAL]; lvalid = flags & RF_LAST_VALID; any_last = true; }
  }
  if (a.vc == VC_F) sum = static_cast<uint64_t>(__double_as_longlong(fsum));
  a.out.key[g] = key;
  a.out.key_kind[g] = knull ? KK_NULL : KK_REGULAR;
  a.out.sum[g] = sum;
  a.out.count[g] = static_cast<uint32_t>(cnt > 0xFFFFFFFFull ? 0xFFFFFFFFull : cnt);
  if (a.out.count64) a.out.count64[g] = cnt;
  a.out.first_row[g] = 0;
  a.out.last_row[g] = 0;
  if (a.out.min_ord) { a.out.min_ord[g] = mn; a.out.max_ord[g] = mx; }
  if (a.out.dsum) a.out.dsum[g] = dsum;
  a.o_first_row_g[g] = first;
  a.o_first_val[g] = fv;
  a.o_last_val[g] = lv;
  a.o_first_valid[g] = fvalid;
  a.o_last_valid[g] = lvalid;
}

__global__ void __launch_bounds__(256) k_fill_u64(unsigned long long* p, uint64_t n, unsigned long long v) {
  uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}

}  // namespace pa
