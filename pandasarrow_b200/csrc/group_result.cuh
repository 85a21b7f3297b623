// Device-side result of stages 1-3: one entry per group, first-appearance order.
#pragma once
#include "common.cuh"

namespace pa {

// key_kind values
constexpr uint8_t KK_REGULAR = 0;   // key[] holds the packed 64-bit key
constexpr uint8_t KK_NULL = 1;      // the null-key group (single-key mode)

struct GroupResult {
  uint64_t* key;        // packed key
  uint8_t* key_kind;    // KK_*
  uint64_t* sum;        // raw bits: double (VC_F) or wrapping (u)int64
  double* dsum;         // sum of values converted to double (VC_I / VC_U mean); may be null for VC_F
  uint32_t* count;      // non-null values
  uint64_t* count64;    // merged (multi-GPU) results only: 64-bit counts, overrides `count` when non-null
  uint32_t* first_row;  // local row index of the first / last row of the group
  uint32_t* last_row;
  uint64_t* min_ord;    // order-mapped min / max over non-null, non-NaN values
  uint64_t* max_ord;
};

// status words shared by the scan / merge kernels and read back by the host
constexpr int ST_OVERFLOW = 0;   // table capacity exceeded -> caller retries on a bigger path
constexpr int ST_NGROUPS = 1;
constexpr int ST_UNSORTED = 2;   // resample: timestamps not sorted
constexpr int ST_COUNTER = 3;    // scratch append counter
constexpr int ST_DENSE_MISS = 4; // low-cardinality dense mode met a key outside its window -> rerun in hash mode
constexpr int ST_PEER_OVERFLOW = 4;  // merge of padded blocks: a source had more groups than a block holds (own status buffer)
constexpr int ST_ABORT = 5;      // some CTA gave up (overflow / dense miss): the others stop early, merge is skipped
constexpr int ST_MODE = 6;       // low-cardinality scan: 1 dense, 2 hash
constexpr int ST_RLOG = 7;       // low-cardinality scan: log2 of the accumulator replication
constexpr int ST_WORDS = 8;

}  // namespace pa
