// C ABI of the B200 group-by path (include/pa_b200.h).  Host-side orchestration only: column
// marshalling from the Arrow C Data Interface, path selection, kernel launches, result export.
// No computation happens on the host and there is no CPU fallback.
#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>
#include <nccl.h>

#include <algorithm>
#include <cstdarg>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pa_b200.h"
#include "emit.cuh"
#include "gtable.cuh"
#include "lowcard.cuh"
#include "merge.cuh"
#include "partition.cuh"
#include "bucketed.cuh"
#include "resample.cuh"
#include "rowids.cuh"
#include "stage2.cuh"
#include "groupings.cuh"
#include "distinct.cuh"
#include "temporal.cuh"
#include "order.cuh"
#include "csort.cuh"
#include "sort.cuh"
#include "strkeys.cuh"

using namespace pa;

namespace {

thread_local std::string g_err;

int set_err(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                             \
  do {                                                                                             \
    cudaError_t e__ = (expr);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return set_err(PA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

#define PA_TRY(expr)        \
  do {                      \
    int rc__ = (expr);      \
    if (rc__ != PA_OK) return rc__; \
  } while (0)

// Stream-ordered device allocation; the pool keeps freed memory so that repeated aggregate
// calls (bench steps) do not pay cudaMalloc.
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  cudaStream_t s = nullptr;
  bool plain = false;     // cudaMalloc / cudaFree instead of the stream-ordered pool: for buffers that outlive the streams they are used on
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept { *this = std::move(o); }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { reset(); p = o.p; bytes = o.bytes; s = o.s; plain = o.plain; o.p = nullptr; o.bytes = 0; }
    return *this;
  }
  ~DevBuf() { reset(); }
  void reset() {
    if (p) { if (plain) cudaFree(p); else cudaFreeAsync(p, s); }
    p = nullptr;
    bytes = 0;
  }
  int alloc(size_t n, cudaStream_t stream) {
    if (n == 0) n = 16;
    if (plain) {
      if (p && bytes >= n) return PA_OK;
      reset();
      CUDA_TRY(cudaMalloc(&p, n));
      bytes = n;
      return PA_OK;
    }
    if (p && s == stream && bytes >= n) return PA_OK;   // recycled handle: the old block is big enough
    reset();
    s = stream;
    static const bool dbg = getenv("PA_DEBUG_ALLOC") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    CUDA_TRY(cudaMallocAsync(&p, n, stream));
    if (dbg) {
      const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      if (ms > 2.0) fprintf(stderr, "[pa alloc] %.1f MB took %.1f ms\n", n / 1048576.0, ms);
    }
    bytes = n;
    return PA_OK;
  }
  template <typename T>
  T* as() const { return static_cast<T*>(p); }
};

struct Column {
  const void* data = nullptr;       // device pointer, already advanced by `offset` elements
  const uint8_t* valid = nullptr;   // device validity bitmap (not advanced) or null
  int64_t bit_off = 0;
  int64_t n = 0;
  int width = 0;                    // bytes per element
  int vc = VC_I;                    // value class (value columns)
  std::string format;               // Arrow format string of the (index) type
  bool is_dict = false;
  bool is_bool = false;             // Arrow boolean (bit-packed) unpacked to one byte per value in own_data
  int64_t dict_len = 0;
  DevBuf own_data, own_valid;       // set when the input lived on the host (or was unpacked)
  DevBuf own_bits;                  // boolean host input: the packed bits on the device
  // utf8 / large_utf8 KEY columns (strkeys.cuh): `data` is the per-row 64-bit hash, the strings stay reachable here
  bool is_str = false;
  StrCol str{};
  DevBuf own_offsets, own_bytes;
  uint64_t str_seed = 0;
};

int parse_format(const char* f, int* width, int* vc) {
  if (!f) return set_err(PA_ERR_INVALID, "schema has no format string");
  switch (f[0]) {
    case 'g': *width = 8; *vc = VC_F; return PA_OK;
    case 'f': *width = 4; *vc = VC_F; return PA_OK;
    case 'l': *width = 8; *vc = VC_I; return PA_OK;
    case 'L': *width = 8; *vc = VC_U; return PA_OK;
    case 'i': *width = 4; *vc = VC_I; return PA_OK;
    case 'I': *width = 4; *vc = VC_U; return PA_OK;
    case 's': *width = 2; *vc = VC_I; return PA_OK;
    case 'S': *width = 2; *vc = VC_U; return PA_OK;
    case 'c': *width = 1; *vc = VC_I; return PA_OK;
    case 'C': *width = 1; *vc = VC_U; return PA_OK;
    case 'b': *width = 1; *vc = VC_U; return PA_OK;   // boolean: unpacked to bytes by load_column
    case 't':
      // tss/tsm/tsu/tsn timestamps, tD* durations, ttu/ttn time64, tdm date64: 64-bit; tdD date32, tts/ttm: 32-bit
      if (f[1] == 's' || f[1] == 'D') { *width = 8; *vc = VC_I; return PA_OK; }
      if (f[1] == 't') { *width = (f[2] == 'u' || f[2] == 'n') ? 8 : 4; *vc = VC_I; return PA_OK; }
      if (f[1] == 'd') { *width = (f[2] == 'm') ? 8 : 4; *vc = VC_I; return PA_OK; }
      break;
    default: break;
  }
  return set_err(PA_ERR_INVALID, "unsupported Arrow format '%s' (fixed-width numeric / temporal only)", f);
}

// Host -> device copy of a column buffer.  Pinned source: one cudaMemcpyAsync at the PCIe rate.  Pageable source
// (ordinary Arrow heap buffers — what pd::DataFrame holds —, an IPC blob, Parquet-decoded columns): the driver would
// bounce it through its own small staging area at a fraction of that rate, so anything large goes through the
// library's two pinned staging buffers, filled by a few host threads while the previous chunk is on the wire.
namespace {
constexpr size_t H2D_CHUNK = 8u << 20;            // bytes per staging buffer (two per worker); PA_H2D_CHUNK_MB overrides
constexpr size_t H2D_STAGED_MIN = 16u << 20;      // smaller copies are not worth the pipeline
constexpr int H2D_MAX_WORKERS = 16;
constexpr int H2D_MAX_BUFS = 4;
struct H2dWorker {
  void* pin[H2D_MAX_BUFS] = {};
  cudaEvent_t done[H2D_MAX_BUFS] = {};
  cudaEvent_t fin = nullptr;
  cudaStream_t stream = nullptr;
};
struct H2dStage {
  std::mutex mu;
  H2dWorker w[H2D_MAX_WORKERS];
  cudaEvent_t start = nullptr;
  bool ok = false, tried = false;
  int workers = 1;
  size_t chunk = H2D_CHUNK;
  int bufs = 2;                                    // staging buffers per worker; PA_H2D_BUFS overrides
};
H2dStage g_h2d_dev[16];   // one per device: streams and events belong to a device

bool h2d_stage_ready(H2dStage& g) {
  if (g.tried) return g.ok;
  g.tried = true;
  if (const char* e = getenv("PA_H2D_STAGED")) { if (atoi(e) == 0) return false; }
  const unsigned hc = std::thread::hardware_concurrency();
  g.workers = static_cast<int>(std::max(1u, std::min(8u, hc ? hc / 2 : 4u)));
  if (const char* e = getenv("PA_H2D_THREADS")) g.workers = std::max(1, std::min(H2D_MAX_WORKERS, atoi(e)));
  if (const char* e = getenv("PA_H2D_BUFS")) g.bufs = std::max(2, std::min(H2D_MAX_BUFS, atoi(e)));
  if (const char* e = getenv("PA_H2D_CHUNK_MB")) g.chunk = static_cast<size_t>(std::max(1, std::min(64, atoi(e)))) << 20;
  if (cudaEventCreateWithFlags(&g.start, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return false; }
  for (int i = 0; i < g.workers; ++i) {
    H2dWorker& w = g.w[i];
    bool good = cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking) == cudaSuccess &&
                cudaEventCreateWithFlags(&w.fin, cudaEventDisableTiming) == cudaSuccess;
    for (int b = 0; b < g.bufs && good; ++b)
      good = cudaHostAlloc(&w.pin[b], g.chunk, cudaHostAllocPortable) == cudaSuccess &&
             cudaEventCreateWithFlags(&w.done[b], cudaEventDisableTiming) == cudaSuccess;
    if (!good) { cudaGetLastError(); return false; }
  }
  g.ok = true;
  return true;
}
}  // namespace

// Every worker thread owns a contiguous slice of the copy, two pinned buffers and a stream: it copies a chunk into
// one buffer while the DMA engine drains the other.  The destination was allocated in stream order on `st`, so the
// worker streams start behind an event on `st`, and `st` continues behind the workers' last copies.
int h2d_copy(void* dst, const void* src, size_t bytes, cudaStream_t st) {
  if (!bytes) return PA_OK;
  bool pageable = false;
  if (bytes >= H2D_STAGED_MIN) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, src) != cudaSuccess) { cudaGetLastError(); pageable = true; }
    else pageable = at.type == cudaMemoryTypeUnregistered;
  }
  if (!pageable) {
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
    return PA_OK;
  }
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  H2dStage& g = g_h2d_dev[dev & 15];
  std::lock_guard<std::mutex> lock(g.mu);
  if (!h2d_stage_ready(g)) {
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
    return PA_OK;
  }
  CUDA_TRY(cudaEventRecord(g.start, st));
  const int T = g.workers;
  const size_t per = ((bytes + T - 1) / T + 4095) & ~static_cast<size_t>(4095);
  std::vector<cudaError_t> err(T, cudaSuccess);
  auto work = [&](int t) {
    H2dWorker& w = g.w[t];
    cudaError_t e = cudaSetDevice(dev);
    const size_t lo = per * t, hi = std::min(bytes, lo + per);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(w.stream, g.start, 0);
    int b = 0;
    for (size_t off = lo; off < hi && e == cudaSuccess; off += g.chunk, b = b + 1 == g.bufs ? 0 : b + 1) {
      const size_t len = std::min(g.chunk, hi - off);
      e = cudaEventSynchronize(w.done[b]);                 // the previous copy out of this buffer has finished
      if (e != cudaSuccess) break;
      memcpy(w.pin[b], static_cast<const char*>(src) + off, len);
      e = cudaMemcpyAsync(static_cast<char*>(dst) + off, w.pin[b], len, cudaMemcpyHostToDevice, w.stream);
      if (e == cudaSuccess) e = cudaEventRecord(w.done[b], w.stream);
    }
    if (e == cudaSuccess) e = cudaEventRecord(w.fin, w.stream);
    err[t] = e;
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < T; ++t) pool.emplace_back(work, t);
  work(0);
  for (auto& th : pool) th.join();
  for (int t = 0; t < T; ++t) {
    if (err[t] != cudaSuccess) return set_err(PA_ERR_CUDA, "staged host-to-device copy failed: %s", cudaGetErrorString(err[t]));
    CUDA_TRY(cudaStreamWaitEvent(st, g.w[t].fin, 0));
  }
  return PA_OK;
}

// Bring one primitive Arrow array onto the device (or borrow it when it already is there).
int load_column(const ArrowDeviceArray* da, const ArrowSchema* sc, cudaStream_t st, int device, Column* out) {
  const ArrowArray& a = da->array;
  if (a.n_buffers < 2) return set_err(PA_ERR_INVALID, "expected a primitive array with 2 buffers, got %lld", (long long)a.n_buffers);
  PA_TRY(parse_format(sc->format, &out->width, &out->vc));
  out->format = sc->format;
  out->is_dict = sc->dictionary != nullptr;
  out->dict_len = (a.dictionary != nullptr) ? a.dictionary->length : 0;
  out->n = a.length;
  const bool has_nulls = a.null_count != 0 && a.buffers[0] != nullptr;
  const char* values = static_cast<const char*>(a.buffers[1]);
  out->is_bool = sc->format[0] == 'b';
  if (out->is_bool) {
    // Bit-packed values -> one byte per value on the device (the kernels then see a uint8 column of 0 / 1).
    if (da->device_type == ARROW_DEVICE_CUDA && da->device_id != device)
      return set_err(PA_ERR_INVALID, "array lives on device %lld, handle on %d", (long long)da->device_id, device);
    const bool on_dev = da->device_type == ARROW_DEVICE_CUDA;
    if (!on_dev && da->device_type != ARROW_DEVICE_CPU && da->device_type != ARROW_DEVICE_CUDA_HOST)
      return set_err(PA_ERR_INVALID, "unsupported device_type %d", (int)da->device_type);
    if (on_dev && da->sync_event) CUDA_TRY(cudaStreamWaitEvent(st, *static_cast<cudaEvent_t*>(da->sync_event), 0));
    const int64_t first_byte = a.offset / 8;
    const size_t nbytes = static_cast<size_t>((a.offset + a.length + 7) / 8 - first_byte);
    const uint8_t* bits = reinterpret_cast<const uint8_t*>(values) + first_byte;
    if (!on_dev) {
      PA_TRY(out->own_bits.alloc(std::max<size_t>(nbytes, 1), st));
      PA_TRY(h2d_copy(out->own_bits.p, bits, nbytes, st));
      bits = out->own_bits.as<uint8_t>();
    }
    PA_TRY(out->own_data.alloc(static_cast<size_t>(std::max<int64_t>(a.length, 1)), st));
    if (a.length > 0) {
      const int grid = static_cast<int>(std::min<int64_t>((a.length + 255) / 256, 148 * 16));
      k_unpack_bool<<<grid, 256, 0, st>>>(bits, a.offset % 8, a.length, out->own_data.as<uint8_t>());
      CUDA_TRY(cudaGetLastError());
    }
    out->data = out->own_data.p;
    if (has_nulls) {
      if (on_dev) {
        out->valid = static_cast<const uint8_t*>(a.buffers[0]);
        out->bit_off = a.offset;
      } else {
        PA_TRY(out->own_valid.alloc(nbytes, st));
        PA_TRY(h2d_copy(out->own_valid.p, static_cast<const uint8_t*>(a.buffers[0]) + first_byte, nbytes, st));
        out->valid = out->own_valid.as<uint8_t>();
        out->bit_off = a.offset % 8;
      }
    }
    return PA_OK;
  }
  if (da->device_type == ARROW_DEVICE_CUDA) {
    if (da->device_id != device) return set_err(PA_ERR_INVALID, "array lives on device %lld, handle on %d", (long long)da->device_id, device);
    out->data = values ? values + a.offset * out->width : nullptr;
    out->valid = has_nulls ? static_cast<const uint8_t*>(a.buffers[0]) : nullptr;
    out->bit_off = a.offset;
    if (da->sync_event) CUDA_TRY(cudaStreamWaitEvent(st, *static_cast<cudaEvent_t*>(da->sync_event), 0));
  } else if (da->device_type == ARROW_DEVICE_CPU || da->device_type == ARROW_DEVICE_CUDA_HOST) {
    const size_t bytes = static_cast<size_t>(a.length) * out->width;
    PA_TRY(out->own_data.alloc(bytes, st));
    PA_TRY(h2d_copy(out->own_data.p, values + a.offset * out->width, bytes, st));
    out->data = out->own_data.p;
    if (has_nulls) {
      // copy the bytes that cover bits [offset, offset+length); keep the sub-byte offset
      const int64_t first_byte = a.offset / 8;
      const size_t nbytes = static_cast<size_t>((a.offset + a.length + 7) / 8 - first_byte);
      PA_TRY(out->own_valid.alloc(nbytes, st));
      PA_TRY(h2d_copy(out->own_valid.p, static_cast<const uint8_t*>(a.buffers[0]) + first_byte, nbytes, st));
      out->valid = out->own_valid.as<uint8_t>();
      out->bit_off = a.offset % 8;
    }
  } else {
    return set_err(PA_ERR_INVALID, "unsupported device_type %d", (int)da->device_type);
  }
  return PA_OK;
}

// utf8 / large_utf8 key column: offsets + bytes onto the device (or borrowed), then one 64-bit hash per row.
int hash_string_key(Column* out, cudaStream_t st, int num_sms) {
  PA_TRY(out->own_data.alloc(static_cast<size_t>(std::max<int64_t>(out->n, 1)) * 8, st));
  if (out->n > 0) {
    const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((out->n + 255) / 256, static_cast<int64_t>(num_sms) * 16)));
    k_str_hash<<<grid, 256, 0, st>>>(out->str, out->valid, out->bit_off, out->n, 0x5851F42D4C957F2Dull + out->str_seed * 0x9E3779B97F4A7C15ull,
                                     out->own_data.as<uint64_t>());
    CUDA_TRY(cudaGetLastError());
  }
  out->data = out->own_data.p;
  return PA_OK;
}

int load_string_key(const ArrowDeviceArray* da, const ArrowSchema* sc, cudaStream_t st, int device, int num_sms, Column* out) {
  const ArrowArray& a = da->array;
  if (a.n_buffers < 3) return set_err(PA_ERR_INVALID, "utf8 array needs 3 buffers (validity, offsets, data), got %lld", (long long)a.n_buffers);
  const bool wide = sc->format[0] == 'U';
  const int ow = wide ? 8 : 4;
  out->format = sc->format;
  out->is_str = true;
  out->width = 8;
  out->vc = VC_U;
  out->n = a.length;
  const bool has_nulls = a.null_count != 0 && a.buffers[0] != nullptr;
  const char* offs = static_cast<const char*>(a.buffers[1]) + a.offset * ow;
  const uint8_t* bytes = static_cast<const uint8_t*>(a.buffers[2]);
  if (da->device_type == ARROW_DEVICE_CUDA) {
    if (da->device_id != device) return set_err(PA_ERR_INVALID, "array lives on device %lld, handle on %d", (long long)da->device_id, device);
    if (da->sync_event) CUDA_TRY(cudaStreamWaitEvent(st, *static_cast<cudaEvent_t*>(da->sync_event), 0));
    out->str.offsets = offs;
    out->str.bytes = bytes;
    out->valid = has_nulls ? static_cast<const uint8_t*>(a.buffers[0]) : nullptr;
    out->bit_off = a.offset;
  } else if (da->device_type == ARROW_DEVICE_CPU || da->device_type == ARROW_DEVICE_CUDA_HOST) {
    const int64_t n = a.length;
    int64_t first = 0, last = 0;
    if (a.buffers[1]) {
      first = wide ? reinterpret_cast<const int64_t*>(offs)[0] : reinterpret_cast<const int32_t*>(offs)[0];
      last = wide ? reinterpret_cast<const int64_t*>(offs)[n] : reinterpret_cast<const int32_t*>(offs)[n];
    }
    PA_TRY(out->own_offsets.alloc(static_cast<size_t>(n + 1) * ow, st));
    if (a.buffers[1]) PA_TRY(h2d_copy(out->own_offsets.p, offs, static_cast<size_t>(n + 1) * ow, st));
    else CUDA_TRY(cudaMemsetAsync(out->own_offsets.p, 0, static_cast<size_t>(n + 1) * ow, st));
    PA_TRY(out->own_bytes.alloc(static_cast<size_t>(std::max<int64_t>(last - first, 1)), st));
    if (last > first) PA_TRY(h2d_copy(out->own_bytes.p, bytes + first, static_cast<size_t>(last - first), st));
    out->str.offsets = out->own_offsets.p;
    out->str.bytes = out->own_bytes.as<uint8_t>() - first;       // offsets stay absolute
    if (has_nulls) {
      const int64_t first_byte = a.offset / 8;
      const size_t nbytes = static_cast<size_t>((a.offset + a.length + 7) / 8 - first_byte);
      PA_TRY(out->own_valid.alloc(nbytes, st));
      PA_TRY(h2d_copy(out->own_valid.p, static_cast<const uint8_t*>(a.buffers[0]) + first_byte, nbytes, st));
      out->valid = out->own_valid.as<uint8_t>();
      out->bit_off = a.offset % 8;
    }
  } else {
    return set_err(PA_ERR_INVALID, "unsupported device_type %d", (int)da->device_type);
  }
  out->str.wide = wide ? 1 : 0;
  return hash_string_key(out, st, num_sms);
}

struct KeyField {
  int width = 8, bits = 64, shift = 0, nullable = 0;
};

struct AggOut {
  uint32_t bit = 0;
  std::string format;
  int width = 8;
  bool nullable = true;
  DevBuf values, valid;
};

}  // namespace

struct pa_groupby {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int num_sms = 148;
  pa_options opt{};
  int64_t n = 0;
  // keys
  std::vector<Column> keys;
  std::vector<KeyField> fields;
  bool packed = false;
  bool str_verified = false;        // utf8 key: the hash grouping has been checked byte for byte (strkeys.cuh)
  DevBuf packed_keys;
  const void* key_data = nullptr;   // what the scan kernels read
  const uint8_t* key_valid = nullptr;
  int64_t key_bit_off = 0;
  int key_width = 8;
  // resample mode (time-bucket specialisation)
  bool resample = false;
  std::string index_format;
  ResampleSpec rs{};
  DevBuf rs_edges, rs_labels;             // calendar rules: bucket edges / labels on the device (rs.edges / rs.labels)
  // group table in first-appearance order
  bool have_groups = false;
  // Sharded step with compact records: the bucketed path may leave its unordered per-bucket group records (BkRec32) as
  // the result of the local pass — the merge orders by GLOBAL first row anyway — instead of ranking and gathering them
  // into the GroupResult (8.5 ms per 100 M groups).  want_unordered: asked for by pa_groupby_sharded_aggregate;
  // unordered: this pass did so (g->G records at unordered_recs; the GroupResult and the outputs were NOT touched).
  bool want_unordered = false, unordered = false;
  const void* unordered_recs = nullptr;
  uint32_t unordered_G = 0;
  uint32_t G = 0;
  uint32_t res_cap = 0;
  DevBuf r_key, r_kind, r_sum, r_dsum, r_count, r_first, r_last, r_min, r_max;
  GroupResult res{};
  // outputs of the last aggregate
  std::vector<AggOut> outs;
  bool last_wide = false;
  int last_vc = VC_I, last_vw = 8;
  std::string last_vfmt = "l";
  // multi-GPU
  int parts_n = 0;
  std::vector<int64_t> parts_counts;
  bool merged = false;
  DevBuf m_count64, m_first_val, m_last_val, m_first_valid, m_last_valid, m_first_row_g;
  // bookkeeping
  DevBuf status;
  int last_path = 0, last_launches = 0;
  int last_mode = 0, last_rlog = 0, last_passes = 0;
  const uint32_t* emit_G_dev = nullptr;   // run_emit: exact group count lives on the device (g->G is an upper bound)
  bool poolable = false;                  // merged handle built by the fused small merge: recycled on destroy
  // deferred (asynchronous) aggregate: kernels are queued, the status words have not been read yet
  bool pending = false;
  Column pending_val;                     // borrowed device pointers of the value column (no ownership)
  bool pending_has_val = false;
  uint32_t pending_mask = 0;
  uint32_t pending_prev_G = 0;
  bool pending_had_groups = false;
  // deferred fused merge of padded blocks (multi-GPU): status not read yet; the caller keeps the blocks alive
  bool pending_merge = false;
  const void* pm_blocks = nullptr;
  int32_t pm_nsrc = 0;
  int64_t pm_block_records = 0;
  uint32_t pm_mask = 0;
  std::string pm_vfmt, pm_kfmt;
  float last_total_ms = 0;
  float stage_ms[4] = {0, 0, 0, 0};
  cudaEvent_t ev[10] = {};                // [6..7]: groupings build, [8..9]: last grouped take
  // group materialisation (groupings.cuh), built on first use; keys are immutable so it never goes stale
  DevBuf grp_order, grp_offsets, grp_dest;
  bool have_groupings = false;
  bool sorted_state = false;              // handle made by pa_sort_create: grp_order / grp_dest hold the sorted row order
  // Scratch of the global-table / resample passes, kept between calls on the handle: returning multi-GB blocks to
  // the stream-ordered pool and asking for them again fragments it (cudaMallocAsync then takes 50-1700 ms per call
  // at 100 M groups); a repeated aggregate on the same handle reuses these without touching the allocator.
  struct Scratch {
    DevBuf table, p_keys, p_vals, p_rows, p_counts, krange, c_first, c_slot, s_first, s_slot, cub_tmp, bnd, bitmap, prefix8, tile_sums;
    // bucketed path (bucketed.cuh): second-level rows, histograms / cursors, sketch, unordered groups
    DevBuf q_keys, q_vals, q_rows, rp_fine, rp_counts1, rp_counts2, rp_ends1, rp_ends2, rp_tiles, rp_hll, rp_next, u_rec, u_first;
  } scr;
  double est_groups = 0;                  // bucketed path: HyperLogLog estimate of the group count
  bool bk_have_hist = false;              // bucketed path: level-1 fine histogram + sketch of this handle's keys exist
  int bk_bits = -1, bk_b1 = -1;           // bucketed path: plan whose offsets / bucket ends are cached in scr.rp_*
};

namespace {

int alloc_result(pa_groupby* g, uint32_t cap, bool wide, bool need_dsum) {
  cudaStream_t st = g->stream;
  if (cap == 0) cap = 1;
  g->res_cap = cap;
  PA_TRY(g->r_key.alloc(sizeof(uint64_t) * cap, st));
  PA_TRY(g->r_kind.alloc(cap, st));
  PA_TRY(g->r_sum.alloc(sizeof(uint64_t) * cap, st));
  PA_TRY(g->r_count.alloc(sizeof(uint32_t) * cap, st));
  PA_TRY(g->r_first.alloc(sizeof(uint32_t) * cap, st));
  PA_TRY(g->r_last.alloc(sizeof(uint32_t) * cap, st));
  g->res = GroupResult{};
  g->res.key = g->r_key.as<uint64_t>();
  g->res.key_kind = g->r_kind.as<uint8_t>();
  g->res.sum = g->r_sum.as<uint64_t>();
  g->res.count = g->r_count.as<uint32_t>();
  g->res.first_row = g->r_first.as<uint32_t>();
  g->res.last_row = g->r_last.as<uint32_t>();
  if (wide) {
    PA_TRY(g->r_min.alloc(sizeof(uint64_t) * cap, st));
    PA_TRY(g->r_max.alloc(sizeof(uint64_t) * cap, st));
    g->res.min_ord = g->r_min.as<uint64_t>();
    g->res.max_ord = g->r_max.as<uint64_t>();
    if (need_dsum) {
      PA_TRY(g->r_dsum.alloc(sizeof(double) * cap, st));
      g->res.dsum = g->r_dsum.as<double>();
    }
  }
  return PA_OK;
}

int run_resample(pa_groupby* g, const Column* val, uint32_t mask, bool wide);

bool is_wide(uint32_t mask, int vc) {
  return (mask & (AGG_MIN | AGG_MAX | AGG_LAST)) || ((mask & AGG_MEAN) && vc != VC_F);
}

// ------------------------------ low-cardinality path ------------------------------
enum LcOutcome { LC_DONE = 0, LC_DENSE_MISS = 1, LC_OVERFLOW = 2 };

bool lc_is_wide(uint32_t mask, int vc) { return is_wide(mask, vc); }

template <int VC, bool WIDE>
int launch_lowcard_t(pa_groupby* g, const LcArgs& a, const LmArgs& m, int grid, bool fast, bool hash_kernel) {
  using Cfg = LcCfg<VC, WIDE>;
  cudaStream_t st = g->stream;
  k_lowcard_prep<<<LC_PREP_GRID, 256, 0, st>>>(a);
  CUDA_TRY(cudaGetLastError());
  if (hash_kernel) {
    using L = LcSmem<VC, WIDE, true>;
    auto scan = fast ? k_lowcard_scan<VC, WIDE, true, false> : k_lowcard_scan<VC, WIDE, false, false>;
    CUDA_TRY(cudaFuncSetAttribute(scan, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L::TOTAL)));
    scan<<<grid, L::WARPS * 32, L::TOTAL, st>>>(a);
  } else {
    using L = LcSmem<VC, WIDE, false>;
    auto scan = fast ? k_lowcard_scan<VC, WIDE, true, true> : k_lowcard_scan<VC, WIDE, false, true>;
    CUDA_TRY(cudaFuncSetAttribute(scan, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L::TOTAL)));
    scan<<<grid, L::WARPS * 32, L::TOTAL, st>>>(a);
  }
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaEventRecord(g->ev[2], st));
  k_lowcard_merge<VC, WIDE><<<(Cfg::GP + 7) / 8, 256, 0, st>>>(m);
  CUDA_TRY(cudaGetLastError());
  k_lowcard_rank<VC, WIDE><<<(Cfg::GP + LR_THREADS / 32 - 1) / (LR_THREADS / 32), LR_THREADS, 0, st>>>(m);
  CUDA_TRY(cudaGetLastError());
  g->last_launches += 4;
  return PA_OK;
}

int run_lowcard(pa_groupby* g, const Column* val, uint32_t mask, bool wide, bool force_hash, LcOutcome* outcome,
                bool deferred = false) {
  cudaStream_t st = g->stream;
  const int grid = std::max<int>(1, g->num_sms - static_cast<int>(std::max<int64_t>(0, std::min<int64_t>(g->opt.sm_reserve, g->num_sms - 1))));
  const int vc = val ? val->vc : VC_I;
  const bool kwide = lc_is_wide(mask, vc);
  const bool dsum = kwide && vc != VC_F;
  const int gp = lc_gmax(vc, kwide) + 2;
  const size_t np = static_cast<size_t>(grid) * gp;
  DevBuf p_sum, p_dsum, p_count, p_first, p_last, p_min, p_max, merged, dir;
  PA_TRY(p_sum.alloc(np * 8, st));
  PA_TRY(p_count.alloc(np * 4, st));
  PA_TRY(p_first.alloc(np * 4, st));
  if (kwide) {
    PA_TRY(p_last.alloc(np * 4, st));
    PA_TRY(p_min.alloc(np * 8, st));
    PA_TRY(p_max.alloc(np * 8, st));
    if (dsum) PA_TRY(p_dsum.alloc(np * 8, st));
  }
  // merged-by-id scratch: sum, dsum, min, max (8 B each) then count, first, last (4 B each)
  const size_t gp8 = static_cast<size_t>(gp) * 8, gp4 = static_cast<size_t>(gp) * 4;
  PA_TRY(merged.alloc(gp8 * 4 + gp4 * 3, st));
  // directory: LcPrep | key_by_id | gt_keys | gt_ids
  const size_t dir_bytes = sizeof(LcPrep) + static_cast<size_t>(LC_GMAX_MAX) * 8 + static_cast<size_t>(LC_GT_CAP) * 12;
  PA_TRY(dir.alloc(dir_bytes, st));
  CUDA_TRY(cudaMemsetAsync(dir.p, 0, sizeof(LcPrep), st));
  CUDA_TRY(cudaMemsetAsync(g->status.p, 0, sizeof(uint32_t) * ST_WORDS, st));
  PA_TRY(alloc_result(g, gp, wide, vc != VC_F));

  LcArgs a{};
  a.keys = g->key_data;
  a.kvalid = g->key_valid;
  a.koff = g->key_bit_off;
  a.kw = g->key_width;
  a.vals = val ? val->data : nullptr;
  a.vvalid = val ? val->valid : nullptr;
  a.voff = val ? val->bit_off : 0;
  a.vw = val ? val->width : 8;
  a.n = g->n;
  a.force_hash = force_hash ? 1 : 0;
  a.hash_rlog = 0;
  if (force_hash) {
    // replicas per id for few scattered keys, from a group count this handle already knows (its keys never change) or
    // the caller's hint; a count that turns out too small overflows to the global path like any other overflow
    const uint64_t known = g->have_groups ? g->G : static_cast<uint64_t>(std::max<int64_t>(g->opt.expected_groups, 0));
    const uint64_t cap = static_cast<uint64_t>(lc_gmax_hash(vc, kwide));
    if (known > 0) while (a.hash_rlog < 5 && ((known + 1) << (a.hash_rlog + 1)) <= cap) ++a.hash_rlog;
  }
  a.agg_mask = mask;
  char* d = dir.as<char>();
  a.dir.prep = reinterpret_cast<LcPrep*>(d);
  a.dir.key_by_id = reinterpret_cast<unsigned long long*>(d + sizeof(LcPrep));
  a.dir.gt_keys = a.dir.key_by_id + LC_GMAX_MAX;
  a.dir.gt_ids = reinterpret_cast<unsigned int*>(a.dir.gt_keys + LC_GT_CAP);
  a.p_sum = p_sum.as<uint64_t>();
  a.p_dsum = p_dsum.as<double>();
  a.p_count = p_count.as<uint32_t>();
  a.p_first = p_first.as<uint32_t>();
  a.p_last = p_last.as<uint32_t>();
  a.p_min = p_min.as<uint64_t>();
  a.p_max = p_max.as<uint64_t>();
  a.status = g->status.as<uint32_t>();
  const bool aligned = (reinterpret_cast<uintptr_t>(a.keys) % 8 == 0) && (reinterpret_cast<uintptr_t>(a.vals) % 8 == 0);
  const bool fast = a.vals && aligned && a.kw == 8 && a.vw == 8 && !a.kvalid && !a.vvalid;
  LmArgs m{};
  m.part = a;
  m.grid = grid;
  m.gp = gp;
  char* mp = merged.as<char>();
  m.m_sum = reinterpret_cast<uint64_t*>(mp);
  m.m_dsum = reinterpret_cast<double*>(mp + gp8);
  m.m_min = reinterpret_cast<uint64_t*>(mp + gp8 * 2);
  m.m_max = reinterpret_cast<uint64_t*>(mp + gp8 * 3);
  m.m_count = reinterpret_cast<uint32_t*>(mp + gp8 * 4);
  m.m_first = m.m_count + gp;
  m.m_last = m.m_first + gp;
  m.out = g->res;
  m.status = a.status;
  CUDA_TRY(cudaEventRecord(g->ev[1], st));
  if (kwide) {
    if (vc == VC_F) PA_TRY((launch_lowcard_t<VC_F, true>(g, a, m, grid, fast, force_hash)));
    else if (vc == VC_I) PA_TRY((launch_lowcard_t<VC_I, true>(g, a, m, grid, fast, force_hash)));
    else PA_TRY((launch_lowcard_t<VC_U, true>(g, a, m, grid, fast, force_hash)));
  } else {
    if (vc == VC_F) PA_TRY((launch_lowcard_t<VC_F, false>(g, a, m, grid, fast, force_hash)));
    else if (vc == VC_I) PA_TRY((launch_lowcard_t<VC_I, false>(g, a, m, grid, fast, force_hash)));
    else PA_TRY((launch_lowcard_t<VC_U, false>(g, a, m, grid, fast, force_hash)));
  }
  CUDA_TRY(cudaEventRecord(g->ev[3], st));
  if (deferred) {   // the scratch buffers above are released in stream order; the status is read by finish_pending()
    g->G = static_cast<uint32_t>(gp);   // upper bound until then
    *outcome = LC_DONE;
    return PA_OK;
  }
  uint32_t h_status[ST_WORDS];
  CUDA_TRY(cudaMemcpyAsync(h_status, g->status.p, sizeof h_status, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (h_status[ST_OVERFLOW]) *outcome = LC_OVERFLOW;
  else if (h_status[ST_DENSE_MISS]) *outcome = LC_DENSE_MISS;
  else { *outcome = LC_DONE; g->G = h_status[ST_NGROUPS]; }
  g->last_mode = static_cast<int>(h_status[ST_MODE]);
  g->last_rlog = static_cast<int>(h_status[ST_RLOG]);
  if (h_status[ST_DENSE_MISS] != 2u) g->last_passes += 1;   // (2 = the dense kernel declined before scanning a row)
  return PA_OK;
}

// Orders G compacted (first_row, slot) pairs by first row WITHOUT a sort (order.cuh): rank query on a bitmap of the
// first rows.  Writes the slots in first-appearance order to `s_slot`.
int order_by_first_row(pa_groupby* g, const uint32_t* c_first, const uint32_t* c_slot, uint32_t G, uint32_t* s_slot) {
  if (G == 0) return PA_OK;
  cudaStream_t st = g->stream;
  const uint64_t n_rows = static_cast<uint64_t>(std::max<int64_t>(g->n, 1));
  const uint64_t nblocks = (n_rows + 32ull * BM_BLOCK_WORDS - 1) / (32ull * BM_BLOCK_WORDS);
  const uint64_t nwords = nblocks * BM_BLOCK_WORDS;
  const uint32_t ntiles = static_cast<uint32_t>((nblocks + SC_TILE - 1) / SC_TILE);
  PA_TRY(g->scr.bitmap.alloc(nwords * 4, st));
  PA_TRY(g->scr.prefix8.alloc(nblocks * 4, st));
  PA_TRY(g->scr.tile_sums.alloc(static_cast<size_t>(ntiles) * 4, st));
  uint32_t* bitmap = g->scr.bitmap.as<uint32_t>();
  uint32_t* prefix8 = g->scr.prefix8.as<uint32_t>();
  uint32_t* tile_sums = g->scr.tile_sums.as<uint32_t>();
  CUDA_TRY(cudaMemsetAsync(bitmap, 0, nwords * 4, st));
  k_bm_set<<<(G + 255) / 256, 256, 0, st>>>(c_first, G, bitmap);
  CUDA_TRY(cudaGetLastError());
  const int cgrid = static_cast<int>(std::min<uint64_t>((nblocks + 255) / 256, static_cast<uint64_t>(g->num_sms) * 16));
  k_bm_count8<<<cgrid, 256, 0, st>>>(bitmap, nblocks, prefix8);
  CUDA_TRY(cudaGetLastError());
  k_scan_tiles<<<ntiles, 256, 0, st>>>(prefix8, nblocks, tile_sums);
  CUDA_TRY(cudaGetLastError());
  k_scan_sums<<<1, 256, 0, st>>>(tile_sums, ntiles);
  CUDA_TRY(cudaGetLastError());
  k_scan_apply<<<ntiles, 256, 0, st>>>(prefix8, nblocks, tile_sums);
  CUDA_TRY(cudaGetLastError());
  k_bm_rank<<<(G + 255) / 256, 256, 0, st>>>(c_first, c_slot, G, bitmap, prefix8, s_slot);
  CUDA_TRY(cudaGetLastError());
  g->last_launches += 6;
  return PA_OK;
}

// ------------------------------ global-table path ------------------------------

// device-wide exclusive scan of a uint32 array in place (order.cuh)
int scan_u32_on(pa_groupby* g, uint32_t* data, uint64_t n, DevBuf* tile_sums) {
  cudaStream_t st = g->stream;
  const uint32_t ntiles = static_cast<uint32_t>((n + SC_TILE - 1) / SC_TILE);
  PA_TRY(tile_sums->alloc(static_cast<size_t>(std::max<uint32_t>(ntiles, 1)) * 4, st));
  if (n == 0) return PA_OK;
  k_scan_tiles<<<ntiles, 256, 0, st>>>(data, n, tile_sums->as<uint32_t>());
  CUDA_TRY(cudaGetLastError());
  k_scan_sums<<<1, 256, 0, st>>>(tile_sums->as<uint32_t>(), ntiles);
  CUDA_TRY(cudaGetLastError());
  k_scan_apply<<<ntiles, 256, 0, st>>>(data, n, tile_sums->as<uint32_t>());
  CUDA_TRY(cudaGetLastError());
  return PA_OK;
}

// HyperLogLog estimate from the registers k_rp_hist1 filled (one key value in 2^HLL_SAMPLE_LOG2 feeds the sketch).
double hll_estimate(const uint32_t* reg) {
  const double m = HLL_M;
  double sum = 0;
  int zeros = 0;
  for (int i = 0; i < HLL_M; ++i) { sum += std::ldexp(1.0, -static_cast<int>(reg[i])); zeros += reg[i] == 0; }
  double e = (0.7213 / (1.0 + 1.079 / m)) * m * m / sum;
  if (e <= 2.5 * m && zeros > 0) e = m * std::log(m / zeros);      // small-range correction (linear counting)
  return e * static_cast<double>(1u << HLL_SAMPLE_LOG2);
}

// One partition level.  At 256 ways and more a 4096-row tile leaves runs of <= 16 rows (128-byte pieces of the key / value
// arrays, 64 of the row numbers): those take 8192-row tiles on one 1024-thread CTA per SM instead of two 512-thread CTAs
// with 4096-row tiles.  Measured per 1 B rows: 100 M groups (256 x 256) 45.0 -> 43.2 ms; at 128 ways the lost overlap
// between the two CTAs costs more than the longer runs give (16 M groups, 128 x 128: 32.0 -> 34.4 ms), so the threshold
// is 2^8.  PA_RP_BIG_TILE_LOG (environment, read once) moves it for A/B runs.
int launch_rp_scatter(pa_groupby* g, const RpArgs& a) {
  static const int big_from = [] { const char* e = std::getenv("PA_RP_BIG_TILE_LOG"); return e ? std::atoi(e) : 8; }();
  cudaStream_t st = g->stream;
  const bool nullable = !a.rows && a.vvalid;          // level 1 of a nullable value column: validity bit -> row-number word
  if (a.log_fan >= big_from) {
    auto kern = nullable ? k_rp_scatter_t<RP_THREADS_BIG, true> : k_rp_scatter_t<RP_THREADS_BIG, false>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(RpSmemBig::TOTAL)));
    kern<<<g->num_sms, RP_THREADS_BIG, RpSmemBig::TOTAL, st>>>(a);
  } else {
    auto kern = nullable ? k_rp_scatter_t<RP_THREADS, true> : k_rp_scatter_t<RP_THREADS, false>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(RpSmem::TOTAL)));
    kern<<<g->num_sms * 2, RP_THREADS, RpSmem::TOTAL, st>>>(a);
  }
  CUDA_TRY(cudaGetLastError());
  return PA_OK;
}

// The bucketed path (bucketed.cuh).  *declined = true: the estimate says a single shared-memory table holds the
// groups, or a bucket overflowed its table — the caller continues on the global-table path.
template <int VC, bool WIDE>
int run_bucketed_t(pa_groupby* g, const Column* val, uint32_t mask, uint64_t* cap_io, bool* declined, bool* table_filled) {
  using T = SmTab<VC, WIDE>;
  using SlotT = typename SlotOf<WIDE>::type;
  *table_filled = false;
  cudaStream_t st = g->stream;
  *declined = false;
  auto& sc = g->scr;
  const int64_t n = g->n;
  const uint64_t* keys = static_cast<const uint64_t*>(g->key_data);
  const uint64_t* vals = val ? static_cast<const uint64_t*>(val->data) : nullptr;
  CUDA_TRY(cudaEventRecord(g->ev[1], st));
  // ---- level 1 histogram (1024 fine buckets per chunk) + sketch ----
  // Everything up to the scatter depends on the KEYS only (the scatter keeps its cursors private), and a handle's keys
  // never change: histograms, offsets and bucket ends are computed by the first pass and reused by every later one.
  const uint32_t nchunks1 = static_cast<uint32_t>((n + RP_CHUNK_ROWS - 1) / RP_CHUNK_ROWS);
  CUDA_TRY(cudaMemsetAsync(g->status.p, 0, sizeof(uint32_t) * ST_WORDS, st));
  RpArgs a1{};
  a1.keys = keys; a1.vals = vals; a1.rows = nullptr; a1.n = n;
  a1.vvalid = val ? val->valid : nullptr; a1.voff = val ? val->bit_off : 0;
  a1.n_parents = 1; a1.n_chunks1 = nchunks1;
  if (!g->bk_have_hist) {
    PA_TRY(sc.rp_fine.alloc(sizeof(uint32_t) * RP_MAX_FAN * static_cast<size_t>(nchunks1), st));
    PA_TRY(sc.rp_hll.alloc(sizeof(uint32_t) * HLL_M, st));
    CUDA_TRY(cudaMemsetAsync(sc.rp_hll.p, 0, sizeof(uint32_t) * HLL_M, st));
    a1.counts = sc.rp_fine.as<unsigned int>(); a1.hll = sc.rp_hll.as<unsigned int>();
    k_rp_hist<true><<<g->num_sms * 2, RH_THREADS, 0, st>>>(a1);
    CUDA_TRY(cudaGetLastError());
    std::vector<uint32_t> reg(HLL_M);
    CUDA_TRY(cudaMemcpyAsync(reg.data(), sc.rp_hll.p, sizeof(uint32_t) * HLL_M, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    g->est_groups = hll_estimate(reg.data());
    g->bk_have_hist = true;
    g->bk_bits = g->bk_b1 = -1;
    g->last_launches += 1;
  }
  const double est = g->est_groups;
  const uint64_t hint = g->opt.expected_groups > 0 ? static_cast<uint64_t>(g->opt.expected_groups) : (g->have_groups ? g->G : 0);
  // plan for the larger of hint and estimate (the sketch is good to a few per cent; a wrong hint must not overflow the tables)
  const uint64_t plan = std::max<uint64_t>(hint, static_cast<uint64_t>(est * 1.05) + 64);
  if (!hint && plan <= static_cast<uint64_t>(T::CAP)) { *declined = true; return PA_OK; }   // the front table (gtable.cuh) holds them
  // Buckets of ~CAP/2 keys on average (the table admits 3/4 CAP; bucket sizes are Poisson around the mean).  Up to
  // 1024 buckets in one level; more = two levels with the bits split evenly (long runs in both scatters).
  const uint64_t per_bucket = T::CAP / 2;
  int bits = 1;
  while (bits < RP_MAX_BITS && (plan >> bits) > per_bucket) ++bits;
  int b1 = bits <= RP_SINGLE_MAX_LOG ? bits : (bits + 1) / 2;
  if (g->opt.bucket_bits > 0) {                         // tuning / test override: level-1 bits | level-2 bits << 8
    b1 = static_cast<int>(g->opt.bucket_bits & 0xFF);
    const int b2o = static_cast<int>((g->opt.bucket_bits >> 8) & 0xFF);
    if (b1 < 1 || b1 > RP_L1_LOG || b2o > RP_L1_LOG) return set_err(PA_ERR_INVALID, "bucket bits override out of range");
    bits = b1 + b2o;
  }
  const int b2 = bits - b1;
  const uint32_t nb = 1u << bits;
  const size_t pad = 4;
  const bool replan = g->bk_bits != bits || g->bk_b1 != b1;    // (first pass, or the group-count hint moved the plan)
  // ---- level 1: fold the fine counts to 2^b1 buckets, scan, bucket ends, scatter ----
  const size_t flat1 = (static_cast<size_t>(nchunks1) << b1) + 1;
  if (replan) {
    PA_TRY(sc.rp_counts1.alloc(sizeof(uint32_t) * flat1, st));
    PA_TRY(sc.rp_ends1.alloc(sizeof(uint32_t) * (1u << b1), st));
    CUDA_TRY(cudaMemsetAsync(sc.rp_counts1.as<uint32_t>() + (flat1 - 1), 0, sizeof(uint32_t), st));
    k_rp_fold<<<static_cast<unsigned>((flat1 - 1 + 255) / 256), 256, 0, st>>>(sc.rp_fine.as<unsigned int>(), nchunks1, b1, sc.rp_counts1.as<unsigned int>());
    CUDA_TRY(cudaGetLastError());
    PA_TRY(scan_u32_on(g, sc.rp_counts1.as<uint32_t>(), flat1, &sc.tile_sums));
    g->last_launches += 4;
  }
  PA_TRY(sc.p_keys.alloc((static_cast<size_t>(n) + pad) * 8, st));
  if (vals) PA_TRY(sc.p_vals.alloc((static_cast<size_t>(n) + pad) * 8, st));
  PA_TRY(sc.p_rows.alloc((static_cast<size_t>(n) + pad) * 4, st));
  a1.shift = 64 - b1; a1.log_fan = b1;
  a1.offsets = sc.rp_counts1.as<unsigned int>();
  a1.out_keys = sc.p_keys.as<uint64_t>(); a1.out_vals = vals ? sc.p_vals.as<uint64_t>() : nullptr; a1.out_rows = sc.p_rows.as<uint32_t>();
  if (replan) {
    k_rp_ends<<<((1u << b1) + 255) / 256, 256, 0, st>>>(a1, sc.rp_ends1.as<unsigned int>());
    CUDA_TRY(cudaGetLastError());
    g->last_launches += 1;
  }
  PA_TRY(launch_rp_scatter(g, a1));
  g->last_launches += 1;
  BkArgs b{};
  b.keys = sc.p_keys.as<uint64_t>(); b.vals = vals ? sc.p_vals.as<uint64_t>() : nullptr; b.rows = sc.p_rows.as<uint32_t>();
  b.bucket_end = sc.rp_ends1.as<unsigned int>();
  b.nullable = (val && val->valid) ? 1 : 0;
  if (b2 > 0) {
    // ---- level 2 inside every level-1 bucket ----
    const int np = 1 << b1;
    const size_t chunks2_max = static_cast<size_t>(n) / RP_CHUNK_ROWS + np + 1;
    const size_t flat2 = (chunks2_max << b2) + 1;
    // (Only level 1 is cached: inside a tile's run the level-1 scatter orders rows by atomic arrival, so the rows on
    // either side of a level-2 chunk boundary differ from pass to pass and the level-2 counts have to be retaken.)
    PA_TRY(sc.rp_tiles.alloc(sizeof(uint32_t) * (np + 1), st));
    PA_TRY(sc.rp_counts2.alloc(sizeof(uint32_t) * flat2, st));
    PA_TRY(sc.rp_ends2.alloc(sizeof(uint32_t) * nb, st));
    CUDA_TRY(cudaMemsetAsync(sc.rp_counts2.p, 0, sizeof(uint32_t) * flat2, st));
    if (replan) {
      k_rp_cprefix<<<1, 1024, 0, st>>>(sc.rp_ends1.as<unsigned int>(), np, sc.rp_tiles.as<unsigned int>());
      CUDA_TRY(cudaGetLastError());
    }
    RpArgs a2{};
    a2.keys = sc.p_keys.as<uint64_t>(); a2.vals = b.vals; a2.rows = sc.p_rows.as<uint32_t>(); a2.n = n;
    a2.shift = 64 - bits; a2.log_fan = b2; a2.n_parents = np;
    a2.parent_end = sc.rp_ends1.as<unsigned int>(); a2.cprefix = sc.rp_tiles.as<unsigned int>();
    a2.counts = sc.rp_counts2.as<unsigned int>();
    k_rp_hist<false><<<g->num_sms * 2, RH_THREADS, 0, st>>>(a2);
    CUDA_TRY(cudaGetLastError());
    PA_TRY(scan_u32_on(g, sc.rp_counts2.as<uint32_t>(), flat2, &sc.tile_sums));
    g->last_launches += 5;
    PA_TRY(sc.q_keys.alloc((static_cast<size_t>(n) + pad) * 8, st));
    if (vals) PA_TRY(sc.q_vals.alloc((static_cast<size_t>(n) + pad) * 8, st));
    PA_TRY(sc.q_rows.alloc((static_cast<size_t>(n) + pad) * 4, st));
    a2.offsets = sc.rp_counts2.as<unsigned int>();
    a2.out_keys = sc.q_keys.as<uint64_t>(); a2.out_vals = vals ? sc.q_vals.as<uint64_t>() : nullptr; a2.out_rows = sc.q_rows.as<uint32_t>();
    k_rp_ends<<<(nb + 255) / 256, 256, 0, st>>>(a2, sc.rp_ends2.as<unsigned int>());
    CUDA_TRY(cudaGetLastError());
    g->last_launches += 1;
    PA_TRY(launch_rp_scatter(g, a2));
    g->last_launches += 1;
    b.keys = sc.q_keys.as<uint64_t>(); b.vals = vals ? sc.q_vals.as<uint64_t>() : nullptr; b.rows = sc.q_rows.as<uint32_t>();
    b.bucket_end = sc.rp_ends2.as<unsigned int>();
  }
  g->bk_bits = bits; g->bk_b1 = b1;
  // ---- bucket aggregation ----
  // Few buckets (they would leave SMs idle, and a bucket is as long as its keys are frequent): ranged mode — equal
  // row shares per CTA, tables flushed into the global table, which the caller then compacts and orders as usual.
  const bool ranged = nb < static_cast<uint32_t>(4 * g->num_sms);
  BkArgs& bref = b;
  bref.ranged = ranged ? 1 : 0;
  bref.n = n;
  if (ranged) {
    uint64_t cap = *cap_io;
    while (cap < plan * 2) cap <<= 1;
    *cap_io = cap;
    const uint64_t nslots = cap + 2;
    PA_TRY(sc.table.alloc(nslots * sizeof(SlotT), st));
    const int init_grid = static_cast<int>(std::min<uint64_t>((nslots + 255) / 256, static_cast<uint64_t>(g->num_sms) * 16));
    k_gtable_init<WIDE><<<init_grid, 256, 0, st>>>(sc.table.as<SlotT>(), nslots);
    CUDA_TRY(cudaGetLastError());
    bref.table = sc.table.p; bref.cap_mask = cap - 1; bref.shift = 64 - __builtin_ctzll(cap);
    bref.n_buckets = nb; bref.part_bits = bits; bref.agg_mask = mask; bref.max_keys = T::MAX_KEYS;
    bref.status = g->status.as<uint32_t>();
    auto kern = bref.nullable ? k_bucket_agg<VC, WIDE, true> : k_bucket_agg<VC, WIDE, false>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(T::TOTAL)));
    kern<<<g->num_sms, BK_THREADS, T::TOTAL, st>>>(bref);
    CUDA_TRY(cudaGetLastError());
    g->last_launches += 2;
    uint32_t h_status[ST_WORDS];
    CUDA_TRY(cudaMemcpyAsync(h_status, g->status.p, sizeof h_status, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (h_status[ST_OVERFLOW]) { *declined = true; return PA_OK; }
    g->last_mode = 5;
    g->last_rlog = bits;
    *table_filled = true;
    return PA_OK;
  }
  uint64_t u_cap64 = std::min<uint64_t>(static_cast<uint64_t>(n) + 1, plan + plan / 4 + 65536);
  u_cap64 = std::min<uint64_t>(u_cap64, static_cast<uint64_t>(nb) * (T::MAX_KEYS + 1));
  const uint32_t u_cap = static_cast<uint32_t>(std::min<uint64_t>(u_cap64, 0xFFFFFFF0ull));
  using Rec = typename BkRecOf<WIDE>::type;
  PA_TRY(sc.u_rec.alloc(static_cast<size_t>(u_cap) * sizeof(Rec), st));
  PA_TRY(sc.u_first.alloc(static_cast<size_t>(u_cap) * 4, st));
  PA_TRY(sc.rp_next.alloc(4, st));
  CUDA_TRY(cudaMemsetAsync(sc.rp_next.p, 0, 4, st));
  b.n_buckets = nb; b.part_bits = bits; b.next_bucket = sc.rp_next.as<unsigned int>();
  b.agg_mask = mask; b.max_keys = T::MAX_KEYS;
  b.u_rec = sc.u_rec.p; b.u_first = sc.u_first.as<uint32_t>();
  b.u_cap = u_cap; b.status = g->status.as<uint32_t>();
  {
    auto kern = b.nullable ? k_bucket_agg<VC, WIDE, true> : k_bucket_agg<VC, WIDE, false>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(T::TOTAL)));
    kern<<<g->num_sms, BK_THREADS, T::TOTAL, st>>>(b);
    CUDA_TRY(cudaGetLastError());
  }
  g->last_launches += 1;
  CUDA_TRY(cudaEventRecord(g->ev[2], st));
  uint32_t h_status[ST_WORDS];
  CUDA_TRY(cudaMemcpyAsync(h_status, g->status.p, sizeof h_status, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (h_status[ST_OVERFLOW]) { *declined = true; return PA_OK; }
  const uint32_t G = h_status[ST_COUNTER];
  if (!WIDE && g->want_unordered) {   // the sharded step exports the records as they are (see pa_groupby::want_unordered)
    g->unordered = true;
    g->unordered_recs = sc.u_rec.p;
    g->unordered_G = G;
    g->last_mode = 5;
    g->last_rlog = bits;
    CUDA_TRY(cudaEventRecord(g->ev[3], st));
    return PA_OK;
  }
  g->G = G;
  PA_TRY(alloc_result(g, G, WIDE, VC != VC_F));
  if (G > 0) {
    // ---- first-appearance order by bitmap rank, written straight into the GroupResult ----
    const uint64_t n_rows = static_cast<uint64_t>(std::max<int64_t>(n, 1));
    const uint64_t nblocks = (n_rows + 32ull * BM_BLOCK_WORDS - 1) / (32ull * BM_BLOCK_WORDS);
    const uint64_t nwords = nblocks * BM_BLOCK_WORDS;
    const uint32_t ntiles = static_cast<uint32_t>((nblocks + SC_TILE - 1) / SC_TILE);
    PA_TRY(sc.bitmap.alloc(nwords * 4, st));
    PA_TRY(sc.prefix8.alloc(nblocks * 4, st));
    PA_TRY(sc.tile_sums.alloc(static_cast<size_t>(std::max<uint32_t>(ntiles, 1)) * 4, st));
    CUDA_TRY(cudaMemsetAsync(sc.bitmap.p, 0, nwords * 4, st));
    k_bm_set<<<(G + 255) / 256, 256, 0, st>>>(b.u_first, G, sc.bitmap.as<uint32_t>());
    CUDA_TRY(cudaGetLastError());
    const int cgrid = static_cast<int>(std::min<uint64_t>((nblocks + 255) / 256, static_cast<uint64_t>(g->num_sms) * 16));
    k_bm_count8<<<cgrid, 256, 0, st>>>(sc.bitmap.as<uint32_t>(), nblocks, sc.prefix8.as<uint32_t>());
    CUDA_TRY(cudaGetLastError());
    k_scan_tiles<<<ntiles, 256, 0, st>>>(sc.prefix8.as<uint32_t>(), nblocks, sc.tile_sums.as<uint32_t>());
    CUDA_TRY(cudaGetLastError());
    k_scan_sums<<<1, 256, 0, st>>>(sc.tile_sums.as<uint32_t>(), ntiles);
    CUDA_TRY(cudaGetLastError());
    k_scan_apply<<<ntiles, 256, 0, st>>>(sc.prefix8.as<uint32_t>(), nblocks, sc.tile_sums.as<uint32_t>());
    CUDA_TRY(cudaGetLastError());
    PA_TRY(sc.s_slot.alloc(static_cast<size_t>(G) * 4, st));
    k_bm_rank<<<(G + 255) / 256, 256, 0, st>>>(b.u_first, nullptr, G, sc.bitmap.as<uint32_t>(), sc.prefix8.as<uint32_t>(), sc.s_slot.as<uint32_t>());
    CUDA_TRY(cudaGetLastError());
    k_bk_gather<WIDE><<<(G + 255) / 256, 256, 0, st>>>(sc.u_rec.p, sc.s_slot.as<uint32_t>(), G, g->res);
    CUDA_TRY(cudaGetLastError());
    g->last_launches += 1;
    g->last_launches += 6;
  }
  g->last_mode = 5;
  g->last_rlog = bits;
  CUDA_TRY(cudaEventRecord(g->ev[3], st));
  return PA_OK;
}

template <int VC, bool WIDE>
int run_global_t(pa_groupby* g, const Column* val, uint32_t mask) {
  using SlotT = typename SlotOf<WIDE>::type;
  cudaStream_t st = g->stream;
  // (an earlier pass on this handle knows the group count exactly: size for it instead of regrowing every call)
  uint64_t want = g->opt.expected_groups > 0 ? static_cast<uint64_t>(g->opt.expected_groups)
                  : (g->have_groups ? std::max<uint64_t>(g->G, 1) : std::min<uint64_t>(static_cast<uint64_t>(g->n), 1ull << 20));
  // Load factor <= 1/2, and never fewer than n / 64 slots (at most 2 M): a small group count
  // hint must not buy a small, crowded table (scattered keys: 30 000 groups in 65 536 slots measured 27 ms
  // per 1 B rows against 21.6 ms in 2 M slots — longer probe sequences, more contended sectors).
  uint64_t cap = 1024;
  const uint64_t cap_floor = std::min<uint64_t>(2ull << 20, static_cast<uint64_t>(g->n) / 64);
  while (cap < want * 2 || cap < cap_floor) cap <<= 1;
  const uint64_t cap_limit = [&] { uint64_t c = 1024; while (c < static_cast<uint64_t>(g->n) * 2) c <<= 1; return c; }();
  if (cap > cap_limit) cap = cap_limit;
  DevBuf& table = g->scr.table;
  const int grid_full = g->num_sms * 6;   // two waves of the 3 resident CTAs per SM
  bool table_filled = false;              // the bucketed path (ranged mode) already left the groups in `table`
  {
    // More groups than one shared-memory table holds (or nobody knows how many): partition into buckets and
    // aggregate those in shared memory (bucketed.cuh); it declines when its own estimate says the front table is enough
    const uint64_t known0 = g->opt.expected_groups > 0 ? static_cast<uint64_t>(g->opt.expected_groups) : (g->have_groups ? g->G : 0);
    // (a nullable VALUE column is fine below 2^31 rows: the validity bit rides in the row-number word of the partition)
    const bool plain = g->key_width == 8 && (!val || val->width == 8) && !g->key_valid && (!val || !val->valid || g->n < (1ll << 31)) &&
                       reinterpret_cast<uintptr_t>(g->key_data) % 32 == 0 && (!val || reinterpret_cast<uintptr_t>(val->data) % 32 == 0);
    if (plain && !g->opt.no_partition && g->n >= (1ll << 22) && (known0 == 0 || known0 > static_cast<uint64_t>(SmTab<VC, WIDE>::CAP))) {
      bool declined = false;
      PA_TRY((run_bucketed_t<VC, WIDE>(g, val, mask, &cap, &declined, &table_filled)));
      if (!declined && !table_filled) return PA_OK;
      if (declined) { g->last_mode = 0; g->last_rlog = 0; }
    }
  }
  if (!table_filled) CUDA_TRY(cudaEventRecord(g->ev[1], st));
  while (!table_filled) {
    const uint64_t nslots = cap + 2;
    PA_TRY(table.alloc(nslots * sizeof(SlotT), st));
    CUDA_TRY(cudaMemsetAsync(g->status.p, 0, sizeof(uint32_t) * ST_WORDS, st));
    const int init_grid = static_cast<int>(std::min<uint64_t>((nslots + 255) / 256, static_cast<uint64_t>(g->num_sms) * 16));
    k_gtable_init<WIDE><<<init_grid, 256, 0, st>>>(table.as<SlotT>(), nslots);
    CUDA_TRY(cudaGetLastError());
    GScanArgs a{};
    a.keys = g->key_data;
    a.kvalid = g->key_valid;
    a.koff = g->key_bit_off;
    a.kw = g->key_width;
    a.vals = val ? val->data : nullptr;
    a.vvalid = val ? val->valid : nullptr;
    a.voff = val ? val->bit_off : 0;
    a.vw = val ? val->width : 8;
    a.n = g->n;
    a.table = table.p;
    a.cap_mask = cap - 1;
    a.shift = 64 - __builtin_ctzll(cap);
    a.status = g->status.as<uint32_t>();
    a.agg_mask = mask;
    const int64_t ntiles = (g->n + 1023) / 1024;
    const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ntiles, grid_full)));
    const bool fast = a.kw == 8 && (!a.vals || a.vw == 8) && !a.kvalid &&
                      reinterpret_cast<uintptr_t>(a.keys) % 32 == 0 && reinterpret_cast<uintptr_t>(a.vals) % 32 == 0;
    // Very many groups (table far larger than L2): reorder the rows by table region first (partition.cuh).
    DevBuf &p_keys = g->scr.p_keys, &p_vals = g->scr.p_vals, &p_rows = g->scr.p_rows, &p_counts = g->scr.p_counts;
    const uint64_t known_g = g->opt.expected_groups > 0 ? static_cast<uint64_t>(g->opt.expected_groups) : (g->have_groups ? g->G : 0);
    if (fast && a.vals && !a.vvalid && known_g >= (2ull << 20) && g->n >= (1ll << 22) && !g->opt.no_partition) {
      const uint64_t table_bytes = nslots * sizeof(SlotT);
      int log_parts = 6;
      while (log_parts < 10 && (table_bytes >> log_parts) > (16ull << 20)) ++log_parts;
      const int parts = 1 << log_parts;
      PA_TRY(p_keys.alloc(static_cast<size_t>(g->n) * 8, st));
      PA_TRY(p_vals.alloc(static_cast<size_t>(g->n) * 8, st));
      PA_TRY(p_rows.alloc(static_cast<size_t>(g->n) * 4, st));
      PA_TRY(p_counts.alloc(sizeof(unsigned long long) * parts, st));
      CUDA_TRY(cudaMemsetAsync(p_counts.p, 0, sizeof(unsigned long long) * parts, st));
      PartArgs pa_{};
      pa_.keys = static_cast<const uint64_t*>(a.keys);
      pa_.vals = static_cast<const uint64_t*>(a.vals);
      pa_.n = g->n;
      pa_.log_parts = log_parts;
      pa_.counts = p_counts.as<unsigned long long>();
      pa_.out_keys = p_keys.as<uint64_t>();
      pa_.out_vals = p_vals.as<uint64_t>();
      pa_.out_rows = p_rows.as<uint32_t>();
      k_part_hist<<<g->num_sms * 2, PH_THREADS, 0, st>>>(pa_);
      CUDA_TRY(cudaGetLastError());
      k_part_offsets<<<1, PT_MAX_PARTS, 0, st>>>(pa_.counts, parts);
      CUDA_TRY(cudaGetLastError());
      CUDA_TRY(cudaFuncSetAttribute(k_part_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(PtSmem::TOTAL)));
      k_part_scatter<<<g->num_sms * 2, PT_THREADS, PtSmem::TOTAL, st>>>(pa_);
      CUDA_TRY(cudaGetLastError());
      g->last_launches += 3;
      a.keys = p_keys.p;
      a.vals = p_vals.p;
      a.rowids = p_rows.as<uint32_t>();
      g->last_rlog = log_parts;   // reported through pa_groupby_last_detail[1] on this path
    }
    // Shared-memory front table unless the caller (or an earlier pass on this handle) says there are
    // far more groups than it holds; it spills to the global table, so it is correct for any input.
    const uint64_t known = g->opt.expected_groups > 0 ? static_cast<uint64_t>(g->opt.expected_groups) : (g->have_groups ? g->G : 0);
    const bool front = known <= static_cast<uint64_t>(SmTab<VC, WIDE>::CAP);   // (dense keys fill every slot; hashed keys 3/4)
    a.sm_max_keys = static_cast<uint32_t>(SmTab<VC, WIDE>::MAX_KEYS);
    DevBuf& krange = g->scr.krange;
    if (front) {
      PA_TRY(krange.alloc(sizeof(KeyRange), st));
      CUDA_TRY(cudaMemsetAsync(krange.p, 0, sizeof(KeyRange), st));
      a.krange = krange.as<KeyRange>();
      k_key_range<<<SM_SAMPLE_GRID, 256, 0, st>>>(a);
      CUDA_TRY(cudaGetLastError());
      g->last_launches += 1;
      auto kern = fast ? k_smemtab_scan<VC, WIDE, true> : k_smemtab_scan<VC, WIDE, false>;
      const size_t smem = SmTab<VC, WIDE>::TOTAL;
      CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      kern<<<g->num_sms, SM_THREADS, smem, st>>>(a);
    } else if (fast) {
      k_gtable_scan<VC, WIDE, true><<<grid, 256, 0, st>>>(a);
    } else {
      k_gtable_scan<VC, WIDE, false><<<grid, 256, 0, st>>>(a);
    }
    g->last_mode = front ? 3 : (a.rowids ? 4 : 0);
    CUDA_TRY(cudaGetLastError());
    g->last_launches += 2;
    uint32_t h_status[ST_WORDS];
    CUDA_TRY(cudaMemcpyAsync(h_status, g->status.p, sizeof h_status, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (!h_status[ST_OVERFLOW]) break;
    if (cap >= cap_limit) return set_err(PA_ERR_CUDA, "global group table overflowed at capacity %llu", (unsigned long long)cap);
    cap = std::min(cap * 8, cap_limit);
  }
  CUDA_TRY(cudaEventRecord(g->ev[2], st));
  // compact occupied slots, order them by first row, gather
  const uint64_t nslots = cap + 2;
  const uint64_t max_groups = std::min<uint64_t>(nslots, static_cast<uint64_t>(g->n) + 2);
  DevBuf &c_first = g->scr.c_first, &c_slot = g->scr.c_slot, &s_slot = g->scr.s_slot;
  PA_TRY(c_first.alloc(max_groups * 4, st));
  PA_TRY(c_slot.alloc(max_groups * 4, st));
  const int cgrid = static_cast<int>(std::min<uint64_t>((nslots + 255) / 256, static_cast<uint64_t>(g->num_sms) * 16));
  k_gtable_compact<WIDE><<<cgrid, 256, 0, st>>>(table.as<SlotT>(), nslots, c_first.as<uint32_t>(), c_slot.as<uint32_t>(), g->status.as<uint32_t>());
  CUDA_TRY(cudaGetLastError());
  g->last_launches += 1;
  uint32_t G = 0;
  CUDA_TRY(cudaMemcpyAsync(&G, g->status.as<uint32_t>() + ST_COUNTER, 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  g->G = G;
  PA_TRY(alloc_result(g, G, WIDE, VC != VC_F));
  if (G > 0) {
    PA_TRY(s_slot.alloc(static_cast<size_t>(G) * 4, st));
    PA_TRY(order_by_first_row(g, c_first.as<uint32_t>(), c_slot.as<uint32_t>(), G, s_slot.as<uint32_t>()));
    k_gtable_gather<WIDE><<<(G + 255) / 256, 256, 0, st>>>(table.as<SlotT>(), cap, s_slot.as<uint32_t>(), G, g->res);
    CUDA_TRY(cudaGetLastError());
    g->last_launches += 1;
  }
  CUDA_TRY(cudaEventRecord(g->ev[3], st));
  return PA_OK;
}

int run_global(pa_groupby* g, const Column* val, uint32_t mask, bool wide) {
  const int vc = val ? val->vc : VC_I;
  if (wide) {
    if (vc == VC_F) return run_global_t<VC_F, true>(g, val, mask);
    if (vc == VC_I) return run_global_t<VC_I, true>(g, val, mask);
    return run_global_t<VC_U, true>(g, val, mask);
  }
  if (vc == VC_F) return run_global_t<VC_F, false>(g, val, mask);
  if (vc == VC_I) return run_global_t<VC_I, false>(g, val, mask);
  return run_global_t<VC_U, false>(g, val, mask);
}

// ------------------------------ emit ------------------------------
const char* sum_format(int vc) { return vc == VC_F ? "g" : (vc == VC_I ? "l" : "L"); }

int run_emit(pa_groupby* g, const Column* val, uint32_t mask) {
  cudaStream_t st = g->stream;
  g->outs.clear();
  const uint32_t G = g->G;
  const size_t words = (static_cast<size_t>(G) + 31) / 32 + 1;
  EmitArgs e{};
  e.r = g->res;
  e.G = G;
  e.G_dev = g->emit_G_dev;
  e.vc = val ? val->vc : VC_I;
  e.vw = val ? val->width : 8;
  e.vals = val ? val->data : nullptr;
  e.vvalid = val ? val->valid : nullptr;
  e.voff = val ? val->bit_off : 0;
  if (g->merged) {
    e.vc = g->last_vc; e.vw = g->last_vw;
    e.m_first_val = g->m_first_val.as<uint64_t>(); e.m_last_val = g->m_last_val.as<uint64_t>();
    e.m_first_valid = g->m_first_valid.as<uint8_t>(); e.m_last_valid = g->m_last_valid.as<uint8_t>();
  }
  auto add = [&](uint32_t bit, const std::string& fmt, int width, bool nullable, void** vptr, uint32_t** bptr) -> int {
    g->outs.emplace_back();
    AggOut& o = g->outs.back();
    o.bit = bit;
    o.format = fmt;
    o.width = width;
    o.nullable = nullable;
    PA_TRY(o.values.alloc(static_cast<size_t>(std::max<uint32_t>(G, 1)) * width, st));
    *vptr = o.values.p;
    if (nullable) {
      PA_TRY(o.valid.alloc(words * 4, st));
      *bptr = o.valid.as<uint32_t>();
    }
    return PA_OK;
  };
  g->outs.reserve(8);
  void* v;
  uint32_t* b;
  const std::string vfmt = g->merged ? g->last_vfmt : (val ? val->format : "l");
  if (mask & AGG_SUM) { PA_TRY(add(AGG_SUM, sum_format(e.vc), 8, true, &v, &b)); e.o_sum = v; e.o_sum_valid = b; }
  if (mask & AGG_MEAN) { PA_TRY(add(AGG_MEAN, "g", 8, true, &v, &b)); e.o_mean = static_cast<double*>(v); e.o_mean_valid = b; }
  if (mask & AGG_COUNT) { PA_TRY(add(AGG_COUNT, "l", 8, false, &v, &b)); e.o_count = static_cast<int64_t*>(v); }
  if (mask & AGG_MIN) { PA_TRY(add(AGG_MIN, vfmt, e.vw, true, &v, &b)); e.o_min = v; e.o_min_valid = b; }
  if (mask & AGG_MAX) { PA_TRY(add(AGG_MAX, vfmt, e.vw, true, &v, &b)); e.o_max = v; e.o_max_valid = b; }
  if (mask & AGG_FIRST) { PA_TRY(add(AGG_FIRST, vfmt, e.vw, true, &v, &b)); e.o_first = v; e.o_first_valid = b; }
  if (mask & AGG_LAST) { PA_TRY(add(AGG_LAST, vfmt, e.vw, true, &v, &b)); e.o_last = v; e.o_last_valid = b; }
  if (G > 0 && mask) {
    k_emit<<<(G + 255) / 256, 256, 0, st>>>(e);
    CUDA_TRY(cudaGetLastError());
    g->last_launches += 1;
  }
  return PA_OK;
}

// ------------------------------ key -> group id lookup, second-stage aggregates ------------------------------
// Read-only open-addressing table key -> rank (first-appearance order) built from the finished GroupResult.
struct RowLookup {
  DevBuf tkeys, tranks, special;
  RowIdArgs a{};
};

int build_row_lookup(pa_groupby* g, RowLookup* lk) {
  cudaStream_t st = g->stream;
  const uint32_t G = g->G;
  uint64_t cap = 1024;
  while (cap < static_cast<uint64_t>(G) * 2) cap <<= 1;
  PA_TRY(lk->tkeys.alloc(cap * 8, st));
  PA_TRY(lk->tranks.alloc(cap * 4, st));
  PA_TRY(lk->special.alloc(8, st));
  const int fgrid = static_cast<int>(std::min<uint64_t>((cap + 255) / 256, static_cast<uint64_t>(g->num_sms) * 16));
  k_fill_u64<<<fgrid, 256, 0, st>>>(lk->tkeys.as<unsigned long long>(), cap, kEmptyKey);
  CUDA_TRY(cudaMemsetAsync(lk->special.p, 0xFF, 8, st));
  RowIdArgs& a = lk->a;
  a.tkeys = lk->tkeys.as<unsigned long long>();
  a.tranks = lk->tranks.as<uint32_t>();
  a.cap_mask = cap - 1;
  a.shift = 64 - __builtin_ctzll(cap);
  a.special = lk->special.as<uint32_t>();
  a.gkey = g->res.key;
  a.gkind = g->res.key_kind;
  a.G = G;
  a.keys = g->key_data;
  a.kvalid = g->key_valid;
  a.koff = g->key_bit_off;
  a.kw = g->key_width;
  a.n = g->n;
  a.resample = g->resample ? 1 : 0;
  a.rs = g->rs;
  a.out = nullptr;
  if (G) {
    k_rowid_build<<<(G + 255) / 256, 256, 0, st>>>(a);
    CUDA_TRY(cudaGetLastError());
  }
  return PA_OK;
}

int verify_string_keys(pa_groupby* g, bool* collided) {
  cudaStream_t st = g->stream;
  *collided = false;
  if (g->n == 0 || g->G == 0) return PA_OK;
  RowLookup lk;
  PA_TRY(build_row_lookup(g, &lk));
  DevBuf flag;
  PA_TRY(flag.alloc(4, st));
  CUDA_TRY(cudaMemsetAsync(flag.p, 0, 4, st));
  const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((g->n + 255) / 256, static_cast<int64_t>(g->num_sms) * 16)));
  k_str_verify<<<grid, 256, 0, st>>>(g->keys[0].str, lk.a, g->res.first_row, flag.as<uint32_t>());
  CUDA_TRY(cudaGetLastError());
  uint32_t h = 0;
  CUDA_TRY(cudaMemcpyAsync(&h, flag.p, 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (h == 2u) return set_err(PA_ERR_STATE, "utf8 keys: a row's hash is missing from the group table");
  *collided = h != 0;
  return PA_OK;
}

// product / variance / stddev (stage2.cuh): the second pass over keys + values, after the ordinary pass has
// produced sum (or the double sum) and count per group.  Appends its outputs to g->outs.
int run_stage2(pa_groupby* g, const Column* val, uint32_t ext) {
  cudaStream_t st = g->stream;
  const uint32_t G = g->G;
  const int vc = val->vc;
  const size_t words = (static_cast<size_t>(G) + 31) / 32 + 1;
  const bool want_m2 = (ext & (AGG_VARIANCE | AGG_STDDEV)) != 0, want_prod = (ext & AGG_PRODUCT) != 0;
  DevBuf mean, m2, prod;
  const size_t gbytes = static_cast<size_t>(std::max<uint32_t>(G, 1)) * 8;
  PA_TRY(mean.alloc(gbytes, st));
  if (want_m2) PA_TRY(m2.alloc(gbytes, st));
  if (want_prod) PA_TRY(prod.alloc(gbytes, st));
  Stage2Emit e{};
  e.r = g->res;
  e.G = G;
  e.vc = vc;
  e.m2 = m2.as<double>();
  e.prod = prod.as<unsigned long long>();
  auto add = [&](uint32_t bit, const std::string& fmt, void** vptr, uint32_t** bptr) -> int {
    g->outs.emplace_back();
    AggOut& o = g->outs.back();
    o.bit = bit;
    o.format = fmt;
    o.width = 8;
    o.nullable = true;
    PA_TRY(o.values.alloc(gbytes, st));
    PA_TRY(o.valid.alloc(words * 4, st));
    *vptr = o.values.p;
    *bptr = o.valid.as<uint32_t>();
    return PA_OK;
  };
  void* v;
  uint32_t* b;
  if (ext & AGG_PRODUCT) { PA_TRY(add(AGG_PRODUCT, sum_format(vc), &v, &b)); e.o_prod = static_cast<unsigned long long*>(v); e.o_prod_valid = b; }
  if (ext & AGG_VARIANCE) { PA_TRY(add(AGG_VARIANCE, "g", &v, &b)); e.o_var = static_cast<double*>(v); e.o_var_valid = b; }
  if (ext & AGG_STDDEV) { PA_TRY(add(AGG_STDDEV, "g", &v, &b)); e.o_std = static_cast<double*>(v); e.o_std_valid = b; }
  if (G == 0) return PA_OK;
  RowLookup lk;
  PA_TRY(build_row_lookup(g, &lk));
  k_stage2_init<<<(G + 255) / 256, 256, 0, st>>>(g->res, G, vc, mean.as<double>(), m2.as<double>(), prod.as<unsigned long long>());
  CUDA_TRY(cudaGetLastError());
  Stage2Args a{};
  a.ids = lk.a;
  a.vals = val->data;
  a.vvalid = val->valid;
  a.voff = val->bit_off;
  a.vw = val->width;
  a.mean = mean.as<double>();
  a.m2 = m2.as<double>();
  a.prod = prod.as<unsigned long long>();
  a.use_smem = G <= S2_SMEM_SLOTS ? 1u : 0u;
  a.rlog = 0;
  // (replicas only while they stay small: the key -> id lookups want the L1 that shared memory takes away —
  //  1000 groups measured 5.7 ms per 1 B rows unreplicated against 7.3 ms with 4 replicas)
  while (a.use_smem && a.rlog < 5 && (static_cast<uint64_t>(G) << (a.rlog + 1)) <= S2_REPL_SLOTS) ++a.rlog;
  if (g->n > 0) {
    const size_t smem = a.use_smem ? (static_cast<size_t>(G) << a.rlog) * 16 : 0;
    auto kern = vc == VC_F ? k_stage2<VC_F> : (vc == VC_I ? k_stage2<VC_I> : k_stage2<VC_U>);
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(S2_SMEM_SLOTS * 16)));
    int per_sm = 1;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, S2_THREADS, smem));
    const int64_t want = (g->n + S2_THREADS - 1) / S2_THREADS;
    const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(want, static_cast<int64_t>(g->num_sms) * std::max(per_sm, 1))));
    kern<<<grid, S2_THREADS, smem, st>>>(a);
    CUDA_TRY(cudaGetLastError());
  }
  k_stage2_emit<<<(G + 255) / 256, 256, 0, st>>>(e);
  CUDA_TRY(cudaGetLastError());
  g->last_launches += 5;
  CUDA_TRY(cudaEventRecord(g->ev[4], st));
  // (mean / m2 / prod and the lookup table are released in stream order when this scope ends)
  return PA_OK;
}

// all / any of a boolean column (GROUPBY_NUMERIC_AGG(all|any, bool), dataframe.cpp:1522-1524): bit-packed Arrow
// booleans from the min / max of the unpacked 0 / 1 bytes.  Appends to g->outs.
int run_bool_emit(pa_groupby* g, uint32_t extb) {
  cudaStream_t st = g->stream;
  const uint32_t G = g->G;
  const size_t words = (static_cast<size_t>(G) + 31) / 32 + 1;
  uint32_t* ptr[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
  int k = 0;
  for (uint32_t bit : {PA_AGG_BOOL_ALL, PA_AGG_BOOL_ANY}) {
    if (extb & bit) {
      g->outs.emplace_back();
      AggOut& o = g->outs.back();
      o.bit = bit;
      o.format = "b";
      o.width = 1;
      o.nullable = true;
      PA_TRY(o.values.alloc(words * 4, st));
      PA_TRY(o.valid.alloc(words * 4, st));
      ptr[k][0] = o.values.as<uint32_t>();
      ptr[k][1] = o.valid.as<uint32_t>();
    }
    ++k;
  }
  if (G > 0) {
    k_emit_bool<<<(G + 255) / 256, 256, 0, st>>>(g->res, G, ptr[0][0], ptr[0][1], ptr[1][0], ptr[1][1]);
    CUDA_TRY(cudaGetLastError());
    g->last_launches += 1;
    CUDA_TRY(cudaEventRecord(g->ev[4], st));
  }
  return PA_OK;
}

// count_distinct (distinct.cuh): ids + value bits -> two stable radix sorts -> boundary count.  Appends to g->outs.
int run_count_distinct(pa_groupby* g, const Column* val) {
  cudaStream_t st = g->stream;
  const uint32_t G = g->G;
  const int64_t n = g->n;
  if (n >= (1ll << 31)) return set_err(PA_ERR_INVALID, "count_distinct sorts with 32-bit positions: at most 2^31 - 1 rows");
  g->outs.emplace_back();
  AggOut& o = g->outs.back();
  o.bit = PA_AGG_COUNT_DISTINCT;
  o.format = "l";
  o.width = 8;
  o.nullable = false;
  PA_TRY(o.values.alloc(static_cast<size_t>(std::max<uint32_t>(G, 1)) * 8, st));
  if (G == 0) return PA_OK;
  DevBuf distinct, id_a, id_b, bits_a, bits_b, tmp;
  PA_TRY(distinct.alloc(static_cast<size_t>(G) * 8, st));
  CUDA_TRY(cudaMemsetAsync(distinct.p, 0, static_cast<size_t>(G) * 8, st));
  if (n > 0) {
    RowLookup lk;
    PA_TRY(build_row_lookup(g, &lk));
    const size_t nn = static_cast<size_t>(n);
    PA_TRY(id_a.alloc(nn * 4, st));
    PA_TRY(id_b.alloc(nn * 4, st));
    PA_TRY(bits_a.alloc(nn * 8, st));
    PA_TRY(bits_b.alloc(nn * 8, st));
    DistinctFillArgs a{};
    a.ids = lk.a;
    a.vals = val->data;
    a.vvalid = val->valid;
    a.voff = val->bit_off;
    a.vw = val->width;
    a.out_bits = bits_a.as<uint64_t>();
    a.out_id = id_a.as<uint32_t>();
    const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, static_cast<int64_t>(g->num_sms) * 16)));
    if (val->vc == VC_F) k_distinct_fill<VC_F><<<grid, 256, 0, st>>>(a);
    else if (val->vc == VC_I) k_distinct_fill<VC_I><<<grid, 256, 0, st>>>(a);
    else k_distinct_fill<VC_U><<<grid, 256, 0, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    // by value (64 bits; narrower columns need fewer), then stably by id (32 bits: kNoGroup must sort last)
    const int vbits = val->vc == VC_F ? 64 : (val->vc == VC_I ? 64 : std::min(64, val->width * 8));
    size_t t1 = 0, t2 = 0;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, t1, bits_a.as<uint64_t>(), bits_b.as<uint64_t>(), id_a.as<uint32_t>(), id_b.as<uint32_t>(),
                                             static_cast<int>(n), 0, vbits, st));
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, t2, id_b.as<uint32_t>(), id_a.as<uint32_t>(), bits_b.as<uint64_t>(), bits_a.as<uint64_t>(),
                                             static_cast<int>(n), 0, 32, st));
    PA_TRY(tmp.alloc(std::max(t1, t2), st));
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp.p, t1, bits_a.as<uint64_t>(), bits_b.as<uint64_t>(), id_a.as<uint32_t>(), id_b.as<uint32_t>(),
                                             static_cast<int>(n), 0, vbits, st));
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp.p, t2, id_b.as<uint32_t>(), id_a.as<uint32_t>(), bits_b.as<uint64_t>(), bits_a.as<uint64_t>(),
                                             static_cast<int>(n), 0, 32, st));
    const int64_t threads = (n + CD_PER_THREAD - 1) / CD_PER_THREAD;
    k_distinct_count<<<static_cast<int>((threads + 255) / 256), 256, 0, st>>>(id_a.as<uint32_t>(), bits_a.as<uint64_t>(), n, distinct.as<unsigned long long>());
    CUDA_TRY(cudaGetLastError());
  }
  k_distinct_emit<<<(G + 255) / 256, 256, 0, st>>>(distinct.as<unsigned long long>(), G, o.values.as<int64_t>());
  CUDA_TRY(cudaGetLastError());
  g->last_launches += 5;
  CUDA_TRY(cudaEventRecord(g->ev[4], st));
  return PA_OK;
}

// ------------------------------ Arrow export ------------------------------
struct ExportPriv {
  void* bufs[3] = {nullptr, nullptr, nullptr};
  const void* ptrs[3] = {nullptr, nullptr, nullptr};
  std::string format;
};

void release_array(ArrowArray* a) {
  auto* p = static_cast<ExportPriv*>(a->private_data);
  if (p) {
    free(p->bufs[0]);
    free(p->bufs[1]);
    free(p->bufs[2]);
    delete p;
  }
  a->release = nullptr;
}
void release_schema(ArrowSchema* s) {
  delete static_cast<std::string*>(s->private_data);
  s->release = nullptr;
}

int export_host(cudaStream_t st, const std::string& format, int width, uint32_t G, const void* d_values,
                const uint32_t* d_valid, ArrowArray* out, ArrowSchema* out_schema) {
  auto* priv = new ExportPriv();
  const size_t vbytes = format == "b" ? (static_cast<size_t>(G) + 31) / 32 * 4 : static_cast<size_t>(G) * width;   // booleans are bit-packed
  priv->bufs[1] = malloc(std::max<size_t>(vbytes, 64));
  const size_t words = (static_cast<size_t>(G) + 31) / 32;
  int64_t null_count = 0;
  if (G > 0) {
    cudaError_t e = cudaMemcpyAsync(priv->bufs[1], d_values, vbytes, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && d_valid) {
      priv->bufs[0] = malloc(std::max<size_t>(words * 4, 64));
      e = cudaMemcpyAsync(priv->bufs[0], d_valid, words * 4, cudaMemcpyDeviceToHost, st);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      free(priv->bufs[0]); free(priv->bufs[1]); delete priv;
      return set_err(PA_ERR_CUDA, "result copy failed: %s", cudaGetErrorString(e));
    }
    if (d_valid) {
      const uint32_t* w = static_cast<const uint32_t*>(priv->bufs[0]);
      int64_t set = 0;
      for (size_t i = 0; i < words; ++i) set += __builtin_popcount(w[i]);
      null_count = static_cast<int64_t>(G) - set;
      if (null_count == 0) { free(priv->bufs[0]); priv->bufs[0] = nullptr; }
    }
  }
  priv->ptrs[0] = priv->bufs[0];
  priv->ptrs[1] = priv->bufs[1];
  memset(out, 0, sizeof *out);
  out->length = G;
  out->null_count = null_count;
  out->offset = 0;
  out->n_buffers = 2;
  out->buffers = priv->ptrs;
  out->release = release_array;
  out->private_data = priv;
  memset(out_schema, 0, sizeof *out_schema);
  auto* f = new std::string(format);
  out_schema->format = f->c_str();
  out_schema->name = "";
  out_schema->flags = ARROW_FLAG_NULLABLE;
  out_schema->release = release_schema;
  out_schema->private_data = f;
  return PA_OK;
}

int ensure_device(pa_groupby* g) {
  CUDA_TRY(cudaSetDevice(g->device));
  return PA_OK;
}

int bits_for(uint64_t n_values) {  // bits needed to represent values 0..n_values-1
  int b = 1;
  while (b < 64 && (1ull << b) < n_values) ++b;
  return b;
}

int setup_keys(pa_groupby* g) {
  const int nk = static_cast<int>(g->keys.size());
  g->fields.assign(nk, KeyField{});
  if (nk == 1) {
    const Column& k = g->keys[0];
    if (k.width != 4 && k.width != 8) return set_err(PA_ERR_INVALID, "key columns must be 32- or 64-bit (format '%s')", k.format.c_str());
    g->packed = false;
    g->key_data = k.data;
    g->key_valid = k.valid;
    g->key_bit_off = k.bit_off;
    g->key_width = k.width;
    g->fields[0].width = k.width;
    return PA_OK;
  }
  if (nk > kMaxKeyCols) return set_err(PA_ERR_INVALID, "at most %d key columns", kMaxKeyCols);
  int shift = 0;
  KeyPackArgs a{};
  for (int j = 0; j < nk; ++j) {
    const Column& k = g->keys[j];
    if (k.width != 4 && k.width != 8) return set_err(PA_ERR_INVALID, "key columns must be 32- or 64-bit (format '%s')", k.format.c_str());
    KeyField& f = g->fields[j];
    f.width = k.width;
    f.bits = (k.is_dict && k.dict_len > 0) ? bits_for(static_cast<uint64_t>(k.dict_len)) : k.width * 8;
    f.nullable = k.valid != nullptr;
    f.shift = shift;
    shift += f.bits + f.nullable;
    a.col[j] = k.data; a.valid[j] = k.valid; a.off[j] = k.bit_off; a.width[j] = k.width;
    a.bits[j] = f.bits; a.shift[j] = f.shift; a.nullable[j] = f.nullable;
  }
  if (shift > 64) return set_err(PA_ERR_NOT_IMPLEMENTED, "composite key needs %d bits; only keys that pack into 64 bits are supported", shift);
  PA_TRY(g->packed_keys.alloc(static_cast<size_t>(g->n) * 8, g->stream));
  a.n_cols = nk;
  a.n = g->n;
  a.out = g->packed_keys.as<uint64_t>();
  const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((g->n + 255) / 256, g->num_sms * 16)));
  k_pack_keys<<<grid, 256, 0, g->stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  g->packed = true;
  g->key_data = g->packed_keys.p;
  g->key_valid = nullptr;
  g->key_bit_off = 0;
  g->key_width = 8;
  return PA_OK;
}

int verify_string_keys(pa_groupby* g, bool* collided);

int aggregate_impl_raw(pa_groupby* g, const Column* val, uint32_t mask, bool deferred, int start_at);

// utf8 keys are grouped by a seeded 64-bit hash of their bytes; the first pass on a handle checks the grouping byte
// for byte against each group's representative and, should two different strings ever share a hash, re-hashes the
// column with another seed and repeats (strkeys.cuh).
int aggregate_impl(pa_groupby* g, const Column* val, uint32_t mask, bool deferred = false, int start_at = 0) {
  const bool str = !g->resample && !g->merged && g->keys.size() == 1 && g->keys[0].is_str;
  if (!str || g->str_verified) return aggregate_impl_raw(g, val, mask, deferred, start_at);
  for (int attempt = 0; attempt < 4; ++attempt) {
    PA_TRY(aggregate_impl_raw(g, val, mask, false, attempt == 0 ? start_at : 0));
    bool collided = false;
    PA_TRY(verify_string_keys(g, &collided));
    if (!collided) { g->str_verified = true; return PA_OK; }
    g->keys[0].str_seed += 1;
    PA_TRY(hash_string_key(&g->keys[0], g->stream, g->num_sms));
    g->key_data = g->keys[0].data;
    g->have_groups = false;
    g->bk_have_hist = false;
    g->G = 0;
    g->outs.clear();
  }
  return set_err(PA_ERR_STATE, "utf8 keys: four differently seeded 64-bit hashes all collided");
}

int aggregate_impl_raw(pa_groupby* g, const Column* val, uint32_t mask, bool deferred, int start_at) {
  cudaStream_t st = g->stream;
  g->last_launches = 0;
  g->last_mode = 0; g->last_rlog = 0; g->last_passes = 0;
  g->unordered = false;
  const int vc = val ? val->vc : VC_I;
  const bool wide = is_wide(mask, vc);
  CUDA_TRY(cudaEventRecord(g->ev[0], st));
  if (g->resample) {
    CUDA_TRY(cudaEventRecord(g->ev[1], st));
    PA_TRY(run_resample(g, val, mask, wide));
    g->last_path = 3;
  } else {
    const int gmax = lc_gmax(vc, lc_is_wide(mask, vc));
    // (an earlier pass on this handle may already have told us how many groups there are)
    const int64_t known_groups = g->opt.expected_groups > 0 ? g->opt.expected_groups : (g->have_groups ? static_cast<int64_t>(g->G) : 0);
    bool try_low = g->opt.path != PA_PATH_GLOBAL && (known_groups <= gmax + 2 || g->opt.path == PA_PATH_LOWCARD);
    // the per-warp count words of the shared-memory path hold 24 bits next to the claim tag (lowcard.cuh): with
    // 148 SMs x >= 12 warps a warp sees at most 2^32 / 1776 = 2.4 M rows; a small MIG slice must not overflow them
    if (g->n / (static_cast<int64_t>(g->num_sms) * 12) >= (1ll << 24) - 1) try_low = false;
    bool done = false;
    if (start_at == 2) try_low = false;                       // (resuming after a deferred pass overflowed)
    if (try_low && deferred && start_at == 0 && !g->opt.lowcard_no_dense) {
      // Deferred: queue the dense-mode pass and the emit without reading the status back.  If the pass turns
      // out to have declined / aborted / overflowed, finish_pending() resumes synchronously from there.
      LcOutcome oc = LC_DONE;
      PA_TRY(run_lowcard(g, val, mask, wide, false, &oc, true));
      g->last_path = PA_PATH_LOWCARD;
      g->have_groups = true;
      g->last_wide = wide;
      g->last_vc = vc;
      g->last_vw = val ? val->width : 8;
      g->last_vfmt = val ? val->format : "l";
      g->parts_n = 0;
      g->emit_G_dev = g->status.as<uint32_t>() + ST_NGROUPS;
      PA_TRY(run_emit(g, val, mask));
      g->emit_G_dev = nullptr;
      CUDA_TRY(cudaEventRecord(g->ev[4], st));
      g->pending = true;
      return PA_OK;
    }
    if (try_low) {
      LcOutcome oc = LC_DONE;
      PA_TRY(run_lowcard(g, val, mask, wide, g->opt.lowcard_no_dense != 0 || start_at == 1, &oc));
      if (oc == LC_DENSE_MISS) PA_TRY(run_lowcard(g, val, mask, wide, true, &oc));   // a key outside the sampled window
      if (oc == LC_DONE) { done = true; g->last_path = PA_PATH_LOWCARD; }
      else if (g->opt.path == PA_PATH_LOWCARD) return set_err(PA_ERR_INVALID, "more groups than the shared-memory path holds (%d)", gmax);
    }
    if (!done) {
      g->last_mode = 0; g->last_rlog = 0; g->last_passes += 1;
      PA_TRY(run_global(g, val, mask, wide));
      g->last_path = PA_PATH_GLOBAL;
    }
  }
  if (g->unordered) {   // sharded step: the records stay where the bucket aggregation left them; GroupResult, outputs
    CUDA_TRY(cudaEventRecord(g->ev[4], st));   // and group count of the handle are as before this pass
    return PA_OK;
  }
  g->have_groups = true;
  g->last_wide = wide;
  g->last_vc = vc;
  g->last_vw = val ? val->width : 8;
  g->last_vfmt = val ? val->format : "l";
  g->parts_n = 0;
  PA_TRY(run_emit(g, val, mask));
  CUDA_TRY(cudaEventRecord(g->ev[4], st));
  return PA_OK;
}

// Completes a deferred aggregate: reads the status words; on success only the group count was missing, otherwise
// the pass is redone synchronously from the stage that failed.
int merge_padded_general(pa_groupby* g, const void* dev_blocks, int32_t n_sources, int64_t block_records, uint32_t agg_mask,
                         const char* value_format, const char* key_format);

int finish_pending(pa_groupby* g) {
  if (g->pending_merge) {
    g->pending_merge = false;
    cudaStream_t st = g->stream;
    CUDA_TRY(cudaSetDevice(g->device));
    uint32_t h_status[ST_WORDS];
    CUDA_TRY(cudaMemcpyAsync(h_status, g->status.p, sizeof h_status, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (h_status[ST_PEER_OVERFLOW]) {
      g->have_groups = false; g->G = 0; g->outs.clear(); g->poolable = false;
      return set_err(PA_ERR_STATE, "a source rank had more groups than a padded block holds; use the counted exchange");
    }
    if (!h_status[ST_OVERFLOW]) {
      g->G = h_status[ST_COUNTER];
      return PA_OK;
    }
    g->outs.clear();                                    // too many records / groups for one CTA: general merge
    g->have_groups = false;
    g->poolable = false;
    return merge_padded_general(g, g->pm_blocks, g->pm_nsrc, g->pm_block_records, g->pm_mask, g->pm_vfmt.c_str(), g->pm_kfmt.c_str());
  }
  if (!g->pending) return PA_OK;
  g->pending = false;
  cudaStream_t st = g->stream;
  CUDA_TRY(cudaSetDevice(g->device));
  uint32_t h_status[ST_WORDS];
  CUDA_TRY(cudaMemcpyAsync(h_status, g->status.p, sizeof h_status, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  const Column* val = g->pending_has_val ? &g->pending_val : nullptr;
  int rc = PA_OK;
  if (!h_status[ST_OVERFLOW] && !h_status[ST_DENSE_MISS]) {
    g->G = h_status[ST_NGROUPS];
    g->last_mode = static_cast<int>(h_status[ST_MODE]);
    g->last_rlog = static_cast<int>(h_status[ST_RLOG]);
    g->last_passes = 1;
  } else {
    const int passes = h_status[ST_DENSE_MISS] == 2u ? 0 : 1;
    g->have_groups = g->pending_had_groups;
    g->G = g->pending_prev_G;
    rc = aggregate_impl(g, val, g->pending_mask, false, h_status[ST_OVERFLOW] ? 2 : 1);
    g->last_passes += passes;
  }
  if (rc == PA_OK && g->pending_had_groups && g->pending_prev_G != g->G)
    rc = set_err(PA_ERR_STATE, "group count changed between passes (%u vs %u)", g->pending_prev_G, g->G);
  return rc;
}

}  // namespace

namespace {

template <int VC, bool WIDE>
int run_resample_t(pa_groupby* g, const Column* val, uint32_t mask) {
  using SlotT = typename SlotOf<WIDE>::type;
  cudaStream_t st = g->stream;
  const int64_t nbins = g->rs.nbins;
  const uint64_t nslots = static_cast<uint64_t>(nbins) + 2;
  DevBuf &table = g->scr.table, &bnd = g->scr.bnd;
  PA_TRY(table.alloc(nslots * sizeof(SlotT), st));
  CUDA_TRY(cudaMemsetAsync(g->status.p, 0, sizeof(uint32_t) * ST_WORDS, st));
  const int init_grid = static_cast<int>(std::min<uint64_t>((nslots + 255) / 256, static_cast<uint64_t>(g->num_sms) * 16));
  k_gtable_init<WIDE><<<init_grid, 256, 0, st>>>(table.as<SlotT>(), nslots);
  CUDA_TRY(cudaGetLastError());
  const int64_t chunk = rs_chunk_rows(g->n);
  const int64_t nchunks = (g->n + chunk - 1) / chunk;
  PA_TRY(bnd.alloc(static_cast<size_t>(std::max<int64_t>(nchunks, 1)) * 2 * sizeof(RsPartial), st));
  if (nchunks > 0) {
    k_rs_init_bnd<<<static_cast<int>((nchunks * 2 + 255) / 256), 256, 0, st>>>(bnd.as<RsPartial>(), nchunks * 2);
    CUDA_TRY(cudaGetLastError());
  }
  RsArgs a{};
  a.ts = static_cast<const int64_t*>(g->key_data);
  a.vals = val ? val->data : nullptr;
  a.vvalid = val ? val->valid : nullptr;
  a.voff = val ? val->bit_off : 0;
  a.vw = val ? val->width : 8;
  a.n = g->n;
  a.spec = g->rs;
  a.table = table.p;
  a.bnd = bnd.as<RsPartial>();
  a.nchunks = nchunks;
  a.chunk = chunk;
  a.status = g->status.as<uint32_t>();
  a.agg_mask = mask;
  if (nchunks > 0) {
    const int64_t warps_per_block = RS_THREADS / 32;
    int per_sm = 1;   // one full wave of resident CTAs: the kernel is a grid-stride loop over chunks
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_resample_scan<VC, WIDE>, RS_THREADS, 0));
    const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((nchunks + warps_per_block - 1) / warps_per_block,
                                                                           static_cast<int64_t>(g->num_sms) * std::max(per_sm, 1))));
    k_resample_scan<VC, WIDE><<<grid, RS_THREADS, 0, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(g->ev[2], st));
    k_resample_fixup<VC, WIDE><<<static_cast<int>((nchunks * 2 + 255) / 256), 256, 0, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    g->last_launches += 4;
  } else {
    CUDA_TRY(cudaEventRecord(g->ev[2], st));
  }
  // occupied buckets -> first-appearance (= time) order -> GroupResult
  const uint64_t max_groups = std::min<uint64_t>(nslots, static_cast<uint64_t>(g->n) + 2);
  DevBuf &c_first = g->scr.c_first, &c_slot = g->scr.c_slot, &s_slot = g->scr.s_slot;
  PA_TRY(c_first.alloc(max_groups * 4, st));
  PA_TRY(c_slot.alloc(max_groups * 4, st));
  const int cgrid = static_cast<int>(std::min<uint64_t>((nslots + 255) / 256, static_cast<uint64_t>(g->num_sms) * 16));
  k_gtable_compact<WIDE><<<cgrid, 256, 0, st>>>(table.as<SlotT>(), nslots, c_first.as<uint32_t>(), c_slot.as<uint32_t>(), g->status.as<uint32_t>());
  CUDA_TRY(cudaGetLastError());
  g->last_launches += 1;
  uint32_t h_status[ST_WORDS];
  CUDA_TRY(cudaMemcpyAsync(h_status, g->status.p, sizeof h_status, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (h_status[ST_UNSORTED]) return set_err(PA_ERR_INVALID, "resample: the index must be sorted ascending (Values falls before first bin / after last bin)");
  const uint32_t G = h_status[ST_COUNTER];
  g->G = G;
  PA_TRY(alloc_result(g, G, WIDE, VC != VC_F));
  if (G > 0) {
    PA_TRY(s_slot.alloc(static_cast<size_t>(G) * 4, st));
    PA_TRY(order_by_first_row(g, c_first.as<uint32_t>(), c_slot.as<uint32_t>(), G, s_slot.as<uint32_t>()));
    k_gtable_gather<WIDE><<<(G + 255) / 256, 256, 0, st>>>(table.as<SlotT>(), static_cast<uint64_t>(nbins), s_slot.as<uint32_t>(), G, g->res);
    CUDA_TRY(cudaGetLastError());
    g->last_launches += 1;
  }
  CUDA_TRY(cudaEventRecord(g->ev[3], st));
  return PA_OK;
}

int run_resample(pa_groupby* g, const Column* val, uint32_t mask, bool wide) {
  const int vc = val ? val->vc : VC_I;
  if (val && val->vc == VC_F && val->width != 4 && val->width != 8) return set_err(PA_ERR_INVALID, "bad float width");
  if (wide) {
    if (vc == VC_F) return run_resample_t<VC_F, true>(g, val, mask);
    if (vc == VC_I) return run_resample_t<VC_I, true>(g, val, mask);
    return run_resample_t<VC_U, true>(g, val, mask);
  }
  if (vc == VC_F) return run_resample_t<VC_F, false>(g, val, mask);
  if (vc == VC_I) return run_resample_t<VC_I, false>(g, val, mask);
  return run_resample_t<VC_U, false>(g, val, mask);
}

int64_t floor_div(int64_t a, int64_t b) {
  int64_t q = a / b;
  if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
  return q;
}

// adjustDatesAnchored (/root/reference/src/resample.cpp:85-178) on integer ticks; tz is not supported.
void anchor_range(int64_t start, int64_t end, int64_t freq, bool closed_right, int origin, int64_t origin_custom,
                  int64_t offset, int64_t ticks_per_day, int64_t* first, int64_t* last) {
  int64_t f = start, l = end, o = 0;
  switch (origin) {
    case 1: o = f; break;                                            // Start
    case 2: o = floor_div(f, ticks_per_day) * ticks_per_day; break;  // StartDay (default)
    case 3: o = l; break;                                            // End
    case 4: o = floor_div(l, ticks_per_day) * ticks_per_day; break;  // EndDay
    case 5: o = origin_custom; break;                                // Custom
    default: o = 0;                                                  // Epoch
  }
  o += offset;
  const int64_t fo = (f - o) % freq, lo = (l - o) % freq;            // truncating %, like the reference
  if (closed_right) {
    f -= (fo > 0) ? fo : freq;
    if (lo > 0) l += freq - lo;
  } else {
    if (fo > 0) f -= fo;
    l += (lo > 0) ? freq - lo : freq;
  }
  *first = f;
  *last = l;
}

}  // namespace

extern "C" {

const char* pa_last_error(void) { return g_err.c_str(); }
int pa_version(void) { return 100; }

void pa_options_init(pa_options* opt) {
  memset(opt, 0, sizeof *opt);
  opt->device = -1;
  opt->path = PA_PATH_AUTO;
}

int pa_device_count(int* out) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { *out = 0; return set_err(PA_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
  *out = n;
  return PA_OK;
}

// Merged (multi-GPU) handles are created and destroyed once per step; their events, status buffer and
// result buffers are recycled through a small pool instead of being re-created (≈ 0.2 ms per step).
static std::mutex g_pool_mutex;
static std::vector<pa_groupby*> g_merge_pool;

static pa_groupby* pool_take(int device, cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  for (size_t i = 0; i < g_merge_pool.size(); ++i) {
    pa_groupby* h = g_merge_pool[i];
    if (h->device == device && h->stream == stream) {
      g_merge_pool.erase(g_merge_pool.begin() + i);
      return h;
    }
  }
  return nullptr;
}

static int handle_init(pa_groupby* g, const pa_options* opt) {
  if (opt) g->opt = *opt; else pa_options_init(&g->opt);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return set_err(PA_ERR_CUDA, "no CUDA device: pandasarrow_b200 has no CPU fallback");
  if (g->opt.device < 0) CUDA_TRY(cudaGetDevice(&g->device)); else g->device = g->opt.device;
  CUDA_TRY(cudaSetDevice(g->device));
  CUDA_TRY(cudaDeviceGetAttribute(&g->num_sms, cudaDevAttrMultiProcessorCount, g->device));
  if (g->opt.cuda_stream) { g->stream = static_cast<cudaStream_t>(g->opt.cuda_stream); g->own_stream = false; }
  else { CUDA_TRY(cudaStreamCreate(&g->stream)); g->own_stream = true; }
  cudaMemPool_t pool;
  CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, g->device));
  uint64_t keep = UINT64_MAX;
  CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  for (auto& e : g->ev) CUDA_TRY(cudaEventCreate(&e));
  PA_TRY(g->status.alloc(sizeof(uint32_t) * ST_WORDS, g->stream));
  return PA_OK;
}

struct HandleDeleter {
  void operator()(pa_groupby* g) const { pa_groupby_destroy(g); }
};
using HandlePtr = std::unique_ptr<pa_groupby, HandleDeleter>;

int pa_groupby_create(const struct ArrowDeviceArray* keys, const struct ArrowSchema* key_schemas, int32_t n_keys,
                      const pa_options* opt, pa_groupby** out) {
  if (!keys || !key_schemas || n_keys < 1 || !out) return set_err(PA_ERR_INVALID, "pa_groupby_create: null argument");
  HandlePtr g(new pa_groupby());
  PA_TRY(handle_init(g.get(), opt));
  g->keys.resize(n_keys);
  for (int i = 0; i < n_keys; ++i) {
    const char* kf = key_schemas[i].format;
    if (kf && (kf[0] == 'u' || kf[0] == 'U') && kf[1] == 0) {
      if (n_keys != 1) return set_err(PA_ERR_NOT_IMPLEMENTED, "utf8 keys inside a composite key: dictionary-encode that column");
      PA_TRY(load_string_key(&keys[i], &key_schemas[i], g->stream, g->device, g->num_sms, &g->keys[i]));
    } else {
      PA_TRY(load_column(&keys[i], &key_schemas[i], g->stream, g->device, &g->keys[i]));
    }
    if (g->keys[i].n != g->keys[0].n) return set_err(PA_ERR_INVALID, "key columns differ in length");
  }
  g->n = g->keys[0].n;
  if (g->n >= 0xFFFFFFFEll) return set_err(PA_ERR_NOT_IMPLEMENTED, "more than 2^32-2 rows per call; shard by row range");
  PA_TRY(setup_keys(g.get()));
  *out = g.release();
  return PA_OK;
}

static int ensure_groups(pa_groupby* g) {
  PA_TRY(finish_pending(g));
  if (g->have_groups) return PA_OK;
  PA_TRY(ensure_device(g));
  return aggregate_impl(g, nullptr, 0);
}

int pa_groupby_num_groups(pa_groupby* g, int64_t* out) {
  if (!g || !out) return set_err(PA_ERR_INVALID, "null argument");
  PA_TRY(ensure_groups(g));
  *out = g->G;
  return PA_OK;
}

static int unique_strings(pa_groupby* g, struct ArrowArray* out, struct ArrowSchema* out_schema);

int pa_groupby_unique(pa_groupby* g, int32_t key_i, struct ArrowArray* out, struct ArrowSchema* out_schema) {
  if (!g || !out || !out_schema) return set_err(PA_ERR_INVALID, "null argument");
  PA_TRY(ensure_groups(g));
  PA_TRY(ensure_device(g));
  if (!g->resample && !g->merged && g->keys.size() == 1 && g->keys[0].is_str) {
    if (key_i != 0) return set_err(PA_ERR_INVALID, "key index %d out of range", key_i);
    return unique_strings(g, out, out_schema);
  }
  if (g->resample) {
    if (key_i != 0) return set_err(PA_ERR_INVALID, "key index %d out of range", key_i);
    return export_host(g->stream, g->index_format, 8, g->G, g->res.key, nullptr, out, out_schema);
  }
  if (key_i < 0 || key_i >= static_cast<int>(g->keys.size())) return set_err(PA_ERR_INVALID, "key index %d out of range", key_i);
  const KeyField& f = g->fields[key_i];
  const uint32_t G = g->G;
  DevBuf vals, valid;
  PA_TRY(vals.alloc(static_cast<size_t>(std::max<uint32_t>(G, 1)) * f.width, g->stream));
  PA_TRY(valid.alloc(((static_cast<size_t>(G) + 31) / 32 + 1) * 4, g->stream));
  if (G > 0) {
    KeyEmitArgs a{};
    a.key = g->res.key; a.key_kind = g->res.key_kind; a.G = G;
    a.width = f.width; a.bits = f.bits; a.shift = f.shift; a.nullable = f.nullable; a.packed = g->packed;
    a.out = vals.p; a.out_valid = valid.as<uint32_t>();
    k_emit_key<<<(G + 255) / 256, 256, 0, g->stream>>>(a);
    CUDA_TRY(cudaGetLastError());
  }
  return export_host(g->stream, g->keys[key_i].format, f.width, G, vals.p, valid.as<uint32_t>(), out, out_schema);
}

static int aggregate_entry(pa_groupby* g, const struct ArrowDeviceArray* values, const struct ArrowSchema* value_schema,
                           uint32_t agg_mask, bool deferred) {
  if (!g || !values || !value_schema) return set_err(PA_ERR_INVALID, "null argument");
  if (agg_mask == 0 || (agg_mask & ~(PA_AGG_ALL | PA_AGG_STAGE2 | PA_AGG_BOOL_ALL | PA_AGG_BOOL_ANY | PA_AGG_COUNT_DISTINCT))) return set_err(PA_ERR_INVALID, "bad aggregate mask 0x%x", agg_mask);
  const uint32_t requested = agg_mask;
  const uint32_t extb = agg_mask & (PA_AGG_BOOL_ALL | PA_AGG_BOOL_ANY);
  const bool want_distinct = (agg_mask & PA_AGG_COUNT_DISTINCT) != 0;
  if (want_distinct && g->merged) return set_err(PA_ERR_NOT_IMPLEMENTED, "count_distinct is not available on merged (multi-GPU) handles");
  // product / variance / stddev: a second pass after the ordinary one, which then has to deliver count (and the mean)
  const uint32_t ext = agg_mask & PA_AGG_STAGE2;
  if (ext && g->merged) return set_err(PA_ERR_NOT_IMPLEMENTED, "product / variance / stddev are not available on merged (multi-GPU) handles");
  if (extb && g->merged) return set_err(PA_ERR_NOT_IMPLEMENTED, "all / any are not available on merged (multi-GPU) handles");
  if (ext || extb || want_distinct) {
    deferred = false;
    // all = (min over the 0 / 1 bytes) != 0, any = (max) != 0: the ordinary pass delivers both
    agg_mask = (agg_mask & PA_AGG_ALL) | PA_AGG_COUNT | ((ext & (PA_AGG_VARIANCE | PA_AGG_STDDEV)) ? PA_AGG_MEAN : 0u) |
               ((extb & PA_AGG_BOOL_ALL) ? PA_AGG_MIN : 0u) | ((extb & PA_AGG_BOOL_ANY) ? PA_AGG_MAX : 0u);
  }
  PA_TRY(ensure_device(g));
  Column val;
  PA_TRY(load_column(values, value_schema, g->stream, g->device, &val));
  if (val.n != g->n) return set_err(PA_ERR_INVALID, "value column has %lld rows, keys have %lld", (long long)val.n, (long long)g->n);
  if (value_schema->dictionary) return set_err(PA_ERR_INVALID, "dictionary-encoded value columns are not aggregatable");
  if (val.is_bool && (requested & ~(PA_AGG_COUNT | PA_AGG_BOOL_ALL | PA_AGG_BOOL_ANY | PA_AGG_COUNT_DISTINCT)))
    return set_err(PA_ERR_INVALID, "boolean columns aggregate with count / all / any only");
  if (extb && !val.is_bool) return set_err(PA_ERR_INVALID, "all / any need a boolean column (arrow::compute has no kernel for other types)");
  uint32_t prev_G = g->G;
  bool had = g->have_groups;
  if (g->pending && deferred && !val.own_data.p && !val.own_valid.p) {
    // A deferred pass whose results nobody asked for is superseded by this one (its buffers are recycled in
    // stream order); had it failed, this pass fails the same way and is redone when it is finished.
    g->pending = false;
    prev_G = g->pending_prev_G;
    had = g->pending_had_groups;
    g->have_groups = had;
    g->G = prev_G;
  } else {
    PA_TRY(finish_pending(g));
    prev_G = g->G;
    had = g->have_groups;
  }
  // deferred mode needs inputs that outlive the call: device-resident columns only
  deferred = deferred && !val.own_data.p && !val.own_valid.p && !g->resample && !g->merged;
  if (deferred) {
    g->pending_val = Column{};
    g->pending_val.data = val.data; g->pending_val.valid = val.valid; g->pending_val.bit_off = val.bit_off;
    g->pending_val.n = val.n; g->pending_val.width = val.width; g->pending_val.vc = val.vc; g->pending_val.format = val.format;
    g->pending_has_val = true;
    g->pending_mask = agg_mask;
    g->pending_prev_G = prev_G;
    g->pending_had_groups = had;
  }
  PA_TRY(aggregate_impl(g, &val, agg_mask, deferred));
  if (g->pending) return PA_OK;
  if (had && prev_G != g->G) return set_err(PA_ERR_STATE, "group count changed between passes (%u vs %u)", prev_G, g->G);
  if (ext) PA_TRY(run_stage2(g, &val, ext));
  if (want_distinct) PA_TRY(run_count_distinct(g, &val));
  if (val.is_bool) {
    // the helper min / max columns are bytes, not Arrow booleans: only what was asked for stays fetchable
    g->outs.erase(std::remove_if(g->outs.begin(), g->outs.end(), [&](const AggOut& o) { return (o.bit & requested) == 0; }), g->outs.end());
    if (extb) PA_TRY(run_bool_emit(g, extb));
  }
  // `val` may own device copies of host data: make sure the kernels reading them are done
  if (val.own_data.p) CUDA_TRY(cudaStreamSynchronize(g->stream));
  return PA_OK;
}

int pa_groupby_aggregate(pa_groupby* g, const struct ArrowDeviceArray* values, const struct ArrowSchema* value_schema,
                         uint32_t agg_mask) {
  return aggregate_entry(g, values, value_schema, agg_mask, false);
}

int pa_groupby_aggregate_async(pa_groupby* g, const struct ArrowDeviceArray* values, const struct ArrowSchema* value_schema,
                               uint32_t agg_mask) {
  return aggregate_entry(g, values, value_schema, agg_mask, true);
}

// NDFrame<T>::sum/mean/min/max/count/first/last/min_max/agg (/root/reference/src/ndframe.cpp:119-241): the whole
// column as ONE group (constant key generated on the device), through the same fused pass.
int pa_column_aggregate(const struct ArrowDeviceArray* values, const struct ArrowSchema* value_schema, uint32_t agg_mask,
                        int32_t skip_nulls, const pa_options* opt, pa_groupby** out) {
  if (!values || !value_schema || !out) return set_err(PA_ERR_INVALID, "pa_column_aggregate: null argument");
  if (agg_mask == 0 || (agg_mask & ~(PA_AGG_ALL | PA_AGG_STAGE2))) return set_err(PA_ERR_INVALID, "bad aggregate mask 0x%x", agg_mask);
  HandlePtr g(new pa_groupby());
  PA_TRY(handle_init(g.get(), opt));
  cudaStream_t st = g->stream;
  const int64_t n = values->array.length;
  if (n >= 0xFFFFFFFEll) return set_err(PA_ERR_NOT_IMPLEMENTED, "more than 2^32-2 rows per call; shard by row range");
  g->keys.resize(1);
  Column& k = g->keys[0];
  k.format = "l"; k.width = 8; k.vc = VC_I; k.n = n;
  PA_TRY(k.own_data.alloc(static_cast<size_t>(std::max<int64_t>(n, 1)) * 8, st));
  CUDA_TRY(cudaMemsetAsync(k.own_data.p, 0, static_cast<size_t>(std::max<int64_t>(n, 1)) * 8, st));
  k.data = k.own_data.p;
  g->n = n;
  PA_TRY(setup_keys(g.get()));
  const uint32_t positional = agg_mask & (PA_AGG_FIRST | PA_AGG_LAST);
  const bool patch = skip_nulls && positional && values->array.null_count != 0 && values->array.buffers[0] != nullptr && n > 0;
  if (!patch) {
    PA_TRY(aggregate_entry(g.get(), values, value_schema, agg_mask, false));
  } else {
    // first / last skip nulls here (arrow's scalar kernels), unlike GroupBy::first/last which are positional
    Column val;
    PA_TRY(load_column(values, value_schema, st, g->device, &val));
    if (val.is_bool) return set_err(PA_ERR_INVALID, "boolean columns aggregate with count / all / any only");
    PA_TRY(aggregate_impl(g.get(), &val, (agg_mask & PA_AGG_ALL) | PA_AGG_FIRST | PA_AGG_LAST));
    DevBuf range;
    PA_TRY(range.alloc(8, st));
    const uint32_t init[2] = {kNoRow, 0u};
    CUDA_TRY(cudaMemcpyAsync(range.p, init, 8, cudaMemcpyHostToDevice, st));
    const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, static_cast<int64_t>(g->num_sms) * 8)));
    k_valid_row_range<<<grid, 256, 0, st>>>(val.valid, val.bit_off, n, range.as<uint32_t>());
    CUDA_TRY(cudaGetLastError());
    if (g->G > 0) {
      k_apply_row_range<<<1, 1, 0, st>>>(range.as<uint32_t>(), g->res.first_row, g->res.last_row);
      CUDA_TRY(cudaGetLastError());
    }
    PA_TRY(run_emit(g.get(), &val, (agg_mask & PA_AGG_ALL) | PA_AGG_FIRST | PA_AGG_LAST));
    if (agg_mask & PA_AGG_STAGE2) PA_TRY(run_stage2(g.get(), &val, agg_mask & PA_AGG_STAGE2));
    CUDA_TRY(cudaStreamSynchronize(st));   // `val` / `range` may own device copies
  }
  *out = g.release();
  return PA_OK;
}

int pa_groupby_fetch(pa_groupby* g, uint32_t agg_bit, struct ArrowArray* out, struct ArrowSchema* out_schema) {
  if (!g || !out || !out_schema) return set_err(PA_ERR_INVALID, "null argument");
  PA_TRY(ensure_device(g));
  PA_TRY(finish_pending(g));
  for (auto& o : g->outs) {
    if (o.bit == agg_bit)
      return export_host(g->stream, o.format, o.width, g->G, o.values.p, o.nullable ? o.valid.as<uint32_t>() : nullptr, out, out_schema);
  }
  return set_err(PA_ERR_STATE, "aggregate 0x%x was not part of the last pa_groupby_aggregate call", agg_bit);
}

int pa_groupby_row_ids(pa_groupby* g, struct ArrowArray* out, struct ArrowSchema* out_schema) {
  if (!g || !out || !out_schema) return set_err(PA_ERR_INVALID, "null argument");
  if (g->merged) return set_err(PA_ERR_STATE, "row ids exist on the rank that holds the rows, not on a merged handle");
  PA_TRY(ensure_groups(g));
  PA_TRY(ensure_device(g));
  cudaStream_t st = g->stream;
  RowLookup lk;
  PA_TRY(build_row_lookup(g, &lk));
  DevBuf ids;
  PA_TRY(ids.alloc(static_cast<size_t>(std::max<int64_t>(g->n, 1)) * 4, st));
  lk.a.out = ids.as<uint32_t>();
  if (g->n) {
    const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((g->n + 255) / 256, static_cast<int64_t>(g->num_sms) * 16)));
    k_rowid_scan<<<grid, 256, 0, st>>>(lk.a);
    CUDA_TRY(cudaGetLastError());
  }
  return export_host(st, "I", 4, static_cast<uint32_t>(g->n), ids.p, nullptr, out, out_schema);
}

namespace {
// device-wide exclusive scan of a uint32 array in place (order.cuh)
int scan_u32(pa_groupby* g, uint32_t* data, uint64_t n, DevBuf* tile_sums) { return scan_u32_on(g, data, n, tile_sums); }

// ids -> stable counting sort by id (csort.cuh) -> (order, dest, offsets), cached on the handle
int ensure_groupings(pa_groupby* g) {
  if (g->have_groupings) return PA_OK;
  if (g->merged) return set_err(PA_ERR_STATE, "groupings exist on the rank that holds the rows, not on a merged handle");
  if (g->n >= (1ll << 31)) return set_err(PA_ERR_INVALID, "groupings use int32 offsets like the reference's ListArray<int32>: at most 2^31 - 1 rows");
  cudaStream_t st = g->stream;
  const uint32_t G = g->G;
  const int64_t n = g->n;
  CUDA_TRY(cudaEventRecord(g->ev[6], st));
  RowLookup lk;
  PA_TRY(build_row_lookup(g, &lk));
  DevBuf ids, keys_a, keys_b, pay_a, pay_b, counts, tile_sums;
  const size_t nb4 = static_cast<size_t>(std::max<int64_t>(n, 1)) * 4;
  PA_TRY(ids.alloc(nb4, st));
  PA_TRY(g->grp_order.alloc(nb4, st));
  PA_TRY(g->grp_dest.alloc(nb4, st));
  PA_TRY(g->grp_offsets.alloc((static_cast<size_t>(G) + 1) * 4, st));
  if (n > 0) {
    const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, static_cast<int64_t>(g->num_sms) * 16)));
    lk.a.out = ids.as<uint32_t>();
    k_rowid_scan<<<grid, 256, 0, st>>>(lk.a);
    CUDA_TRY(cudaGetLastError());
    int bits = 1;
    while (bits < 32 && (1ull << bits) < static_cast<uint64_t>(G)) ++bits;
    const int passes = (bits + CS_BITS - 1) / CS_BITS;
    const int64_t ntiles = (n + CS_TILE - 1) / CS_TILE;
    const int nb = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ntiles, static_cast<int64_t>(g->num_sms) * 4)));
    const int64_t chunk = ((ntiles + nb - 1) / nb) * CS_TILE;
    PA_TRY(counts.alloc(sizeof(uint32_t) * CS_R * static_cast<size_t>(nb), st));
    if (passes > 1) {   // ping-pong buffers; the last pass writes its payload to grp_order and the sorted ids to keys_a / keys_b
      PA_TRY(keys_a.alloc(nb4, st));
      PA_TRY(pay_a.alloc(nb4, st));
      PA_TRY(keys_b.alloc(nb4, st));
      if (passes > 2) PA_TRY(pay_b.alloc(nb4, st));
    }
    static bool attr_set = false;
    if (!attr_set) {
      CUDA_TRY(cudaFuncSetAttribute(k_cs_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(CS_SMEM)));
      attr_set = true;
    }
    const uint32_t* in_keys = ids.as<uint32_t>();
    const uint32_t* in_pay = nullptr;
    for (int p = 0; p < passes; ++p) {
      const bool last = p == passes - 1;
      CsArgs a{};
      a.keys = in_keys; a.payload = in_pay; a.n = n; a.shift = p * CS_BITS; a.nb = nb; a.chunk = chunk;
      a.counts = counts.as<uint32_t>();
      DevBuf& ok = (p % 2 == 0) ? keys_a : keys_b;
      DevBuf& op = (p % 2 == 0) ? pay_a : pay_b;
      a.out_keys = last ? (passes > 1 ? ok.as<uint32_t>() : nullptr) : ok.as<uint32_t>();
      a.out_payload = last ? g->grp_order.as<uint32_t>() : op.as<uint32_t>();
      a.out_dest = last ? g->grp_dest.as<uint32_t>() : nullptr;
      k_cs_hist<<<nb, CS_THREADS, 0, st>>>(a);
      CUDA_TRY(cudaGetLastError());
      PA_TRY(scan_u32(g, counts.as<uint32_t>(), static_cast<uint64_t>(CS_R) * nb, &tile_sums));
      k_cs_scatter<<<nb, CS_THREADS, CS_SMEM, st>>>(a);
      CUDA_TRY(cudaGetLastError());
      if (last) {
        if (passes == 1) {
          k_cs_offsets<<<(G + 1 + 255) / 256, 256, 0, st>>>(counts.as<uint32_t>(), nb, G, n, g->grp_offsets.as<int32_t>());
        } else {
          k_group_offsets<<<grid, 256, 0, st>>>(a.out_keys, n, G, g->grp_offsets.as<int32_t>());
        }
        CUDA_TRY(cudaGetLastError());
      }
      in_keys = a.out_keys;
      in_pay = a.out_payload;
    }
  } else {
    CUDA_TRY(cudaMemsetAsync(g->grp_offsets.p, 0, (static_cast<size_t>(G) + 1) * 4, st));
  }
  CUDA_TRY(cudaEventRecord(g->ev[7], st));
  CUDA_TRY(cudaEventRecord(g->ev[8], st));
  CUDA_TRY(cudaEventRecord(g->ev[9], st));
  CUDA_TRY(cudaStreamSynchronize(st));   // the sort buffers above are released here
  g->have_groupings = true;
  return PA_OK;
}
}  // namespace

// unique() of a utf8 / large_utf8 key: the bytes of every group's first row, in group order (int32 offsets; a result
// of more than 2 GiB of distinct strings is refused).
static int unique_strings(pa_groupby* g, struct ArrowArray* out, struct ArrowSchema* out_schema) {
  cudaStream_t st = g->stream;
  const uint32_t G = g->G;
  const Column& k = g->keys[0];
  DevBuf offs, tile_sums, bytes, valid;
  PA_TRY(offs.alloc(static_cast<size_t>(G + 1) * 4, st));
  k_str_lengths<<<(G + 1 + 255) / 256, 256, 0, st>>>(k.str, g->res.first_row, g->res.key_kind, G, offs.as<uint32_t>());
  CUDA_TRY(cudaGetLastError());
  uint64_t total64 = 0;
  {
    DevBuf tot;
    PA_TRY(tot.alloc(8, st));
    CUDA_TRY(cudaMemsetAsync(tot.p, 0, 8, st));
    if (G) k_sum_u32<<<std::min<uint32_t>((G + 255) / 256, 1024u), 256, 0, st>>>(offs.as<uint32_t>(), G, tot.as<unsigned long long>());
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(&total64, tot.p, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  if (total64 > 0x7FFFFFFFull) return set_err(PA_ERR_NOT_IMPLEMENTED, "distinct utf8 keys total %llu bytes; int32 offsets hold 2 GiB", (unsigned long long)total64);
  PA_TRY(scan_u32(g, offs.as<uint32_t>(), static_cast<uint64_t>(G) + 1, &tile_sums));
  PA_TRY(bytes.alloc(std::max<uint64_t>(total64, 1), st));
  PA_TRY(valid.alloc(((static_cast<size_t>(G) + 31) / 32 + 1) * 4, st));
  if (G) {
    k_str_gather<<<(static_cast<uint64_t>(G) * 32 + 255) / 256, 256, 0, st>>>(k.str, g->res.first_row, g->res.key_kind, G, offs.as<uint32_t>(), bytes.as<uint8_t>());
    CUDA_TRY(cudaGetLastError());
    k_kind_validity<<<(G + 31) / 32 * 32 / 256 + 1, 256, 0, st>>>(g->res.key_kind, G, valid.as<uint32_t>());
    CUDA_TRY(cudaGetLastError());
  }
  auto* priv = new ExportPriv();
  const size_t words = (static_cast<size_t>(G) + 31) / 32;
  priv->bufs[0] = malloc(std::max<size_t>(words * 4, 64));
  priv->bufs[1] = malloc(std::max<size_t>((static_cast<size_t>(G) + 1) * 4, 64));
  priv->bufs[2] = malloc(std::max<size_t>(total64, 64));
  cudaError_t e = cudaMemcpyAsync(priv->bufs[1], offs.p, (static_cast<size_t>(G) + 1) * 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && G) e = cudaMemcpyAsync(priv->bufs[0], valid.p, words * 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && total64) e = cudaMemcpyAsync(priv->bufs[2], bytes.p, total64, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    free(priv->bufs[0]); free(priv->bufs[1]); free(priv->bufs[2]); delete priv;
    return set_err(PA_ERR_CUDA, "result copy failed: %s", cudaGetErrorString(e));
  }
  int64_t set = 0;
  for (size_t i = 0; i < words; ++i) set += __builtin_popcount(static_cast<const uint32_t*>(priv->bufs[0])[i]);
  const int64_t null_count = static_cast<int64_t>(G) - set;
  if (null_count == 0) { free(priv->bufs[0]); priv->bufs[0] = nullptr; }
  const bool large = k.str.wide != 0;      // large_utf8 in, large_utf8 out (int64 offsets, widened on the host: G + 1 numbers)
  if (large) {
    int64_t* wide = static_cast<int64_t*>(malloc(std::max<size_t>((static_cast<size_t>(G) + 1) * 8, 64)));
    for (size_t i = 0; i <= G; ++i) wide[i] = static_cast<const uint32_t*>(priv->bufs[1])[i];
    free(priv->bufs[1]);
    priv->bufs[1] = wide;
  }
  for (int i = 0; i < 3; ++i) priv->ptrs[i] = priv->bufs[i];
  memset(out, 0, sizeof *out);
  out->length = G;
  out->null_count = null_count;
  out->n_buffers = 3;
  out->buffers = priv->ptrs;
  out->release = release_array;
  out->private_data = priv;
  memset(out_schema, 0, sizeof *out_schema);
  auto* f = new std::string(large ? "U" : "u");
  out_schema->format = f->c_str();
  out_schema->name = "";
  out_schema->flags = ARROW_FLAG_NULLABLE;
  out_schema->release = release_schema;
  out_schema->private_data = f;
  return PA_OK;
}

int pa_groupby_groupings(pa_groupby* g, struct ArrowArray* offsets, struct ArrowSchema* offsets_schema,
                         struct ArrowArray* rows, struct ArrowSchema* rows_schema) {
  if (!g || !offsets || !offsets_schema) return set_err(PA_ERR_INVALID, "null argument");
  if ((rows == nullptr) != (rows_schema == nullptr)) return set_err(PA_ERR_INVALID, "rows and rows_schema go together");
  PA_TRY(ensure_groups(g));
  PA_TRY(ensure_device(g));
  PA_TRY(ensure_groupings(g));
  PA_TRY(export_host(g->stream, "i", 4, g->G + 1, g->grp_offsets.p, nullptr, offsets, offsets_schema));
  if (rows) PA_TRY(export_host(g->stream, "i", 4, static_cast<uint32_t>(g->n), g->grp_order.p, nullptr, rows, rows_schema));
  return PA_OK;
}

int pa_groupby_take_grouped(pa_groupby* g, const struct ArrowDeviceArray* column, const struct ArrowSchema* schema,
                            struct ArrowArray* out, struct ArrowSchema* out_schema) {
  if (!g || !column || !schema || !out || !out_schema) return set_err(PA_ERR_INVALID, "null argument");
  PA_TRY(ensure_groups(g));
  PA_TRY(ensure_device(g));
  PA_TRY(ensure_groupings(g));
  cudaStream_t st = g->stream;
  Column col;
  PA_TRY(load_column(column, schema, st, g->device, &col));
  if (col.n != g->n) return set_err(PA_ERR_INVALID, "column has %lld rows, keys have %lld", (long long)col.n, (long long)g->n);
  if (schema->dictionary) return set_err(PA_ERR_INVALID, "dictionary-encoded columns: gather the indices, keep the dictionary");
  const int64_t n = g->n;
  const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, static_cast<int64_t>(g->num_sms) * 16)));
  const size_t words = (static_cast<size_t>(n) + 31) / 32 + 1;
  DevBuf vals, valid;
  if (col.is_bool) {
    // bit-packed values: gathered bit by bit from the packed input (load_column keeps it in own_bits for host
    // inputs; device inputs are read in place)
    const ArrowArray& arr = column->array;
    const uint8_t* bits = col.own_bits.p ? col.own_bits.as<uint8_t>() : static_cast<const uint8_t*>(arr.buffers[1]);
    const int64_t boff = col.own_bits.p ? arr.offset % 8 : arr.offset;
    PA_TRY(vals.alloc(words * 4, st));
    if (col.valid) PA_TRY(valid.alloc(words * 4, st));
    if (n > 0) {
      CUDA_TRY(cudaEventRecord(g->ev[8], st));
      k_take_bits<<<grid, 256, 0, st>>>(bits, boff, g->grp_order.as<uint32_t>(), n, vals.as<uint32_t>());
      CUDA_TRY(cudaGetLastError());
      if (col.valid) {
        k_take_bits<<<grid, 256, 0, st>>>(col.valid, col.bit_off, g->grp_order.as<uint32_t>(), n, valid.as<uint32_t>());
        CUDA_TRY(cudaGetLastError());
      }
      CUDA_TRY(cudaEventRecord(g->ev[9], st));
    }
    return export_host(st, "b", 1, static_cast<uint32_t>(n), vals.p, col.valid ? valid.as<uint32_t>() : nullptr, out, out_schema);
  }
  PA_TRY(vals.alloc(static_cast<size_t>(std::max<int64_t>(n, 1)) * col.width, st));
  if (col.valid) {
    PA_TRY(valid.alloc(words * 4, st));
    CUDA_TRY(cudaMemsetAsync(valid.p, 0, words * 4, st));
  }
  if (n > 0) {
    TakeScatterArgs a{};
    a.col = col.data;
    a.valid = col.valid;
    a.bit_off = col.bit_off;
    a.width = col.width;
    a.dest = g->grp_dest.as<uint32_t>();
    a.n = n;
    a.out = vals.p;
    a.out_valid = col.valid ? valid.as<uint32_t>() : nullptr;
    CUDA_TRY(cudaEventRecord(g->ev[8], st));
    k_take_scatter<<<grid, 256, 0, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(g->ev[9], st));
  }
  return export_host(st, col.format, col.width, static_cast<uint32_t>(n), vals.p, col.valid ? valid.as<uint32_t>() : nullptr, out, out_schema);
}

int pa_groupby_groupings_timing(pa_groupby* g, double* build_ms, double* take_ms) {
  if (!g) return set_err(PA_ERR_INVALID, "null argument");
  if (!g->have_groupings) return set_err(PA_ERR_STATE, "no groupings were built on this handle");
  PA_TRY(ensure_device(g));
  CUDA_TRY(cudaEventSynchronize(g->ev[9]));
  float t = 0;
  if (build_ms) { CUDA_TRY(cudaEventElapsedTime(&t, g->ev[6], g->ev[7])); *build_ms = t; }
  if (take_ms) { CUDA_TRY(cudaEventElapsedTime(&t, g->ev[8], g->ev[9])); *take_ms = t; }
  return PA_OK;
}

int pa_groupby_first_rows(pa_groupby* g, struct ArrowArray* out, struct ArrowSchema* out_schema) {
  if (!g || !out || !out_schema) return set_err(PA_ERR_INVALID, "null argument");
  PA_TRY(ensure_groups(g));
  PA_TRY(ensure_device(g));
  if (g->merged) return export_host(g->stream, "L", 8, g->G, g->m_first_row_g.p, nullptr, out, out_schema);
  PA_TRY(export_host(g->stream, "I", 4, g->G, g->res.first_row, nullptr, out, out_schema));
  // widen to uint64 and add the shard's row base on the host (G values)
  auto* priv = static_cast<ExportPriv*>(out->private_data);
  const uint32_t* src = static_cast<const uint32_t*>(priv->bufs[1]);
  uint64_t* wide = static_cast<uint64_t*>(malloc(std::max<size_t>(static_cast<size_t>(g->G) * 8, 64)));
  for (uint32_t i = 0; i < g->G; ++i) wide[i] = static_cast<uint64_t>(g->opt.row_base) + src[i];
  free(priv->bufs[1]);
  priv->bufs[1] = wide;
  priv->ptrs[1] = wide;
  delete static_cast<std::string*>(out_schema->private_data);
  auto* f = new std::string("L");
  out_schema->format = f->c_str();
  out_schema->private_data = f;
  return PA_OK;
}

int pa_groupby_last_timing(pa_groupby* g, double* total_ms, double stage_ms[4]) {
  if (!g) return set_err(PA_ERR_INVALID, "null argument");
  PA_TRY(ensure_device(g));
  PA_TRY(finish_pending(g));
  CUDA_TRY(cudaEventSynchronize(g->ev[4]));
  float t = 0;
  CUDA_TRY(cudaEventElapsedTime(&t, g->ev[0], g->ev[4]));
  if (total_ms) *total_ms = t;
  if (stage_ms) {
    float a = 0, b = 0, c = 0, d = 0;
    CUDA_TRY(cudaEventElapsedTime(&a, g->ev[0], g->ev[1]));
    CUDA_TRY(cudaEventElapsedTime(&b, g->ev[1], g->ev[2]));
    CUDA_TRY(cudaEventElapsedTime(&c, g->ev[2], g->ev[3]));
    CUDA_TRY(cudaEventElapsedTime(&d, g->ev[3], g->ev[4]));
    stage_ms[0] = a; stage_ms[1] = b; stage_ms[2] = c; stage_ms[3] = d;
  }
  return PA_OK;
}

int pa_groupby_last_path(pa_groupby* g, int32_t* path, int32_t* kernel_launches) {
  if (!g) return set_err(PA_ERR_INVALID, "null argument");
  PA_TRY(finish_pending(g));
  if (path) *path = g->last_path;
  if (kernel_launches) *kernel_launches = g->last_launches;
  return PA_OK;
}

int pa_groupby_last_detail(pa_groupby* g, int32_t detail[4]) {
  if (!g || !detail) return set_err(PA_ERR_INVALID, "null argument");
  PA_TRY(finish_pending(g));
  detail[0] = g->last_mode; detail[1] = g->last_rlog; detail[2] = g->last_passes; detail[3] = 0;
  return PA_OK;
}

int pa_groupby_sync(pa_groupby* g) {
  if (!g) return set_err(PA_ERR_INVALID, "null argument");
  PA_TRY(ensure_device(g));
  PA_TRY(finish_pending(g));
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  return PA_OK;
}

void pa_groupby_destroy(pa_groupby* g) {
  if (!g) return;
  if (g->merged && !g->own_stream && g->poolable) {
    // recycle: nothing is freed, so no synchronisation is needed either (stream order protects the buffers)
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (g_merge_pool.size() < 4) {
      g->outs.clear();
      g->have_groups = false;
      g->pending_merge = false;
      g->G = 0;
      g_merge_pool.push_back(g);
      return;
    }
  }
  cudaSetDevice(g->device);
  cudaStreamSynchronize(g->stream);
  g->outs.clear();
  cudaStream_t st = g->stream;
  const bool own = g->own_stream;
  for (auto& e : g->ev) if (e) cudaEventDestroy(e);
  delete g;   // DevBufs free on `st`
  if (own) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
}

int pa_resample_create(const struct ArrowDeviceArray* index, const struct ArrowSchema* index_schema, int64_t freq_ns,
                       int32_t closed_right, int32_t label_right, int32_t origin, int64_t origin_custom_ns,
                       int64_t offset_ns, const pa_options* opt, pa_groupby** out) {
  if (!index || !index_schema || !out) return set_err(PA_ERR_INVALID, "pa_resample_create: null argument");
  if (freq_ns <= 0) return set_err(PA_ERR_INVALID, "FREQ must be positive");   // core.cpp:319-322
  HandlePtr g(new pa_groupby());
  PA_TRY(handle_init(g.get(), opt));
  g->keys.resize(1);
  Column& ix = g->keys[0];
  PA_TRY(load_column(index, index_schema, g->stream, g->device, &ix));
  const char* f = index_schema->format;
  const bool is_ts = f && f[0] == 't' && f[1] == 's';
  if (!(is_ts || (f && f[0] == 'l')) || ix.width != 8)
    return set_err(PA_ERR_INVALID, "axis must be a TimestampArray but got array of type '%s'", f ? f : "?");  // resample.cpp:213-216
  if (ix.valid) return set_err(PA_ERR_NOT_IMPLEMENTED, "resample: null timestamps are not supported");
  int64_t ticks_per_day = 86400LL * 1000000000LL;
  int64_t ns_per_tick = 1;
  if (is_ts) {
    switch (f[2]) {
      case 's': ticks_per_day = 86400LL; ns_per_tick = 1000000000LL; break;
      case 'm': ticks_per_day = 86400LL * 1000; ns_per_tick = 1000000LL; break;
      case 'u': ticks_per_day = 86400LL * 1000000; ns_per_tick = 1000LL; break;
      default: break;
    }
  }
  // freq / offset / custom origin arrive in nanoseconds; the bucket arithmetic runs in index ticks
  if (ns_per_tick != 1) {
    if (freq_ns % ns_per_tick || offset_ns % ns_per_tick || (origin == 5 && origin_custom_ns % ns_per_tick))
      return set_err(PA_ERR_INVALID, "resample: freq / offset / origin are not whole multiples of the index unit ('%s')", f);
    freq_ns /= ns_per_tick;
    offset_ns /= ns_per_tick;
    origin_custom_ns /= ns_per_tick;
  }
  g->n = ix.n;
  if (g->n >= 0xFFFFFFFELL) return set_err(PA_ERR_NOT_IMPLEMENTED, "more than 2^32-2 rows per call; shard by row range");
  g->resample = true;
  g->index_format = f;
  g->fields.assign(1, KeyField{});
  g->key_data = ix.data;
  g->key_width = 8;
  g->rs = ResampleSpec{};
  g->rs.freq = freq_ns;
  g->rs.closed_right = closed_right != 0;
  g->rs.label_off = label_right ? freq_ns : 0;
  if (g->n > 0) {
    int64_t ends[2];
    CUDA_TRY(cudaMemcpyAsync(&ends[0], ix.data, 8, cudaMemcpyDeviceToHost, g->stream));
    CUDA_TRY(cudaMemcpyAsync(&ends[1], static_cast<const int64_t*>(ix.data) + (g->n - 1), 8, cudaMemcpyDeviceToHost, g->stream));
    CUDA_TRY(cudaStreamSynchronize(g->stream));
    if (ends[1] < ends[0]) return set_err(PA_ERR_INVALID, "resample: the index must be sorted ascending");
    int64_t first, last;
    anchor_range(ends[0], ends[1], freq_ns, closed_right != 0, origin, origin_custom_ns, offset_ns, ticks_per_day, &first, &last);
    if (first >= last) return set_err(PA_ERR_INVALID, "start date has to be less than end date");  // core.cpp:314-317
    g->rs.first = first;
    g->rs.nbins = (last - first) / freq_ns;
    if (g->n < g->rs.nbins) return set_err(PA_ERR_NOT_IMPLEMENTED, "upSampling is not implemented.");  // resample.h:102-105
  }
  *out = g.release();
  return PA_OK;
}

// pd::resample with a DateOffset rule (makeGroupInfo's DateOffset branch, /root/reference/src/resample.cpp:248-267;
// DateOffset::add, core.cpp:12-60; date_range + switchFunction, core.cpp:175-265; adjustBinEdges, resample.cpp:180-200).
// The host computes the O(#buckets) binner / edge / label arrays; the rows are reduced by the same sorted-run kernel
// as the fixed-width rules, which looks a bucket up in the edge array only where a new run of rows begins.
namespace {
int64_t days_from_civil(int64_t y, int m, int d) {       // proleptic Gregorian, days since 1970-01-01
  y -= m <= 2;
  const int64_t era = (y >= 0 ? y : y - 399) / 400;
  const int64_t yoe = y - era * 400;
  const int64_t doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
  const int64_t doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
  return era * 146097 + doe - 719468;
}
void civil_from_days(int64_t z, int64_t* y, int* m, int* d) {
  z += 719468;
  const int64_t era = (z >= 0 ? z : z - 146096) / 146097;
  const int64_t doe = z - era * 146097;
  const int64_t yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
  const int64_t doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
  const int64_t mp = (5 * doy + 2) / 153;
  *d = static_cast<int>(doy - (153 * mp + 2) / 5 + 1);
  *m = static_cast<int>(mp < 10 ? mp + 3 : mp - 9);
  *y = yoe + era * 400 + (*m <= 2);
}
// DateOffset::add for the rule types date_range accepts; the day of the month only survives for Day / WeekStart,
// the others snap to the first day of a month, so boost's end-of-month rules never come into play.
int64_t offset_add(int64_t day, int type, int64_t k) {
  if (type == PA_OFFSET_DAY) return day + k;
  if (type == PA_OFFSET_WEEK_START) return day + 7 * k;
  int64_t y; int m, d;
  civil_from_days(day, &y, &m, &d);
  int64_t ym = y * 12 + (m - 1);
  if (type == PA_OFFSET_MONTH_START) ym += k;
  else if (type == PA_OFFSET_QUARTER_START) { ym += 3 * k; ym = floor_div(ym, 3) * 3; }
  else /* YEAR_START */ { ym += 12 * k; ym = floor_div(ym, 12) * 12; }
  return days_from_civil(floor_div(ym, 12), static_cast<int>(ym - floor_div(ym, 12) * 12) + 1, 1);
}
}  // namespace

int pa_resample_create_calendar(const struct ArrowDeviceArray* index, const struct ArrowSchema* index_schema, int32_t offset_type,
                                int32_t multiplier, int32_t closed_right, int32_t label_right, const pa_options* opt,
                                pa_groupby** out) {
  if (!index || !index_schema || !out) return set_err(PA_ERR_INVALID, "pa_resample_create_calendar: null argument");
  if (multiplier < 1) return set_err(PA_ERR_INVALID, "FREQ must be >= 1");                                   // core.cpp:186-189
  switch (offset_type) {
    case PA_OFFSET_DAY: case PA_OFFSET_WEEK_START: case PA_OFFSET_MONTH_START: case PA_OFFSET_QUARTER_START: case PA_OFFSET_YEAR_START: break;
    case PA_OFFSET_MONTH_END: return set_err(PA_ERR_NOT_IMPLEMENTED, "MonthEnd not supported use arrow month().groupby()");       // core.cpp:243-260
    case PA_OFFSET_QUARTER_END: return set_err(PA_ERR_NOT_IMPLEMENTED, "QuarterEnd not supported use arrow quarter().groupby()");
    case PA_OFFSET_WEEK_END: return set_err(PA_ERR_NOT_IMPLEMENTED, "WeekEnd not supported use arrow weeks().groupby()");
    case PA_OFFSET_YEAR_END: return set_err(PA_ERR_NOT_IMPLEMENTED, "YearEnd not supported use arrow year().groupby()");
    default: return set_err(PA_ERR_INVALID, "unknown DateOffset type %d", offset_type);
  }
  if (!closed_right) return set_err(PA_ERR_NOT_IMPLEMENTED, "closed_left is not currently supported by DateOffset");   // resample.cpp:256-258
  HandlePtr g(new pa_groupby());
  PA_TRY(handle_init(g.get(), opt));
  g->keys.resize(1);
  Column& ix = g->keys[0];
  PA_TRY(load_column(index, index_schema, g->stream, g->device, &ix));
  const char* f = index_schema->format;
  if (!(f && f[0] == 't' && f[1] == 's') || ix.width != 8)
    return set_err(PA_ERR_INVALID, "axis must be a TimestampArray but got array of type '%s'", f ? f : "?");   // resample.cpp:213-216
  if (f[2] != 'n') return set_err(PA_ERR_NOT_IMPLEMENTED, "resample by a DateOffset rule needs a timestamp[ns] index (the reference's bins and labels are nanoseconds)");
  if (ix.valid) return set_err(PA_ERR_NOT_IMPLEMENTED, "resample: null timestamps are not supported");
  g->n = ix.n;
  if (g->n >= 0xFFFFFFFELL) return set_err(PA_ERR_NOT_IMPLEMENTED, "more than 2^32-2 rows per call; shard by row range");
  g->resample = true;
  g->index_format = f;
  g->fields.assign(1, KeyField{});
  g->key_data = ix.data;
  g->key_width = 8;
  g->rs = ResampleSpec{};
  g->rs.closed_right = 1;
  if (g->n > 0) {
    constexpr int64_t kDay = 86400LL * 1000000000LL;
    int64_t ends[2];
    CUDA_TRY(cudaMemcpyAsync(&ends[0], ix.data, 8, cudaMemcpyDeviceToHost, g->stream));
    CUDA_TRY(cudaMemcpyAsync(&ends[1], static_cast<const int64_t*>(ix.data) + (g->n - 1), 8, cudaMemcpyDeviceToHost, g->stream));
    CUDA_TRY(cudaStreamSynchronize(g->stream));
    if (ends[1] < ends[0]) return set_err(PA_ERR_INVALID, "resample: the index must be sorted ascending");
    const int64_t first_day = floor_div(ends[0], kDay), last_day = floor_div(ends[1], kDay);
    const int64_t start = offset_add(first_day, offset_type, -static_cast<int64_t>(multiplier));
    const int64_t stop = offset_add(last_day, offset_type, multiplier);
    if (start >= stop) return set_err(PA_ERR_INVALID, "start date has to be less than end date");              // core.cpp:181-184
    if (offset_type == PA_OFFSET_QUARTER_START) {
      int64_t y; int m, d;
      civil_from_days(start, &y, &m, &d);
      if (m / 3 != 0) return set_err(PA_ERR_INVALID, "A quarter freq requires month is on a quarter, +/- with DateOffset");   // core.cpp:247-250
    }
    // date_range(first - freq, last + freq, freq): the iterator steps `multiplier` units from `start` while <= stop
    std::vector<int64_t> binner;
    for (int64_t i = 0;; ++i) {
      int64_t day;
      if (offset_type == PA_OFFSET_DAY) day = start + i * multiplier;
      else if (offset_type == PA_OFFSET_WEEK_START) day = start + 7 * i * multiplier;
      else day = offset_add(start, offset_type, i * multiplier);
      if (day > stop) break;
      binner.push_back(day * kDay);
      if (binner.size() > (1u << 28)) return set_err(PA_ERR_INVALID, "resample: too many buckets");
    }
    std::vector<int64_t> edges(binner);
    if (!(offset_type == PA_OFFSET_DAY && multiplier == 1)) {                                                 // adjustBinEdges
      for (auto& e : edges) e += kDay - 1;
      if (edges.size() >= 2 && edges[edges.size() - 2] > ends[1]) { edges.pop_back(); binner.pop_back(); }
    }
    if (edges.size() < 2) return set_err(PA_ERR_INVALID, "Invalid length for values or for binner");
    if (ends[0] < edges.front()) return set_err(PA_ERR_INVALID, "Values falls before first bin");             // resample.cpp:33-41
    if (ends[1] > edges.back()) return set_err(PA_ERR_INVALID, "Values falls after last bin");
    const int64_t nbins = static_cast<int64_t>(edges.size()) - 1;
    if (g->n < nbins) return set_err(PA_ERR_NOT_IMPLEMENTED, "upSampling is not implemented.");               // resample.h:102-105
    std::vector<int64_t> labels(static_cast<size_t>(nbins));
    for (int64_t i = 0; i < nbins; ++i) labels[i] = binner[static_cast<size_t>(i) + (label_right ? 1 : 0)];
    // bin i = (edges[i], edges[i+1]] on ts = [edges[i], edges[i+1]) on ts' = ts - 1; a value equal to edges[0] passes
    // the reference's range check and lands in bin 0
    edges[0] -= 1;
    PA_TRY(g->rs_edges.alloc(edges.size() * 8, g->stream));
    PA_TRY(g->rs_labels.alloc(labels.size() * 8, g->stream));
    CUDA_TRY(cudaMemcpyAsync(g->rs_edges.p, edges.data(), edges.size() * 8, cudaMemcpyHostToDevice, g->stream));
    CUDA_TRY(cudaMemcpyAsync(g->rs_labels.p, labels.data(), labels.size() * 8, cudaMemcpyHostToDevice, g->stream));
    CUDA_TRY(cudaStreamSynchronize(g->stream));
    g->rs.first = edges[0];
    g->rs.freq = kDay;
    g->rs.nbins = nbins;
    g->rs.edges = g->rs_edges.as<int64_t>();
    g->rs.labels = g->rs_labels.as<int64_t>();
  }
  *out = g.release();
  return PA_OK;
}

// DataFrame::downsample (/root/reference/src/dataframe.cpp:1265-1290): per-row label = Floor/CeilTemporal(index)
// computed on the device (temporal.cuh), then the ordinary hash group-by on the labels.
int pa_downsample_create(const struct ArrowDeviceArray* index, const struct ArrowSchema* index_schema, int32_t multiple,
                         char unit, int32_t closed_label_right, int32_t week_starts_monday, int32_t calendar_based_origin,
                         const pa_options* opt, pa_groupby** out) {
  if (!index || !index_schema || !out) return set_err(PA_ERR_INVALID, "pa_downsample_create: null argument");
  if (multiple < 1) return set_err(PA_ERR_INVALID, "Invalid time offset: multiple must be >= 1");
  const char* units = "NULSTHDWMQY";
  if (!unit || !strchr(units, unit)) return set_err(PA_ERR_INVALID, "Invalid time offset unit '%c'", unit ? unit : '?');
  const char* f = index_schema->format;
  if (!f || f[0] != 't' || f[1] != 's') return set_err(PA_ERR_INVALID, "axis must be a TimestampArray but got array of type '%s'", f ? f : "?");
  HandlePtr g(new pa_groupby());
  PA_TRY(handle_init(g.get(), opt));
  cudaStream_t st = g->stream;
  Column ix;
  PA_TRY(load_column(index, index_schema, st, g->device, &ix));
  int64_t tps = 1000000000LL;
  switch (f[2]) {
    case 's': tps = 1; break;
    case 'm': tps = 1000; break;
    case 'u': tps = 1000000; break;
    default: break;
  }
  const bool calendar = unit == 'W' || unit == 'M' || unit == 'Q' || unit == 'Y';
  // the reference's W / M / Q / Y branch subtracts one day and re-types the labels as timestamp[ns]
  // (dataframe.cpp:1279-1287); other resolutions would change meaning there
  if (calendar && tps != 1000000000LL) return set_err(PA_ERR_NOT_IMPLEMENTED, "downsample by W / M / Q / Y needs a timestamp[ns] index");
  TemporalSpec spec{};
  spec.multiple = multiple;
  spec.unit = unit;
  spec.ceil = closed_label_right ? 1 : 0;
  spec.week_starts_monday = week_starts_monday ? 1 : 0;
  spec.calendar_origin = calendar_based_origin ? 1 : 0;
  spec.ticks_per_sec = tps;
  spec.post_shift = calendar ? -86400LL * tps : 0;
  g->keys.resize(1);
  Column& k = g->keys[0];
  k.format = calendar ? "tsn:" : f;
  k.width = 8; k.vc = VC_I; k.n = ix.n;
  PA_TRY(k.own_data.alloc(static_cast<size_t>(std::max<int64_t>(ix.n, 1)) * 8, st));
  if (ix.n > 0) {
    const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((ix.n + 255) / 256, static_cast<int64_t>(g->num_sms) * 16)));
    k_temporal_labels<<<grid, 256, 0, st>>>(static_cast<const int64_t*>(ix.data), ix.valid, ix.bit_off, ix.n, spec, k.own_data.as<int64_t>());
    CUDA_TRY(cudaGetLastError());
  }
  k.data = k.own_data.p;
  if (ix.valid) {   // null timestamps stay null labels: their own group, like any null key
    if (ix.own_valid.p) k.own_valid = std::move(ix.own_valid);
    k.valid = ix.valid;
    k.bit_off = ix.bit_off;
  }
  g->n = ix.n;
  if (g->n >= 0xFFFFFFFEll) return set_err(PA_ERR_NOT_IMPLEMENTED, "more than 2^32-2 rows per call; shard by row range");
  PA_TRY(setup_keys(g.get()));
  CUDA_TRY(cudaStreamSynchronize(st));   // `ix` may own the device copy of a host index
  *out = g.release();
  return PA_OK;
}

// ---- ingest: one host Arrow array -> device-resident ArrowDeviceArray ----
// What DataFrame::readBinary / readParquet (dataframe.cpp:757-791, 646-683) produce are host Arrow arrays inside an IPC
// blob or a decoded table; this puts a column on the device ONCE (pageable memory through the pinned staging pipeline of
// h2d_copy) and hands back an ArrowDeviceArray(CUDA) that every other entry point uses in place, so that the
// constructor and every aggregate after it stop paying PCIe.
namespace {
struct DevArrayPriv {
  void* owned[3] = {nullptr, nullptr, nullptr};
  const void* bufs[3] = {nullptr, nullptr, nullptr};
};
void dev_array_release(struct ArrowArray* a) {
  if (!a || !a->release) return;
  auto* p = static_cast<DevArrayPriv*>(a->private_data);
  if (p) {
    for (void* q : p->owned) if (q) cudaFree(q);
    delete p;
  }
  a->release = nullptr;
}
}  // namespace

int pa_column_to_device(const struct ArrowDeviceArray* host, const struct ArrowSchema* schema, const pa_options* opt,
                        struct ArrowDeviceArray* out) {
  if (!host || !schema || !out || !schema->format) return set_err(PA_ERR_INVALID, "pa_column_to_device: null argument");
  if (host->device_type != ARROW_DEVICE_CPU && host->device_type != ARROW_DEVICE_CUDA_HOST)
    return set_err(PA_ERR_INVALID, "pa_column_to_device: the input must be a host array");
  if (schema->dictionary) return set_err(PA_ERR_NOT_IMPLEMENTED, "pa_column_to_device: dictionary-encoded columns (ingest the indices)");
  const ArrowArray& a = host->array;
  const char f0 = schema->format[0];
  const bool is_str = f0 == 'u' || f0 == 'U';
  const bool is_bool = f0 == 'b';
  int width = 0, vc = 0;
  if (!is_str) PA_TRY(parse_format(schema->format, &width, &vc));
  if (a.n_buffers < (is_str ? 3 : 2)) return set_err(PA_ERR_INVALID, "unexpected buffer count %lld for format '%s'", (long long)a.n_buffers, schema->format);
  int device = 0;
  if (opt && opt->device >= 0) device = opt->device; else CUDA_TRY(cudaGetDevice(&device));
  CUDA_TRY(cudaSetDevice(device));
  cudaStream_t st = opt && opt->cuda_stream ? static_cast<cudaStream_t>(opt->cuda_stream) : nullptr;
  std::unique_ptr<DevArrayPriv> p(new DevArrayPriv());
  struct Guard { DevArrayPriv* p; ~Guard() { if (p) for (void* q : p->owned) if (q) cudaFree(q); } } guard{p.get()};
  const int64_t k = a.offset % 8, start = a.offset - k, n = a.length;
  const bool has_nulls = a.null_count != 0 && a.buffers[0] != nullptr;
  const size_t bit_bytes = static_cast<size_t>((k + n + 7) / 8);
  if (has_nulls) {
    CUDA_TRY(cudaMalloc(&p->owned[0], std::max<size_t>(bit_bytes, 16)));
    PA_TRY(h2d_copy(p->owned[0], static_cast<const uint8_t*>(a.buffers[0]) + start / 8, bit_bytes, st));
    p->bufs[0] = p->owned[0];
  }
  if (is_str) {
    const int ow = f0 == 'U' ? 8 : 4;
    const char* offs = static_cast<const char*>(a.buffers[1]) + start * ow;
    const size_t obytes = static_cast<size_t>(k + n + 1) * ow;
    CUDA_TRY(cudaMalloc(&p->owned[1], std::max<size_t>(obytes, 16)));
    int64_t first = 0, last = 0;
    if (a.buffers[1]) {
      PA_TRY(h2d_copy(p->owned[1], offs, obytes, st));
      first = ow == 8 ? reinterpret_cast<const int64_t*>(offs)[0] : reinterpret_cast<const int32_t*>(offs)[0];
      last = ow == 8 ? reinterpret_cast<const int64_t*>(offs)[k + n] : reinterpret_cast<const int32_t*>(offs)[k + n];
    } else {
      CUDA_TRY(cudaMemsetAsync(p->owned[1], 0, obytes, st));
    }
    CUDA_TRY(cudaMalloc(&p->owned[2], static_cast<size_t>(std::max<int64_t>(last - first, 16))));
    if (last > first) PA_TRY(h2d_copy(p->owned[2], static_cast<const uint8_t*>(a.buffers[2]) + first, static_cast<size_t>(last - first), st));
    p->bufs[1] = p->owned[1];
    p->bufs[2] = static_cast<const uint8_t*>(p->owned[2]) - first;     // offsets stay absolute
  } else if (is_bool) {
    CUDA_TRY(cudaMalloc(&p->owned[1], std::max<size_t>(bit_bytes, 16)));
    if (a.buffers[1]) PA_TRY(h2d_copy(p->owned[1], static_cast<const uint8_t*>(a.buffers[1]) + start / 8, bit_bytes, st));
    p->bufs[1] = p->owned[1];
  } else {
    const size_t bytes = static_cast<size_t>(k + n) * width;
    CUDA_TRY(cudaMalloc(&p->owned[1], std::max<size_t>(bytes, 16)));
    if (a.buffers[1]) PA_TRY(h2d_copy(p->owned[1], static_cast<const char*>(a.buffers[1]) + start * width, bytes, st));
    p->bufs[1] = p->owned[1];
  }
  CUDA_TRY(cudaStreamSynchronize(st));
  memset(out, 0, sizeof *out);
  out->array.length = n;
  out->array.null_count = has_nulls ? a.null_count : 0;
  out->array.offset = k;
  out->array.n_buffers = is_str ? 3 : 2;
  out->array.n_children = 0;
  out->array.buffers = p->bufs;
  out->array.release = dev_array_release;
  out->device_id = device;
  out->device_type = ARROW_DEVICE_CUDA;
  out->sync_event = nullptr;
  guard.p = nullptr;
  out->array.private_data = p.release();
  return PA_OK;
}

// ---- stable argsort of one column (sort.cuh): Series::sort / argsort, DataFrame::sort_index / sort_values ----
// The result lives on a pa_groupby handle in "sorted" state: its row order / inverse order are what a groupings build
// leaves there, so pa_groupby_take_grouped takes any column into sorted order with the scatter-shaped kernel.
}  // extern "C"
template <int VC>
static int sort_build_t(pa_groupby* g, const Column& col, bool descending) {
  cudaStream_t st = g->stream;
  const int64_t n = g->n;
  const size_t nb4 = static_cast<size_t>(std::max<int64_t>(n, 1)) * 4, nb8 = nb4 * 2;
  PA_TRY(g->grp_order.alloc(nb4, st));
  PA_TRY(g->grp_dest.alloc(nb4, st));
  if (n == 0) return PA_OK;
  DevBuf keys_a, keys_b, pay_a, pay_b, cls, hist, counts, tile_sums;
  PA_TRY(keys_a.alloc(nb8, st));
  PA_TRY(cls.alloc(static_cast<size_t>(n), st));
  PA_TRY(hist.alloc(sizeof(uint32_t) * (LS_DIGITS * CS_R + 4), st));
  CUDA_TRY(cudaMemsetAsync(hist.p, 0, sizeof(uint32_t) * (LS_DIGITS * CS_R + 4), st));
  SortKeyArgs ka{};
  ka.vals = col.data; ka.valid = col.valid; ka.voff = col.bit_off; ka.n = n; ka.vw = col.width; ka.descending = descending ? 1 : 0;
  ka.keys = keys_a.as<uint64_t>(); ka.cls = cls.as<uint8_t>(); ka.hist = hist.as<uint32_t>(); ka.flags = hist.as<uint32_t>() + LS_DIGITS * CS_R;
  const int kgrid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((n + 511) / 512, static_cast<int64_t>(g->num_sms) * 4)));
  k_sort_keys<VC><<<kgrid, 512, 0, st>>>(ka);
  CUDA_TRY(cudaGetLastError());
  std::vector<uint32_t> h(LS_DIGITS * CS_R + 4);
  CUDA_TRY(cudaMemcpyAsync(h.data(), hist.p, h.size() * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  // passes: every digit that is not constant over the column, then the class byte when NaNs / nulls exist
  std::vector<int> passes;
  for (int d = 0; d < LS_DIGITS; ++d) {
    bool constant = false;
    for (int b = 0; b < CS_R; ++b) if (h[d * CS_R + b] == static_cast<uint64_t>(n)) { constant = true; break; }
    if (!constant) passes.push_back(d);
  }
  if (h[LS_DIGITS * CS_R] & 1u) passes.push_back(-1);
  const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, static_cast<int64_t>(g->num_sms) * 16)));
  if (passes.empty()) {
    k_ls_identity<<<grid, 256, 0, st>>>(g->grp_order.as<uint32_t>(), g->grp_dest.as<uint32_t>(), n);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(st));
    return PA_OK;
  }
  const int64_t ntiles = (n + CS_TILE - 1) / CS_TILE;
  const int nb = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ntiles, static_cast<int64_t>(g->num_sms) * 4)));
  const int64_t chunk = ((ntiles + nb - 1) / nb) * CS_TILE;
  PA_TRY(counts.alloc(sizeof(uint32_t) * CS_R * static_cast<size_t>(nb), st));
  if (passes.size() > 1) {
    PA_TRY(keys_b.alloc(nb8, st));
    PA_TRY(pay_a.alloc(nb4, st));
    if (passes.size() > 2) PA_TRY(pay_b.alloc(nb4, st));
  }
  CUDA_TRY(cudaFuncSetAttribute(k_ls_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(CS_SMEM)));
  const uint64_t* in_keys = keys_a.as<uint64_t>();
  const uint32_t* in_pay = nullptr;
  for (size_t p = 0; p < passes.size(); ++p) {
    const bool last = p + 1 == passes.size();
    LsArgs a{};
    a.keys = in_keys; a.payload = in_pay; a.n = n; a.nb = nb; a.chunk = chunk;
    a.cls = passes[p] < 0 ? cls.as<uint8_t>() : nullptr;
    a.shift = passes[p] < 0 ? 0 : passes[p] * CS_BITS;
    a.counts = counts.as<uint32_t>();
    DevBuf& ok = (p % 2 == 0) ? keys_b : keys_a;          // (keys_a holds the input of pass 0)
    DevBuf& op = (p % 2 == 0) ? pay_a : pay_b;
    a.out_keys = last ? nullptr : ok.as<uint64_t>();
    a.out_payload = last ? g->grp_order.as<uint32_t>() : op.as<uint32_t>();
    a.out_dest = last ? g->grp_dest.as<uint32_t>() : nullptr;
    k_ls_hist<<<nb, CS_THREADS, 0, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    PA_TRY(scan_u32_on(g, counts.as<uint32_t>(), static_cast<uint64_t>(CS_R) * nb, &tile_sums));
    k_ls_scatter<<<nb, CS_THREADS, CS_SMEM, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    in_keys = a.out_keys;
    in_pay = a.out_payload;
  }
  CUDA_TRY(cudaStreamSynchronize(st));   // the ping-pong buffers are released here
  return PA_OK;
}
extern "C" {

int pa_sort_create(const struct ArrowDeviceArray* values, const struct ArrowSchema* schema, int32_t ascending,
                   const pa_options* opt, pa_groupby** out) {
  if (!values || !schema || !out) return set_err(PA_ERR_INVALID, "pa_sort_create: null argument");
  if (schema->dictionary) return set_err(PA_ERR_NOT_IMPLEMENTED, "sorting dictionary-encoded columns");
  HandlePtr g(new pa_groupby());
  PA_TRY(handle_init(g.get(), opt));
  cudaStream_t st = g->stream;
  g->keys.resize(1);
  Column& col = g->keys[0];
  PA_TRY(load_column(values, schema, st, g->device, &col));
  if (col.is_str || col.is_bool) return set_err(PA_ERR_NOT_IMPLEMENTED, "sorting '%s' columns (numeric and temporal types only)", schema->format);
  g->n = col.n;
  if (g->n >= (1ll << 31)) return set_err(PA_ERR_INVALID, "sort uses 32-bit row numbers: at most 2^31 - 1 rows");
  CUDA_TRY(cudaEventRecord(g->ev[6], st));
  int rc;
  if (col.vc == VC_F) rc = sort_build_t<VC_F>(g.get(), col, !ascending);
  else if (col.vc == VC_I) rc = sort_build_t<VC_I>(g.get(), col, !ascending);
  else rc = sort_build_t<VC_U>(g.get(), col, !ascending);
  PA_TRY(rc);
  CUDA_TRY(cudaEventRecord(g->ev[7], st));
  CUDA_TRY(cudaEventRecord(g->ev[8], st));
  CUDA_TRY(cudaEventRecord(g->ev[9], st));
  CUDA_TRY(cudaStreamSynchronize(st));
  g->sorted_state = true;
  g->have_groups = true;
  g->G = 0;
  g->have_groupings = true;
  *out = g.release();
  return PA_OK;
}

int pa_sort_indices(pa_groupby* g, struct ArrowArray* out, struct ArrowSchema* out_schema) {
  if (!g || !out || !out_schema) return set_err(PA_ERR_INVALID, "null argument");
  if (!g->sorted_state) return set_err(PA_ERR_STATE, "not a handle made by pa_sort_create");
  cudaStream_t st = g->stream;
  CUDA_TRY(cudaSetDevice(g->device));
  DevBuf wide;
  PA_TRY(wide.alloc(static_cast<size_t>(std::max<int64_t>(g->n, 1)) * 8, st));
  if (g->n > 0) {
    const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((g->n + 255) / 256, static_cast<int64_t>(g->num_sms) * 16)));
    k_ls_widen<<<grid, 256, 0, st>>>(g->grp_order.as<uint32_t>(), wide.as<uint64_t>(), g->n);
    CUDA_TRY(cudaGetLastError());
  }
  return export_host(st, "L", 8, static_cast<uint32_t>(g->n), wide.p, nullptr, out, out_schema);
}

// ---- multi-GPU partial export / merge ----
int pa_groupby_partials_count(pa_groupby* g, int32_t n_parts, int64_t* counts_host) {
  if (!g || !counts_host || n_parts < 1 || n_parts > 64) return set_err(PA_ERR_INVALID, "bad argument (1 <= n_parts <= 64)");
  PA_TRY(finish_pending(g));
  if (!g->have_groups || g->merged) return set_err(PA_ERR_STATE, "partials need a finished local aggregate");
  if (g->keys.size() != 1 && !g->resample) return set_err(PA_ERR_NOT_IMPLEMENTED, "multi-GPU merge of composite keys");
  if (!g->resample && g->keys[0].is_str) return set_err(PA_ERR_NOT_IMPLEMENTED, "multi-GPU merge of utf8 keys: dictionary-encode the column against a shared dictionary");
  PA_TRY(ensure_device(g));
  cudaStream_t st = g->stream;
  DevBuf counts;
  PA_TRY(counts.alloc(sizeof(uint64_t) * n_parts, st));
  CUDA_TRY(cudaMemsetAsync(counts.p, 0, sizeof(uint64_t) * n_parts, st));
  PartialsArgs a{};
  a.r = g->res; a.G = g->G; a.nparts = n_parts; a.counts = counts.as<unsigned long long>();
  if (g->G) {
    k_partials_count<<<(g->G + 255) / 256, 256, 0, st>>>(a);
    CUDA_TRY(cudaGetLastError());
  }
  std::vector<uint64_t> h(n_parts);
  CUDA_TRY(cudaMemcpyAsync(h.data(), counts.p, sizeof(uint64_t) * n_parts, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  g->parts_n = n_parts;
  g->parts_counts.assign(h.begin(), h.end());
  for (int i = 0; i < n_parts; ++i) counts_host[i] = static_cast<int64_t>(h[i]);
  return PA_OK;
}

// `compact`: REC_WORDS_COMPACT-word records (merge.cuh) — the sharded step's own format for narrow aggregate sets
static int partials_export(pa_groupby* g, int32_t n_parts, void* dev_records, int64_t capacity_records, bool compact) {
  if (!g || !dev_records) return set_err(PA_ERR_INVALID, "null argument");
  if (g->parts_n != n_parts) return set_err(PA_ERR_STATE, "call pa_groupby_partials_count(n_parts=%d) first", n_parts);
  if (capacity_records < static_cast<int64_t>(g->G)) return set_err(PA_ERR_INVALID, "record buffer too small");
  PA_TRY(ensure_device(g));
  cudaStream_t st = g->stream;
  std::vector<uint64_t> prefix(n_parts, 0);
  for (int i = 1; i < n_parts; ++i) prefix[i] = prefix[i - 1] + static_cast<uint64_t>(g->parts_counts[i - 1]);
  DevBuf cursor;
  PA_TRY(cursor.alloc(sizeof(uint64_t) * n_parts, st));
  CUDA_TRY(cudaMemcpyAsync(cursor.p, prefix.data(), sizeof(uint64_t) * n_parts, cudaMemcpyHostToDevice, st));
  PartialsArgs a{};
  a.r = g->res; a.G = g->G; a.nparts = n_parts; a.row_base = g->opt.row_base;
  a.vw = g->last_vw; a.wide = g->last_wide;
  a.compact = compact;
  for (auto& o : g->outs) {
    if (o.bit == AGG_FIRST) { a.first_vals = o.values.p; a.first_valid = o.valid.as<uint32_t>(); }
    if (o.bit == AGG_LAST) { a.last_vals = o.values.p; a.last_valid = o.valid.as<uint32_t>(); }
  }
  a.cursor = cursor.as<unsigned long long>();
  a.records = static_cast<uint64_t*>(dev_records);
  if (g->G) {
    k_partials_scatter<<<(g->G + 255) / 256, 256, 0, st>>>(a);
    CUDA_TRY(cudaGetLastError());
  }
  CUDA_TRY(cudaStreamSynchronize(st));   // `prefix` and `cursor` die here; records are complete for the caller's collective
  return PA_OK;
}

int pa_groupby_partials_export(pa_groupby* g, int32_t n_parts, void* dev_records, int64_t capacity_records) {
  return partials_export(g, n_parts, dev_records, capacity_records, false);
}

// Common part of the two merge entry points.  `d_off` = device array [n_sources + 1], exclusive prefix
// of the per-source record counts (the exact total is d_off[n_sources]); nrec_max = host-side upper bound.
// Scratch of one merge (key table, per-source index, compaction / sort buffers).  A communicator keeps one between
// steps (pa_groupby_sharded_aggregate) so that a step allocates nothing but its result.
struct MergeScratch { DevBuf tkeys, idx, m_first, m_slot, s_first, s_slot, cub_tmp; uint64_t last_slots = 0; };

static int merge_build(pa_groupby* g, const void* dev_records, const uint64_t* d_off, int32_t n_sources, uint64_t nrec_max,
                       uint32_t agg_mask, const char* value_format, const char* key_format, MergeScratch* keep = nullptr,
                       uint64_t row_bound = 0, bool compact = false, uint64_t distinct_hint = 0) {
  cudaStream_t st = g->stream;
  g->merged = true;
  g->keys.resize(1);
  int kw = 8, kvc = VC_I;
  PA_TRY(parse_format(key_format, &kw, &kvc));
  g->keys[0].format = key_format;
  g->keys[0].width = kw;
  g->fields.assign(1, KeyField{});
  g->fields[0].width = kw;
  g->index_format = key_format;
  PA_TRY(parse_format(value_format, &g->last_vw, &g->last_vc));
  g->last_vfmt = value_format;
  if (nrec_max >= 0xFFFFFFFFull) return set_err(PA_ERR_NOT_IMPLEMENTED, "more than 2^32-1 partial records per rank");
  MergeScratch local_scratch;
  MergeScratch& ms = keep ? *keep : local_scratch;
  DevBuf &tkeys = ms.tkeys, &idx = ms.idx, &m_first = ms.m_first, &m_slot = ms.m_slot, &s_first = ms.s_first, &s_slot = ms.s_slot, &cub_tmp = ms.cub_tmp;
  PA_TRY(m_first.alloc(std::max<uint64_t>(nrec_max, 1) * 8, st));
  PA_TRY(m_slot.alloc(std::max<uint64_t>(nrec_max, 1) * 4, st));
  MergeArgs a{};
  a.records = static_cast<const uint64_t*>(dev_records);
  a.src_offset = d_off;
  a.nsrc = n_sources; a.nrec = nrec_max; a.compact = compact;
  a.status = g->status.as<uint32_t>();
  a.m_first_row = m_first.as<uint64_t>(); a.m_slot = m_slot.as<uint32_t>();
  a.vc = g->last_vc;
  // Join table: 2^k >= 2 x the number of distinct keys.  A source holds every key at most once, so the count lies
  // between the largest source and the sum; sources that saw (nearly) the same keys — row-range shards of one column —
  // are the common case, so the table is sized by `distinct_hint` (the largest source; 0 = unknown) first: 8 sources of
  // 12.5 M records fill 2^25 slots instead of 2^28 (index array 1 GB instead of 8.6 GB per step).  Too small shows as a
  // probe sequence beyond max_probe; then the merge is redone with the table every key set fits.
  uint32_t h_status[ST_WORDS];
  uint64_t sized_for = distinct_hint ? std::min<uint64_t>(std::max<uint64_t>(distinct_hint + distinct_hint / 4, 512), std::max<uint64_t>(nrec_max, 512)) : std::max<uint64_t>(nrec_max, 512);
  for (;;) {
    uint64_t cap = 1024;
    while (cap < sized_for * 2) cap <<= 1;
    const uint64_t nslots = cap + 2;
    ms.last_slots = nslots;
    PA_TRY(tkeys.alloc(nslots * 8, st));
    PA_TRY(idx.alloc(nslots * n_sources * 4, st));
    const int fgrid = static_cast<int>(std::min<uint64_t>((nslots + 255) / 256, static_cast<uint64_t>(g->num_sms) * 16));
    k_fill_u64<<<fgrid, 256, 0, st>>>(tkeys.as<unsigned long long>(), nslots, kEmptyKey);
    CUDA_TRY(cudaMemsetAsync(idx.p, 0xFF, nslots * n_sources * 4, st));
    a.tkeys = tkeys.as<unsigned long long>(); a.cap_mask = cap - 1; a.idx = idx.as<uint32_t>();
    a.max_probe = sized_for >= nrec_max ? cap : std::min<uint64_t>(cap, 2048);
    CUDA_TRY(cudaEventRecord(g->ev[1], st));
    if (nrec_max) {
      k_merge_insert<<<static_cast<int>((nrec_max + 255) / 256), 256, 0, st>>>(a);
      CUDA_TRY(cudaGetLastError());
    }
    k_merge_compact<<<static_cast<int>((nslots + MC_THREADS - 1) / MC_THREADS), MC_THREADS, 0, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(g->ev[2], st));
    CUDA_TRY(cudaMemcpyAsync(h_status, g->status.p, sizeof h_status, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (h_status[ST_PEER_OVERFLOW]) return set_err(PA_ERR_STATE, "a source rank had more groups than a padded block holds; use the counted exchange");
    if (h_status[ST_OVERFLOW] && sized_for < nrec_max) {   // the hint was too small: every key set fits a table sized by the record count
      sized_for = nrec_max;
      CUDA_TRY(cudaMemsetAsync(a.status + ST_OVERFLOW, 0, sizeof(uint32_t), st));
      CUDA_TRY(cudaMemsetAsync(a.status + ST_COUNTER, 0, sizeof(uint32_t), st));
      continue;
    }
    if (h_status[ST_OVERFLOW]) return set_err(PA_ERR_CUDA, "merge table overflow");
    break;
  }
  const uint32_t G = h_status[ST_COUNTER];
  g->G = G;
  const bool wide = is_wide(agg_mask, g->last_vc);
  g->last_wide = wide;
  PA_TRY(alloc_result(g, G, true, true));
  PA_TRY(g->m_count64.alloc(static_cast<size_t>(std::max<uint32_t>(G, 1)) * 8, st));
  g->res.count64 = g->m_count64.as<uint64_t>();
  PA_TRY(g->m_first_val.alloc(static_cast<size_t>(std::max<uint32_t>(G, 1)) * 8, st));
  PA_TRY(g->m_first_row_g.alloc(static_cast<size_t>(std::max<uint32_t>(G, 1)) * 8, st));
  PA_TRY(g->m_last_val.alloc(static_cast<size_t>(std::max<uint32_t>(G, 1)) * 8, st));
  PA_TRY(g->m_first_valid.alloc(std::max<uint32_t>(G, 1), st));
  PA_TRY(g->m_last_valid.alloc(std::max<uint32_t>(G, 1), st));
  if (G) {
    PA_TRY(s_first.alloc(static_cast<size_t>(G) * 8, st));
    PA_TRY(s_slot.alloc(static_cast<size_t>(G) * 4, st));
    size_t tmp_bytes = 0;
    // (global first rows are below row_bound when the caller knows the shards: sort only the bits that can differ)
    int end_bit = 64;
    if (row_bound) { end_bit = 1; while (end_bit < 64 && (row_bound >> end_bit)) ++end_bit; }
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, m_first.as<uint64_t>(), s_first.as<uint64_t>(), m_slot.as<uint32_t>(), s_slot.as<uint32_t>(), static_cast<int>(G), 0, end_bit, st));
    PA_TRY(cub_tmp.alloc(tmp_bytes, st));
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(cub_tmp.p, tmp_bytes, m_first.as<uint64_t>(), s_first.as<uint64_t>(), m_slot.as<uint32_t>(), s_slot.as<uint32_t>(), static_cast<int>(G), 0, end_bit, st));
    a.order = s_slot.as<uint32_t>();
    a.G = G;
    a.out = g->res;
    a.o_first_val = g->m_first_val.as<uint64_t>(); a.o_last_val = g->m_last_val.as<uint64_t>();
    a.o_first_valid = g->m_first_valid.as<uint8_t>(); a.o_last_valid = g->m_last_valid.as<uint8_t>();
    a.o_first_row_g = g->m_first_row_g.as<uint64_t>();
    k_merge_fold<<<(G + 255) / 256, 256, 0, st>>>(a);
    CUDA_TRY(cudaGetLastError());
  }
  CUDA_TRY(cudaEventRecord(g->ev[3], st));
  g->have_groups = true;
  g->last_launches = 4;
  g->last_path = 4;
  PA_TRY(run_emit(g, nullptr, agg_mask));
  CUDA_TRY(cudaEventRecord(g->ev[4], st));
  CUDA_TRY(cudaStreamSynchronize(st));   // local scratch (table, idx, sort buffers) is released after this point
  return PA_OK;
}

int pa_merge_create(const void* dev_records, const int64_t* counts_by_source, int32_t n_sources, uint32_t agg_mask,
                    const char* value_format, const char* key_format, const pa_options* opt, pa_groupby** out) {
  if (!counts_by_source || n_sources < 1 || !value_format || !key_format || !out) return set_err(PA_ERR_INVALID, "null argument");
  if (agg_mask & ~PA_AGG_ALL) return set_err(PA_ERR_INVALID, "bad aggregate mask");
  HandlePtr g(new pa_groupby());
  PA_TRY(handle_init(g.get(), opt));
  cudaStream_t st = g->stream;
  std::vector<uint64_t> off(n_sources + 1, 0);
  for (int i = 0; i < n_sources; ++i) {
    if (counts_by_source[i] < 0) return set_err(PA_ERR_INVALID, "negative record count");
    off[i + 1] = off[i] + static_cast<uint64_t>(counts_by_source[i]);
  }
  const uint64_t nrec = off[n_sources];
  if (nrec && !dev_records) return set_err(PA_ERR_INVALID, "null record buffer");
  DevBuf d_off;
  PA_TRY(d_off.alloc(sizeof(uint64_t) * (n_sources + 1), st));
  CUDA_TRY(cudaMemcpyAsync(d_off.p, off.data(), sizeof(uint64_t) * (n_sources + 1), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemsetAsync(g->status.p, 0, sizeof(uint32_t) * ST_WORDS, st));
  CUDA_TRY(cudaEventRecord(g->ev[0], st));
  PA_TRY(merge_build(g.get(), dev_records, d_off.as<uint64_t>(), n_sources, nrec, agg_mask, value_format, key_format, nullptr, 0, false,
                     static_cast<uint64_t>(*std::max_element(counts_by_source, counts_by_source + n_sources))));
  *out = g.release();
  return PA_OK;
}

int pa_groupby_partials_export_padded(pa_groupby* g, int32_t n_parts, void* dev_blocks, int64_t block_records) {
  if (!g || !dev_blocks || n_parts < 1 || n_parts > 64 || block_records < 1) return set_err(PA_ERR_INVALID, "bad argument (1 <= n_parts <= 64)");
  if (!g->have_groups || g->merged) return set_err(PA_ERR_STATE, "partials need a finished local aggregate");
  if (g->keys.size() != 1 && !g->resample) return set_err(PA_ERR_NOT_IMPLEMENTED, "multi-GPU merge of composite keys");
  if (!g->resample && g->keys[0].is_str) return set_err(PA_ERR_NOT_IMPLEMENTED, "multi-GPU merge of utf8 keys: dictionary-encode the column against a shared dictionary");
  PA_TRY(ensure_device(g));
  PartialsArgs a{};
  a.r = g->res; a.G = g->G; a.nparts = n_parts; a.row_base = g->opt.row_base;
  a.vw = g->last_vw; a.wide = g->last_wide;
  for (auto& o : g->outs) {
    if (o.bit == AGG_FIRST) { a.first_vals = o.values.p; a.first_valid = o.valid.as<uint32_t>(); }
    if (o.bit == AGG_LAST) { a.last_vals = o.values.p; a.last_valid = o.valid.as<uint32_t>(); }
  }
  a.records = static_cast<uint64_t*>(dev_blocks);
  if (g->pending) {   // deferred local aggregate: group count and failure flag are still on the device
    a.G_dev = g->status.as<uint32_t>() + ST_NGROUPS;
    a.abort_dev = g->status.as<uint32_t>() + ST_ABORT;
  }
  k_partials_pack_padded<<<1, 1024, 0, g->stream>>>(a, static_cast<uint64_t>(block_records));
  CUDA_TRY(cudaGetLastError());
  return PA_OK;   // stream ordered: no host synchronisation
}

int pa_merge_create_padded(const void* dev_blocks, int32_t n_sources, int64_t block_records, uint32_t agg_mask,
                           const char* value_format, const char* key_format, const pa_options* opt, pa_groupby** out) {
  if (!dev_blocks || n_sources < 1 || n_sources > 64 || block_records < 1 || !value_format || !key_format || !out)
    return set_err(PA_ERR_INVALID, "bad argument (1 <= n_sources <= 64)");
  if (agg_mask & ~PA_AGG_ALL) return set_err(PA_ERR_INVALID, "bad aggregate mask");
  HandlePtr g;
  {
    int dev = opt && opt->device >= 0 ? opt->device : -1;
    if (dev < 0) CUDA_TRY(cudaGetDevice(&dev));
    pa_groupby* recycled = (opt && opt->cuda_stream) ? pool_take(dev, static_cast<cudaStream_t>(opt->cuda_stream)) : nullptr;
    if (recycled) {
      g.reset(recycled);
      g->opt = *opt;
      CUDA_TRY(cudaSetDevice(g->device));
    } else {
      g.reset(new pa_groupby());
      PA_TRY(handle_init(g.get(), opt));
    }
  }
  g->poolable = false;
  cudaStream_t st = g->stream;
  const uint64_t nrec_max = static_cast<uint64_t>(n_sources) * static_cast<uint64_t>(block_records);
  CUDA_TRY(cudaMemsetAsync(g->status.p, 0, sizeof(uint32_t) * ST_WORDS, st));
  CUDA_TRY(cudaEventRecord(g->ev[0], st));
  // ---- few records: the whole merge is one single-CTA kernel + emit, one synchronisation ----
  {
    pa_groupby* h = g.get();
    h->merged = true;
    h->keys.resize(1);
    int kw = 8, kvc = VC_I;
    PA_TRY(parse_format(key_format, &kw, &kvc));
    h->keys[0].format = key_format;
    h->keys[0].width = kw;
    h->fields.assign(1, KeyField{});
    h->fields[0].width = kw;
    h->index_format = key_format;
    PA_TRY(parse_format(value_format, &h->last_vw, &h->last_vc));
    h->last_vfmt = value_format;
    h->last_wide = is_wide(agg_mask, h->last_vc);
    const uint32_t cap_out = static_cast<uint32_t>(std::min<uint64_t>(nrec_max, MS_MAX_GROUPS)) + 2;
    PA_TRY(alloc_result(h, cap_out, true, true));
    PA_TRY(h->m_count64.alloc(static_cast<size_t>(cap_out) * 8, st));
    h->res.count64 = h->m_count64.as<uint64_t>();
    PA_TRY(h->m_first_val.alloc(static_cast<size_t>(cap_out) * 8, st));
    PA_TRY(h->m_first_row_g.alloc(static_cast<size_t>(cap_out) * 8, st));
    PA_TRY(h->m_last_val.alloc(static_cast<size_t>(cap_out) * 8, st));
    PA_TRY(h->m_first_valid.alloc(cap_out, st));
    PA_TRY(h->m_last_valid.alloc(cap_out, st));
    MergeSmallArgs m{};
    m.blocks = static_cast<const uint64_t*>(dev_blocks);
    m.nsrc = static_cast<uint32_t>(n_sources);
    m.block_records = static_cast<uint64_t>(block_records);
    m.vc = h->last_vc;
    m.out = h->res;
    m.o_first_val = h->m_first_val.as<uint64_t>(); m.o_last_val = h->m_last_val.as<uint64_t>();
    m.o_first_valid = h->m_first_valid.as<uint8_t>(); m.o_last_valid = h->m_last_valid.as<uint8_t>();
    m.o_first_row_g = h->m_first_row_g.as<uint64_t>();
    m.status = h->status.as<uint32_t>();
    static bool attr_set[64] = {};
    if (!attr_set[h->device & 63]) {
      CUDA_TRY(cudaFuncSetAttribute(k_merge_small, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(MS_SMEM_BYTES)));
      attr_set[h->device & 63] = true;
    }
    CUDA_TRY(cudaEventRecord(h->ev[1], st));
    k_merge_small<<<1, MS_THREADS, MS_SMEM_BYTES, st>>>(m);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(h->ev[2], st));
    CUDA_TRY(cudaEventRecord(h->ev[3], st));
    h->have_groups = true;
    h->last_launches = 1;
    h->last_path = 4;
    h->G = cap_out;                                     // upper bound for the emit grid and the output buffers
    h->emit_G_dev = h->status.as<uint32_t>() + ST_COUNTER;
    PA_TRY(run_emit(h, nullptr, agg_mask));
    h->emit_G_dev = nullptr;
    CUDA_TRY(cudaEventRecord(h->ev[4], st));
    // the status words (group count, overflow markers) are read by the first call that needs them
    h->pending_merge = true;
    h->poolable = true;
    h->pm_blocks = dev_blocks; h->pm_nsrc = n_sources; h->pm_block_records = block_records; h->pm_mask = agg_mask;
    h->pm_vfmt = value_format; h->pm_kfmt = key_format;
    *out = g.release();
    return PA_OK;
  }
}

extern "C++" {
namespace {
int merge_padded_general(pa_groupby* g, const void* dev_blocks, int32_t n_sources, int64_t block_records, uint32_t agg_mask,
                         const char* value_format, const char* key_format) {
  cudaStream_t st = g->stream;
  const uint64_t nrec_max = static_cast<uint64_t>(n_sources) * static_cast<uint64_t>(block_records);
  DevBuf d_off, records;
  PA_TRY(d_off.alloc(sizeof(uint64_t) * (n_sources + 1), st));
  PA_TRY(records.alloc(nrec_max * PA_PARTIAL_WORDS * 8, st));
  CUDA_TRY(cudaMemsetAsync(g->status.p, 0, sizeof(uint32_t) * ST_WORDS, st));
  k_merge_unpad<<<n_sources, 256, 0, st>>>(static_cast<const uint64_t*>(dev_blocks), static_cast<uint32_t>(n_sources),
                                           static_cast<uint64_t>(block_records), records.as<uint64_t>(), d_off.as<uint64_t>(),
                                           g->status.as<uint32_t>());
  CUDA_TRY(cudaGetLastError());
  return merge_build(g, records.p, d_off.as<uint64_t>(), n_sources, nrec_max, agg_mask, value_format, key_format);
}
}  // namespace
}  // extern "C++"

// ---- multi-GPU through the C ABI: communicator + the whole sharded step (SURVEY.md §8b "multi-GPU variants taking a
// communicator handle", §8e) ----
// ---- host frames larger than the device (or than 2^32-2 rows): aggregate chunk by chunk, merge the partials ----
// The row-range shards of the multi-GPU path, executed one after the other on ONE device: every chunk of the host
// columns is copied (staged, §h2d_copy), aggregated with row_base = its first row, and exported as partial records;
// the owner-side merge then joins the chunks' records exactly as it joins ranks' — sources folded in chunk (= row)
// order, groups ordered by global first row.  Device memory in use: one chunk of keys and values + the records.
int pa_groupby_aggregate_chunked(const struct ArrowDeviceArray* keys, const struct ArrowSchema* key_schema,
                                 const struct ArrowDeviceArray* values, const struct ArrowSchema* value_schema,
                                 uint32_t agg_mask, int64_t chunk_rows, const pa_options* opt, pa_groupby** merged_out) {
  if (!keys || !key_schema || !values || !value_schema || !merged_out) return set_err(PA_ERR_INVALID, "null argument");
  if (agg_mask == 0 || (agg_mask & ~PA_AGG_ALL)) return set_err(PA_ERR_NOT_IMPLEMENTED, "chunked aggregates: sum / mean / count / min / max / first / last");
  if (keys->device_type == ARROW_DEVICE_CUDA || values->device_type == ARROW_DEVICE_CUDA)
    return set_err(PA_ERR_INVALID, "chunked aggregation is for HOST columns (device columns already fit the device)");
  const int64_t n = keys->array.length;
  if (values->array.length != n) return set_err(PA_ERR_INVALID, "key and value columns differ in length");
  if (chunk_rows <= 0) chunk_rows = 1ll << 30;
  if (chunk_rows >= 0xFFFFFFFEll) return set_err(PA_ERR_INVALID, "chunk_rows must be below 2^32 - 2");
  const int64_t n_chunks = std::max<int64_t>(1, (n + chunk_rows - 1) / chunk_rows);
  if (n_chunks > 64) return set_err(PA_ERR_INVALID, "%lld chunks: at most 64 (raise chunk_rows)", (long long)n_chunks);
  pa_options o;
  if (opt) o = *opt; else pa_options_init(&o);
  const int64_t base0 = o.row_base;
  DevBuf records;                      // all chunks' records, chunk after chunk
  records.plain = true;
  std::vector<int64_t> counts(static_cast<size_t>(n_chunks), 0);
  uint64_t total = 0;
  for (int64_t c = 0; c < n_chunks; ++c) {
    const int64_t lo = c * chunk_rows, len = std::min(chunk_rows, n - lo);
    ArrowDeviceArray kc = *keys, vc = *values;           // views: same buffers, shifted window
    kc.array.offset += lo; kc.array.length = len;
    vc.array.offset += lo; vc.array.length = len;
    if (kc.array.null_count > 0) kc.array.null_count = -1;
    if (vc.array.null_count > 0) vc.array.null_count = -1;
    kc.array.release = nullptr; vc.array.release = nullptr;
    o.row_base = base0 + lo;
    pa_groupby* h = nullptr;
    PA_TRY(pa_groupby_create(&kc, key_schema, 1, &o, &h));
    HandlePtr guard(h);
    PA_TRY(pa_groupby_aggregate(h, &vc, value_schema, agg_mask));
    int64_t cnt = 0;
    PA_TRY(pa_groupby_partials_count(h, 1, &cnt));
    counts[static_cast<size_t>(c)] = cnt;
    // grow the record buffer (plain allocation: it outlives the per-chunk handles and their streams)
    const size_t need = (total + static_cast<uint64_t>(cnt)) * PA_PARTIAL_WORDS * 8;
    if (need > records.bytes) {
      DevBuf bigger;
      bigger.plain = true;
      PA_TRY(bigger.alloc(std::max<size_t>(need * 2, 1u << 20), nullptr));
      if (total) CUDA_TRY(cudaMemcpy(bigger.p, records.p, total * PA_PARTIAL_WORDS * 8, cudaMemcpyDeviceToDevice));
      records = std::move(bigger);
    }
    if (cnt) PA_TRY(pa_groupby_partials_export(h, 1, static_cast<char*>(records.p) + total * PA_PARTIAL_WORDS * 8, cnt));
    PA_TRY(pa_groupby_sync(h));
    total += static_cast<uint64_t>(cnt);
  }
  o.row_base = base0;
  const std::string kfmt = key_schema->format;
  PA_TRY(pa_merge_create(records.p, counts.data(), static_cast<int32_t>(n_chunks), agg_mask, value_schema->format, kfmt.c_str(), &o, merged_out));
  CUDA_TRY(cudaDeviceSynchronize());   // `records` is released on return
  return PA_OK;
}

struct pa_comm {
  ncclComm_t comm = nullptr;
  int world = 1, rank = 0, device = 0;
  bool own = false;
  DevBuf send, recv, d_counts, d_all, d_cursor;     // kept between steps (grow only)
  MergeScratch merge;
  double phase_ms[5] = {0, 0, 0, 0, 0};   // last step: local pass, count + export, exchange, merge, total
  cudaEvent_t ev[6] = {};
  uint64_t h_bound = 0;                   // end of this rank's row shard (staging for the async copy)
  int last_record_bytes = 0, last_unordered = 0;   // pa_comm_last_exchange
  int64_t last_table_slots = 0;
};

#define NCCL_TRY(expr)                                                                                          \
  do {                                                                                                          \
    ncclResult_t r__ = (expr);                                                                                  \
    if (r__ != ncclSuccess) return set_err(PA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, ncclGetErrorString(r__), __FILE__, __LINE__); \
  } while (0)

int pa_comm_unique_id(void* out_id, int64_t capacity_bytes) {
  if (!out_id || capacity_bytes < static_cast<int64_t>(sizeof(ncclUniqueId))) return set_err(PA_ERR_INVALID, "id buffer must hold %zu bytes", sizeof(ncclUniqueId));
  ncclUniqueId id;
  NCCL_TRY(ncclGetUniqueId(&id));
  memcpy(out_id, &id, sizeof id);
  return PA_OK;
}

static int comm_finish(pa_comm* c) {
  // The communicator outlives the handles (and their streams) it is used with: its scratch does not belong to any
  // stream's allocation order.
  for (DevBuf* b : {&c->send, &c->recv, &c->d_counts, &c->d_all, &c->d_cursor, &c->merge.tkeys, &c->merge.idx, &c->merge.m_first, &c->merge.m_slot,
                    &c->merge.s_first, &c->merge.s_slot, &c->merge.cub_tmp})
    b->plain = true;
  for (auto& e : c->ev) CUDA_TRY(cudaEventCreate(&e));
  return PA_OK;
}

int pa_comm_create(const void* id, int32_t world, int32_t rank, int32_t device, pa_comm** out) {
  if (!id || !out || world < 1 || world > 64 || rank < 0 || rank >= world) return set_err(PA_ERR_INVALID, "pa_comm_create: bad argument (1 <= world <= 64)");
  std::unique_ptr<pa_comm> c(new pa_comm());
  c->world = world; c->rank = rank; c->device = device; c->own = true;
  CUDA_TRY(cudaSetDevice(device));
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof uid);
  NCCL_TRY(ncclCommInitRank(&c->comm, world, uid, rank));
  PA_TRY(comm_finish(c.get()));
  *out = c.release();
  return PA_OK;
}

int pa_comm_adopt(void* nccl_comm, int32_t world, int32_t rank, int32_t device, pa_comm** out) {
  if (!nccl_comm || !out || world < 1 || world > 64 || rank < 0 || rank >= world) return set_err(PA_ERR_INVALID, "pa_comm_adopt: bad argument");
  std::unique_ptr<pa_comm> c(new pa_comm());
  c->comm = static_cast<ncclComm_t>(nccl_comm);
  c->world = world; c->rank = rank; c->device = device; c->own = false;
  CUDA_TRY(cudaSetDevice(device));
  PA_TRY(comm_finish(c.get()));
  *out = c.release();
  return PA_OK;
}

void pa_comm_destroy(pa_comm* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (auto& e : c->ev) if (e) cudaEventDestroy(e);
  if (c->own && c->comm) ncclCommDestroy(c->comm);
  delete c;
}

int pa_comm_last_exchange(pa_comm* c, int32_t* record_bytes, int32_t* unordered_export, int64_t* merge_table_slots) {
  if (!c) return set_err(PA_ERR_INVALID, "null argument");
  if (record_bytes) *record_bytes = c->last_record_bytes;
  if (unordered_export) *unordered_export = c->last_unordered;
  if (merge_table_slots) *merge_table_slots = c->last_table_slots;
  return PA_OK;
}

int pa_comm_last_phases(pa_comm* c, double phase_ms[5]) {
  if (!c || !phase_ms) return set_err(PA_ERR_INVALID, "null argument");
  for (int i = 0; i < 5; ++i) phase_ms[i] = c->phase_ms[i];
  return PA_OK;
}

// One multi-GPU step for this rank: local stages 1-3 on its row shard -> partial records bucketed by owner =
// hash(key) % world -> counts (ncclAllGather) and records (grouped ncclSend / ncclRecv: NCCL 2.27 has no all-to-all
// primitive) over NVLink -> owner-side merge in source-rank order.  Everything runs on the handle's stream.
int pa_groupby_sharded_aggregate(pa_groupby* g, pa_comm* c, const struct ArrowDeviceArray* values,
                                 const struct ArrowSchema* value_schema, uint32_t agg_mask, pa_groupby** merged_out) {
  if (!g || !c || !values || !value_schema || !merged_out) return set_err(PA_ERR_INVALID, "null argument");
  if (agg_mask == 0 || (agg_mask & ~PA_AGG_ALL)) return set_err(PA_ERR_NOT_IMPLEMENTED, "sharded aggregates: sum / mean / count / min / max / first / last");
  if (g->device != c->device) return set_err(PA_ERR_INVALID, "handle lives on device %d, communicator on %d", g->device, c->device);
  if (g->keys.size() != 1 && !g->resample) return set_err(PA_ERR_NOT_IMPLEMENTED, "multi-GPU merge of composite keys");
  if (!g->resample && g->keys[0].is_str) return set_err(PA_ERR_NOT_IMPLEMENTED, "multi-GPU merge of utf8 keys: dictionary-encode the column against a shared dictionary");
  PA_TRY(ensure_device(g));
  cudaStream_t st = g->stream;
  const int W = c->world;
  CUDA_TRY(cudaEventRecord(c->ev[0], st));
  // 1. local pass.  sum / mean of floats / count need one 32-byte sector per group (compact records, merge.cuh),
  // everything else the full record; with compact records the bucketed path may hand over its unordered group records
  // as they are (the merge orders by global first row anyway).
  int vw_ = 8, vc_ = VC_I;
  PA_TRY(parse_format(value_schema->format, &vw_, &vc_));
  const bool compact = !is_wide(agg_mask, vc_) && !(agg_mask & (AGG_FIRST | AGG_LAST));
  g->want_unordered = compact && !g->resample;
  const int local_rc = aggregate_entry(g, values, value_schema, agg_mask, false);
  g->want_unordered = false;
  if (local_rc != PA_OK) return local_rc;
  const bool unordered = g->unordered;
  g->unordered = false;
  const uint32_t G_local = unordered ? g->unordered_G : g->G;
  if (unordered && g->opt.expected_groups <= 0) g->opt.expected_groups = std::max<int64_t>(G_local, 1);   // (the keys of a handle never change)
  CUDA_TRY(cudaEventRecord(c->ev[1], st));
  // 2. counts per owner, all ranks' counts to everybody
  // (every rank also tells the others where its row shard ends: the merge orders by GLOBAL first row and only has
  // to sort the bits below the largest one)
  const int W1 = W + 1;
  PA_TRY(c->d_counts.alloc(sizeof(uint64_t) * W1, st));
  PA_TRY(c->d_all.alloc(sizeof(uint64_t) * W1 * W, st));
  CUDA_TRY(cudaMemsetAsync(c->d_counts.p, 0, sizeof(uint64_t) * W1, st));
  c->h_bound = static_cast<uint64_t>(g->opt.row_base) + static_cast<uint64_t>(g->n);
  CUDA_TRY(cudaMemcpyAsync(c->d_counts.as<uint64_t>() + W, &c->h_bound, sizeof(uint64_t), cudaMemcpyHostToDevice, st));
  if (unordered) {
    if (G_local) {
      k_bkrec_count<<<(G_local + 255) / 256, 256, 0, st>>>(static_cast<const BkRec32*>(g->unordered_recs), G_local, static_cast<uint32_t>(W), c->d_counts.as<unsigned long long>());
      CUDA_TRY(cudaGetLastError());
    }
  } else {
    PartialsArgs a{};
    a.r = g->res; a.G = g->G; a.nparts = W; a.counts = c->d_counts.as<unsigned long long>();
    if (g->G) {
      k_partials_count<<<(g->G + 255) / 256, 256, 0, st>>>(a);
      CUDA_TRY(cudaGetLastError());
    }
  }
  NCCL_TRY(ncclAllGather(c->d_counts.p, c->d_all.p, W1, ncclUint64, c->comm, st));
  std::vector<uint64_t> all(static_cast<size_t>(W1) * W);
  CUDA_TRY(cudaMemcpyAsync(all.data(), c->d_all.p, sizeof(uint64_t) * W1 * W, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  std::vector<int64_t> send_cnt(W), recv_cnt(W);
  uint64_t send_total = 0, recv_total = 0, row_bound = 1;
  for (int p = 0; p < W; ++p) {
    send_cnt[p] = static_cast<int64_t>(all[static_cast<size_t>(c->rank) * W1 + p]);
    recv_cnt[p] = static_cast<int64_t>(all[static_cast<size_t>(p) * W1 + c->rank]);
    send_total += send_cnt[p];
    recv_total += recv_cnt[p];
    row_bound = std::max(row_bound, all[static_cast<size_t>(p) * W1 + W]);
  }
  // 3. export the records grouped by owner
  const size_t RW = static_cast<size_t>(rec_words(compact));
  PA_TRY(c->send.alloc(std::max<uint64_t>(send_total, 1) * RW * 8, st));
  PA_TRY(c->recv.alloc(std::max<uint64_t>(recv_total, 1) * RW * 8, st));
  if (unordered) {
    std::vector<uint64_t> prefix(W, 0);
    for (int i = 1; i < W; ++i) prefix[i] = prefix[i - 1] + static_cast<uint64_t>(send_cnt[i - 1]);
    PA_TRY(c->d_cursor.alloc(sizeof(uint64_t) * W, st));
    CUDA_TRY(cudaMemcpyAsync(c->d_cursor.p, prefix.data(), sizeof(uint64_t) * W, cudaMemcpyHostToDevice, st));
    if (G_local) {
      k_bkrec_scatter<<<(G_local + 255) / 256, 256, 0, st>>>(static_cast<const BkRec32*>(g->unordered_recs), G_local, static_cast<uint32_t>(W),
                                                             g->opt.row_base, c->d_cursor.as<unsigned long long>(), c->send.as<uint64_t>());
      CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaStreamSynchronize(st));   // `prefix` dies here
  } else {
    g->parts_n = W;
    g->parts_counts.assign(send_cnt.begin(), send_cnt.end());
    PA_TRY(partials_export(g, W, c->send.p, static_cast<int64_t>(std::max<uint64_t>(send_total, g->G)), compact));
  }
  CUDA_TRY(cudaEventRecord(c->ev[2], st));
  // 4. all-to-all of the records
  NCCL_TRY(ncclGroupStart());
  {
    uint64_t so = 0, ro = 0;
    for (int p = 0; p < W; ++p) {
      if (send_cnt[p]) NCCL_TRY(ncclSend(c->send.as<uint64_t>() + so * RW, static_cast<size_t>(send_cnt[p]) * RW, ncclUint64, p, c->comm, st));
      if (recv_cnt[p]) NCCL_TRY(ncclRecv(c->recv.as<uint64_t>() + ro * RW, static_cast<size_t>(recv_cnt[p]) * RW, ncclUint64, p, c->comm, st));
      so += send_cnt[p];
      ro += recv_cnt[p];
    }
  }
  NCCL_TRY(ncclGroupEnd());
  CUDA_TRY(cudaEventRecord(c->ev[3], st));
  // 5. owner-side merge
  const std::string kfmt = g->resample ? g->index_format : g->keys[0].format;
  HandlePtr m(new pa_groupby());
  pa_options mo;
  pa_options_init(&mo);
  mo.device = g->device;
  mo.cuda_stream = st;
  PA_TRY(handle_init(m.get(), &mo));
  std::vector<uint64_t> off(W + 1, 0);
  for (int i = 0; i < W; ++i) off[i + 1] = off[i] + static_cast<uint64_t>(recv_cnt[i]);
  DevBuf d_off;
  PA_TRY(d_off.alloc(sizeof(uint64_t) * (W + 1), st));
  CUDA_TRY(cudaMemcpyAsync(d_off.p, off.data(), sizeof(uint64_t) * (W + 1), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemsetAsync(m->status.p, 0, sizeof(uint32_t) * ST_WORDS, st));
  CUDA_TRY(cudaEventRecord(m->ev[0], st));
  PA_TRY(merge_build(m.get(), c->recv.p, d_off.as<uint64_t>(), W, recv_total, agg_mask, value_schema->format, kfmt.c_str(), &c->merge, row_bound, compact,
                      static_cast<uint64_t>(*std::max_element(recv_cnt.begin(), recv_cnt.end()))));
  CUDA_TRY(cudaEventRecord(c->ev[4], st));
  CUDA_TRY(cudaEventSynchronize(c->ev[4]));
  float t = 0;
  for (int i = 0; i < 4; ++i) { CUDA_TRY(cudaEventElapsedTime(&t, c->ev[i], c->ev[i + 1])); c->phase_ms[i] = t; }
  CUDA_TRY(cudaEventElapsedTime(&t, c->ev[0], c->ev[4]));
  c->phase_ms[4] = t;
  c->last_record_bytes = static_cast<int>(RW * 8);
  c->last_unordered = unordered ? 1 : 0;
  c->last_table_slots = static_cast<int64_t>(c->merge.last_slots);
  *merged_out = m.release();
  return PA_OK;
}

// ---- synthetic generator ----
static int synth_grid(int64_t n) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, static_cast<int64_t>(sms) * 16)));
}
int pa_synth_keys_i64(void* dev_out, int64_t n, int64_t first_row, uint64_t n_groups, uint64_t seed, void* cuda_stream) {
  if (!dev_out || n_groups == 0) return set_err(PA_ERR_INVALID, "bad argument");
  k_synth_keys<<<synth_grid(n), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(static_cast<int64_t*>(dev_out), n, first_row, n_groups, seed);
  CUDA_TRY(cudaGetLastError());
  return PA_OK;
}
int pa_synth_vals_f64(void* dev_out, int64_t n, int64_t first_row, uint64_t seed, void* cuda_stream) {
  if (!dev_out) return set_err(PA_ERR_INVALID, "bad argument");
  k_synth_vals<<<synth_grid(n), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(static_cast<double*>(dev_out), n, first_row, seed);
  CUDA_TRY(cudaGetLastError());
  return PA_OK;
}
int pa_synth_validity(void* dev_bitmap_out, int64_t n, int64_t first_row, uint64_t seed, uint32_t null_every, void* cuda_stream) {
  if (!dev_bitmap_out || null_every == 0) return set_err(PA_ERR_INVALID, "bad argument");
  k_synth_validity<<<synth_grid((n + 7) / 8), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(static_cast<uint8_t*>(dev_bitmap_out), n, first_row, seed, null_every);
  CUDA_TRY(cudaGetLastError());
  return PA_OK;
}
int pa_synth_timestamps(void* dev_out, int64_t n, int64_t first_row, int64_t t0_ns, int64_t step_ns, uint64_t seed, void* cuda_stream) {
  if (!dev_out || step_ns <= 0) return set_err(PA_ERR_INVALID, "bad argument");
  k_synth_ts<<<synth_grid(n), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(static_cast<int64_t*>(dev_out), n, first_row, t0_ns, step_ns, seed);
  CUDA_TRY(cudaGetLastError());
  return PA_OK;
}

}  // extern "C"
