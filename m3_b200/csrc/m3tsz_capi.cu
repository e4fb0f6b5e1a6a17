// m3tsz_capi.cu -- the C ABI declared in include/m3tsz_b200.h.
//
// Thin host layer: argument validation, launch bookkeeping, and the *_host
// variants (H2D copy -> kernel -> D2H copy through context-owned scratch).
// There is deliberately no CPU codec in this library: without a CUDA device
// m3tsz_ctx_create fails with M3TSZ_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "m3tsz_common.cuh"
#include "m3tsz_kernels.h"

#include "m3tsz_ctx.h"

using namespace m3tsz;

using namespace m3tsz::host;

extern "C" {

int m3tsz_version(void) { return M3TSZ_B200_VERSION; }

const char *m3tsz_status_string(int st) {
  switch (st) {
    case M3TSZ_OK: return "ok";
    case M3TSZ_ERR_EOF: return "EOF";
    case M3TSZ_ERR_ENCODER_CLOSED: return "encoder is closed";
    case M3TSZ_ERR_NO_DATAPOINTS: return "encoder has no encoded datapoints";
    case M3TSZ_ERR_DOD_OVERFLOW: return "deltaOfDelta value overflows 32 bits";
    case M3TSZ_ERR_NO_TIME_SCHEME: return "time encoding scheme doesn't exist for unit";
    case M3TSZ_ERR_UNRECOGNIZED_UNIT: return "unrecognized time unit";
    case M3TSZ_ERR_INVALID_MULT: return "supplied multiplier is invalid";
    case M3TSZ_ERR_ANNOTATION_LEN: return "expected annotation length to be >= 0";
    case M3TSZ_ERR_ANNOTATION_SHORT: return "expected to read annotation bytes, but got end of stream";
    case M3TSZ_ERR_ITER_CLOSED: return "iterator is closed";
    case M3TSZ_ERR_VARINT_OVERFLOW: return "binary: varint overflows a 64-bit integer";
    case M3TSZ_ERR_UNEXPECTED_EOF: return "unexpected EOF";
    case M3TSZ_ERR_OUT_OF_ORDER: return "values are out of order from inner iterator";
    case M3TSZ_ERR_TOO_MANY_ITERATORS: return "too many replicas / readers for one series";
    case M3TSZ_ERR_CHECKSUM_MISMATCH: return "checksum does not match expected checksum";
    case M3TSZ_ERR_CAPACITY: return "output capacity exceeded";
    case M3TSZ_ERR_INVALID_ARG: return "invalid argument";
    case M3TSZ_ERR_CUDA: return "CUDA error";
    case M3TSZ_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU fallback)";
    case M3TSZ_ERR_STREAM_TOO_LARGE: return "stream too large";
    default: return "unknown status";
  }
}

int m3tsz_ctx_create(int device, m3tsz_ctx **out) {
  if (!out) return M3TSZ_ERR_INVALID_ARG;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    (void)cudaGetLastError();
    return M3TSZ_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= n) return M3TSZ_ERR_INVALID_ARG;
  m3tsz_ctx *ctx = new (std::nothrow) m3tsz_ctx();
  if (!ctx) return M3TSZ_ERR_INVALID_ARG;
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev, cudaEventDisableTiming) != cudaSuccess ||
      cudaMalloc(&ctx->d_flag, 2 * sizeof(int32_t)) != cudaSuccess) {
    (void)cudaGetLastError();
    delete ctx;
    return M3TSZ_ERR_CUDA;
  }
  *out = ctx;
  return M3TSZ_OK;
}

void m3tsz_ctx_destroy(m3tsz_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  for (auto &sc : ctx->s)
    if (sc.ptr) cudaFree(sc.ptr);
  if (ctx->d_flag) cudaFree(ctx->d_flag);
  for (int i = 0; i < 2; i++)
    if (ctx->h_stage[i]) cudaFreeHost(ctx->h_stage[i]);
  if (ctx->ev) cudaEventDestroy(ctx->ev);
  if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char *m3tsz_last_cuda_error(const m3tsz_ctx *ctx) { return ctx ? ctx->last_error : ""; }
uint64_t m3tsz_ctx_launch_count(const m3tsz_ctx *ctx) { return ctx ? ctx->launches : 0; }

uint64_t m3tsz_encode_bound_units(uint64_t n, int per_datapoint_units) {
  // 64-bit start + per datapoint <= 68 (timestamp: '1111' + 64-bit delta-of-delta) + 80 (value:
  // 3 control bits + 13-bit int header + 64 bits) bits + 11-bit end-of-stream marker, plus the
  // write-ahead guard the kernel keeps.  With per-datapoint units a datapoint may also carry a
  // time-unit marker (11 + 8 bits) and then a raw 64-bit delta-of-delta: 83 bits of timestamp.
  const uint64_t per_dp = per_datapoint_units ? (83 + 80) : (68 + 80);
  uint64_t bits = 64 + n * per_dp + 11;
  uint64_t bytes = (bits + 7) / 8 + 64;
  return (bytes + 15) & ~15ull;
}
uint64_t m3tsz_encode_bound(uint64_t n) { return m3tsz_encode_bound_units(n, 0); }

// --------------------------------------------------------------------------
int m3tsz_decode_batch_ex(m3tsz_ctx *ctx, const m3tsz_options *opts, const uint8_t *d_streams,
                          uint64_t streams_bytes, const uint64_t *d_offsets, uint64_t n_series,
                          int64_t *d_ts, double *d_val, uint64_t max_points, uint32_t *d_n_points,
                          int32_t *d_status, uint8_t *d_unit, m3tsz_annotation_ref *d_ann,
                          const m3tsz_decode_extras *ex, void *stream) {
  if (!ctx || !valid_opts(opts)) return M3TSZ_ERR_INVALID_ARG;
  if (n_series == 0) return M3TSZ_OK;
  if (!d_streams || !d_offsets || !d_ts || !d_val || max_points == 0 ||
      max_points >= (1ull << 27) || ((uintptr_t)d_streams & 15u) ||
      streams_bytes >= (1ull << 34))  // 32-bit word / byte offsets inside the kernel; split larger batches
    return M3TSZ_ERR_INVALID_ARG;
  if (ex && ex->d_events && (!ex->d_event_count || ex->events_capacity == 0)) return M3TSZ_ERR_INVALID_ARG;
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return set_cuda_error(ctx, guard.err, "cudaSetDevice");
  DecodeParams p;
  memset(&p, 0, sizeof(p));
  p.streams = d_streams;
  p.streams_bytes = streams_bytes;
  p.offsets = d_offsets;
  p.n_series = n_series;
  p.default_unit = opts->default_time_unit;
  p.ts = d_ts;
  p.val = d_val;
  p.cap = max_points;
  p.n_points = d_n_points;
  p.status = d_status;
  p.unit_out = d_unit;
  p.ann_out = d_ann;
  if (ex) {
    p.lengths = ex->d_lengths;
    p.unit_first_out = ex->d_unit_first;
    if (ex->d_events) {
      p.events = ex->d_events;
      p.events_capacity = ex->events_capacity;
      p.event_count = reinterpret_cast<unsigned long long *>(ex->d_event_count);
    }
  }
  CK(launch_decode(p, opts->int_optimized != 0, (ex && ex->point_major) ? 3 : 0, (cudaStream_t)stream));
  ctx->launches++;
  return M3TSZ_OK;
}

int m3tsz_decode_batch(m3tsz_ctx *ctx, const m3tsz_options *opts, const uint8_t *d_streams,
                       uint64_t streams_bytes, const uint64_t *d_offsets, uint64_t n_series,
                       int64_t *d_ts, double *d_val, uint64_t max_points, uint32_t *d_n_points,
                       int32_t *d_status, uint8_t *d_unit, m3tsz_annotation_ref *d_ann,
                       void *stream) {
  return m3tsz_decode_batch_ex(ctx, opts, d_streams, streams_bytes, d_offsets, n_series, d_ts, d_val,
                               max_points, d_n_points, d_status, d_unit, d_ann, nullptr, stream);
}

static int downsample_impl(m3tsz_ctx *ctx, const m3tsz_options *opts, const uint8_t *d_streams,
                           uint64_t streams_bytes, const uint64_t *d_offsets, const uint64_t *d_lengths,
                           uint64_t n_series,
                           int64_t range_start_ns, int64_t window_ns, uint32_t n_windows, double *d_sum,
                           int64_t *d_count, double *d_min, double *d_max, double *d_last,
                           int64_t *d_last_at, bool want_last, uint32_t *d_n_points, int32_t *d_status,
                           void *stream) {
  if (!ctx || !valid_opts(opts)) return M3TSZ_ERR_INVALID_ARG;
  if (n_series == 0) return M3TSZ_OK;
  if (!d_streams || !d_offsets || !d_sum || !d_count || !d_min || !d_max || window_ns <= 0 ||
      n_windows == 0 || n_windows > 0x7fffffffu || ((uintptr_t)d_streams & 15u) ||
      streams_bytes >= (1ull << 34) || (want_last && !d_last))
    return M3TSZ_ERR_INVALID_ARG;
  // range_start + n_windows*window must not overflow int64
  if ((__int128)range_start_ns + (__int128)n_windows * (__int128)window_ns > (__int128)INT64_MAX)
    return M3TSZ_ERR_INVALID_ARG;
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return set_cuda_error(ctx, guard.err, "cudaSetDevice");
  if (want_last && !d_last_at) {  // lastAt lives in context scratch
    void *sc = nullptr;
    int rc = ensure(ctx, 12, (size_t)n_series * n_windows * 8, &sc);
    if (rc) return rc;
    d_last_at = (int64_t *)sc;
  }
  DecodeParams p;
  memset(&p, 0, sizeof(p));
  p.streams = d_streams;
  p.streams_bytes = streams_bytes;
  p.offsets = d_offsets;
  p.lengths = d_lengths;
  p.n_series = n_series;
  p.default_unit = opts->default_time_unit;
  p.range_start = range_start_ns;
  p.window = window_ns;
  p.n_windows = n_windows;
  p.ds_sum = d_sum;
  p.ds_count = d_count;
  p.ds_min = d_min;
  p.ds_max = d_max;
  p.ds_last = d_last;
  p.ds_last_at = d_last_at;
  p.n_points = d_n_points;
  p.status = d_status;
  CK(launch_decode(p, opts->int_optimized != 0, want_last ? 2 : 1, (cudaStream_t)stream));
  ctx->launches++;
  return M3TSZ_OK;
}

int m3tsz_decode_downsample_batch(m3tsz_ctx *ctx, const m3tsz_options *opts,
                                  const uint8_t *d_streams, uint64_t streams_bytes,
                                  const uint64_t *d_offsets, uint64_t n_series,
                                  int64_t range_start_ns, int64_t window_ns, uint32_t n_windows,
                                  double *d_sum, int64_t *d_count, double *d_min, double *d_max,
                                  uint32_t *d_n_points, int32_t *d_status, void *stream) {
  return downsample_impl(ctx, opts, d_streams, streams_bytes, d_offsets, nullptr, n_series, range_start_ns,
                         window_ns, n_windows, d_sum, d_count, d_min, d_max, nullptr, nullptr, false,
                         d_n_points, d_status, stream);
}

int m3tsz_decode_downsample_last_batch(m3tsz_ctx *ctx, const m3tsz_options *opts,
                                       const uint8_t *d_streams, uint64_t streams_bytes,
                                       const uint64_t *d_offsets, uint64_t n_series,
                                       int64_t range_start_ns, int64_t window_ns, uint32_t n_windows,
                                       double *d_sum, int64_t *d_count, double *d_min, double *d_max,
                                       double *d_last, int64_t *d_last_at, uint32_t *d_n_points,
                                       int32_t *d_status, void *stream) {
  return downsample_impl(ctx, opts, d_streams, streams_bytes, d_offsets, nullptr, n_series, range_start_ns,
                         window_ns, n_windows, d_sum, d_count, d_min, d_max, d_last, d_last_at, true,
                         d_n_points, d_status, stream);
}

int m3tsz_encode_batch_ex(m3tsz_ctx *ctx, const m3tsz_options *opts, const int64_t *d_ts,
                          const double *d_val, uint64_t n_series, uint64_t points_stride,
                          const uint32_t *d_n_points, const int64_t *d_start, int32_t unit,
                          const uint8_t *d_units, const uint64_t *d_ann_series_off,
                          const m3tsz_annotation_entry *d_ann_entries, const uint8_t *d_ann_bytes,
                          uint8_t *d_out, uint64_t out_stride, uint64_t *d_out_len, int32_t *d_status,
                          const m3tsz_encode_extras *extras, void *stream) {
  if (!ctx || !valid_opts(opts)) return M3TSZ_ERR_INVALID_ARG;
  if (n_series == 0) return M3TSZ_OK;
  if (!d_ts || !d_val || !d_start || !d_out || !d_out_len || points_stride > 0xffffffffull ||
      (out_stride & 15u) || out_stride == 0 || out_stride > (1ull << 33) ||
      ((uintptr_t)d_out & 15u) || ((uintptr_t)d_ts & 7u) || ((uintptr_t)d_val & 7u))
    return M3TSZ_ERR_INVALID_ARG;
  if (d_ann_series_off && (!d_ann_entries || !d_ann_bytes)) return M3TSZ_ERR_INVALID_ARG;
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return set_cuda_error(ctx, guard.err, "cudaSetDevice");
  EncodeParams p;
  memset(&p, 0, sizeof(p));
  p.ts = d_ts;
  p.val = d_val;
  p.n_series = n_series;
  p.points_stride = points_stride;
  p.n_points = d_n_points;
  p.start = d_start;
  p.unit = unit;
  p.units = d_units;
  p.ann_series_off = d_ann_series_off;
  p.ann_entries = d_ann_entries;
  p.ann_bytes = d_ann_bytes;
  p.default_unit = opts->default_time_unit;
  p.out = d_out;
  p.out_stride = out_stride;
  p.out_len = d_out_len;
  p.status = d_status;
  if (extras) {
    p.last_value = extras->d_last_value;
    p.out_bits = extras->d_out_bits;
    if (extras->point_major_input) {
      if (d_units || d_ann_series_off) return M3TSZ_ERR_INVALID_ARG;
      p.in_mode = 1;
    }
  }
  CK(launch_encode(p, opts->int_optimized != 0, (cudaStream_t)stream));
  ctx->launches++;
  return M3TSZ_OK;
}

int m3tsz_encode_batch(m3tsz_ctx *ctx, const m3tsz_options *opts, const int64_t *d_ts,
                       const double *d_val, uint64_t n_series, uint64_t points_stride,
                       const uint32_t *d_n_points, const int64_t *d_start, int32_t unit,
                       const uint8_t *d_units, const uint64_t *d_ann_series_off,
                       const m3tsz_annotation_entry *d_ann_entries, const uint8_t *d_ann_bytes,
                       uint8_t *d_out, uint64_t out_stride, uint64_t *d_out_len, int32_t *d_status,
                       void *stream) {
  return m3tsz_encode_batch_ex(ctx, opts, d_ts, d_val, n_series, points_stride, d_n_points, d_start, unit,
                               d_units, d_ann_series_off, d_ann_entries, d_ann_bytes, d_out, out_stride,
                               d_out_len, d_status, nullptr, stream);
}

int m3tsz_encode_batch_packed(m3tsz_ctx *ctx, const m3tsz_options *opts, const int64_t *d_ts,
                              const double *d_val, uint64_t n_series, uint64_t points_stride,
                              const uint32_t *d_n_points, const int64_t *d_start, int32_t unit,
                              const uint8_t *d_units, const uint64_t *d_ann_series_off,
                              const m3tsz_annotation_entry *d_ann_entries, const uint8_t *d_ann_bytes,
                              uint64_t slot_bytes, uint32_t align, uint8_t *d_packed,
                              uint64_t packed_capacity, uint64_t *d_offsets, uint64_t *d_out_len,
                              int32_t *d_status, uint64_t *d_total_bytes, void *stream) {
  return m3tsz_encode_batch_packed_ex(ctx, opts, d_ts, d_val, n_series, points_stride, d_n_points, d_start, unit,
                                      d_units, d_ann_series_off, d_ann_entries, d_ann_bytes, slot_bytes, align,
                                      d_packed, packed_capacity, d_offsets, d_out_len, d_status, d_total_bytes,
                                      nullptr, stream);
}

int m3tsz_encode_batch_packed_ex(m3tsz_ctx *ctx, const m3tsz_options *opts, const int64_t *d_ts,
                                 const double *d_val, uint64_t n_series, uint64_t points_stride,
                                 const uint32_t *d_n_points, const int64_t *d_start, int32_t unit,
                                 const uint8_t *d_units, const uint64_t *d_ann_series_off,
                                 const m3tsz_annotation_entry *d_ann_entries, const uint8_t *d_ann_bytes,
                                 uint64_t slot_bytes, uint32_t align, uint8_t *d_packed,
                                 uint64_t packed_capacity, uint64_t *d_offsets, uint64_t *d_out_len,
                                 int32_t *d_status, uint64_t *d_total_bytes, const m3tsz_encode_extras *extras,
                                 void *stream) {
  if (!ctx || !valid_opts(opts) || !d_total_bytes) return M3TSZ_ERR_INVALID_ARG;
  if (!(align == 1 || align == 4 || align == 8 || align == 16 || align == 32 || align == 64))
    return M3TSZ_ERR_INVALID_ARG;
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return set_cuda_error(ctx, guard.err, "cudaSetDevice");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaMemsetAsync(d_total_bytes, 0, sizeof(uint64_t), st));
  if (n_series == 0) return M3TSZ_OK;
  if (!d_ts || !d_val || !d_start || !d_packed || !d_offsets || !d_out_len ||
      points_stride > 0xffffffffull || ((uintptr_t)d_packed & 63u) || ((uintptr_t)d_ts & 7u) ||
      ((uintptr_t)d_val & 7u))
    return M3TSZ_ERR_INVALID_ARG;
  if (d_ann_series_off && (!d_ann_entries || !d_ann_bytes)) return M3TSZ_ERR_INVALID_ARG;
  if (slot_bytes == 0) slot_bytes = m3tsz_encode_bound_units(points_stride, d_units != nullptr);
  slot_bytes = (slot_bytes + 15) & ~15ull;
  if (slot_bytes > (1ull << 33)) return M3TSZ_ERR_INVALID_ARG;
  const uint64_t slots = encode_packed_scratch_slots(n_series);
  if (slots == 0) return set_cuda_error(ctx, cudaGetLastError(), "encode_packed_scratch_slots");
  void *scratch = nullptr, *ctr = nullptr;
  int rc;
  if ((rc = ensure(ctx, 36, slots * slot_bytes, &scratch))) return rc;
  if ((rc = ensure(ctx, 37, 64, &ctr))) return rc;
  CK(cudaMemsetAsync(ctr, 0, 64, st));
  EncodeParams p;
  memset(&p, 0, sizeof(p));
  p.ts = d_ts;
  p.val = d_val;
  p.n_series = n_series;
  p.points_stride = points_stride;
  p.n_points = d_n_points;
  p.start = d_start;
  p.unit = unit;
  p.units = d_units;
  p.ann_series_off = d_ann_series_off;
  p.ann_entries = d_ann_entries;
  p.ann_bytes = d_ann_bytes;
  p.default_unit = opts->default_time_unit;
  p.out = (uint8_t *)scratch;
  p.out_stride = slot_bytes;
  p.out_len = d_out_len;
  p.status = d_status;
  p.packed = d_packed;
  p.packed_capacity = packed_capacity;
  p.packed_off = d_offsets;
  p.packed_cursor = reinterpret_cast<unsigned long long *>(d_total_bytes);
  p.batch_counter = reinterpret_cast<unsigned long long *>(ctr);
  p.align = align;
  if (extras) {
    p.last_value = extras->d_last_value;
    p.out_bits = extras->d_out_bits;
    if (extras->point_major_input) {
      if (d_units || d_ann_series_off) return M3TSZ_ERR_INVALID_ARG;
      p.in_mode = 1;
    }
  }
  {
    // optional start-up phase spread of the persistent warps (tuning knob; measured on a B200 at
    // 1M x 1440: 0 / 250 / 500 / 800 ns per datapoint -> 13.0 / 13.2 / 13.4 / 13.6 ms, so it is off:
    // the copy phase costs warp residency, not DRAM contention -- profiles/r02_decode_history.md)
    static const long ns_per_dp = [] {
      const char *e = getenv("M3TSZ_ENC_STAGGER_NS_PER_DP");
      return e ? atol(e) : 0L;
    }();
    const uint64_t n_batches = (n_series + 31) / 32;
    if (ns_per_dp > 0 && n_batches >= 3 * (slots / 32)) p.stagger_ns = points_stride * (uint64_t)ns_per_dp;
  }
  CK(launch_encode(p, opts->int_optimized != 0, st));
  ctx->launches++;
  return M3TSZ_OK;
}

int m3tsz_compact_streams(m3tsz_ctx *ctx, const uint8_t *d_slots, uint64_t slot_stride,
                          const uint64_t *d_len, uint64_t n_series, uint32_t align,
                          uint8_t *d_packed, uint64_t packed_capacity, uint64_t *d_offsets,
                          void *stream) {
  if (!ctx || !d_offsets || (n_series && (!d_slots || !d_len || !d_packed))) return M3TSZ_ERR_INVALID_ARG;
  if (!(align == 1 || align == 4 || align == 8 || align == 16 || align == 32 || align == 64)) return M3TSZ_ERR_INVALID_ARG;
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return set_cuda_error(ctx, guard.err, "cudaSetDevice");
  cudaStream_t st = (cudaStream_t)stream;
  size_t tmp_bytes = compact_scan_tmp_bytes(n_series);
  void *tmp = nullptr;
  int rc = ensure(ctx, 15, tmp_bytes, &tmp);
  if (rc) return rc;
  CK(cudaMemsetAsync(ctx->d_flag, 0, sizeof(int32_t), st));
  CK(launch_compact(d_slots, slot_stride, d_len, n_series, align, d_packed, packed_capacity,
                    d_offsets, tmp, tmp_bytes, ctx->d_flag, st));
  ctx->launches += 3;
  int32_t flag = 0;
  CK(cudaMemcpyAsync(&flag, ctx->d_flag, sizeof(flag), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return flag ? M3TSZ_ERR_CAPACITY : M3TSZ_OK;
}

int m3tsz_prom_convert_batch(m3tsz_ctx *ctx, const int64_t *d_ts, const double *d_val, uint64_t cap,
                             const uint32_t *d_n_points, uint64_t n_series, int64_t resolution_ns,
                             const uint8_t *d_handle_resets, double value_decrease_tolerance,
                             int64_t tolerance_until_ns, int64_t *d_ts_ms_out, double *d_val_out,
                             uint64_t out_cap, uint32_t *d_n_out, int32_t *d_status, void *stream) {
  if (!ctx) return M3TSZ_ERR_INVALID_ARG;
  if (n_series == 0) return M3TSZ_OK;
  if (!d_ts || !d_val || !d_n_points || !d_ts_ms_out || !d_val_out || !d_n_out || cap == 0 || out_cap == 0 ||
      resolution_ns < 0 || (d_handle_resets && resolution_ns == 0) ||
      value_decrease_tolerance != value_decrease_tolerance)
    return M3TSZ_ERR_INVALID_ARG;
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return set_cuda_error(ctx, guard.err, "cudaSetDevice");
  PromParams p;
  memset(&p, 0, sizeof(p));
  p.ts = d_ts;
  p.val = d_val;
  p.cap = cap;
  p.n_points = d_n_points;
  p.n_series = n_series;
  p.resolution = resolution_ns;
  p.handle_resets = d_handle_resets;
  p.tolerance = value_decrease_tolerance;
  p.tolerance_until = tolerance_until_ns;
  p.ts_out = d_ts_ms_out;
  p.val_out = d_val_out;
  p.out_cap = out_cap;
  p.n_out = d_n_out;
  p.status = d_status;
  CK(launch_prom(p, (cudaStream_t)stream));
  ctx->launches++;
  return M3TSZ_OK;
}

int m3tsz_aggregate_tiles_batch(m3tsz_ctx *ctx, const m3tsz_options *opts, const uint8_t *d_streams,
                                uint64_t streams_bytes, const uint64_t *d_offsets,
                                const uint64_t *d_lengths, uint64_t n_series, int64_t start_ns,
                                int64_t step_ns, uint32_t n_windows, int32_t agg_type, int32_t out_unit,
                                uint32_t align, uint8_t *d_packed, uint64_t packed_capacity,
                                uint64_t *d_out_offsets, uint64_t *d_out_len, int32_t *d_status,
                                uint32_t *d_n_tiles, uint64_t *d_total_bytes, void *stream) {
  if (!ctx || !valid_opts(opts) || !d_total_bytes || !d_status) return M3TSZ_ERR_INVALID_ARG;
  if (!(agg_type == M3TSZ_AGG_LAST || agg_type == M3TSZ_AGG_MIN || agg_type == M3TSZ_AGG_MAX ||
        agg_type == M3TSZ_AGG_MEAN || agg_type == M3TSZ_AGG_COUNT || agg_type == M3TSZ_AGG_SUM))
    return M3TSZ_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (n_series == 0) {
    DeviceGuard guard(ctx->device);
    CK(cudaMemsetAsync(d_total_bytes, 0, sizeof(uint64_t), st));
    return M3TSZ_OK;
  }
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return set_cuda_error(ctx, guard.err, "cudaSetDevice");
  // context scratch: window-major aggregates, then series-major tiles
  const size_t wb = (size_t)n_series * n_windows * 8;
  void *d_sum, *d_cnt, *d_min, *d_max, *d_last, *d_lat, *d_sst, *d_np, *d_es = nullptr;
  int rc;
  if ((rc = ensure(ctx, 8, wb, &d_sum))) return rc;
  if ((rc = ensure(ctx, 9, wb, &d_cnt))) return rc;
  if ((rc = ensure(ctx, 10, wb, &d_min))) return rc;
  if ((rc = ensure(ctx, 11, wb, &d_max))) return rc;
  if ((rc = ensure(ctx, 12, wb, &d_lat))) return rc;
  if ((rc = ensure(ctx, 38, wb, &d_last))) return rc;
  if ((rc = ensure(ctx, 4, n_series * 4, &d_np))) return rc;
  if ((rc = ensure(ctx, 5, n_series * 4, &d_sst))) return rc;
  // 1. fused decode + Gauge windows (with `last`)
  rc = downsample_impl(ctx, opts, d_streams, streams_bytes, d_offsets, d_lengths, n_series, start_ns, step_ns,
                       n_windows, (double *)d_sum, (int64_t *)d_cnt, (double *)d_min, (double *)d_max,
                       (double *)d_last, (int64_t *)d_lat, true, (uint32_t *)d_np, (int32_t *)d_sst, st);
  if (rc) return rc;
  // 2. re-encode straight from the window-major aggregates (the encoder's input stage skips empty
  //    windows, stamps the window end and takes Gauge.ValueOf(agg_type)) into the packed buffer
  {
    if (!(align == 1 || align == 4 || align == 8 || align == 16 || align == 32 || align == 64))
      return M3TSZ_ERR_INVALID_ARG;
    if (!d_packed || !d_out_offsets || !d_out_len || ((uintptr_t)d_packed & 63u)) return M3TSZ_ERR_INVALID_ARG;
    CK(cudaMemsetAsync(d_total_bytes, 0, sizeof(uint64_t), st));
    const uint64_t slot_bytes = m3tsz_encode_bound(n_windows);
    const uint64_t slots = encode_packed_scratch_slots(n_series);
    if (slots == 0) return set_cuda_error(ctx, cudaGetLastError(), "encode_packed_scratch_slots");
    void *scratch = nullptr, *ctr = nullptr;
    if ((rc = ensure(ctx, 36, slots * slot_bytes, &scratch))) return rc;
    if ((rc = ensure(ctx, 37, 64, &ctr))) return rc;
    CK(cudaMemsetAsync(ctr, 0, 64, st));
    // every series' encoder starts at the tile range's start (the target block's start)
    if ((rc = ensure(ctx, 39, n_series * 8, &d_es))) return rc;
    CK(launch_fill_i64((int64_t *)d_es, start_ns, n_series, st));
    EncodeParams p;
    memset(&p, 0, sizeof(p));
    p.n_series = n_series;
    p.points_stride = n_windows;
    p.start = (const int64_t *)d_es;
    p.unit = out_unit;
    p.default_unit = opts->default_time_unit;
    p.out = (uint8_t *)scratch;
    p.out_stride = slot_bytes;
    p.out_len = d_out_len;
    p.status = d_status;
    p.packed = d_packed;
    p.packed_capacity = packed_capacity;
    p.packed_off = d_out_offsets;
    p.packed_cursor = reinterpret_cast<unsigned long long *>(d_total_bytes);
    p.batch_counter = reinterpret_cast<unsigned long long *>(ctr);
    p.align = align;
    p.in_mode = 2;
    p.tile_sum = (const double *)d_sum;
    p.tile_count = (const int64_t *)d_cnt;
    p.tile_min = (const double *)d_min;
    p.tile_max = (const double *)d_max;
    p.tile_last = (const double *)d_last;
    p.tile_src_status = (const int32_t *)d_sst;
    p.tile_start = start_ns;
    p.tile_step = step_ns;
    p.agg_type = agg_type;
    p.n_tiles_out = d_n_tiles;
    CK(launch_encode(p, opts->int_optimized != 0, st));
    ctx->launches += 2;
  }
  CK(launch_tiles_status((const int32_t *)d_sst, d_status, n_series, st));
  ctx->launches++;
  return M3TSZ_OK;
}

int m3tsz_merge_series_batch_ex(m3tsz_ctx *ctx, const int64_t *d_ts, const double *d_val, uint64_t cap,
                                const uint32_t *d_n_points, const int32_t *d_seq_status,
                                const uint64_t *d_slice_off, const uint64_t *d_replica_off,
                                const uint64_t *d_series_off, uint64_t n_series, int64_t start_ns,
                                int64_t end_ns, int32_t strategy, int64_t *d_ts_out, double *d_val_out,
                                uint64_t out_cap, uint32_t *d_n_out, int32_t *d_status, uint64_t n_seq,
                                int32_t point_major_in, int32_t point_major_out, void *stream) {
  if (!ctx) return M3TSZ_ERR_INVALID_ARG;
  if (n_series == 0) return M3TSZ_OK;
  if (!d_ts || !d_val || !d_n_points || !d_slice_off || !d_replica_off || !d_series_off || !d_ts_out ||
      !d_val_out || !d_n_out || !d_status || cap == 0 || out_cap == 0 || out_cap > 0xffffffffull ||
      strategy < 0 || strategy > 3)
    return M3TSZ_ERR_INVALID_ARG;
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return set_cuda_error(ctx, guard.err, "cudaSetDevice");
  MergeParams p;
  memset(&p, 0, sizeof(p));
  p.ts = d_ts;
  p.val = d_val;
  p.cap = cap;
  p.n_points = d_n_points;
  p.seq_status = d_seq_status;
  p.slice_off = d_slice_off;
  p.replica_off = d_replica_off;
  p.series_off = d_series_off;
  p.n_series = n_series;
  p.start = start_ns;
  p.end = end_ns;
  p.strategy = strategy;
  p.ts_out = d_ts_out;
  p.val_out = d_val_out;
  p.out_cap = out_cap;
  p.n_out = d_n_out;
  p.status = d_status;
  if (point_major_in && n_seq == 0) return M3TSZ_ERR_INVALID_ARG;
  p.in_seq_stride = point_major_in ? 1 : cap;
  p.in_pt_stride = point_major_in ? n_seq : 1;
  p.out_seq_stride = point_major_out ? 1 : out_cap;
  p.out_pt_stride = point_major_out ? n_series : 1;
  CK(launch_merge(p, (cudaStream_t)stream));
  ctx->launches++;
  return M3TSZ_OK;
}

int m3tsz_merge_series_batch(m3tsz_ctx *ctx, const int64_t *d_ts, const double *d_val, uint64_t cap,
                             const uint32_t *d_n_points, const int32_t *d_seq_status,
                             const uint64_t *d_slice_off, const uint64_t *d_replica_off,
                             const uint64_t *d_series_off, uint64_t n_series, int64_t start_ns,
                             int64_t end_ns, int32_t strategy, int64_t *d_ts_out, double *d_val_out,
                             uint64_t out_cap, uint32_t *d_n_out, int32_t *d_status, void *stream) {
  return m3tsz_merge_series_batch_ex(ctx, d_ts, d_val, cap, d_n_points, d_seq_status, d_slice_off, d_replica_off,
                                     d_series_off, n_series, start_ns, end_ns, strategy, d_ts_out, d_val_out,
                                     out_cap, d_n_out, d_status, 0, 0, 0, stream);
}

int m3tsz_checksum_batch(m3tsz_ctx *ctx, const uint8_t *d_streams, uint64_t streams_bytes,
                         const uint64_t *d_offsets, const uint64_t *d_lengths, uint64_t n_series,
                         const uint32_t *d_expected, uint32_t *d_checksums, int32_t *d_status,
                         void *stream) {
  if (!ctx) return M3TSZ_ERR_INVALID_ARG;
  if (n_series == 0) return M3TSZ_OK;
  if (!d_streams || !d_offsets || (!d_checksums && !d_status) || (d_expected && !d_status))
    return M3TSZ_ERR_INVALID_ARG;
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return set_cuda_error(ctx, guard.err, "cudaSetDevice");
  ChecksumParams p;
  memset(&p, 0, sizeof(p));
  p.streams = d_streams;
  p.streams_bytes = streams_bytes;
  p.offsets = d_offsets;
  p.lengths = d_lengths;
  p.n_series = n_series;
  p.expected = d_expected;
  p.out = d_checksums;
  p.status = d_status;
  CK(launch_checksum(p, (cudaStream_t)stream));
  ctx->launches++;
  return M3TSZ_OK;
}

// --------------------------------------------------------------------------
// host-buffer variants
// --------------------------------------------------------------------------
static int m3tsz_decode_batch_host_impl(m3tsz_ctx *ctx, const m3tsz_options *opts, const uint8_t *h_streams,
                            uint64_t streams_bytes, const uint64_t *h_offsets, uint64_t n_series,
                            int64_t *h_ts, double *h_val, uint64_t max_points,
                            uint32_t *h_n_points, int32_t *h_status, uint8_t *h_unit,
                            m3tsz_annotation_ref *h_ann) {
  if (!ctx || !valid_opts(opts)) return M3TSZ_ERR_INVALID_ARG;
  if (n_series == 0) return M3TSZ_OK;
  if (!h_streams || !h_offsets || !h_ts || !h_val || max_points == 0) return M3TSZ_ERR_INVALID_ARG;
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return set_cuda_error(ctx, guard.err, "cudaSetDevice");
  void *d_streams, *d_off, *d_ts, *d_val, *d_n, *d_st, *d_unit = nullptr, *d_ann = nullptr;
  int rc;
  if ((rc = ensure(ctx, 0, streams_bytes + 16, &d_streams))) return rc;
  if ((rc = ensure(ctx, 1, (n_series + 1) * 8, &d_off))) return rc;
  if ((rc = ensure(ctx, 2, n_series * max_points * 8, &d_ts))) return rc;
  if ((rc = ensure(ctx, 3, n_series * max_points * 8, &d_val))) return rc;
  if ((rc = ensure(ctx, 4, n_series * 4, &d_n))) return rc;
  if ((rc = ensure(ctx, 5, n_series * 4, &d_st))) return rc;
  if (h_unit && (rc = ensure(ctx, 6, n_series, &d_unit))) return rc;
  if (h_ann && (rc = ensure(ctx, 7, n_series * sizeof(m3tsz_annotation_ref), &d_ann))) return rc;
  // Chunked pipeline over two streams: the H2D copy of chunk c+1 overlaps the
  // kernel and the D2H copy of chunk c (PCIe is full duplex).
  cudaStream_t sts[2] = {ctx->stream, ctx->stream2};
  CK(cudaMemcpyAsync(d_off, h_offsets, (n_series + 1) * 8, cudaMemcpyHostToDevice, sts[0]));
  CK(cudaEventRecord(ctx->ev, sts[0]));
  CK(cudaStreamWaitEvent(sts[1], ctx->ev, 0));
  const uint64_t per_series = streams_bytes / n_series + max_points * 16;
  const uint64_t ch = pick_chunk(n_series, per_series);
  int k = 0;
  for (uint64_t c0 = 0; c0 < n_series; c0 += ch, k++) {
    const uint64_t c1 = (c0 + ch < n_series) ? c0 + ch : n_series, n = c1 - c0;
    cudaStream_t st = sts[k & 1];
    const uint64_t b0 = h_offsets[c0], b1 = h_offsets[c1];
    if (b1 < b0 || b1 > streams_bytes) return M3TSZ_ERR_INVALID_ARG;
    if (b1 > b0)
      CK(cudaMemcpyAsync((uint8_t *)d_streams + b0, h_streams + b0, b1 - b0, cudaMemcpyHostToDevice, st));
    rc = m3tsz_decode_batch(ctx, opts, (const uint8_t *)d_streams, streams_bytes,
                            (const uint64_t *)d_off + c0, n, (int64_t *)d_ts + c0 * max_points,
                            (double *)d_val + c0 * max_points, max_points, (uint32_t *)d_n + c0,
                            (int32_t *)d_st + c0, d_unit ? (uint8_t *)d_unit + c0 : nullptr,
                            d_ann ? (m3tsz_annotation_ref *)d_ann + c0 : nullptr, st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h_ts + c0 * max_points, (int64_t *)d_ts + c0 * max_points, n * max_points * 8,
                       cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_val + c0 * max_points, (double *)d_val + c0 * max_points, n * max_points * 8,
                       cudaMemcpyDeviceToHost, st));
    if (h_n_points) CK(cudaMemcpyAsync(h_n_points + c0, (uint32_t *)d_n + c0, n * 4, cudaMemcpyDeviceToHost, st));
    if (h_status) CK(cudaMemcpyAsync(h_status + c0, (int32_t *)d_st + c0, n * 4, cudaMemcpyDeviceToHost, st));
    if (h_unit) CK(cudaMemcpyAsync(h_unit + c0, (uint8_t *)d_unit + c0, n, cudaMemcpyDeviceToHost, st));
    if (h_ann)
      CK(cudaMemcpyAsync(h_ann + c0, (m3tsz_annotation_ref *)d_ann + c0, n * sizeof(m3tsz_annotation_ref),
                         cudaMemcpyDeviceToHost, st));
  }
  CK(cudaStreamSynchronize(sts[0]));
  CK(cudaStreamSynchronize(sts[1]));
  return M3TSZ_OK;
}

int m3tsz_decode_batch_host(m3tsz_ctx *ctx, const m3tsz_options *opts, const uint8_t *h_streams,
                            uint64_t streams_bytes, const uint64_t *h_offsets, uint64_t n_series,
                            int64_t *h_ts, double *h_val, uint64_t max_points,
                            uint32_t *h_n_points, int32_t *h_status, uint8_t *h_unit,
                            m3tsz_annotation_ref *h_ann) {
  const int rc = m3tsz_decode_batch_host_impl(ctx, opts, h_streams, streams_bytes, h_offsets, n_series, h_ts, h_val, max_points, h_n_points, h_status, h_unit, h_ann);
  if (ctx && ctx->stream) {  // never return with copies from / into the caller's buffers in flight
    DeviceGuard guard(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->stream2);
  }
  return rc;
}

static int m3tsz_decode_downsample_batch_host_impl(m3tsz_ctx *ctx, const m3tsz_options *opts,
                                       const uint8_t *h_streams, uint64_t streams_bytes,
                                       const uint64_t *h_offsets, uint64_t n_series,
                                       int64_t range_start_ns, int64_t window_ns,
                                       uint32_t n_windows, double *h_sum, int64_t *h_count,
                                       double *h_min, double *h_max, uint32_t *h_n_points,
                                       int32_t *h_status) {
  if (!ctx || !valid_opts(opts)) return M3TSZ_ERR_INVALID_ARG;
  if (n_series == 0) return M3TSZ_OK;
  if (!h_streams || !h_offsets || !h_sum || !h_count || !h_min || !h_max) return M3TSZ_ERR_INVALID_ARG;
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return set_cuda_error(ctx, guard.err, "cudaSetDevice");
  cudaStream_t st = ctx->stream;
  const size_t wb = (size_t)n_series * n_windows * 8;
  void *d_streams, *d_off, *d_sum, *d_cnt, *d_min, *d_max, *d_n, *d_st;
  int rc;
  if ((rc = ensure(ctx, 0, streams_bytes + 16, &d_streams))) return rc;
  if ((rc = ensure(ctx, 1, (n_series + 1) * 8, &d_off))) return rc;
  if ((rc = ensure(ctx, 8, wb, &d_sum))) return rc;
  if ((rc = ensure(ctx, 9, wb, &d_cnt))) return rc;
  if ((rc = ensure(ctx, 10, wb, &d_min))) return rc;
  if ((rc = ensure(ctx, 11, wb, &d_max))) return rc;
  if ((rc = ensure(ctx, 4, n_series * 4, &d_n))) return rc;
  if ((rc = ensure(ctx, 5, n_series * 4, &d_st))) return rc;
  CK(cudaMemcpyAsync(d_streams, h_streams, streams_bytes, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_off, h_offsets, (n_series + 1) * 8, cudaMemcpyHostToDevice, st));
  rc = m3tsz_decode_downsample_batch(ctx, opts, (const uint8_t *)d_streams, streams_bytes,
                                     (const uint64_t *)d_off, n_series, range_start_ns, window_ns,
                                     n_windows, (double *)d_sum, (int64_t *)d_cnt, (double *)d_min,
                                     (double *)d_max, (uint32_t *)d_n, (int32_t *)d_st, st);
  if (rc) return rc;
  CK(cudaMemcpyAsync(h_sum, d_sum, wb, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h_count, d_cnt, wb, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h_min, d_min, wb, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h_max, d_max, wb, cudaMemcpyDeviceToHost, st));
  if (h_n_points) CK(cudaMemcpyAsync(h_n_points, d_n, n_series * 4, cudaMemcpyDeviceToHost, st));
  if (h_status) CK(cudaMemcpyAsync(h_status, d_st, n_series * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return M3TSZ_OK;
}

int m3tsz_decode_downsample_batch_host(m3tsz_ctx *ctx, const m3tsz_options *opts,
                                       const uint8_t *h_streams, uint64_t streams_bytes,
                                       const uint64_t *h_offsets, uint64_t n_series,
                                       int64_t range_start_ns, int64_t window_ns,
                                       uint32_t n_windows, double *h_sum, int64_t *h_count,
                                       double *h_min, double *h_max, uint32_t *h_n_points,
                                       int32_t *h_status) {
  const int rc = m3tsz_decode_downsample_batch_host_impl(ctx, opts, h_streams, streams_bytes, h_offsets, n_series, range_start_ns, window_ns, n_windows, h_sum, h_count, h_min, h_max, h_n_points, h_status);
  if (ctx && ctx->stream) {  // never return with copies from / into the caller's buffers in flight
    DeviceGuard guard(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->stream2);
  }
  return rc;
}

static int m3tsz_encode_batch_host_impl(m3tsz_ctx *ctx, const m3tsz_options *opts, const int64_t *h_ts,
                            const double *h_val, uint64_t n_series, uint64_t points_stride,
                            const uint32_t *h_n_points, const int64_t *h_start, int32_t unit,
                            const uint8_t *h_units, const uint64_t *h_ann_series_off,
                            const m3tsz_annotation_entry *h_ann_entries,
                            const uint8_t *h_ann_bytes, uint64_t ann_bytes_len, uint32_t align,
                            uint8_t *h_packed, uint64_t packed_capacity, uint64_t *h_offsets,
                            uint64_t *h_out_len, int32_t *h_status) {
  if (!ctx || !valid_opts(opts)) return M3TSZ_ERR_INVALID_ARG;
  if (!h_offsets) return M3TSZ_ERR_INVALID_ARG;
  h_offsets[0] = 0;
  if (n_series == 0) return M3TSZ_OK;
  if (!h_ts || !h_val || !h_start || !h_packed) return M3TSZ_ERR_INVALID_ARG;
  if (!(align == 1 || align == 4 || align == 8 || align == 16 || align == 32 || align == 64)) return M3TSZ_ERR_INVALID_ARG;
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return set_cuda_error(ctx, guard.err, "cudaSetDevice");
  cudaStream_t sts[2] = {ctx->stream, ctx->stream2};
  const uint64_t ann_total = h_ann_series_off ? ann_bytes_len + 16 * h_ann_series_off[n_series] : 0;
  const uint64_t out_stride = m3tsz_encode_bound_units(points_stride, h_units != nullptr) + ((ann_total + 15) & ~15ull);
  const uint64_t ch = pick_chunk(n_series, points_stride * 16 + points_stride * 8);
  const size_t row_bytes = (size_t)points_stride * 8;
  int rc;
  // whole-batch device inputs that are small: start, n_points, annotations
  void *d_start, *d_np = nullptr, *d_aoff = nullptr, *d_aent = nullptr, *d_abytes = nullptr;
  if ((rc = ensure(ctx, 1, (n_series + 1) * 8, &d_start))) return rc;
  CK(cudaMemcpyAsync(d_start, h_start, n_series * 8, cudaMemcpyHostToDevice, sts[0]));
  if (h_n_points) {
    if ((rc = ensure(ctx, 4, n_series * 4, &d_np))) return rc;
    CK(cudaMemcpyAsync(d_np, h_n_points, n_series * 4, cudaMemcpyHostToDevice, sts[0]));
  }
  if (h_ann_series_off) {
    if (!h_ann_entries || (!h_ann_bytes && ann_bytes_len)) return M3TSZ_ERR_INVALID_ARG;
    const uint64_t n_ent = h_ann_series_off[n_series];
    if ((rc = ensure(ctx, 7, (n_series + 1) * 8, &d_aoff))) return rc;
    if ((rc = ensure(ctx, 13, n_ent * sizeof(m3tsz_annotation_entry), &d_aent))) return rc;
    if ((rc = ensure(ctx, 14, ann_bytes_len, &d_abytes))) return rc;
    CK(cudaMemcpyAsync(d_aoff, h_ann_series_off, (n_series + 1) * 8, cudaMemcpyHostToDevice, sts[0]));
    if (n_ent)
      CK(cudaMemcpyAsync(d_aent, h_ann_entries, n_ent * sizeof(m3tsz_annotation_entry),
                         cudaMemcpyHostToDevice, sts[0]));
    if (ann_bytes_len) CK(cudaMemcpyAsync(d_abytes, h_ann_bytes, ann_bytes_len, cudaMemcpyHostToDevice, sts[0]));
  }
  CK(cudaEventRecord(ctx->ev, sts[0]));
  CK(cudaStreamWaitEvent(sts[1], ctx->ev, 0));

  // per-stream chunk buffers
  struct Lane {
    void *ts, *val, *units, *out, *len, *st, *packed, *off, *tmp;
    uint64_t c0, c1;
    bool busy;
  } ln[2];
  const size_t tmp_bytes = compact_scan_tmp_bytes(ch);
  // every start is rounded up to `align`: budget the aligned worst case per stream
  const uint64_t packed_stride = (out_stride + align - 1) & ~(uint64_t)(align - 1);
  for (int i = 0; i < 2; i++) {
    const int b = 16 + i * 10;
    ln[i].units = nullptr;
    ln[i].busy = false;
    if ((rc = ensure(ctx, b + 0, ch * row_bytes, &ln[i].ts))) return rc;
    if ((rc = ensure(ctx, b + 1, ch * row_bytes, &ln[i].val))) return rc;
    if (h_units && (rc = ensure(ctx, b + 2, ch * points_stride, &ln[i].units))) return rc;
    if ((rc = ensure(ctx, b + 3, ch * out_stride, &ln[i].out))) return rc;
    if ((rc = ensure(ctx, b + 4, ch * 8, &ln[i].len))) return rc;
    if ((rc = ensure(ctx, b + 5, ch * 4, &ln[i].st))) return rc;
    if ((rc = ensure(ctx, b + 6, ch * packed_stride + 64, &ln[i].packed))) return rc;
    if ((rc = ensure(ctx, b + 7, (ch + 1) * 8, &ln[i].off))) return rc;
    if ((rc = ensure(ctx, b + 8, tmp_bytes, &ln[i].tmp))) return rc;
    if ((rc = ensure_stage(ctx, i, (ch + 1) * 8 + 8))) return rc;
  }
  CK(cudaMemsetAsync(ctx->d_flag, 0, 2 * sizeof(int32_t), sts[0]));
  CK(cudaStreamSynchronize(sts[0]));  // the flags are zero before either lane's kernels can set them

  auto issue = [&](int i, uint64_t c0) -> int {
    Lane &L = ln[i];
    cudaStream_t st = sts[i];
    L.c0 = c0;
    L.c1 = (c0 + ch < n_series) ? c0 + ch : n_series;
    L.busy = true;
    const uint64_t n = L.c1 - c0;
    CK(cudaMemcpyAsync(L.ts, h_ts + c0 * points_stride, n * row_bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(L.val, h_val + c0 * points_stride, n * row_bytes, cudaMemcpyHostToDevice, st));
    if (h_units)
      CK(cudaMemcpyAsync(L.units, h_units + c0 * points_stride, n * points_stride, cudaMemcpyHostToDevice, st));
    int r = m3tsz_encode_batch(ctx, opts, (const int64_t *)L.ts, (const double *)L.val, n, points_stride,
                               d_np ? (const uint32_t *)d_np + c0 : nullptr, (const int64_t *)d_start + c0,
                               unit, (const uint8_t *)L.units, d_aoff ? (const uint64_t *)d_aoff + c0 : nullptr,
                               (const m3tsz_annotation_entry *)d_aent, (const uint8_t *)d_abytes,
                               (uint8_t *)L.out, out_stride, (uint64_t *)L.len, (int32_t *)L.st, st);
    if (r) return r;
    CK(launch_compact((const uint8_t *)L.out, out_stride, (const uint64_t *)L.len, n, align,
                      (uint8_t *)L.packed, ch * packed_stride + 64, (uint64_t *)L.off, L.tmp, tmp_bytes,
                      ctx->d_flag + i, st));
    ctx->launches += 3;
    CK(cudaMemcpyAsync(ctx->h_stage[i], L.off, (n + 1) * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync((uint8_t *)ctx->h_stage[i] + (ch + 1) * 8, ctx->d_flag + i, 4, cudaMemcpyDeviceToHost, st));
    if (h_out_len) CK(cudaMemcpyAsync(h_out_len + c0, L.len, n * 8, cudaMemcpyDeviceToHost, st));
    if (h_status) CK(cudaMemcpyAsync(h_status + c0, L.st, n * 4, cudaMemcpyDeviceToHost, st));
    return M3TSZ_OK;
  };
  uint64_t host_base = 0;
  auto finish = [&](int i) -> int {
    Lane &L = ln[i];
    if (!L.busy) return M3TSZ_OK;
    L.busy = false;
    CK(cudaStreamSynchronize(sts[i]));  // offsets of this chunk are on the host now
    const uint64_t n = L.c1 - L.c0;
    const uint64_t *co = (const uint64_t *)ctx->h_stage[i];
    const uint64_t total = co[n];
    int32_t overflow = 0;
    memcpy(&overflow, (const uint8_t *)ctx->h_stage[i] + (ch + 1) * 8, 4);
    if (overflow || host_base + total > packed_capacity) return M3TSZ_ERR_CAPACITY;
    if (total)
      CK(cudaMemcpyAsync(h_packed + host_base, L.packed, total, cudaMemcpyDeviceToHost, sts[i]));
    for (uint64_t j = 0; j <= n; j++) h_offsets[L.c0 + j] = host_base + co[j];
    host_base += total;
    host_base = (host_base + align - 1) & ~(uint64_t)(align - 1);
    return M3TSZ_OK;
  };
  int k = 0;
  if ((rc = issue(0, 0))) return rc;
  for (uint64_t c0 = 0; c0 < n_series; c0 += ch, k++) {
    const uint64_t next = c0 + ch;
    if (next < n_series) {
      // the other lane's previous chunk has been finished already (its D2H is queued
      // on its own stream, so reusing its buffers is ordered)
      if ((rc = issue((k + 1) & 1, next))) return rc;
    }
    if ((rc = finish(k & 1))) return rc;
  }
  CK(cudaStreamSynchronize(sts[0]));
  CK(cudaStreamSynchronize(sts[1]));
  // CSR convention: offsets[n] is the end of the last stream; un-pad the final alignment
  return M3TSZ_OK;
}

int m3tsz_encode_batch_host(m3tsz_ctx *ctx, const m3tsz_options *opts, const int64_t *h_ts,
                            const double *h_val, uint64_t n_series, uint64_t points_stride,
                            const uint32_t *h_n_points, const int64_t *h_start, int32_t unit,
                            const uint8_t *h_units, const uint64_t *h_ann_series_off,
                            const m3tsz_annotation_entry *h_ann_entries,
                            const uint8_t *h_ann_bytes, uint64_t ann_bytes_len, uint32_t align,
                            uint8_t *h_packed, uint64_t packed_capacity, uint64_t *h_offsets,
                            uint64_t *h_out_len, int32_t *h_status) {
  const int rc = m3tsz_encode_batch_host_impl(ctx, opts, h_ts, h_val, n_series, points_stride, h_n_points, h_start, unit, h_units, h_ann_series_off, h_ann_entries, h_ann_bytes, ann_bytes_len, align, h_packed, packed_capacity, h_offsets, h_out_len, h_status);
  if (ctx && ctx->stream) {  // never return with copies from / into the caller's buffers in flight
    DeviceGuard guard(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->stream2);
  }
  return rc;
}


// --------------------------------------------------------------------------
// Fetch path with host buffers: only compressed bytes go up, only the merged result
// comes back (the decoded replicas never cross PCIe).
int m3tsz_fetch_batch_host(m3tsz_ctx *ctx, const m3tsz_options *opts, const uint8_t *h_streams,
                           uint64_t streams_bytes, const uint64_t *h_offsets, uint64_t n_seq,
                           const uint64_t *h_slice_off, const uint64_t *h_replica_off,
                           const uint64_t *h_series_off, uint64_t n_series, uint64_t max_points,
                           int64_t start_ns, int64_t end_ns, int32_t strategy, int64_t *h_ts_out,
                           double *h_val_out, uint64_t out_cap, uint32_t *h_n_out, int32_t *h_status) {
  if (!ctx || !valid_opts(opts)) return M3TSZ_ERR_INVALID_ARG;
  if (n_series == 0) return M3TSZ_OK;
  if (!h_streams || !h_offsets || !h_slice_off || !h_replica_off || !h_series_off || !h_ts_out || !h_val_out ||
      !h_n_out || !h_status || max_points == 0 || out_cap == 0 || n_seq == 0 || strategy < 0 || strategy > 3)
    return M3TSZ_ERR_INVALID_ARG;
  const uint64_t n_rep = h_series_off[n_series], n_slice = h_replica_off[n_rep];
  if (h_slice_off[n_slice] != n_seq || h_series_off[0] != 0 || h_replica_off[0] != 0 || h_slice_off[0] != 0)
    return M3TSZ_ERR_INVALID_ARG;
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return set_cuda_error(ctx, guard.err, "cudaSetDevice");
  cudaStream_t sts[2] = {ctx->stream, ctx->stream2};
  // series per chunk: ~8 chunks, whole sequences of whole series
  uint64_t ch = (n_series + 7) / 8;
  {
    const uint64_t per_series = (streams_bytes + (n_seq * max_points * 16)) / n_series + out_cap * 16;
    uint64_t min_ch = (32ull << 20) / (per_series ? per_series : 1);
    if (min_ch < 1024) min_ch = 1024;
    if (ch < min_ch) ch = min_ch;
    if (ch > n_series) ch = n_series;
  }
  // the largest chunk in sequences / slices / replicas
  uint64_t max_q = 0, max_k = 0, max_r = 0;
  for (uint64_t s0 = 0; s0 < n_series; s0 += ch) {
    const uint64_t s1 = s0 + ch < n_series ? s0 + ch : n_series;
    const uint64_t r0 = h_series_off[s0], r1 = h_series_off[s1];
    if (r1 < r0 || r1 > n_rep) return M3TSZ_ERR_INVALID_ARG;
    const uint64_t k0 = h_replica_off[r0], k1 = h_replica_off[r1];
    if (k1 < k0 || k1 > n_slice) return M3TSZ_ERR_INVALID_ARG;
    const uint64_t q0 = h_slice_off[k0], q1 = h_slice_off[k1];
    if (q1 < q0 || q1 > n_seq) return M3TSZ_ERR_INVALID_ARG;
    if (q1 - q0 > max_q) max_q = q1 - q0;
    if (k1 - k0 > max_k) max_k = k1 - k0;
    if (r1 - r0 > max_r) max_r = r1 - r0;
  }
  void *d_streams, *d_off;
  int rc;
  if ((rc = ensure(ctx, 0, streams_bytes + 16, &d_streams))) return rc;
  if ((rc = ensure(ctx, 1, (n_seq + 1) * 8, &d_off))) return rc;
  struct Lane {
    void *ts, *val, *n, *st, *meta, *ots, *oval, *on, *ost;
  } ln[2];
  const size_t meta_words = (max_k + 1) + (max_r + 1) + (ch + 1);
  for (int i = 0; i < 2; i++) {
    const int b = 16 + i * 10;
    if ((rc = ensure(ctx, b + 0, max_q * max_points * 8, &ln[i].ts))) return rc;
    if ((rc = ensure(ctx, b + 1, max_q * max_points * 8, &ln[i].val))) return rc;
    if ((rc = ensure(ctx, b + 2, max_q * 4, &ln[i].n))) return rc;
    if ((rc = ensure(ctx, b + 3, max_q * 4, &ln[i].st))) return rc;
    if ((rc = ensure(ctx, b + 4, meta_words * 8, &ln[i].meta))) return rc;
    if ((rc = ensure(ctx, b + 5, ch * out_cap * 8, &ln[i].ots))) return rc;
    if ((rc = ensure(ctx, b + 6, ch * out_cap * 8, &ln[i].oval))) return rc;
    if ((rc = ensure(ctx, b + 7, ch * 4, &ln[i].on))) return rc;
    if ((rc = ensure(ctx, b + 8, ch * 4, &ln[i].ost))) return rc;
    if ((rc = ensure_stage(ctx, i, meta_words * 8))) return rc;
  }
  auto fail = [&](int code) {  // never return with copies into the caller's buffers in flight
    cudaStreamSynchronize(sts[0]);
    cudaStreamSynchronize(sts[1]);
    return code;
  };
  CK(cudaMemcpyAsync(d_off, h_offsets, (n_seq + 1) * 8, cudaMemcpyHostToDevice, sts[0]));
  CK(cudaEventRecord(ctx->ev, sts[0]));
  CK(cudaStreamWaitEvent(sts[1], ctx->ev, 0));
  int k = 0;
  for (uint64_t s0 = 0; s0 < n_series; s0 += ch, k++) {
    const int i = k & 1;
    cudaStream_t st = sts[i];
    Lane &L = ln[i];
    const uint64_t s1 = s0 + ch < n_series ? s0 + ch : n_series, ns = s1 - s0;
    const uint64_t r0 = h_series_off[s0], r1 = h_series_off[s1];
    const uint64_t k0 = h_replica_off[r0], k1 = h_replica_off[r1];
    const uint64_t q0 = h_slice_off[k0], q1 = h_slice_off[k1], nq = q1 - q0;
    const uint64_t b0 = h_offsets[q0], b1 = h_offsets[q1];
    if (b1 < b0 || b1 > streams_bytes) return fail(M3TSZ_ERR_INVALID_ARG);
    if (cudaStreamSynchronize(st) != cudaSuccess) return fail(M3TSZ_ERR_CUDA);  // lane buffers + staging are free again
    // iterator structure of the chunk, rebased to the chunk's first sequence / slice / replica
    uint64_t *m = (uint64_t *)ctx->h_stage[i];
    uint64_t *m_slice = m, *m_rep = m + (max_k + 1), *m_ser = m_rep + (max_r + 1);
    for (uint64_t j = k0; j <= k1; j++) m_slice[j - k0] = h_slice_off[j] - q0;
    for (uint64_t j = r0; j <= r1; j++) m_rep[j - r0] = h_replica_off[j] - k0;
    for (uint64_t j = s0; j <= s1; j++) m_ser[j - s0] = h_series_off[j] - r0;
    if (cudaMemcpyAsync(L.meta, m, meta_words * 8, cudaMemcpyHostToDevice, st) != cudaSuccess)
      return fail(M3TSZ_ERR_CUDA);
    if (b1 > b0 && cudaMemcpyAsync((uint8_t *)d_streams + b0, h_streams + b0, b1 - b0, cudaMemcpyHostToDevice, st) !=
                       cudaSuccess)
      return fail(M3TSZ_ERR_CUDA);
    if (nq) {
      rc = m3tsz_decode_batch(ctx, opts, (const uint8_t *)d_streams, streams_bytes, (const uint64_t *)d_off + q0, nq,
                              (int64_t *)L.ts, (double *)L.val, max_points, (uint32_t *)L.n, (int32_t *)L.st,
                              nullptr, nullptr, st);
      if (rc) return fail(rc);
    }
    const uint64_t *dm = (const uint64_t *)L.meta;
    rc = m3tsz_merge_series_batch(ctx, (const int64_t *)L.ts, (const double *)L.val, max_points,
                                  (const uint32_t *)L.n, (const int32_t *)L.st, dm, dm + (max_k + 1),
                                  dm + (max_k + 1) + (max_r + 1), ns, start_ns, end_ns, strategy, (int64_t *)L.ots,
                                  (double *)L.oval, out_cap, (uint32_t *)L.on, (int32_t *)L.ost, st);
    if (rc) return fail(rc);
    if (cudaMemcpyAsync(h_ts_out + s0 * out_cap, L.ots, ns * out_cap * 8, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaMemcpyAsync(h_val_out + s0 * out_cap, L.oval, ns * out_cap * 8, cudaMemcpyDeviceToHost, st) !=
            cudaSuccess ||
        cudaMemcpyAsync(h_n_out + s0, L.on, ns * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaMemcpyAsync(h_status + s0, L.ost, ns * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess)
      return fail(M3TSZ_ERR_CUDA);
  }
  if (cudaStreamSynchronize(sts[0]) != cudaSuccess || cudaStreamSynchronize(sts[1]) != cudaSuccess)
    return set_cuda_error(ctx, cudaGetLastError(), "m3tsz_fetch_batch_host");
  return M3TSZ_OK;
}

}  // extern "C"
