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

using namespace m3tsz;

struct Scratch {
  void *ptr = nullptr;
  size_t bytes = 0;
};

struct m3tsz_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;  // used by the *_host entry points
  uint64_t launches = 0;
  char last_error[256] = {0};
  Scratch s[16];
  int32_t *d_flag = nullptr;
};

namespace {

int set_cuda_error(m3tsz_ctx *ctx, cudaError_t e, const char *where) {
  if (ctx) snprintf(ctx->last_error, sizeof(ctx->last_error), "%s: %s", where, cudaGetErrorString(e));
  return M3TSZ_ERR_CUDA;
}

#define CK(call)                                                 \
  do {                                                           \
    cudaError_t _e = (call);                                     \
    if (_e != cudaSuccess) return set_cuda_error(ctx, _e, #call); \
  } while (0)

int ensure(m3tsz_ctx *ctx, int slot, size_t bytes, void **out) {
  Scratch &sc = ctx->s[slot];
  if (bytes == 0) bytes = 16;
  if (sc.bytes < bytes) {
    if (sc.ptr) CK(cudaFree(sc.ptr));
    sc.ptr = nullptr;
    sc.bytes = 0;
    size_t want = bytes + bytes / 8;  // grow-only with slack
    cudaError_t e = cudaMalloc(&sc.ptr, want);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      want = bytes;
      CK(cudaMalloc(&sc.ptr, want));
    }
    sc.bytes = want;
  }
  *out = sc.ptr;
  return M3TSZ_OK;
}

bool valid_opts(const m3tsz_options *o) {
  return o && (o->int_optimized == 0 || o->int_optimized == 1) && o->default_time_unit >= 0 &&
         o->default_time_unit <= 8;
}

}  // namespace

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
      cudaMalloc(&ctx->d_flag, sizeof(int32_t)) != cudaSuccess) {
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
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char *m3tsz_last_cuda_error(const m3tsz_ctx *ctx) { return ctx ? ctx->last_error : ""; }
uint64_t m3tsz_ctx_launch_count(const m3tsz_ctx *ctx) { return ctx ? ctx->launches : 0; }

uint64_t m3tsz_encode_bound(uint64_t n) {
  // 64-bit start + per datapoint <= 68 (timestamp) + 80 (value) bits + 11-bit
  // end-of-stream marker, plus the write-ahead guard the kernel keeps.
  uint64_t bits = 64 + n * 148 + 11;
  uint64_t bytes = (bits + 7) / 8 + 64;
  return (bytes + 15) & ~15ull;
}

// --------------------------------------------------------------------------
int m3tsz_decode_batch(m3tsz_ctx *ctx, const m3tsz_options *opts, const uint8_t *d_streams,
                       uint64_t streams_bytes, const uint64_t *d_offsets, uint64_t n_series,
                       int64_t *d_ts, double *d_val, uint64_t max_points, uint32_t *d_n_points,
                       int32_t *d_status, uint8_t *d_unit, m3tsz_annotation_ref *d_ann,
                       void *stream) {
  if (!ctx || !valid_opts(opts)) return M3TSZ_ERR_INVALID_ARG;
  if (n_series == 0) return M3TSZ_OK;
  if (!d_streams || !d_offsets || !d_ts || !d_val || max_points == 0 ||
      max_points > 0xffffffffull || ((uintptr_t)d_streams & 15u) ||
      streams_bytes >= (1ull << 34))  // 32-bit word indices inside the kernel; split larger batches
    return M3TSZ_ERR_INVALID_ARG;
  CK(cudaSetDevice(ctx->device));
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
  CK(launch_decode(p, opts->int_optimized != 0, false, (cudaStream_t)stream));
  ctx->launches++;
  return M3TSZ_OK;
}

int m3tsz_decode_downsample_batch(m3tsz_ctx *ctx, const m3tsz_options *opts,
                                  const uint8_t *d_streams, uint64_t streams_bytes,
                                  const uint64_t *d_offsets, uint64_t n_series,
                                  int64_t range_start_ns, int64_t window_ns, uint32_t n_windows,
                                  double *d_sum, int64_t *d_count, double *d_min, double *d_max,
                                  uint32_t *d_n_points, int32_t *d_status, void *stream) {
  if (!ctx || !valid_opts(opts)) return M3TSZ_ERR_INVALID_ARG;
  if (n_series == 0) return M3TSZ_OK;
  if (!d_streams || !d_offsets || !d_sum || !d_count || !d_min || !d_max || window_ns <= 0 ||
      n_windows == 0 || ((uintptr_t)d_streams & 15u) || streams_bytes >= (1ull << 34))
    return M3TSZ_ERR_INVALID_ARG;
  // range_start + n_windows*window must not overflow int64
  if ((__int128)range_start_ns + (__int128)n_windows * (__int128)window_ns > (__int128)INT64_MAX)
    return M3TSZ_ERR_INVALID_ARG;
  CK(cudaSetDevice(ctx->device));
  DecodeParams p;
  memset(&p, 0, sizeof(p));
  p.streams = d_streams;
  p.streams_bytes = streams_bytes;
  p.offsets = d_offsets;
  p.n_series = n_series;
  p.default_unit = opts->default_time_unit;
  p.range_start = range_start_ns;
  p.window = window_ns;
  p.n_windows = n_windows;
  p.ds_sum = d_sum;
  p.ds_count = d_count;
  p.ds_min = d_min;
  p.ds_max = d_max;
  p.n_points = d_n_points;
  p.status = d_status;
  CK(launch_decode(p, opts->int_optimized != 0, true, (cudaStream_t)stream));
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
  if (!ctx || !valid_opts(opts)) return M3TSZ_ERR_INVALID_ARG;
  if (n_series == 0) return M3TSZ_OK;
  if (!d_ts || !d_val || !d_start || !d_out || !d_out_len || points_stride > 0xffffffffull ||
      (out_stride & 15u) || out_stride == 0 || out_stride > (1ull << 33) ||
      ((uintptr_t)d_out & 15u) || ((uintptr_t)d_ts & 7u) || ((uintptr_t)d_val & 7u))
    return M3TSZ_ERR_INVALID_ARG;
  if (d_ann_series_off && (!d_ann_entries || !d_ann_bytes)) return M3TSZ_ERR_INVALID_ARG;
  CK(cudaSetDevice(ctx->device));
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
  CK(launch_encode(p, opts->int_optimized != 0, (cudaStream_t)stream));
  ctx->launches++;
  return M3TSZ_OK;
}

int m3tsz_compact_streams(m3tsz_ctx *ctx, const uint8_t *d_slots, uint64_t slot_stride,
                          const uint64_t *d_len, uint64_t n_series, uint32_t align,
                          uint8_t *d_packed, uint64_t packed_capacity, uint64_t *d_offsets,
                          void *stream) {
  if (!ctx || !d_offsets || (n_series && (!d_slots || !d_len || !d_packed))) return M3TSZ_ERR_INVALID_ARG;
  if (!(align == 1 || align == 4 || align == 8 || align == 16)) return M3TSZ_ERR_INVALID_ARG;
  CK(cudaSetDevice(ctx->device));
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

// --------------------------------------------------------------------------
// host-buffer variants
// --------------------------------------------------------------------------
int m3tsz_decode_batch_host(m3tsz_ctx *ctx, const m3tsz_options *opts, const uint8_t *h_streams,
                            uint64_t streams_bytes, const uint64_t *h_offsets, uint64_t n_series,
                            int64_t *h_ts, double *h_val, uint64_t max_points,
                            uint32_t *h_n_points, int32_t *h_status, uint8_t *h_unit,
                            m3tsz_annotation_ref *h_ann) {
  if (!ctx || !valid_opts(opts)) return M3TSZ_ERR_INVALID_ARG;
  if (n_series == 0) return M3TSZ_OK;
  if (!h_streams || !h_offsets || !h_ts || !h_val || max_points == 0) return M3TSZ_ERR_INVALID_ARG;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
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
  CK(cudaMemcpyAsync(d_streams, h_streams, streams_bytes, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_off, h_offsets, (n_series + 1) * 8, cudaMemcpyHostToDevice, st));
  rc = m3tsz_decode_batch(ctx, opts, (const uint8_t *)d_streams, streams_bytes,
                          (const uint64_t *)d_off, n_series, (int64_t *)d_ts, (double *)d_val,
                          max_points, (uint32_t *)d_n, (int32_t *)d_st, (uint8_t *)d_unit,
                          (m3tsz_annotation_ref *)d_ann, st);
  if (rc) return rc;
  CK(cudaMemcpyAsync(h_ts, d_ts, n_series * max_points * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h_val, d_val, n_series * max_points * 8, cudaMemcpyDeviceToHost, st));
  if (h_n_points) CK(cudaMemcpyAsync(h_n_points, d_n, n_series * 4, cudaMemcpyDeviceToHost, st));
  if (h_status) CK(cudaMemcpyAsync(h_status, d_st, n_series * 4, cudaMemcpyDeviceToHost, st));
  if (h_unit) CK(cudaMemcpyAsync(h_unit, d_unit, n_series, cudaMemcpyDeviceToHost, st));
  if (h_ann)
    CK(cudaMemcpyAsync(h_ann, d_ann, n_series * sizeof(m3tsz_annotation_ref), cudaMemcpyDeviceToHost, st));
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
  if (!ctx || !valid_opts(opts)) return M3TSZ_ERR_INVALID_ARG;
  if (n_series == 0) return M3TSZ_OK;
  if (!h_streams || !h_offsets || !h_sum || !h_count || !h_min || !h_max) return M3TSZ_ERR_INVALID_ARG;
  CK(cudaSetDevice(ctx->device));
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

int m3tsz_encode_batch_host(m3tsz_ctx *ctx, const m3tsz_options *opts, const int64_t *h_ts,
                            const double *h_val, uint64_t n_series, uint64_t points_stride,
                            const uint32_t *h_n_points, const int64_t *h_start, int32_t unit,
                            const uint8_t *h_units, const uint64_t *h_ann_series_off,
                            const m3tsz_annotation_entry *h_ann_entries,
                            const uint8_t *h_ann_bytes, uint64_t ann_bytes_len, uint32_t align,
                            uint8_t *h_packed, uint64_t packed_capacity, uint64_t *h_offsets,
                            uint64_t *h_out_len, int32_t *h_status) {
  if (!ctx || !valid_opts(opts)) return M3TSZ_ERR_INVALID_ARG;
  if (n_series == 0) return M3TSZ_OK;
  if (!h_ts || !h_val || !h_start || !h_packed || !h_offsets) return M3TSZ_ERR_INVALID_ARG;
  if (!(align == 1 || align == 4 || align == 8 || align == 16)) return M3TSZ_ERR_INVALID_ARG;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t nb = (size_t)n_series * points_stride * 8;
  uint64_t ann_total = h_ann_series_off ? ann_bytes_len + 16 * h_ann_series_off[n_series] : 0;
  const uint64_t out_stride = m3tsz_encode_bound(points_stride) + ((ann_total + 15) & ~15ull);
  void *d_ts, *d_val, *d_np = nullptr, *d_start, *d_units = nullptr, *d_aoff = nullptr,
                      *d_aent = nullptr, *d_abytes = nullptr, *d_out, *d_len, *d_st, *d_packed, *d_off;
  int rc;
  if ((rc = ensure(ctx, 2, nb, &d_ts))) return rc;
  if ((rc = ensure(ctx, 3, nb, &d_val))) return rc;
  if ((rc = ensure(ctx, 1, (n_series + 1) * 8, &d_start))) return rc;
  if ((rc = ensure(ctx, 8, n_series * out_stride, &d_out))) return rc;
  if ((rc = ensure(ctx, 12, n_series * 8, &d_len))) return rc;
  if ((rc = ensure(ctx, 5, n_series * 4, &d_st))) return rc;
  if ((rc = ensure(ctx, 0, packed_capacity + 16, &d_packed))) return rc;
  if ((rc = ensure(ctx, 9, (n_series + 1) * 8, &d_off))) return rc;
  CK(cudaMemcpyAsync(d_ts, h_ts, nb, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_val, h_val, nb, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_start, h_start, n_series * 8, cudaMemcpyHostToDevice, st));
  if (h_n_points) {
    if ((rc = ensure(ctx, 4, n_series * 4, &d_np))) return rc;
    CK(cudaMemcpyAsync(d_np, h_n_points, n_series * 4, cudaMemcpyHostToDevice, st));
  }
  if (h_units) {
    if ((rc = ensure(ctx, 6, n_series * points_stride, &d_units))) return rc;
    CK(cudaMemcpyAsync(d_units, h_units, n_series * points_stride, cudaMemcpyHostToDevice, st));
  }
  if (h_ann_series_off) {
    if (!h_ann_entries || (!h_ann_bytes && ann_bytes_len)) return M3TSZ_ERR_INVALID_ARG;
    const uint64_t n_ent = h_ann_series_off[n_series];
    if ((rc = ensure(ctx, 7, (n_series + 1) * 8, &d_aoff))) return rc;
    if ((rc = ensure(ctx, 13, n_ent * sizeof(m3tsz_annotation_entry), &d_aent))) return rc;
    if ((rc = ensure(ctx, 14, ann_bytes_len, &d_abytes))) return rc;
    CK(cudaMemcpyAsync(d_aoff, h_ann_series_off, (n_series + 1) * 8, cudaMemcpyHostToDevice, st));
    if (n_ent)
      CK(cudaMemcpyAsync(d_aent, h_ann_entries, n_ent * sizeof(m3tsz_annotation_entry),
                         cudaMemcpyHostToDevice, st));
    if (ann_bytes_len) CK(cudaMemcpyAsync(d_abytes, h_ann_bytes, ann_bytes_len, cudaMemcpyHostToDevice, st));
  }
  rc = m3tsz_encode_batch(ctx, opts, (const int64_t *)d_ts, (const double *)d_val, n_series,
                          points_stride, (const uint32_t *)d_np, (const int64_t *)d_start, unit,
                          (const uint8_t *)d_units, (const uint64_t *)d_aoff,
                          (const m3tsz_annotation_entry *)d_aent, (const uint8_t *)d_abytes,
                          (uint8_t *)d_out, out_stride, (uint64_t *)d_len, (int32_t *)d_st, st);
  if (rc) return rc;
  rc = m3tsz_compact_streams(ctx, (const uint8_t *)d_out, out_stride, (const uint64_t *)d_len,
                             n_series, align, (uint8_t *)d_packed, packed_capacity,
                             (uint64_t *)d_off, st);
  if (rc) return rc;
  CK(cudaMemcpyAsync(h_offsets, d_off, (n_series + 1) * 8, cudaMemcpyDeviceToHost, st));
  if (h_out_len) CK(cudaMemcpyAsync(h_out_len, d_len, n_series * 8, cudaMemcpyDeviceToHost, st));
  if (h_status) CK(cudaMemcpyAsync(h_status, d_st, n_series * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const uint64_t total = h_offsets[n_series];
  if (total > packed_capacity) return M3TSZ_ERR_CAPACITY;
  CK(cudaMemcpyAsync(h_packed, d_packed, total, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return M3TSZ_OK;
}

}  // extern "C"
