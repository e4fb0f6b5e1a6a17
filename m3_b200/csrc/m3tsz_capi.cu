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
  cudaStream_t stream = nullptr;   // used by the *_host entry points
  cudaStream_t stream2 = nullptr;  // second lane of the chunked H2D / kernel / D2H pipeline
  cudaEvent_t ev = nullptr;
  uint64_t launches = 0;
  char last_error[256] = {0};
  Scratch s[40];
  int32_t *d_flag = nullptr;  // [2]
  void *h_stage[2] = {nullptr, nullptr};  // pinned staging for per-chunk offsets
  size_t h_stage_bytes[2] = {0, 0};
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

int ensure_stage(m3tsz_ctx *ctx, int i, size_t bytes) {
  if (ctx->h_stage_bytes[i] >= bytes) return M3TSZ_OK;
  if (ctx->h_stage[i]) CK(cudaFreeHost(ctx->h_stage[i]));
  ctx->h_stage[i] = nullptr;
  ctx->h_stage_bytes[i] = 0;
  CK(cudaMallocHost(&ctx->h_stage[i], bytes));
  ctx->h_stage_bytes[i] = bytes;
  return M3TSZ_OK;
}

// Series per pipeline chunk: ~8 chunks per call, but never tiny ones.
uint64_t pick_chunk(uint64_t n_series, uint64_t bytes_per_series) {
  uint64_t ch = (n_series + 7) / 8;
  uint64_t min_ch = (16ull << 20) / (bytes_per_series ? bytes_per_series : 1);
  if (min_ch < 2048) min_ch = 2048;
  if (ch < min_ch) ch = min_ch;
  if (ch > n_series) ch = n_series;
  ch = (ch + 127) & ~127ull;  // whole thread blocks
  return ch;
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
      max_points >= (1ull << 27) || ((uintptr_t)d_streams & 15u) ||
      streams_bytes >= (1ull << 34))  // 32-bit word / byte offsets inside the kernel; split larger batches
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
  if (!(align == 1 || align == 4 || align == 8 || align == 16 || align == 32 || align == 64)) return M3TSZ_ERR_INVALID_ARG;
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

int m3tsz_merge_series_batch(m3tsz_ctx *ctx, const int64_t *d_ts, const double *d_val, uint64_t cap,
                             const uint32_t *d_n_points, const int32_t *d_seq_status,
                             const uint64_t *d_slice_off, const uint64_t *d_replica_off,
                             const uint64_t *d_series_off, uint64_t n_series, int64_t start_ns,
                             int64_t end_ns, int32_t strategy, int64_t *d_ts_out, double *d_val_out,
                             uint64_t out_cap, uint32_t *d_n_out, int32_t *d_status, void *stream) {
  if (!ctx) return M3TSZ_ERR_INVALID_ARG;
  if (n_series == 0) return M3TSZ_OK;
  if (!d_ts || !d_val || !d_n_points || !d_slice_off || !d_replica_off || !d_series_off || !d_ts_out ||
      !d_val_out || !d_n_out || !d_status || cap == 0 || out_cap == 0 || out_cap > 0xffffffffull ||
      strategy < 0 || strategy > 3)
    return M3TSZ_ERR_INVALID_ARG;
  CK(cudaSetDevice(ctx->device));
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
  CK(launch_merge(p, (cudaStream_t)stream));
  ctx->launches++;
  return M3TSZ_OK;
}

int m3tsz_checksum_batch(m3tsz_ctx *ctx, const uint8_t *d_streams, uint64_t streams_bytes,
                         const uint64_t *d_offsets, const uint64_t *d_lengths, uint64_t n_series,
                         const uint32_t *d_expected, uint32_t *d_checksums, int32_t *d_status,
                         void *stream) {
  if (!ctx) return M3TSZ_ERR_INVALID_ARG;
  if (n_series == 0) return M3TSZ_OK;
  if (!d_streams || !d_offsets || (!d_checksums && !d_status) || (d_expected && !d_status))
    return M3TSZ_ERR_INVALID_ARG;
  CK(cudaSetDevice(ctx->device));
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
int m3tsz_decode_batch_host(m3tsz_ctx *ctx, const m3tsz_options *opts, const uint8_t *h_streams,
                            uint64_t streams_bytes, const uint64_t *h_offsets, uint64_t n_series,
                            int64_t *h_ts, double *h_val, uint64_t max_points,
                            uint32_t *h_n_points, int32_t *h_status, uint8_t *h_unit,
                            m3tsz_annotation_ref *h_ann) {
  if (!ctx || !valid_opts(opts)) return M3TSZ_ERR_INVALID_ARG;
  if (n_series == 0) return M3TSZ_OK;
  if (!h_streams || !h_offsets || !h_ts || !h_val || max_points == 0) return M3TSZ_ERR_INVALID_ARG;
  CK(cudaSetDevice(ctx->device));
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
  if (!h_offsets) return M3TSZ_ERR_INVALID_ARG;
  h_offsets[0] = 0;
  if (n_series == 0) return M3TSZ_OK;
  if (!h_ts || !h_val || !h_start || !h_packed) return M3TSZ_ERR_INVALID_ARG;
  if (!(align == 1 || align == 4 || align == 8 || align == 16 || align == 32 || align == 64)) return M3TSZ_ERR_INVALID_ARG;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t sts[2] = {ctx->stream, ctx->stream2};
  const uint64_t ann_total = h_ann_series_off ? ann_bytes_len + 16 * h_ann_series_off[n_series] : 0;
  const uint64_t out_stride = m3tsz_encode_bound(points_stride) + ((ann_total + 15) & ~15ull);
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
    if ((rc = ensure(ctx, b + 6, ch * out_stride + 16, &ln[i].packed))) return rc;
    if ((rc = ensure(ctx, b + 7, (ch + 1) * 8, &ln[i].off))) return rc;
    if ((rc = ensure(ctx, b + 8, tmp_bytes, &ln[i].tmp))) return rc;
    if ((rc = ensure_stage(ctx, i, (ch + 1) * 8))) return rc;
  }
  CK(cudaMemsetAsync(ctx->d_flag, 0, 2 * sizeof(int32_t), sts[0]));

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
                      (uint8_t *)L.packed, ch * out_stride + 16, (uint64_t *)L.off, L.tmp, tmp_bytes,
                      ctx->d_flag + i, st));
    ctx->launches += 3;
    CK(cudaMemcpyAsync(ctx->h_stage[i], L.off, (n + 1) * 8, cudaMemcpyDeviceToHost, st));
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
    if (host_base + total > packed_capacity) return M3TSZ_ERR_CAPACITY;
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

}  // extern "C"
