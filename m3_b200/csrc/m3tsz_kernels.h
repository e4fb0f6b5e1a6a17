// m3tsz_kernels.h -- internal launch interface between the C-ABI layer
// (m3tsz_capi.cu) and the kernels (m3tsz_decode.cu, m3tsz_encode.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/m3tsz_b200.h"

namespace m3tsz {

struct DecodeParams {
  const uint8_t *streams;
  uint64_t streams_bytes;
  const uint64_t *offsets;
  const uint64_t *lengths;  // optional [n_series]: stream s = [offsets[s], offsets[s] + lengths[s])
  uint64_t n_series;
  int default_unit;
  // plain decode outputs (series-major [n_series][cap])
  int64_t *ts;
  double *val;
  uint64_t cap;
  // fused downsample outputs (window-major [n_windows][n_series])
  int64_t range_start, window;
  uint32_t n_windows;
  double *ds_sum;
  int64_t *ds_count;
  double *ds_min;
  double *ds_max;
  double *ds_last;      // mode 2
  int64_t *ds_last_at;  // mode 2 (lastAt of every window; needed to re-open a window)
  // per-series outputs
  uint32_t *n_points;
  int32_t *status;
  uint8_t *unit_out;
  m3tsz_annotation_ref *ann_out;
  uint8_t *unit_first_out;
  // per-datapoint unit / annotation events (optional)
  m3tsz_dp_event *events;
  uint64_t events_capacity;
  unsigned long long *event_count;
};

// mode: 0 plain decode, 1 fused downsample (sum/count/min/max), 2 fused downsample + last
cudaError_t launch_decode(const DecodeParams &p, bool int_optimized, int mode, cudaStream_t stream);

struct EncodeParams {
  const int64_t *ts;
  const double *val;
  uint64_t n_series;
  uint64_t points_stride;
  const uint32_t *n_points;  // optional
  const int64_t *start;
  int unit;
  const uint8_t *units;  // optional per-datapoint units
  const uint64_t *ann_series_off;
  const m3tsz_annotation_entry *ann_entries;
  const uint8_t *ann_bytes;
  int default_unit;
  uint8_t *out;
  uint64_t out_stride;
  uint64_t *out_len;
  int32_t *status;
  uint64_t *out_bits;   // optional [n_series]: stream length in bits (incl. the EOS marker)
  double *last_value;  // optional [n_series]: Encoder.LastEncoded().Value incl. the scaled-int quirk
  // packed mode (out = per-warp scratch slot sets, see encode_kernel)
  uint8_t *packed;
  uint64_t packed_capacity;
  uint64_t *packed_off;                // [n_series] start of every stream in `packed`
  unsigned long long *packed_cursor;   // bytes allocated so far (zero on entry)
  unsigned long long *batch_counter;   // work counter (zero on entry)
  uint32_t align;
  uint64_t stagger_ns;  // start-up phase spread of the persistent warps (0 = none)
  // input stage (see encode_kernel): 0 series-major ts/val, 1 point-major ts/val, 2 Gauge aggregates
  int in_mode;
  const double *tile_sum, *tile_min, *tile_max, *tile_last;  // window-major [points_stride][n_series]
  const int64_t *tile_count;
  const int32_t *tile_src_status;  // optional [n_series]
  int64_t tile_start, tile_step;
  int agg_type;
  uint32_t *n_tiles_out;  // optional [n_series]
};

uint64_t encode_packed_resident_blocks();
uint64_t encode_packed_scratch_slots(uint64_t n_series);  // slots of out_stride bytes the packed mode needs
cudaError_t launch_encode(const EncodeParams &p, bool int_optimized, cudaStream_t stream);

// iterator layer above the codec (m3tsz_merge.cu)
struct MergeParams {
  const int64_t *ts;  // decoded reader sequences [n_seq][cap]
  const double *val;
  uint64_t cap;
  const uint32_t *n_points;   // [n_seq]
  const int32_t *seq_status;  // optional [n_seq]
  const uint64_t *slice_off, *replica_off, *series_off;
  uint64_t n_series;
  int64_t start, end;
  int strategy;
  int64_t *ts_out;  // [n_series][out_cap]
  double *val_out;
  uint64_t out_cap;
  uint32_t *n_out;
  int32_t *status;
  // element strides: series-major [seq][cap] = (cap, 1); point-major [cap][n_seq] = (1, n_seq)
  uint64_t in_seq_stride, in_pt_stride, out_seq_stride, out_pt_stride;
};
cudaError_t launch_merge(const MergeParams &p, cudaStream_t stream);

struct ChecksumParams {
  const uint8_t *streams;
  uint64_t streams_bytes;
  const uint64_t *offsets;  // CSR [n_series + 1]
  const uint64_t *lengths;  // optional [n_series]: exact sizes when the starts are padded
  uint64_t n_series;
  const uint32_t *expected;  // optional: index-entry DataChecksum per stream
  uint32_t *out;             // optional: Adler-32 per stream
  int32_t *status;           // optional: OK / CHECKSUM_MISMATCH / INVALID_ARG / STREAM_TOO_LARGE
};
cudaError_t launch_checksum(const ChecksumParams &p, cudaStream_t stream);

// Prometheus conversion epilogue (m3tsz_query.cu)
struct PromParams {
  const int64_t *ts;  // [n_series][cap]
  const double *val;
  uint64_t cap;
  const uint32_t *n_points;
  uint64_t n_series;
  int64_t resolution;
  const uint8_t *handle_resets;  // optional [n_series]
  double tolerance;
  int64_t tolerance_until;
  int64_t *ts_out;  // [n_series][out_cap] milliseconds
  double *val_out;
  uint64_t out_cap;
  uint32_t *n_out;
  int32_t *status;  // optional
};
cudaError_t launch_prom(const PromParams &p, cudaStream_t stream);

// tile aggregation glue (m3tsz_query.cu)
cudaError_t launch_fill_i64(int64_t *dst, int64_t v, uint64_t n, cudaStream_t stream);
cudaError_t launch_tiles_status(const int32_t *src_status, int32_t *status, uint64_t n, cudaStream_t stream);

// exclusive scan of aligned lengths + gather into a packed buffer
cudaError_t launch_compact(const uint8_t *slots, uint64_t slot_stride, const uint64_t *len,
                           uint64_t n_series, uint32_t align, uint8_t *packed,
                           uint64_t packed_capacity, uint64_t *offsets, void *scan_tmp,
                           size_t scan_tmp_bytes, int32_t *overflow_flag, cudaStream_t stream);
size_t compact_scan_tmp_bytes(uint64_t n_series);

}  // namespace m3tsz
