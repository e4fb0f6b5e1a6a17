/*
 * m3tsz_b200.h -- C ABI of the B200-native M3TSZ batch codec (libm3tsz_b200.so).
 *
 * This is the drop-in boundary for the reference's M3TSZ path.  The reference
 * (m3db/m3, 100 % Go) has no FFI today: the codec sits behind the Go interfaces
 *   encoding.Encoder         src/dbnode/encoding/types.go:39-91
 *   encoding.ReaderIterator  src/dbnode/encoding/types.go:180-203
 *   encoding.Decoder         src/dbnode/encoding/types.go:342-345
 * installed through the pool allocator closures at
 *   src/dbnode/server/server.go:1780-1800 (and the sites in SURVEY.md §1).
 * A cgo shim implements those interfaces over the entry points
 * below; each entry point cites the reference function(s) it replaces.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every function
 * returns an int status (M3TSZ_OK == 0) and never throws or aborts; the caller
 * owns every buffer; the library owns only scratch held by an explicit
 * m3tsz_ctx.  "d_" pointers are CUDA device pointers, "h_" pointers are host
 * pointers.  `stream` is a cudaStream_t passed as void* (NULL = default stream).
 * Device entry points are asynchronous on `stream`; host entry points
 * synchronise before returning.
 *
 * Bit-exactness contract: encoded streams are byte-identical to what the
 * reference m3tsz encoder produces for the same (start, datapoints, units,
 * annotations); decoded (timestamp, value) pairs are bit-identical to the
 * reference iterator's.  See DESIGN.md for the (documented) behaviour on
 * corrupt / truncated streams.
 */
#ifndef M3TSZ_B200_H
#define M3TSZ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define M3TSZ_B200_VERSION 100 /* 0.1.0 */

/* xtime.Unit, src/x/time/unit.go:30-42 */
enum {
  M3TSZ_UNIT_NONE = 0,
  M3TSZ_UNIT_SECOND = 1,
  M3TSZ_UNIT_MILLISECOND = 2,
  M3TSZ_UNIT_MICROSECOND = 3,
  M3TSZ_UNIT_NANOSECOND = 4,
  M3TSZ_UNIT_MINUTE = 5,
  M3TSZ_UNIT_HOUR = 6,
  M3TSZ_UNIT_DAY = 7,
  M3TSZ_UNIT_YEAR = 8
};

/* Status codes.  1..12 mirror the reference's error values for this path. */
enum {
  M3TSZ_OK = 0,
  M3TSZ_ERR_EOF = 1,               /* io.EOF from IStream (istream.go:86-92): truncated stream */
  M3TSZ_ERR_ENCODER_CLOSED = 2,    /* errEncoderClosed, m3tsz/encoder.go:37 */
  M3TSZ_ERR_NO_DATAPOINTS = 3,     /* errNoEncodedDatapoints, m3tsz/encoder.go:38 */
  M3TSZ_ERR_DOD_OVERFLOW = 4,      /* "deltaOfDelta value %d %s overflows 32 bits", timestamp_encoder.go:219 */
  M3TSZ_ERR_NO_TIME_SCHEME = 5,    /* errNoTimeSchemaForUnit, timestamp_iterator.go:33 */
  M3TSZ_ERR_UNRECOGNIZED_UNIT = 6, /* errUnrecognizedTimeUnit, src/x/time/unit.go:45 */
  M3TSZ_ERR_INVALID_MULT = 7,      /* errInvalidMultiplier, m3tsz/m3tsz.go:69 */
  M3TSZ_ERR_ANNOTATION_LEN = 8,    /* errUnexpectedAnnotationLength, timestamp_iterator.go:34 */
  M3TSZ_ERR_ANNOTATION_SHORT = 9,  /* errAnnotationTooFewBytes, timestamp_iterator.go:35 */
  M3TSZ_ERR_ITER_CLOSED = 10,      /* errClosed, m3tsz/iterator.go:33 */
  M3TSZ_ERR_VARINT_OVERFLOW = 11,  /* Go encoding/binary errOverflow */
  M3TSZ_ERR_UNEXPECTED_EOF = 12,   /* io.ErrUnexpectedEOF */
  M3TSZ_ERR_OUT_OF_ORDER = 13,     /* errOutOfOrderIterator, encoding/iterators.go:229-236 */
  M3TSZ_ERR_TOO_MANY_ITERATORS = 14, /* > 12 replicas per series or readers per block slice */
  M3TSZ_ERR_CHECKSUM_MISMATCH = 15, /* errSeekChecksumMismatch, persist/fs/seek.go:50-51, read.go:395-397 */
  /* library-level conditions */
  M3TSZ_ERR_CAPACITY = 100,        /* more datapoints / bytes than the caller's buffer holds */
  M3TSZ_ERR_INVALID_ARG = 101,
  M3TSZ_ERR_CUDA = 102,            /* a CUDA runtime call failed; see m3tsz_last_cuda_error */
  M3TSZ_ERR_NO_DEVICE = 103,       /* no usable CUDA device: the library has NO CPU fallback */
  M3TSZ_ERR_STREAM_TOO_LARGE = 104 /* a single stream exceeds 256 MiB */
};

/* encoding.Options subset that changes the bitstream (encoding/options.go:31-73)
 * + the intOptimized constructor flag (m3tsz/encoder.go:64-69, iterator.go:67-71). */
typedef struct m3tsz_options {
  int32_t int_optimized;     /* m3tsz.DefaultIntOptimizationEnabled = 1 (m3tsz/m3tsz.go:30) */
  int32_t default_time_unit; /* encoding.Options.DefaultTimeUnit(), default M3TSZ_UNIT_SECOND */
} m3tsz_options;

typedef struct m3tsz_ctx m3tsz_ctx;

/* Library / context ------------------------------------------------------ */
int m3tsz_version(void);
const char *m3tsz_status_string(int status);
/* Creates a context bound to CUDA device `device`.  Fails with
 * M3TSZ_ERR_NO_DEVICE when there is no GPU: there is no CPU code path. */
int m3tsz_ctx_create(int device, m3tsz_ctx **out);
void m3tsz_ctx_destroy(m3tsz_ctx *ctx);
const char *m3tsz_last_cuda_error(const m3tsz_ctx *ctx);
/* Number of kernel launches this context has issued (bench "gpu_launches"). */
uint64_t m3tsz_ctx_launch_count(const m3tsz_ctx *ctx);

/* First annotation seen in a stream (ReaderIterator.Current()'s ts.Annotation,
 * encoding/types.go:184-187): bit offset of its first payload byte from the
 * start of the stream, its length in bytes, and how many annotations the
 * stream holds in total. */
typedef struct m3tsz_annotation_ref {
  uint64_t bit_offset;
  uint32_t length;
  uint32_t count;
} m3tsz_annotation_ref;

/* Per-datapoint unit and annotation (ReaderIterator.Current() returns
 * (datapoint, unit, annotation) for EVERY datapoint, encoding/types.go:184-187,
 * m3tsz/iterator.go:229-231).  Both change rarely, so the decoder reports them as
 * an event table instead of two dense [n_series][max_points] arrays:
 *   M3TSZ_EVENT_TIME_UNIT  : from datapoint dp_index on, the unit in force is `unit`
 *                            (a time-unit marker, timestamp_iterator.go:118-134; the unit of
 *                            datapoint 0 is reported per series in d_unit_first)
 *   M3TSZ_EVENT_ANNOTATION : datapoint dp_index carries an annotation of `length` bytes whose
 *                            first payload bit is `bit_offset` bits after the start of the stream
 *                            (timestamp_iterator.go:328-356); datapoints without an event have none.
 * Events of one series appear in stream order but interleaved with other series';
 * events whose dp_index >= n_points[series] belong to a datapoint that failed to decode. */
enum { M3TSZ_EVENT_TIME_UNIT = 1, M3TSZ_EVENT_ANNOTATION = 2 };
typedef struct m3tsz_dp_event {
  uint64_t series;
  uint32_t dp_index;
  uint16_t kind;
  uint16_t unit;
  uint64_t bit_offset;
  uint32_t length;
  uint32_t reserved;
} m3tsz_dp_event;

/* Optional inputs / outputs of m3tsz_decode_batch_ex (all may be NULL / 0). */
typedef struct m3tsz_decode_extras {
  const uint64_t *d_lengths; /* [n_series] exact stream sizes: stream s = [d_offsets[s], +d_lengths[s]);
                                lets the streams sit anywhere in d_streams (index-entry (Offset, Size),
                                persist/schema/types.go:70-78); d_offsets then needs n_series entries */
  uint8_t *d_unit_first;     /* [n_series] unit in force at the first datapoint */
  m3tsz_dp_event *d_events;  /* event table, events_capacity entries */
  uint64_t events_capacity;
  uint64_t *d_event_count;   /* required with d_events; must be zero on entry; receives the number of
                                events produced (may exceed events_capacity: re-run with a larger table) */
  int32_t point_major;       /* != 0: d_ts / d_val are POINT-major, [max_points][n_series] (datapoint i of
                                series s at i * n_series + s) -- the step-major layout the query engine's
                                step iterators consume (src/query/storage/m3/encoded_step_iterator_generic.go),
                                like the downsample outputs.  Every decode step then stores 32 consecutive
                                elements per warp (coalesced 256-byte rows) instead of 32 separate 32-byte
                                sectors; rows >= n_points[s] of a series are left untouched. */
  int32_t reserved;
} m3tsz_decode_extras;

/* ------------------------------------------------------------------------
 * Batch decode.  Replaces, for n_series independent streams, the loop
 *   it := m3tsz.NewReaderIterator(reader, intOptimized, opts)   m3tsz/iterator.go:67-78
 *   for it.Next() { dp, unit, ann := it.Current() }             m3tsz/iterator.go:81-106,229-231
 * i.e. the per-series bodies of src/query/storage/prom_converter.go:65-110 and
 * src/query/storage/m3/encoded_series_iterator.go:94-120.
 *
 * d_streams : all streams concatenated (16-byte aligned base); stream s is
 *             bytes [d_offsets[s], d_offsets[s+1]) = ts.Segment head||tail.
 * d_ts/d_val: [n_series][max_points] (series-major); d_val receives float64.
 * d_n_points: datapoints decoded per series (may exceed max_points =>
 *             status M3TSZ_ERR_CAPACITY, first max_points are stored).
 * d_status  : per-series iterator Err() (M3TSZ_OK on a clean end-of-stream).
 * d_unit    : optional [n_series], time unit in force at the last datapoint.
 * d_ann     : optional [n_series], first annotation reference.
 * Limits (M3TSZ_ERR_INVALID_ARG beyond): streams_bytes < 2^34, max_points < 2^27,
 * one stream < 2^28 bytes (M3TSZ_ERR_STREAM_TOO_LARGE for that series); split
 * larger batches.  Stream starts may have any alignment; 64-byte aligned starts
 * (m3tsz_compact_streams with align = 64) decode fastest.  Outputs are stored as
 * whole 32-byte sectors when d_ts / d_val are 32-byte aligned and max_points is a
 * multiple of 4 (any other shape is stored row by row: same result, slower).
 * ---------------------------------------------------------------------- */
int m3tsz_decode_batch(m3tsz_ctx *ctx, const m3tsz_options *opts, const uint8_t *d_streams,
                       uint64_t streams_bytes, const uint64_t *d_offsets, uint64_t n_series,
                       int64_t *d_ts, double *d_val, uint64_t max_points, uint32_t *d_n_points,
                       int32_t *d_status, uint8_t *d_unit, m3tsz_annotation_ref *d_ann,
                       void *stream);

/* m3tsz_decode_batch + the optional per-datapoint unit / annotation events and
 * non-CSR stream placement (extras may be NULL). */
int m3tsz_decode_batch_ex(m3tsz_ctx *ctx, const m3tsz_options *opts, const uint8_t *d_streams,
                          uint64_t streams_bytes, const uint64_t *d_offsets, uint64_t n_series,
                          int64_t *d_ts, double *d_val, uint64_t max_points, uint32_t *d_n_points,
                          int32_t *d_status, uint8_t *d_unit, m3tsz_annotation_ref *d_ann,
                          const m3tsz_decode_extras *extras, void *stream);

/* Same call with HOST buffers (pageable or pinned): copies the streams and
 * offsets to the device, decodes, copies the results back, synchronises. */
int m3tsz_decode_batch_host(m3tsz_ctx *ctx, const m3tsz_options *opts, const uint8_t *h_streams,
                            uint64_t streams_bytes, const uint64_t *h_offsets, uint64_t n_series,
                            int64_t *h_ts, double *h_val, uint64_t max_points,
                            uint32_t *h_n_points, int32_t *h_status, uint8_t *h_unit,
                            m3tsz_annotation_ref *h_ann);

/* ------------------------------------------------------------------------
 * Batch encode.  Replaces, for n_series independent series, the loop
 *   enc.Reset(start, capacity, schema)                          m3tsz/encoder.go:262-279
 *   for each dp { enc.Encode(dp, unit, annotation) }            m3tsz/encoder.go:90-110
 *   seg := enc.Discard()                                        m3tsz/encoder.go:374-381
 * at the batch re-encode sites src/dbnode/storage/series/buffer.go:1543-1573,
 * 583-610, src/dbnode/persist/fs/merger.go:333-356.
 *
 * d_ts/d_val : [n_series][points_stride] series-major inputs.
 * d_n_points : optional [n_series]; NULL => every series has points_stride points.
 * d_start    : [n_series] encoder start (block start), ns.
 * unit       : time unit passed to every Encode call; d_units (optional,
 *              [n_series][points_stride] bytes) overrides it per datapoint.
 * annotations: optional sparse list: entries [d_ann_series_off[s], d_ann_series_off[s+1])
 *              belong to series s, sorted by dp_index; bytes in d_ann_bytes.
 * d_out      : [n_series][out_stride] bytes (out_stride % 16 == 0); series s
 *              gets its final stream (head||tail incl. end-of-stream marker)
 *              at d_out + s*out_stride, length d_out_len[s].
 * d_status   : per-series first Encode() error (encoding stops at that point;
 *              d_out_len then covers the datapoints before the error, like the
 *              reference's numEncoded), or M3TSZ_ERR_CAPACITY.
 * ---------------------------------------------------------------------- */
typedef struct m3tsz_annotation_entry {
  uint32_t dp_index;
  uint32_t length;
  uint64_t byte_offset; /* into d_ann_bytes */
} m3tsz_annotation_entry;

int m3tsz_encode_batch(m3tsz_ctx *ctx, const m3tsz_options *opts, const int64_t *d_ts,
                       const double *d_val, uint64_t n_series, uint64_t points_stride,
                       const uint32_t *d_n_points, const int64_t *d_start, int32_t unit,
                       const uint8_t *d_units, const uint64_t *d_ann_series_off,
                       const m3tsz_annotation_entry *d_ann_entries, const uint8_t *d_ann_bytes,
                       uint8_t *d_out, uint64_t out_stride, uint64_t *d_out_len, int32_t *d_status,
                       void *stream);

/* m3tsz_encode_batch + optional per-series encoder state that the reference's accessors
 * expose (extras may be NULL):
 *   d_last_value: Encoder.LastEncoded().Value (encoder.go:305-319) -- the float in float mode,
 *                 else the encoder's intVal: the SCALED integer in int mode and 0 when
 *                 int_optimized == 0 (the reference never sets isFloat there);
 *   point_major_input: see the struct;
 *   d_out_bits  : stream length in bits incl. the end-of-stream marker, before the zero
 *                 padding (splits the bytes into ts.Segment head / tail: the tail is the
 *                 last ceil((pos + 11) / 8) bytes, pos = ((bits - 12) mod 8) + 1,
 *                 scheme.go:198-211). */
typedef struct m3tsz_encode_extras {
  double *d_last_value;
  uint64_t *d_out_bits;
  int32_t point_major_input; /* != 0: d_ts / d_val are POINT-major, [points_stride][n_series] (datapoint i of
                                series s at i * n_series + s; the layout m3tsz_decode_batch_ex writes with
                                extras.point_major): every encode step reads 32 consecutive elements per warp,
                                no shared-memory staging.  Per-datapoint units and annotations are not
                                supported with it (M3TSZ_ERR_INVALID_ARG). */
  int32_t reserved;
} m3tsz_encode_extras;
int m3tsz_encode_batch_ex(m3tsz_ctx *ctx, const m3tsz_options *opts, const int64_t *d_ts,
                          const double *d_val, uint64_t n_series, uint64_t points_stride,
                          const uint32_t *d_n_points, const int64_t *d_start, int32_t unit,
                          const uint8_t *d_units, const uint64_t *d_ann_series_off,
                          const m3tsz_annotation_entry *d_ann_entries, const uint8_t *d_ann_bytes,
                          uint8_t *d_out, uint64_t out_stride, uint64_t *d_out_len, int32_t *d_status,
                          const m3tsz_encode_extras *extras, void *stream);

/* Worst-case stream bytes for n points without annotations (rounded up to 16),
 * for a batch-wide unit (m3tsz_encode_bound) or with per-datapoint units
 * (d_units given: every datapoint may add a time-unit marker and a raw 64-bit
 * delta-of-delta, timestamp_encoder.go:130-164,197-205). */
uint64_t m3tsz_encode_bound(uint64_t n_points);
uint64_t m3tsz_encode_bound_units(uint64_t n_points, int per_datapoint_units);

/* ------------------------------------------------------------------------
 * Batch encode straight into ONE packed buffer (the fileset data-file layout,
 * src/dbnode/persist/fs/write.go: concatenated segments + index entries that
 * carry (Offset, Size), persist/schema/types.go:70-78) -- no per-series slots,
 * no separate compaction pass.  Same inputs as m3tsz_encode_batch.  Stream s is
 *   d_packed[d_offsets[s] .. d_offsets[s] + d_out_len[s])
 * with every start rounded up to `align` bytes (1,4,8,16,32,64; >= 16 copies
 * fastest, 64 decodes fastest).  Streams are placed in completion order, NOT in
 * series order: d_offsets has n_series entries and is not monotonic; decode with
 * m3tsz_decode_batch_ex (extras.d_lengths = d_out_len).  *d_total_bytes (device)
 * receives the bytes used; a series that does not fit packed_capacity gets
 * M3TSZ_ERR_CAPACITY and length 0.  slot_bytes = per-series staging bound (0 =
 * m3tsz_encode_bound_units(points_stride, d_units != NULL); add the annotation
 * bytes of the largest series when annotations are given).  d_packed must be
 * 64-byte aligned.  The context keeps the staging slots (resident warps x 32 x
 * slot_bytes, ~2 GB for 1440-point series on a B200), reused by every call.
 * ---------------------------------------------------------------------- */
int m3tsz_encode_batch_packed(m3tsz_ctx *ctx, const m3tsz_options *opts, const int64_t *d_ts,
                              const double *d_val, uint64_t n_series, uint64_t points_stride,
                              const uint32_t *d_n_points, const int64_t *d_start, int32_t unit,
                              const uint8_t *d_units, const uint64_t *d_ann_series_off,
                              const m3tsz_annotation_entry *d_ann_entries, const uint8_t *d_ann_bytes,
                              uint64_t slot_bytes, uint32_t align, uint8_t *d_packed,
                              uint64_t packed_capacity, uint64_t *d_offsets, uint64_t *d_out_len,
                              int32_t *d_status, uint64_t *d_total_bytes, void *stream);
/* ... with the optional extras of m3tsz_encode_batch_ex: point-major inputs
 * (extras->point_major_input: d_ts / d_val are [points_stride][n_series], no per-datapoint
 * units / annotations), LastEncoded values, stream bit lengths. */
int m3tsz_encode_batch_packed_ex(m3tsz_ctx *ctx, const m3tsz_options *opts, const int64_t *d_ts,
                                 const double *d_val, uint64_t n_series, uint64_t points_stride,
                                 const uint32_t *d_n_points, const int64_t *d_start, int32_t unit,
                                 const uint8_t *d_units, const uint64_t *d_ann_series_off,
                                 const m3tsz_annotation_entry *d_ann_entries, const uint8_t *d_ann_bytes,
                                 uint64_t slot_bytes, uint32_t align, uint8_t *d_packed,
                                 uint64_t packed_capacity, uint64_t *d_offsets, uint64_t *d_out_len,
                                 int32_t *d_status, uint64_t *d_total_bytes,
                                 const m3tsz_encode_extras *extras, void *stream);

/* Packs the per-series slots written by m3tsz_encode_batch into one contiguous
 * buffer (the fileset data-file layout, src/dbnode/persist/fs/write.go): fills
 * d_offsets[n_series+1] (exclusive prefix sum of d_out_len, each start rounded
 * up to `align` bytes, align in {1,4,8,16,32,64}) and copies the bytes.  64-byte
 * aligned starts let the decoder's lanes read their staging rings in phase (fewer
 * shared-memory bank conflicts); any alignment decodes to the same result. */
int m3tsz_compact_streams(m3tsz_ctx *ctx, const uint8_t *d_slots, uint64_t slot_stride,
                          const uint64_t *d_len, uint64_t n_series, uint32_t align,
                          uint8_t *d_packed, uint64_t packed_capacity, uint64_t *d_offsets,
                          void *stream);

/* Host-buffer encode: copies the inputs to the device, encodes, packs the
 * streams on the device (m3tsz_compact_streams, `align`-byte aligned starts)
 * and copies back ONE contiguous buffer + CSR offsets -- the fileset data-file
 * layout -- so that only compressed bytes cross PCIe.  h_offsets has
 * n_series+1 entries; stream s is h_packed[h_offsets[s] .. +h_out_len[s]). */
int m3tsz_encode_batch_host(m3tsz_ctx *ctx, const m3tsz_options *opts, const int64_t *h_ts,
                            const double *h_val, uint64_t n_series, uint64_t points_stride,
                            const uint32_t *h_n_points, const int64_t *h_start, int32_t unit,
                            const uint8_t *h_units, const uint64_t *h_ann_series_off,
                            const m3tsz_annotation_entry *h_ann_entries,
                            const uint8_t *h_ann_bytes, uint64_t ann_bytes_len, uint32_t align,
                            uint8_t *h_packed, uint64_t packed_capacity, uint64_t *h_offsets,
                            uint64_t *h_out_len, int32_t *h_status);

/* ------------------------------------------------------------------------
 * Fused decode + downsample (BASELINE config 4).  Decodes every stream and
 * folds each datapoint into fixed windows with the reference aggregator's
 * Gauge arithmetic: aggregation.Gauge.updateTotals
 * (src/aggregator/aggregation/gauge.go:73-106: count++ always; NaN skipped for
 * sum/min/max; min/max start NaN; sum added in arrival order), window index =
 * floor((ts - range_start) / window) as timestamp.Truncate(resolution) does
 * (src/aggregator/aggregator/generic_elem.go:220).  Datapoints outside
 * [range_start, range_start + n_windows*window) are ignored.
 *
 * Outputs are WINDOW-major: element (series s, window w) at [w*n_series + s]
 * (step-major, the layout the query engine's step iterators consume,
 * src/query/storage/m3/encoded_step_iterator_generic.go).
 * ---------------------------------------------------------------------- */
int m3tsz_decode_downsample_batch(m3tsz_ctx *ctx, const m3tsz_options *opts,
                                  const uint8_t *d_streams, uint64_t streams_bytes,
                                  const uint64_t *d_offsets, uint64_t n_series,
                                  int64_t range_start_ns, int64_t window_ns, uint32_t n_windows,
                                  double *d_sum, int64_t *d_count, double *d_min, double *d_max,
                                  uint32_t *d_n_points, int32_t *d_status, void *stream);

/* Same, plus Gauge.Last() per window (gauge.go:73-84: the value with the latest
 * timestamp, first arrival among equal timestamps; 0 for an empty window) in
 * d_last, and optionally Gauge.LastAt() (ns; 0 for an empty window) in d_last_at
 * (NULL: the context keeps it in scratch -- it is what lets an out-of-order
 * datapoint re-open a committed window exactly).  Gauge.Mean() = sum / count
 * (0 when count == 0) is left to the consumer (gauge.go:117-122). */
int m3tsz_decode_downsample_last_batch(m3tsz_ctx *ctx, const m3tsz_options *opts,
                                       const uint8_t *d_streams, uint64_t streams_bytes,
                                       const uint64_t *d_offsets, uint64_t n_series,
                                       int64_t range_start_ns, int64_t window_ns, uint32_t n_windows,
                                       double *d_sum, int64_t *d_count, double *d_min, double *d_max,
                                       double *d_last, int64_t *d_last_at, uint32_t *d_n_points,
                                       int32_t *d_status, void *stream);

int m3tsz_decode_downsample_batch_host(m3tsz_ctx *ctx, const m3tsz_options *opts,
                                       const uint8_t *h_streams, uint64_t streams_bytes,
                                       const uint64_t *h_offsets, uint64_t n_series,
                                       int64_t range_start_ns, int64_t window_ns,
                                       uint32_t n_windows, double *h_sum, int64_t *h_count,
                                       double *h_min, double *h_max, uint32_t *h_n_points,
                                       int32_t *h_status);

/* ------------------------------------------------------------------------
 * Segment checksums (fileset ingestion).  The fileset data file
 * is the concatenated streams and every index entry carries (Offset, Size,
 * DataChecksum) (src/dbnode/persist/schema/types.go:70-78) -- i.e. this library's
 * CSR layout plus the Adler-32 that the reference verifies on every read:
 *   ts.Segment.CalculateChecksum  src/dbnode/ts/segment.go:60-76  (head then tail)
 *   digest.Checksum               src/dbnode/digest/digest.go:36-38
 *   check                         src/dbnode/persist/fs/read.go:395-397, seek.go:370-373
 * Computes the Adler-32 of every stream [d_offsets[s], d_offsets[s+1]) -- or
 * [d_offsets[s], d_offsets[s] + d_lengths[s]) when d_lengths is given (index-entry
 * (Offset, Size) addressing: any placement and order, d_offsets then needs only
 * n_series entries) -- into d_checksums (optional) and, when d_expected (the index entries'
 * DataChecksum) is given, sets d_status[s] to M3TSZ_ERR_CHECKSUM_MISMATCH where
 * it differs.  One of d_checksums / d_status is required.
 * ---------------------------------------------------------------------- */
int m3tsz_checksum_batch(m3tsz_ctx *ctx, const uint8_t *d_streams, uint64_t streams_bytes,
                         const uint64_t *d_offsets, const uint64_t *d_lengths, uint64_t n_series,
                         const uint32_t *d_expected, uint32_t *d_checksums, int32_t *d_status,
                         void *stream);

/* ------------------------------------------------------------------------
 * Series merge: the iterator layer directly above the codec.
 * For every series, merges the DECODED streams of its replicas / blocks exactly
 * like the reference's seriesIterator over multiReaderIterators:
 *   iterators            src/dbnode/encoding/iterators.go:56-262
 *   multiReaderIterator  src/dbnode/encoding/multi_reader_iterator.go:62-155
 *   seriesIterator       src/dbnode/encoding/series_iterator.go:74-83,129-215
 * i.e. k-way timestamp merge, equal-timestamp strategy (0 last-pushed = default,
 * 1 highest value, 2 lowest value, 3 highest frequency; encoding/iterators_types.go),
 * removal of consecutive equal timestamps, [start_ns, end_ns) filter (both 0 = no
 * filter), errOutOfOrderIterator, propagation of reader (decode) errors.
 *
 * Inputs are the outputs of m3tsz_decode_batch over n_seq streams:
 *   sequence q = (d_ts + q*cap, d_val + q*cap, d_n_points[q], d_seq_status[q])
 *   slice k    = sequences [d_slice_off[k], d_slice_off[k+1])  readers of one block
 *   replica r  = slices    [d_replica_off[r], d_replica_off[r+1])  in block order
 *   series s   = replicas  [d_series_off[s], d_series_off[s+1])
 * Outputs: merged datapoints [n_series][out_cap], count, status per series.
 * ---------------------------------------------------------------------- */
int m3tsz_merge_series_batch(m3tsz_ctx *ctx, const int64_t *d_ts, const double *d_val, uint64_t cap,
                             const uint32_t *d_n_points, const int32_t *d_seq_status,
                             const uint64_t *d_slice_off, const uint64_t *d_replica_off,
                             const uint64_t *d_series_off, uint64_t n_series, int64_t start_ns,
                             int64_t end_ns, int32_t strategy, int64_t *d_ts_out, double *d_val_out,
                             uint64_t out_cap, uint32_t *d_n_out, int32_t *d_status, void *stream);

/* ------------------------------------------------------------------------
 * Fetch with HOST buffers: decode + series merge in one call, so that only the
 * COMPRESSED replica streams go up over PCIe and only the MERGED series come back
 * (the decoded replicas, 16 B per datapoint and replica, never leave the device).
 * Same structure arguments as m3tsz_merge_series_batch, over the streams
 * h_offsets[q] .. h_offsets[q+1] of sequence q (sequences of one series must be
 * consecutive: the call works through chunks of whole series on two CUDA streams,
 * H2D / decode / merge / D2H overlapped).  max_points = decode capacity per
 * sequence; outputs [n_series][out_cap] + h_n_out + h_status as m3tsz_merge_series_batch.
 * ---------------------------------------------------------------------- */
int m3tsz_fetch_batch_host(m3tsz_ctx *ctx, const m3tsz_options *opts, const uint8_t *h_streams,
                           uint64_t streams_bytes, const uint64_t *h_offsets, uint64_t n_seq,
                           const uint64_t *h_slice_off, const uint64_t *h_replica_off,
                           const uint64_t *h_series_off, uint64_t n_series, uint64_t max_points,
                           int64_t start_ns, int64_t end_ns, int32_t strategy, int64_t *h_ts_out,
                           double *h_val_out, uint64_t out_cap, uint32_t *h_n_out, int32_t *h_status);

/* ------------------------------------------------------------------------
 * Prometheus conversion epilogue: iteratorToPromResult,
 * src/query/storage/prom_converter.go:42-120, over decoded / merged series that are
 * already in HBM -- the last per-datapoint host loop of a fetch:
 *   - timestamps ns -> ms (TimeToPromTimestamp, converter.go:388-391, truncating);
 *   - value_decrease_tolerance > 0: a value that dips below its predecessor by less
 *     than the tolerance, before tolerance_until_ns, is replaced by the predecessor
 *     (the replaced value is the next predecessor), :68-72;
 *   - d_handle_resets[s] != 0 (the host sets it when the series' first annotation says
 *     OpenMetricsHandleValueResets and maxResolution >= the normalisation threshold,
 *     :74-82): one sample per resolution_ns window holding the reset-aware cumulative
 *     sum, stamped with the window's last datapoint, :84-98,113-118.
 * Inputs [n_series][cap] + d_n_points; outputs [n_series][out_cap] + d_n_out (the
 * number of samples; > out_cap => d_status M3TSZ_ERR_CAPACITY, first out_cap stored).
 * d_handle_resets and d_status may be NULL.  In-place (outputs == inputs, out_cap ==
 * cap) is allowed.
 * ---------------------------------------------------------------------- */
int m3tsz_prom_convert_batch(m3tsz_ctx *ctx, const int64_t *d_ts, const double *d_val, uint64_t cap,
                             const uint32_t *d_n_points, uint64_t n_series, int64_t resolution_ns,
                             const uint8_t *d_handle_resets, double value_decrease_tolerance,
                             int64_t tolerance_until_ns, int64_t *d_ts_ms_out, double *d_val_out,
                             uint64_t out_cap, uint32_t *d_n_out, int32_t *d_status, void *stream);

/* ------------------------------------------------------------------------
 * Tile aggregation: the compute of a storage.TileAggregator
 * (src/dbnode/storage/types.go:1444-1472, AggregateTilesOptions{Start, End, Step};
 * the open-source default is a no-op, storage/options.go:949-951).  For every
 * source stream: decode, fold the datapoints of [start_ns, start_ns + n_windows *
 * step_ns) into step-sized windows with the aggregator's Gauge
 * (aggregation/gauge.go:73-106), emit one datapoint per NON-EMPTY window --
 * timestamp = the window's end boundary (aggregator/list.go:541-543), value =
 * Gauge.ValueOf(agg_type) (gauge.go:144-165) -- and re-encode them as the target
 * namespace's M3TSZ stream (encoder start = start_ns, unit = out_unit), straight
 * into one packed buffer (layout of m3tsz_encode_batch_packed: d_out_offsets[s],
 * d_out_len[s], *d_total_bytes).  d_lengths may be NULL (CSR offsets).
 * d_status[s]: the source stream's decode error if any (no output for that series),
 * else the encoder's status.  d_n_tiles (optional): datapoints written per series.
 * ---------------------------------------------------------------------- */
enum { /* aggregation.Type ids, src/metrics/aggregation/type.go:31-38 */
  M3TSZ_AGG_LAST = 1,
  M3TSZ_AGG_MIN = 2,
  M3TSZ_AGG_MAX = 3,
  M3TSZ_AGG_MEAN = 4,
  M3TSZ_AGG_COUNT = 6,
  M3TSZ_AGG_SUM = 7
};
int m3tsz_aggregate_tiles_batch(m3tsz_ctx *ctx, const m3tsz_options *opts, const uint8_t *d_streams,
                                uint64_t streams_bytes, const uint64_t *d_offsets,
                                const uint64_t *d_lengths, uint64_t n_series, int64_t start_ns,
                                int64_t step_ns, uint32_t n_windows, int32_t agg_type, int32_t out_unit,
                                uint32_t align, uint8_t *d_packed, uint64_t packed_capacity,
                                uint64_t *d_out_offsets, uint64_t *d_out_len, int32_t *d_status,
                                uint32_t *d_n_tiles, uint64_t *d_total_bytes, void *stream);

/* ------------------------------------------------------------------------
 * Per-series streaming handles: what a cgo shim binds METHOD FOR METHOD to the
 * reference's interfaces (INTEGRATION.md §2 is that shim):
 *   encoding.Encoder         src/dbnode/encoding/types.go:39-91   -> m3tsz_encoder_*
 *   encoding.ReaderIterator  types.go:180-203, Decoder :342-345   -> m3tsz_iter_*
 *   EncoderPool / ReaderIteratorPool  encoder_pool.go:27-48, iterator_pool.go:27-47 -> *_pool_*
 * The handles keep host state only; bitstreams are produced / consumed by the
 * batch kernels (one series per launch) lazily: an encoder when Len / Stream /
 * Discard / LastEncoded is called, an iterator on its first Next().  Handles of one
 * context serialise on the context's scratch; a handle itself is single-threaded like
 * the reference's objects.  Error conventions are the reference's:
 *   Encode on a closed encoder -> M3TSZ_ERR_ENCODER_CLOSED; a second/millisecond
 *   delta-of-delta that overflows 32 bits -> M3TSZ_ERR_DOD_OVERFLOW; invalid unit ->
 *   M3TSZ_ERR_UNRECOGNIZED_UNIT (the failing datapoint is dropped as a whole);
 *   LastEncoded / LastAnnotationChecksum on an empty encoder -> M3TSZ_ERR_NO_DATAPOINTS;
 *   iterator Err(): sticky status, visible after the Next() that failed; closed ->
 *   M3TSZ_ERR_ITER_CLOSED.
 * ---------------------------------------------------------------------- */
typedef struct m3tsz_encoder m3tsz_encoder;
typedef struct m3tsz_iter m3tsz_iter;
typedef struct m3tsz_encoder_pool m3tsz_encoder_pool;
typedef struct m3tsz_iter_pool m3tsz_iter_pool;

/* m3tsz.NewEncoder(start, nil, intOptimized, opts), m3tsz/encoder.go:64-85 */
int m3tsz_encoder_create(m3tsz_ctx *ctx, const m3tsz_options *opts, int64_t start_ns, m3tsz_encoder **out);
void m3tsz_encoder_destroy(m3tsz_encoder *enc);
/* Reset(start, capacity, schema), encoder.go:262-279 (re-opens a closed encoder) */
int m3tsz_encoder_reset(m3tsz_encoder *enc, int64_t start_ns, uint64_t capacity);
/* Encode(dp, unit, annotation), encoder.go:90-110 */
int m3tsz_encoder_encode(m3tsz_encoder *enc, int64_t ts_ns, double value, int32_t unit,
                         const uint8_t *annotation, uint64_t annotation_len);
/* the delta-of-delta (in time units) of the last Encode that returned M3TSZ_ERR_DOD_OVERFLOW: the
 * value the reference formats into its error ("deltaOfDelta value %d %s overflows 32 bits") */
int64_t m3tsz_encoder_failed_dod(const m3tsz_encoder *enc);
uint64_t m3tsz_encoder_num_encoded(const m3tsz_encoder *enc); /* NumEncoded, :299-302 */
/* LastEncoded, :305-319: PrevTime and -- the reference's quirk, kept -- the float value in
 * float mode, else the encoder's intVal: the SCALED integer in int mode, 0 when int_optimized == 0 */
int m3tsz_encoder_last_encoded(m3tsz_encoder *enc, int64_t *ts_ns, double *value);
/* LastAnnotationChecksum, :321-327: XXH64 (cespare/xxhash/v2) of the last annotation written */
int m3tsz_encoder_last_annotation_checksum(const m3tsz_encoder *enc, uint64_t *checksum);
int m3tsz_encoder_empty(const m3tsz_encoder *enc);              /* Empty, :330-332 */
int m3tsz_encoder_len(m3tsz_encoder *enc, uint64_t *len);        /* Len, :336-354 (with the tail) */
/* Stream(ctx), :282-297: copies head||tail into buf; *len == 0 <=> (nil, false).  *tail_len
 * (optional) = bytes of the ts.Segment tail, so head = buf[0 .. len - tail_len). */
int m3tsz_encoder_stream(m3tsz_encoder *enc, uint8_t *buf, uint64_t cap, uint64_t *len, uint64_t *tail_len);
int m3tsz_encoder_close(m3tsz_encoder *enc);                    /* Close, :357-370 (idempotent) */
/* Discard, :374-381 (segment + Close) / DiscardReset, :385-392 (segment + Reset) */
int m3tsz_encoder_discard(m3tsz_encoder *enc, uint8_t *buf, uint64_t cap, uint64_t *len, uint64_t *tail_len);
int m3tsz_encoder_discard_reset(m3tsz_encoder *enc, int64_t start_ns, uint64_t capacity, uint8_t *buf,
                                uint64_t cap, uint64_t *len, uint64_t *tail_len);

/* m3tsz.NewReaderIterator(nil, intOptimized, opts), m3tsz/iterator.go:67-78 */
int m3tsz_iter_create(m3tsz_ctx *ctx, const m3tsz_options *opts, m3tsz_iter **out);
void m3tsz_iter_destroy(m3tsz_iter *it);
/* Reset(reader, schema), iterator.go:253-263: the handle copies the stream bytes */
int m3tsz_iter_reset(m3tsz_iter *it, const uint8_t *data, uint64_t len);
int m3tsz_iter_next(m3tsz_iter *it); /* Next, :81-106: 1 while a datapoint is current, else 0 */
/* Current, :229-231: datapoint, the unit in force AT THIS datapoint, and its annotation
 * (valid until the next call on this iterator; NULL / 0 when the datapoint has none) */
int m3tsz_iter_current(const m3tsz_iter *it, int64_t *ts_ns, double *value, int32_t *unit,
                       const uint8_t **annotation, uint64_t *annotation_len);
int m3tsz_iter_err(const m3tsz_iter *it); /* Err, :234-236 */
int m3tsz_iter_close(m3tsz_iter *it);     /* Close, :267-278 */

/* EncoderPool.Init(alloc) / Get() / Put via Close(), encoder_pool.go:27-48 */
int m3tsz_encoder_pool_create(m3tsz_ctx *ctx, const m3tsz_options *opts, uint64_t size,
                              m3tsz_encoder_pool **out);
int m3tsz_encoder_pool_get(m3tsz_encoder_pool *pool, m3tsz_encoder **out);
void m3tsz_encoder_pool_destroy(m3tsz_encoder_pool *pool);
/* ReaderIteratorPool, iterator_pool.go:27-47 (the pooled iterators are created with a nil reader) */
int m3tsz_iter_pool_create(m3tsz_ctx *ctx, const m3tsz_options *opts, uint64_t size, m3tsz_iter_pool **out);
int m3tsz_iter_pool_get(m3tsz_iter_pool *pool, m3tsz_iter **out);
void m3tsz_iter_pool_destroy(m3tsz_iter_pool *pool);

/* ------------------------------------------------------------------------
 * Fetch-side exchange (BASELINE config 5): series shard trivially, so
 * encode / decode need no collective; only a query that spans shards needs every
 * shard's DECODED blocks on every GPU.  m3tsz_allgather_decoded decodes the first
 * gather_series streams of the local shard chunk by chunk and all-gathers chunk k-1
 * over NCCL / NVLink while chunk k+1 is decoded: every chunk is decoded straight into this
 * rank's block of the gathered arrays and gathered IN PLACE (one grouped ncclAllGather per
 * chunk, no staging copies).  Every rank passes the same gather_series / max_points /
 * chunk_series (gather_series a multiple of chunk_series; 0 = 32768).  Outputs (on every
 * rank); the datapoint arrays are CHUNK-major, the layout the collective writes without a
 * second pass:
 *   d_ts_all / d_val_all : [n_chunks][n_ranks][chunk_series][max_points]; series s of rank r
 *                          sits at (((s / C) * n_ranks + r) * C + s % C) * max_points, C = chunk_series
 *   d_n_points_all / d_status_all : [n_ranks][gather_series]
 * nccl_comm is an ncclComm_t of n_ranks ranks (the caller's, or one made with
 * m3tsz_nccl_comm_create from an id broadcast out of band).  NCCL is resolved at run
 * time (dlopen of the libnccl.so.2 the process has loaded); without it the calls
 * return M3TSZ_ERR_NO_DEVICE.  d_lengths may be NULL (CSR offsets).  Asynchronous:
 * `stream` continues after the pipeline has finished.
 * ---------------------------------------------------------------------- */
int m3tsz_nccl_unique_id(uint8_t *out128);
int m3tsz_nccl_comm_create(m3tsz_ctx *ctx, const uint8_t *unique_id128, int n_ranks, int rank, void **comm);
int m3tsz_nccl_comm_destroy(void *comm);
int m3tsz_allgather_decoded(m3tsz_ctx *ctx, const m3tsz_options *opts, void *nccl_comm, int n_ranks,
                            const uint8_t *d_streams, uint64_t streams_bytes, const uint64_t *d_offsets,
                            const uint64_t *d_lengths, uint64_t n_series, uint64_t gather_series,
                            uint64_t max_points, uint64_t chunk_series, int64_t *d_ts_all, double *d_val_all,
                            uint32_t *d_n_points_all, int32_t *d_status_all, void *stream);

/* m3tsz_merge_series_batch with POINT-major arrays: point_major_in: d_ts / d_val are
 * [cap][n_seq] (datapoint i of sequence q at i * n_seq + q, what m3tsz_decode_batch_ex writes with
 * extras.point_major); point_major_out: outputs are [out_cap][n_series].  With replicas that
 * advance in step (the normal RF = 3 fetch) the threads of a warp then read and write
 * consecutive elements. */
int m3tsz_merge_series_batch_ex(m3tsz_ctx *ctx, const int64_t *d_ts, const double *d_val, uint64_t cap,
                                const uint32_t *d_n_points, const int32_t *d_seq_status,
                                const uint64_t *d_slice_off, const uint64_t *d_replica_off,
                                const uint64_t *d_series_off, uint64_t n_series, int64_t start_ns,
                                int64_t end_ns, int32_t strategy, int64_t *d_ts_out, double *d_val_out,
                                uint64_t out_cap, uint32_t *d_n_out, int32_t *d_status, uint64_t n_seq,
                                int32_t point_major_in, int32_t point_major_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* M3TSZ_B200_H */
