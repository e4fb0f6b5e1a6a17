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
  M3TSZ_ERR_TOO_MANY_ITERATORS = 14, /* > 8 replicas per series or readers per block slice */
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

/* Worst-case stream bytes for n points without annotations (rounded up to 16). */
uint64_t m3tsz_encode_bound(uint64_t n_points);

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
 * [d_offsets[s], d_offsets[s] + d_lengths[s]) when d_lengths is given (padded
 * starts) -- into d_checksums (optional) and, when d_expected (the index entries'
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

#ifdef __cplusplus
}
#endif
#endif /* M3TSZ_B200_H */
