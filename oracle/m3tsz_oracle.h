/*
 * m3tsz_oracle.h -- CPU restatement of the reference M3TSZ codec.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle: a plain-C restatement
 * of the algorithm in m3db/m3 `src/dbnode/encoding/m3tsz` (Go).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may call into it.  The product (m3_b200/, libm3tsz_b200.so) never links,
 * imports or executes anything in this directory.
 *
 * Parity pinning: every golden vector the reference's own tests hold for this
 * path is transcribed into tests/golden/ and checked by tests/test_oracle_*.py
 * (see tests/golden/README.md for file:line provenance).  The reference itself
 * is Go and cannot be built in this image (no Go toolchain, no module cache),
 * so there is no oracle/_ref; the goldens are the pin.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference/src/dbnode/encoding unless noted).
 */
#ifndef M3TSZ_ORACLE_H
#define M3TSZ_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* xtime.Unit enum, src/x/time/unit.go:30-42 */
enum {
  M3O_UNIT_NONE = 0,
  M3O_UNIT_SECOND = 1,
  M3O_UNIT_MILLISECOND = 2,
  M3O_UNIT_MICROSECOND = 3,
  M3O_UNIT_NANOSECOND = 4,
  M3O_UNIT_MINUTE = 5,
  M3O_UNIT_HOUR = 6,
  M3O_UNIT_DAY = 7,
  M3O_UNIT_YEAR = 8,
  M3O_UNIT_COUNT = 9
};

/* error codes (0 = nil error) */
enum {
  M3O_OK = 0,
  M3O_ERR_EOF = 1,               /* io.EOF from IStream, istream.go:86-92 */
  M3O_ERR_ENCODER_CLOSED = 2,    /* errEncoderClosed, m3tsz/encoder.go:37 */
  M3O_ERR_NO_DATAPOINTS = 3,     /* errNoEncodedDatapoints, m3tsz/encoder.go:38 */
  M3O_ERR_DOD_OVERFLOW = 4,      /* "deltaOfDelta value %d %s overflows 32 bits", m3tsz/timestamp_encoder.go:219 */
  M3O_ERR_NO_TIME_SCHEME = 5,    /* errNoTimeSchemaForUnit, m3tsz/timestamp_iterator.go:33 */
  M3O_ERR_UNRECOGNIZED_UNIT = 6, /* errUnrecognizedTimeUnit, src/x/time/unit.go:45 */
  M3O_ERR_INVALID_MULT = 7,      /* errInvalidMultiplier, m3tsz/m3tsz.go:69 */
  M3O_ERR_ANNOTATION_LEN = 8,    /* errUnexpectedAnnotationLength, m3tsz/timestamp_iterator.go:34 */
  M3O_ERR_ANNOTATION_SHORT = 9,  /* errAnnotationTooFewBytes, m3tsz/timestamp_iterator.go:35 */
  M3O_ERR_ITER_CLOSED = 10,      /* errClosed, m3tsz/iterator.go:33 */
  M3O_ERR_VARINT_OVERFLOW = 11,  /* encoding/binary errOverflow (Go stdlib ReadUvarint) */
  M3O_ERR_UNEXPECTED_EOF = 12,   /* io.ErrUnexpectedEOF (Go stdlib ReadUvarint, i>0) */
  M3O_ERR_OUT_OF_ORDER = 13,     /* errOutOfOrderIterator, encoding/iterators_types.go / iterators.go:229-236 */
  M3O_ERR_TOO_MANY_ITERATORS = 14, /* more than 12 iterators at one level (restatement limit) */
  M3O_ERR_CHECKSUM = 15           /* errSeekChecksumMismatch / errReadChecksum..., persist/fs/read.go:395-397 */
};

/* ---- bit I/O (ostream.go / istream.go) exposed for the golden bit-I/O tests ---- */
typedef struct m3o_ostream m3o_ostream;
m3o_ostream *m3o_ostream_new(void);
void m3o_ostream_free(m3o_ostream *os);
void m3o_ostream_reset(m3o_ostream *os);
void m3o_ostream_write_bit(m3o_ostream *os, int bit);
void m3o_ostream_write_byte(m3o_ostream *os, uint8_t b);
void m3o_ostream_write_bytes(m3o_ostream *os, const uint8_t *p, size_t n);
void m3o_ostream_write_bits(m3o_ostream *os, uint64_t v, int nbits);
/* RawBytes(): returns length, *pos = bits used in last byte (0 if empty) */
size_t m3o_ostream_raw(const m3o_ostream *os, const uint8_t **data, int *pos);

typedef struct m3o_istream m3o_istream;
m3o_istream *m3o_istream_new(const uint8_t *data, size_t len);
void m3o_istream_free(m3o_istream *is);
int m3o_istream_read_bits(m3o_istream *is, int nbits, uint64_t *out);
int m3o_istream_peek_bits(m3o_istream *is, int nbits, uint64_t *out);
int m3o_istream_remaining_bits_in_current_byte(const m3o_istream *is);

/* ---- helpers (encoding.go:29-49, m3tsz/m3tsz.go:78-127) ---- */
int m3o_num_sig(uint64_t v);
void m3o_leading_trailing_zeros(uint64_t v, int *lz, int *tz);
int64_t m3o_sign_extend(uint64_t v, int nbits);
/* returns error code; outputs val, mult, is_float */
int m3o_convert_to_int_float(double v, int cur_max_mult, double *val, int *mult, int *is_float);
double m3o_convert_from_int_float(double val, int mult);
int m3o_initial_time_unit(int64_t start_ns, int unit);
uint64_t m3o_xxh64(const uint8_t *p, size_t n); /* cespare/xxhash/v2 Sum64 (XXH64 seed 0) */

/* ---- field-level writers for the golden field tests (encoder_test.go:54-205) ---- */
int m3o_write_dod_unit_unchanged(m3o_ostream *os, int64_t prev_delta, int64_t cur_delta, int unit);
void m3o_write_dod_unit_changed(m3o_ostream *os, int64_t prev_delta, int64_t cur_delta);
void m3o_write_xor(m3o_ostream *os, uint64_t prev_xor, uint64_t cur_xor);

/* ---- encoder (m3tsz/encoder.go) ---- */
typedef struct m3o_encoder m3o_encoder;
/* NewEncoder(start, nil, intOptimized, opts) with opts.DefaultTimeUnit = default_unit */
m3o_encoder *m3o_encoder_new(int64_t start_ns, int int_optimized, int default_unit);
void m3o_encoder_free(m3o_encoder *e);
void m3o_encoder_reset(m3o_encoder *e, int64_t start_ns);
int m3o_encoder_encode(m3o_encoder *e, int64_t ts_ns, double value, int unit,
                       const uint8_t *ann, size_t ann_len);
int m3o_encoder_num_encoded(const m3o_encoder *e);
int m3o_encoder_last_encoded(const m3o_encoder *e, int64_t *ts_ns, double *value);
int m3o_encoder_last_annotation_checksum(const m3o_encoder *e, uint64_t *sum);
size_t m3o_encoder_len(const m3o_encoder *e);
int m3o_encoder_empty(const m3o_encoder *e);
/* Stream(): copies head+tail into out (cap bytes); returns total length (0 => (nil,false)) */
size_t m3o_encoder_stream(const m3o_encoder *e, uint8_t *out, size_t cap);
size_t m3o_encoder_raw(const m3o_encoder *e, const uint8_t **data, int *pos);
void m3o_encoder_close(m3o_encoder *e);

/* ---- reader iterator (m3tsz/iterator.go) ---- */
typedef struct m3o_iter m3o_iter;
m3o_iter *m3o_iter_new(const uint8_t *data, size_t len, int int_optimized, int default_unit);
void m3o_iter_free(m3o_iter *it);
void m3o_iter_reset(m3o_iter *it, const uint8_t *data, size_t len);
int m3o_iter_next(m3o_iter *it); /* 1 = true */
void m3o_iter_current(const m3o_iter *it, int64_t *ts_ns, double *value, int *unit,
                      const uint8_t **ann, size_t *ann_len);
int m3o_iter_err(const m3o_iter *it);
int m3o_iter_done(const m3o_iter *it);
/* test hooks mirroring iterator_test.go's white-box cases */
void m3o_iter_set_float_state(m3o_iter *it, uint64_t prev_bits, uint64_t prev_xor);
void m3o_iter_get_float_state(const m3o_iter *it, uint64_t *prev_bits, uint64_t *prev_xor);
void m3o_iter_read_next_value(m3o_iter *it);
void m3o_iter_set_ts_state(m3o_iter *it, int unit, int64_t prev_delta);
int m3o_iter_read_next_timestamp(m3o_iter *it);
int m3o_iter_read_first_timestamp(m3o_iter *it);
int64_t m3o_iter_prev_time_delta(const m3o_iter *it);
int m3o_iter_read_annotation(m3o_iter *it, const uint8_t **ann, size_t *ann_len);
int m3o_iter_read_time_unit(m3o_iter *it, int *unit, int *changed);

/* ---- whole-series convenience (one Encode loop / one Next loop) ---- */
/* Encodes n datapoints (constant unit, no annotations) and writes the final
 * segment (head+tail, = Discard() bytes).  Returns byte length, or -(err) on
 * error, or -1000 if cap too small. */
int64_t m3o_encode_series(const int64_t *ts, const double *vals, size_t n, int64_t start_ns,
                          int unit, int int_optimized, int default_unit, uint8_t *out,
                          size_t cap);
/* Decodes a stream.  Returns number of datapoints, *err = final iterator error. */
int64_t m3o_decode_series(const uint8_t *data, size_t len, int int_optimized, int default_unit,
                          int64_t *ts_out, double *val_out, size_t cap, int *err);

/* ---- multi-threaded batch (one series per task; mirrors the one-goroutine-per-series
 * fan-out at src/query/storage/prom_converter.go:170-215).  Used as the CPU baseline. ---- */
/* streams: concatenated, off[S+1] byte offsets; outputs fixed stride `cap` points */
int m3o_decode_batch(const uint8_t *streams, const uint64_t *off, size_t n_series,
                     int int_optimized, int default_unit, int64_t *ts_out, double *val_out,
                     size_t cap, uint32_t *n_points, int32_t *status, int n_threads);
/* ts/vals: [S][n_per] ; out: [S][out_stride] bytes; out_len[S] */
int m3o_encode_batch(const int64_t *ts, const double *vals, size_t n_series, size_t n_per,
                     const int64_t *start_ns, int unit, int int_optimized, int default_unit,
                     uint8_t *out, size_t out_stride, uint64_t *out_len, int32_t *status,
                     int n_threads);

/* ---- downsample oracle: aggregation.Gauge semantics
 * (src/aggregator/aggregation/gauge.go:45-106; window = ts.Truncate(resolution),
 * src/aggregator/aggregator/generic_elem.go:220).  Windows are indexed relative
 * to `range_start_ns` (must be window-aligned): w = floor((ts - range_start)/window).
 * Datapoints outside [0, n_windows) are ignored.  Outputs per window:
 * sum, count, min, max, last.  Initial: sum 0, count 0, min/max NaN, last 0. ---- */
void m3o_downsample_series(const int64_t *ts, const double *vals, size_t n, int64_t range_start_ns,
                           int64_t window_ns, size_t n_windows, double *sum, int64_t *count,
                           double *min, double *max, double *last);

/* ---- iterator layer above the codec (m3tsz_merge_oracle.c):
 * iterators / multiReaderIterator / seriesIterator over decoded reader sequences.
 * Sequence q = (ts + q*cap, val + q*cap, n_points[q], seq_status[q]); slice k = sequences
 * [slice_off[k], slice_off[k+1]) (readers of one block); replica r = slices
 * [replica_off[r], replica_off[r+1]) in block order; series s = replicas
 * [series_off[s], series_off[s+1]).  start/end: [start, end) filter (both 0 = none);
 * strategy: 0 last-pushed, 1 highest value, 2 lowest value, 3 highest frequency. ---- */
int64_t m3o_series_merge(const int64_t *ts, const double *val, uint64_t cap, const uint32_t *n_points,
                         const int32_t *seq_status, const uint64_t *slice_off,
                         const uint64_t *replica_off, uint64_t rep0, uint64_t rep1, int64_t start,
                         int64_t end, int strategy, int64_t *ts_out, double *val_out, uint64_t out_cap,
                         int32_t *status);
void m3o_series_merge_batch(const int64_t *ts, const double *val, uint64_t cap, const uint32_t *n_points,
                            const int32_t *seq_status, const uint64_t *slice_off,
                            const uint64_t *replica_off, const uint64_t *series_off, uint64_t n_series,
                            int64_t start, int64_t end, int strategy, int64_t *ts_out, double *val_out,
                            uint64_t out_cap, uint32_t *n_out, int32_t *status);

/* ---- segment checksum (m3tsz_segment_oracle.c): Adler-32 of
 * a stream = ts.Segment.CalculateChecksum (src/dbnode/ts/segment.go:60-76). ---- */
uint32_t m3o_adler32(const uint8_t *data, size_t n);
void m3o_adler32_batch(const uint8_t *streams, const uint64_t *offsets, uint64_t n_series,
                       const uint32_t *expected, uint32_t *out, int32_t *status);

/* ---- query-side consumers of the decode path (m3tsz_query_oracle.c) ---- */
/* iteratorToPromResult, src/query/storage/prom_converter.go:42-120 */
size_t m3o_prom_convert_series(const int64_t *ts, const double *vals, size_t n, int64_t resolution_ns,
                               int handle_resets, double value_decrease_tolerance,
                               int64_t tolerance_until_ns, int64_t *ts_ms_out, double *val_out,
                               size_t out_cap);
/* Gauge.ValueOf (gauge.go:144-165); agg_type = aggregation.Type id (1 Last, 2 Min, 3 Max, 4 Mean, 6 Count, 7 Sum) */
double m3o_gauge_value_of(int agg_type, double sum, int64_t count, double min, double max, double last);
/* decode -> Gauge per Step window -> one datapoint per non-empty window at the window end */
size_t m3o_aggregate_tiles_series(const int64_t *ts, const double *vals, size_t n, int64_t start_ns,
                                  int64_t step_ns, size_t n_windows, int agg_type, int64_t *ts_out,
                                  double *val_out);

#ifdef __cplusplus
}
#endif
#endif /* M3TSZ_ORACLE_H */
