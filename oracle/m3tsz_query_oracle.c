/*
 * m3tsz_query_oracle.c -- TEST INFRASTRUCTURE (CPU oracle), never linked into or
 * called by the product library.  Restates the two consumers directly after the
 * decode path:
 *
 *  (1) the Prometheus conversion epilogue, iteratorToPromResult,
 *      /root/reference/src/query/storage/prom_converter.go:42-120
 *      (TimeToPromTimestamp: src/query/storage/converter.go:388-391);
 *      pinned by the tables of prom_converter_test.go:319-440 (counter
 *      normalisation) and :444-500 (value decrease tolerance), transcribed into
 *      tests/golden/m3tsz_goldens.json by scripts/make_goldens.py.
 *
 *  (2) tile aggregation = what a storage.TileAggregator
 *      (src/dbnode/storage/types.go:1444-1472; open-source default is a no-op,
 *      storage/options.go:949-951) does with one series: decode the source block,
 *      fold the datapoints of [Start, End) into Step-sized windows with the
 *      aggregator's Gauge (src/aggregator/aggregation/gauge.go:73-165), emit one
 *      datapoint per NON-EMPTY window stamped with the window's end boundary
 *      (standardMetricTimestampNanos, src/aggregator/aggregator/list.go:541-543)
 *      and re-encode with the M3TSZ encoder.  The aggregation arithmetic is pinned by
 *      gauge_test.go (see tests/test_oracle_goldens.py); the encoder by the m3tsz goldens.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "m3tsz_oracle.h"

/* Go integer division truncates toward zero, like C. */

/* prom_converter.go:65-118.  Returns the number of samples written (<= out_cap
 * are stored; the return value may exceed out_cap). */
size_t m3o_prom_convert_series(const int64_t *ts, const double *vals, size_t n, int64_t resolution_ns,
                               int handle_resets, double value_decrease_tolerance,
                               int64_t tolerance_until_ns, int64_t *ts_ms_out, double *val_out,
                               size_t out_cap) {
  int first = 1;
  double cumulative = 0.0;
  int64_t prev_t = 0;
  double prev_v = 0.0;
  size_t n_out = 0;
  for (size_t i = 0; i < n; i++) {
    int64_t t = ts[i];
    double v = vals[i];
    if (value_decrease_tolerance > 0 && t < tolerance_until_ns) { /* :68-72 */
      if (!first && v < prev_v && v > prev_v * (1 - value_decrease_tolerance)) v = prev_v;
    }
    if (handle_resets) { /* :84-98 */
      if (resolution_ns != 0 && t / resolution_ns != prev_t / resolution_ns && !first) {
        if (n_out < out_cap) {
          ts_ms_out[n_out] = prev_t / 1000000;
          val_out[n_out] = cumulative;
        }
        n_out++;
      }
      if (v < prev_v)
        cumulative += v; /* counter reset */
      else
        cumulative += v - prev_v;
    } else { /* :99-104 */
      if (n_out < out_cap) {
        ts_ms_out[n_out] = t / 1000000;
        val_out[n_out] = v;
      }
      n_out++;
    }
    prev_t = t;
    prev_v = v;
    first = 0;
  }
  if (handle_resets && !first) { /* :113-118; handleResets is only ever set inside the loop */
    if (n_out < out_cap) {
      ts_ms_out[n_out] = prev_t / 1000000;
      val_out[n_out] = cumulative;
    }
    n_out++;
  }
  return n_out;
}

/* Gauge.ValueOf, gauge.go:144-165, for the types the tile path supports
 * (aggregation.Type ids, src/metrics/aggregation/type.go:31-38). */
double m3o_gauge_value_of(int agg_type, double sum, int64_t count, double min, double max, double last) {
  switch (agg_type) {
    case 1: return last;
    case 2: return min;
    case 3: return max;
    case 4: return count == 0 ? 0.0 : sum / (double)count; /* Mean, gauge.go:117-122 */
    case 6: return (double)count;
    case 7: return sum;
    default: return 0.0;
  }
}

/* Aggregates one decoded series into tiles.  Returns the number of output
 * datapoints (one per non-empty window, ts = window end). */
size_t m3o_aggregate_tiles_series(const int64_t *ts, const double *vals, size_t n, int64_t start_ns,
                                  int64_t step_ns, size_t n_windows, int agg_type, int64_t *ts_out,
                                  double *val_out) {
  double *sum = (double *)malloc(sizeof(double) * 5 * (n_windows ? n_windows : 1));
  int64_t *count = (int64_t *)malloc(sizeof(int64_t) * (n_windows ? n_windows : 1));
  double *mn = sum + n_windows, *mx = mn + n_windows, *last = mx + n_windows;
  m3o_downsample_series(ts, vals, n, start_ns, step_ns, n_windows, sum, count, mn, mx, last);
  size_t k = 0;
  for (size_t w = 0; w < n_windows; w++) {
    if (count[w] == 0) continue;
    ts_out[k] = start_ns + (int64_t)(w + 1) * step_ns;
    val_out[k] = m3o_gauge_value_of(agg_type, sum[w], count[w], mn[w], mx[w], last[w]);
    k++;
  }
  free(sum);
  free(count);
  return k;
}
