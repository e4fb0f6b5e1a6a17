/*
 * m3tsz_merge_oracle.c -- CPU restatement of the iterator layer directly above
 * the codec.  TEST INFRASTRUCTURE ONLY (same rules as
 * m3tsz_oracle.h).
 *
 * Restates, over already-decoded reader sequences:
 *   iterators            src/dbnode/encoding/iterators.go:56-262
 *   multiReaderIterator  src/dbnode/encoding/multi_reader_iterator.go:62-155,186-196
 *   seriesIterator       src/dbnode/encoding/series_iterator.go:74-83,129-215
 * A "reader sequence" stands for one ReaderIterator: n datapoints, then
 * Next()==false with Err()==err (the decode status of that stream).
 *
 * Pinned by the reference's table tests (iterators_test.go, multi_reader_iterator_test.go,
 * series_iterator_test.go), transcribed in tests/test_merge_oracle.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "m3tsz_oracle.h"

#define MAXK 12 /* sort.Slice is an insertion sort up to 12 elements (Go sort: maxInsertion) */
#define TIME_MAX INT64_MAX

typedef struct member member;
struct member {
  int (*next)(member *);
  void (*current)(member *, int64_t *ts, double *val);
  int (*err)(member *);
};

/* ---- iterators (iterators.go) ---- */
typedef struct {
  member *values[MAXK];
  int n_values;
  member *earliest[MAXK];
  int n_earliest;
  int64_t earliest_at;
  int64_t filter_start, filter_end;
  int filtering;
  int strategy;
} iters_t;

static void iters_reset(iters_t *it) { /* :238-253 */
  it->n_values = 0;
  it->n_earliest = 0;
  it->earliest_at = TIME_MAX;
}

static int is_nan(double v) { return v != v; }

static void iters_current(iters_t *it, int64_t *ts, double *val) { /* :60-113 */
  int n = it->n_earliest;
  if (it->strategy == 1 || it->strategy == 2 || it->strategy == 3) {
    double key[MAXK];
    for (int a = 0; a < n; a++) {
      int64_t t;
      double v;
      it->earliest[a]->current(it->earliest[a], &t, &v);
      key[a] = v;
    }
    int freq[MAXK];
    if (it->strategy == 3) {
      for (int a = 0; a < n; a++) {
        int f = 0;
        if (!is_nan(key[a]))
          for (int b = 0; b < n; b++)
            if (key[b] == key[a]) f++;
        freq[a] = f; /* a NaN key is never found again in a Go map: frequency 0 */
      }
    }
    /* sort.Slice on <= 12 elements == insertion sort (stable) */
    for (int a = 1; a < n; a++) {
      for (int b = a; b > 0; b--) {
        int less;
        if (it->strategy == 1)
          less = key[b] < key[b - 1];
        else if (it->strategy == 2)
          less = key[b] > key[b - 1];
        else
          less = freq[b] < freq[b - 1];
        if (!less) break;
        member *tm = it->earliest[b];
        it->earliest[b] = it->earliest[b - 1];
        it->earliest[b - 1] = tm;
        double tk = key[b];
        key[b] = key[b - 1];
        key[b - 1] = tk;
        if (it->strategy == 3) {
          int tf = freq[b];
          freq[b] = freq[b - 1];
          freq[b - 1] = tf;
        }
      }
    }
  }
  it->earliest[n - 1]->current(it->earliest[n - 1], ts, val);
}

static void iters_try_add_earliest(iters_t *it, member *m) { /* :129-143 */
  int64_t t;
  double v;
  m->current(m, &t, &v);
  if (t == it->earliest_at) {
    it->earliest[it->n_earliest++] = m;
  } else if (t < it->earliest_at) {
    it->n_earliest = 0;
    it->earliest[it->n_earliest++] = m;
    it->earliest_at = t;
  }
}

static int iters_move_to_filter_next(iters_t *it, member *m) { /* :145-163 */
  int next = 1;
  while (next) {
    int64_t t;
    double v;
    m->current(m, &t, &v);
    if (t < it->filter_start) {
      next = m->next(m);
      continue;
    }
    if (t >= it->filter_end) {
      next = 0;
      break;
    }
    break;
  }
  return next;
}

static int iters_push(iters_t *it, member *m) { /* :119-127 */
  if (it->filtering && !iters_move_to_filter_next(it, m)) return 0;
  it->values[it->n_values++] = m;
  iters_try_add_earliest(it, m);
  return 1;
}

/* returns next (0/1); *err_out receives the error */
static int iters_move_to_valid_next(iters_t *it, int *err_out) { /* :165-227 */
  int64_t prev_at = it->earliest_at;
  int n = it->n_values;
  *err_out = 0;
  for (int e = 0; e < it->n_earliest; e++) {
    member *m = it->earliest[e];
    int next = m->next(m);
    if (next && it->filtering) next = iters_move_to_filter_next(it, m);
    int err = m->err(m);
    if (err) {
      iters_reset(it);
      *err_out = err;
      return 0;
    }
    if (next) continue;
    int idx = -1;
    for (int k = 0; k < n; k++)
      if (it->values[k] == m) {
        idx = k;
        break;
      }
    it->values[idx] = it->values[n - 1];
    it->values[n - 1] = NULL;
    n--;
    it->n_values = n;
  }
  it->n_earliest = 0;
  if (n == 0) {
    iters_reset(it);
    return 0;
  }
  it->earliest_at = TIME_MAX;
  for (int k = 0; k < it->n_values; k++) iters_try_add_earliest(it, it->values[k]);
  if (it->filtering) {
    int in_filter = it->earliest_at < it->filter_end && it->earliest_at >= it->filter_start;
    if (!in_filter) return iters_move_to_valid_next(it, err_out);
  }
  if (it->earliest_at < prev_at) { /* validateNext :229-236 */
    iters_reset(it);
    *err_out = M3O_ERR_OUT_OF_ORDER;
    return 0;
  }
  return 1;
}

/* ---- a decoded reader sequence as a ReaderIterator ---- */
typedef struct {
  member m;
  const int64_t *ts;
  const double *val;
  int64_t n, idx;
  int err, cur_err;
} seq_t;

static int seq_next(member *mm) {
  seq_t *s = (seq_t *)mm;
  if (s->cur_err) return 0;
  if (s->idx + 1 >= s->n) {
    s->cur_err = s->err; /* the stream's decode error surfaces when it runs out */
    return 0;
  }
  s->idx++;
  return 1;
}
static void seq_current(member *mm, int64_t *ts, double *val) {
  seq_t *s = (seq_t *)mm;
  int64_t i = s->idx < 0 ? 0 : s->idx;
  if (s->n == 0) {
    *ts = 0;
    *val = 0;
    return;
  }
  *ts = s->ts[i];
  *val = s->val[i];
}
static int seq_err(member *mm) { return ((seq_t *)mm)->cur_err; }

/* ---- multiReaderIterator over slices of reader sequences ---- */
typedef struct {
  member m;
  iters_t iters;
  seq_t *seqs;               /* sequences of this series; seqs[0] is global sequence seq_base */
  uint64_t seq_base;
  const uint64_t *slice_off; /* slice k = seqs [slice_off[k], slice_off[k+1]) */
  uint64_t slice_cur, slice_end;
  int slices_open; /* slicesIter != nil */
  int err, first_next;
} mri_t;

static int mri_has_next(mri_t *it) { /* :83-92 */
  return !it->err && (it->iters.n_values > 0 || it->slices_open);
}

static void mri_move_iterators_to_next(mri_t *it) { /* :136-155 */
  for (;;) {
    int64_t prev = it->iters.earliest_at;
    int err;
    int next = iters_move_to_valid_next(&it->iters, &err);
    if (!it->err && err) {
      it->err = err;
      return;
    }
    if (err || !next) return;
    if (it->iters.earliest_at != prev) return;
  }
}

static void mri_move_to_next(mri_t *it) { /* :94-134 */
  if (it->iters.n_values > 0) mri_move_iterators_to_next(it);
  if (it->iters.n_values > 0 || it->err) return;
  if (!it->slices_open) return;
  if (it->slice_cur >= it->slice_end) { /* slicesIter.Next() == false */
    it->slices_open = 0;
    return;
  }
  uint64_t k = it->slice_cur++;
  for (uint64_t r = it->slice_off[k]; r < it->slice_off[k + 1]; r++) {
    seq_t *s = &it->seqs[r - it->seq_base];
    if (s->m.next(&s->m)) {
      iters_push(&it->iters, &s->m);
    } else {
      int err = s->m.err(&s->m);
      if (!it->err && err) it->err = err;
    }
  }
  if (it->iters.n_values == 0 && !it->err) mri_move_to_next(it);
}

static int mri_next(member *mm) { /* :62-73 */
  mri_t *it = (mri_t *)mm;
  if (!it->first_next) {
    if (!mri_has_next(it)) return 0;
    mri_move_to_next(it);
  }
  it->first_next = 0;
  return mri_has_next(it);
}
static void mri_current(member *mm, int64_t *ts, double *val) {
  mri_t *it = (mri_t *)mm;
  if (it->iters.n_earliest == 0) {
    *ts = 0;
    *val = 0;
    return;
  }
  iters_current(&it->iters, ts, val);
}
static int mri_err(member *mm) { return ((mri_t *)mm)->err; }

static void mri_reset(mri_t *it, seq_t *seqs, uint64_t seq_base, const uint64_t *slice_off, uint64_t s0,
                      uint64_t s1) { /* :177-186 */
  it->m.next = mri_next;
  it->m.current = mri_current;
  it->m.err = mri_err;
  memset(&it->iters, 0, sizeof(it->iters));
  iters_reset(&it->iters);
  it->seqs = seqs;
  it->seq_base = seq_base;
  it->slice_off = slice_off;
  it->slice_cur = s0;
  it->slice_end = s1;
  it->slices_open = 1;
  it->err = 0;
  it->first_next = 1;
  mri_move_to_next(it);
}

/* ---- seriesIterator over replicas (each a multiReaderIterator) ----
 * Returns the number of datapoints produced; *status = final Err(). */
int64_t m3o_series_merge(const int64_t *ts, const double *val, uint64_t cap, const uint32_t *n_points,
                         const int32_t *seq_status, const uint64_t *slice_off,
                         const uint64_t *replica_off, uint64_t rep0, uint64_t rep1, int64_t start,
                         int64_t end, int strategy, int64_t *ts_out, double *val_out, uint64_t out_cap,
                         int32_t *status) {
  uint64_t n_rep = rep1 - rep0;
  if (n_rep > MAXK) {
    *status = M3O_ERR_TOO_MANY_ITERATORS;
    return 0;
  }
  uint64_t q0 = slice_off[replica_off[rep0]], q1 = slice_off[replica_off[rep1]];
  seq_t *seqs = (seq_t *)calloc(q1 - q0 + 1, sizeof(seq_t));
  for (uint64_t q = q0; q < q1; q++) {
    seq_t *s = &seqs[q - q0];
    s->m.next = seq_next;
    s->m.current = seq_current;
    s->m.err = seq_err;
    s->ts = ts + q * cap;
    s->val = val + q * cap;
    s->n = n_points[q] < cap ? n_points[q] : cap;
    s->idx = -1;
    s->err = seq_status ? seq_status[q] : 0;
  }
  mri_t *reps = (mri_t *)calloc(n_rep + 1, sizeof(mri_t));
  iters_t top; /* seriesIterator.iters, series_iterator.go:129-170 */
  memset(&top, 0, sizeof(top));
  iters_reset(&top);
  top.strategy = strategy;
  if (start != 0 && end != 0) {
    top.filtering = 1;
    top.filter_start = start;
    top.filter_end = end;
  }
  int err = 0;
  for (uint64_t r = 0; r < n_rep; r++) {
    mri_t *it = &reps[r];
    /* slices of this replica must not hold more than MAXK readers (restatement limit, not a reference error:
     * only THIS condition stops the loop -- an error of an earlier replica does not, see below) */
    int too_many = 0;
    for (uint64_t k = replica_off[rep0 + r]; k < replica_off[rep0 + r + 1]; k++)
      if (slice_off[k + 1] - slice_off[k] > MAXK) too_many = 1;
    if (too_many) {
      err = M3O_ERR_TOO_MANY_ITERATORS;
      break;
    }
    mri_reset(it, seqs, q0, slice_off, replica_off[rep0 + r], replica_off[rep0 + r + 1]);
    if (!it->m.next(&it->m) || !iters_push(&top, &it->m)) {
      /* seriesIterator.Reset, series_iterator.go:157-168: `it.err = replica.Err()` -- every failing replica
       * overwrites the error of the ones before it, and the loop goes on to the remaining replicas */
      if (it->m.err(&it->m)) err = it->m.err(&it->m);
      continue;
    }
  }
  int64_t n_out = 0;
  int first_next = 1;
  for (;;) { /* Next(): series_iterator.go:74-83 */
    if (!first_next) {
      if (err || top.n_values == 0) break;
      for (;;) { /* moveToNext :196-215 */
        int64_t prev = top.earliest_at;
        int e2;
        int next = iters_move_to_valid_next(&top, &e2);
        if (e2) {
          err = e2;
          break;
        }
        if (!next) break;
        if (top.earliest_at != prev) break;
      }
    }
    first_next = 0;
    if (err || top.n_values == 0) break;
    int64_t t;
    double v;
    iters_current(&top, &t, &v);
    if ((uint64_t)n_out < out_cap) {
      ts_out[n_out] = t;
      val_out[n_out] = v;
    }
    n_out++;
  }
  *status = err;
  free(reps);
  free(seqs);
  return n_out;
}

void m3o_series_merge_batch(const int64_t *ts, const double *val, uint64_t cap, const uint32_t *n_points,
                            const int32_t *seq_status, const uint64_t *slice_off,
                            const uint64_t *replica_off, const uint64_t *series_off, uint64_t n_series,
                            int64_t start, int64_t end, int strategy, int64_t *ts_out, double *val_out,
                            uint64_t out_cap, uint32_t *n_out, int32_t *status) {
  for (uint64_t s = 0; s < n_series; s++) {
    int32_t st = 0;
    int64_t n = m3o_series_merge(ts, val, cap, n_points, seq_status, slice_off, replica_off, series_off[s],
                                 series_off[s + 1], start, end, strategy, ts_out + s * out_cap,
                                 val_out + s * out_cap, out_cap, &st);
    n_out[s] = (uint32_t)n;
    status[s] = st;
  }
}
