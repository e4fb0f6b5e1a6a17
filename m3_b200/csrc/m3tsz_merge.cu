// m3tsz_merge.cu -- the iterator layer directly above the codec, on device
// (SURVEY.md §8f row N1): for every series of a fetch, merge the decoded
// streams of its replicas / blocks exactly like the reference's
//   iterators            src/dbnode/encoding/iterators.go:56-262
//   multiReaderIterator  src/dbnode/encoding/multi_reader_iterator.go:62-155,186-196
//   seriesIterator       src/dbnode/encoding/series_iterator.go:74-83,129-215
// (k-way timestamp merge, equal-timestamp strategy incl. the swap-removal order
// of `values`, consecutive-duplicate removal, [start,end) filter, out-of-order
// error, reader errors).
//
// Mapping: one THREAD per series; the per-series state machine is small and
// sequential, the decoded inputs are read through L1 (each thread walks its own
// arrays front to back).  Two kernels: a register-resident fast path for the
// shape a fetch normally has, and the general restatement for everything else
// (run only on the series the fast path gave up on).
#include "m3tsz_common.cuh"
#include "m3tsz_kernels.h"

namespace m3tsz {

constexpr int MRG_K = 12;  // max replicas per series and max readers per slice: up to 12 elements Go's sort.Slice
                           // (the reference's tie order, iterators.go) is a stable insertion sort; beyond that it is
                           // pdqsort and the order of equal timestamps is an implementation detail of the Go runtime
constexpr int64_t kTimeMax = 0x7fffffffffffffffLL;

struct Iters {  // iterators.go:45-59 (ids instead of interface values)
  uint8_t values[MRG_K];
  uint8_t earliest[MRG_K];
  int n_values, n_earliest;
  int64_t earliest_at;
  __device__ void reset() {  // :238-253
    n_values = 0;
    n_earliest = 0;
    earliest_at = kTimeMax;
  }
};

struct Filter {
  bool on;
  int64_t start, end;
};

// ---- generic `iterators` algorithms over a member set M (next / current / err by id) ----
template <class M>
__device__ void try_add_earliest(Iters &it, M &m, int id) {  // :129-143
  int64_t t;
  double v;
  m.current(id, t, v);
  if (t == it.earliest_at) {
    it.earliest[it.n_earliest++] = (uint8_t)id;
  } else if (t < it.earliest_at) {
    it.n_earliest = 0;
    it.earliest[it.n_earliest++] = (uint8_t)id;
    it.earliest_at = t;
  }
}

template <class M>
__device__ bool move_to_filter_next(const Filter &f, M &m, int id) {  // :145-163
  bool next = true;
  while (next) {
    int64_t t;
    double v;
    m.current(id, t, v);
    if (t < f.start) {
      next = m.next(id);
      continue;
    }
    if (t >= f.end) {
      next = false;
      break;
    }
    break;
  }
  return next;
}

template <class M>
__device__ bool iters_push(Iters &it, const Filter &f, M &m, int id) {  // :119-127
  if (f.on && !move_to_filter_next(f, m, id)) return false;
  it.values[it.n_values++] = (uint8_t)id;
  try_add_earliest(it, m, id);
  return true;
}

template <class M>
__device__ bool move_to_valid_next(Iters &it, const Filter &f, M &m, int &err_out) {  // :165-236
  err_out = 0;
  for (;;) {  // the reference recurses when the new earliest falls outside the filter
    const int64_t prev_at = it.earliest_at;
    int n = it.n_values;
    for (int e = 0; e < it.n_earliest; e++) {
      const int id = it.earliest[e];
      bool next = m.next(id);
      if (next && f.on) next = move_to_filter_next(f, m, id);
      const int err = m.err(id);
      if (err) {
        it.reset();
        err_out = err;
        return false;
      }
      if (next) continue;
      int idx = 0;
      for (int k = 0; k < n; k++)
        if (it.values[k] == id) {
          idx = k;
          break;
        }
      it.values[idx] = it.values[n - 1];  // swap the tail in, shrink by one
      n--;
      it.n_values = n;
    }
    it.n_earliest = 0;
    if (n == 0) {
      it.reset();
      return false;
    }
    it.earliest_at = kTimeMax;
    for (int k = 0; k < it.n_values; k++) try_add_earliest(it, m, it.values[k]);
    if (f.on && !(it.earliest_at < f.end && it.earliest_at >= f.start)) continue;
    if (it.earliest_at < prev_at) {  // validateNext
      it.reset();
      err_out = M3TSZ_ERR_OUT_OF_ORDER;
      return false;
    }
    return true;
  }
}

// current() with the equal-timestamp strategy (:60-113); sort.Slice on <= 12
// elements is an insertion sort, which also reorders `earliest` in place.
template <class M>
__device__ void iters_current(Iters &it, int strategy, M &m, int64_t &t_out, double &v_out) {
  const int n = it.n_earliest;
  if (strategy != 0 && n > 1) {
    double key[MRG_K];
    int freq[MRG_K];
    for (int a = 0; a < n; a++) {
      int64_t t;
      m.current(it.earliest[a], t, key[a]);
    }
    if (strategy == 3) {
      for (int a = 0; a < n; a++) {
        int fq = 0;
        if (key[a] == key[a])
          for (int b = 0; b < n; b++)
            if (key[b] == key[a]) fq++;
        freq[a] = fq;
      }
    }
    for (int a = 1; a < n; a++) {
      for (int b = a; b > 0; b--) {
        bool less;
        if (strategy == 1)
          less = key[b] < key[b - 1];
        else if (strategy == 2)
          less = key[b] > key[b - 1];
        else
          less = freq[b] < freq[b - 1];
        if (!less) break;
        const uint8_t tm = it.earliest[b];
        it.earliest[b] = it.earliest[b - 1];
        it.earliest[b - 1] = tm;
        const double tk = key[b];
        key[b] = key[b - 1];
        key[b - 1] = tk;
        const int tf = freq[b];
        freq[b] = freq[b - 1];
        freq[b - 1] = tf;
      }
    }
  }
  m.current(it.earliest[n - 1], t_out, v_out);
}

// ---- level 0: a decoded reader sequence == one ReaderIterator ----
struct SeqCursor {
  uint64_t q;  // global sequence index
  int32_t idx, n;
  int32_t err, cur_err;
};

// ---- level 1: multiReaderIterator (one replica): slices of reader sequences ----
struct Mri {
  Iters iters;
  SeqCursor cur[MRG_K];  // readers of the slice being merged
  uint64_t slice_cur, slice_end;
  bool slices_open, first_next;
  int err;
};

struct ReaderSet {  // member set of an Mri: its current slice's readers
  Mri &r;
  const MergeParams &p;
  __device__ bool next(int id) {
    SeqCursor &c = r.cur[id];
    if (c.cur_err) return false;
    if (c.idx + 1 >= c.n) {
      c.cur_err = c.err;  // the stream's decode error surfaces when it runs out
      return false;
    }
    c.idx++;
    return true;
  }
  __device__ void current(int id, int64_t &t, double &v) {
    const SeqCursor &c = r.cur[id];
    if (c.n == 0) {
      t = 0;
      v = 0.0;
      return;
    }
    const uint64_t o = c.q * p.in_seq_stride + (uint64_t)(c.idx < 0 ? 0 : c.idx) * p.in_pt_stride;
    t = p.ts[o];
    v = p.val[o];
  }
  __device__ int err(int id) { return r.cur[id].cur_err; }
};

__device__ bool mri_has_next(const Mri &r) { return !r.err && (r.iters.n_values > 0 || r.slices_open); }

__device__ void mri_move_to_next(Mri &r, const MergeParams &p) {  // multi_reader_iterator.go:94-155
  const Filter nofilter = {false, 0, 0};
  ReaderSet rs{r, p};
  for (;;) {
    if (r.iters.n_values > 0) {  // moveIteratorsToNext :136-155
      for (;;) {
        const int64_t prev = r.iters.earliest_at;
        int err;
        const bool next = move_to_valid_next(r.iters, nofilter, rs, err);
        if (!r.err && err) {
          r.err = err;
          break;
        }
        if (err || !next) break;
        if (r.iters.earliest_at != prev) break;
      }
    }
    if (r.iters.n_values > 0 || r.err) return;
    if (!r.slices_open) return;
    if (r.slice_cur >= r.slice_end) {
      r.slices_open = false;
      return;
    }
    const uint64_t k = r.slice_cur++;
    const uint64_t q0 = p.slice_off[k], q1 = p.slice_off[k + 1];
    if (q1 - q0 > (uint64_t)MRG_K) {
      r.err = M3TSZ_ERR_TOO_MANY_ITERATORS;
      return;
    }
    for (uint64_t q = q0; q < q1; q++) {
      const int id = (int)(q - q0);
      SeqCursor &c = r.cur[id];
      c.q = q;
      c.idx = -1;
      const uint32_t np = p.n_points[q];
      c.n = (int32_t)(np < p.cap ? np : p.cap);
      c.err = p.seq_status ? p.seq_status[q] : 0;
      c.cur_err = 0;
      if (rs.next(id)) {
        iters_push(r.iters, nofilter, rs, id);
      } else {
        const int e = rs.err(id);
        if (!r.err && e) r.err = e;
      }
    }
    if (r.iters.n_values == 0 && !r.err) continue;  // nothing added: next slice
    return;
  }
}

struct ReplicaSet {  // member set of the series iterator: its replicas
  Mri *reps;
  const MergeParams &p;
  __device__ bool next(int id) {  // multiReaderIterator.Next :62-73
    Mri &r = reps[id];
    if (!r.first_next) {
      if (!mri_has_next(r)) return false;
      mri_move_to_next(r, p);
    }
    r.first_next = false;
    return mri_has_next(r);
  }
  __device__ void current(int id, int64_t &t, double &v) {
    Mri &r = reps[id];
    if (r.iters.n_earliest == 0) {
      t = 0;
      v = 0.0;
      return;
    }
    ReaderSet rs{r, p};
    iters_current(r.iters, 0, rs, t, v);  // a multiReaderIterator's own iterators use the default strategy
  }
  __device__ int err(int id) { return reps[id].err; }
};

// ---------------------------------------------------------------------------
// Fast path: the shape a fetch normally has.  <= MRG_FAST_R replicas per series,
// every slice holds at most one reader, every reader decoded without error, the
// default (last pushed) equal-timestamp strategy, and every replica's
// concatenated blocks are STRICTLY increasing in time.  Under those conditions
// the three nested iterators reduce to a k-way merge of increasing sequences
// whose value on a tie comes from the member that sits last in `values`
// (iterators.go:60-74 with the swap-removal order of :178-196) -- restated here
// with the whole state in registers.  Anything else (several readers in a slice,
// a decode error, a duplicate or out-of-order timestamp, another strategy) makes
// the thread give up and mark the series for the general kernel, which restarts
// it from scratch.
// ---------------------------------------------------------------------------
constexpr int MRG_FAST_R = 4;  // replicas per series the fast kernel is instantiated for
constexpr int32_t MRG_REDO = -1;  // status marker: "run the general kernel on this series"

template <int R>
struct FastState {
  uint64_t slice_cur[R], slice_end[R], base[R];
  int32_t idx[R], n[R];
  int64_t t[R];
  bool bail;
};

// moves replica r to its next datapoint inside [f.start, f.end) (when the filter
// is on); false = exhausted (or past the end of the range: same thing to the caller)
template <int R>
__device__ __forceinline__ bool fast_advance(FastState<R> &st, const MergeParams &p, const Filter &f, int r,
                                             bool have_prev) {
  for (;;) {
    bool got = false;
    int64_t tn = 0;
    if (st.idx[r] + 1 < st.n[r]) {
      st.idx[r]++;
      tn = p.ts[st.base[r] + (uint64_t)st.idx[r] * p.in_pt_stride];
      got = true;
    } else {
      while (st.slice_cur[r] < st.slice_end[r]) {
        const uint64_t k = st.slice_cur[r]++;
        const uint64_t q0 = p.slice_off[k], q1 = p.slice_off[k + 1];
        if (q1 == q0) continue;
        if (q1 - q0 > 1 || (p.seq_status && p.seq_status[q0] != 0)) {
          st.bail = true;
          return false;
        }
        const uint32_t np = p.n_points[q0];
        const int32_t n = (int32_t)(np < p.cap ? np : p.cap);
        if (n == 0) continue;
        st.base[r] = q0 * p.in_seq_stride;
        st.idx[r] = 0;
        st.n[r] = n;
        tn = p.ts[st.base[r]];
        got = true;
        break;
      }
    }
    if (!got) return false;
    if (have_prev && tn <= st.t[r]) {  // duplicate / out of order: the general path decides
      st.bail = true;
      return false;
    }
    st.t[r] = tn;
    have_prev = true;
    if (f.on) {
      if (tn < f.start) continue;
      if (tn >= f.end) return false;
    }
    return true;
  }
}

template <int R>
__device__ void merge_fast(const MergeParams &p, uint64_t s, uint64_t rep0) {
  FastState<R> st;
  st.bail = false;
  Filter f;
  f.on = (p.start != 0 && p.end != 0);
  f.start = p.start;
  f.end = p.end;
  uint32_t order = 0;  // `values`: member ids by position, 4 bits each
  int nv = 0;
#pragma unroll
  for (int r = 0; r < R; r++) {
    st.slice_cur[r] = p.replica_off[rep0 + r];
    st.slice_end[r] = p.replica_off[rep0 + r + 1];
    st.base[r] = 0;
    st.idx[r] = 0;
    st.n[r] = 0;
    st.t[r] = 0;
    if (fast_advance<R>(st, p, f, r, false)) {
      order |= (uint32_t)r << (4 * nv);
      nv++;
    }
  }
  int64_t *ts_out = p.ts_out + s * p.out_seq_stride;
  double *val_out = p.val_out + s * p.out_seq_stride;
  uint32_t n_out = 0;
  while (nv > 0 && !st.bail) {
    // earliest timestamp and its members in position order; the value of the last one
    int64_t tmin = kTimeMax;
#pragma unroll
    for (int r = 0; r < R; r++) {
      bool in = false;
      for (int q = 0; q < nv; q++) in = in || (((order >> (4 * q)) & 15u) == (uint32_t)r);
      if (in && st.t[r] < tmin) tmin = st.t[r];
    }
    uint32_t e_ids = 0;
    int ne = 0;
    uint64_t win_addr = 0;
    for (int q = 0; q < nv; q++) {
      const uint32_t id = (order >> (4 * q)) & 15u;
#pragma unroll
      for (int r = 0; r < R; r++) {
        if ((uint32_t)r == id && st.t[r] == tmin) {
          e_ids |= id << (4 * ne);
          ne++;
          win_addr = st.base[r] + (uint64_t)st.idx[r] * p.in_pt_stride;
        }
      }
    }
    if (n_out < p.out_cap) {
      ts_out[(uint64_t)n_out * p.out_pt_stride] = tmin;
      val_out[(uint64_t)n_out * p.out_pt_stride] = p.val[win_addr];
    }
    n_out++;
    // advance the earliest members in list order; an exhausted member's place in
    // `values` is taken by the tail (iterators.go:188-192)
    for (int k = 0; k < ne; k++) {
      const uint32_t id = (e_ids >> (4 * k)) & 15u;
      bool alive = true;
#pragma unroll
      for (int r = 0; r < R; r++)
        if ((uint32_t)r == id) alive = fast_advance<R>(st, p, f, r, true);
      if (st.bail) break;
      if (!alive) {
        int at = 0;
        for (int q = 0; q < nv; q++)
          if (((order >> (4 * q)) & 15u) == id) {
            at = q;
            break;
          }
        const uint32_t last = (order >> (4 * (nv - 1))) & 15u;
        order = (order & ~(15u << (4 * at))) | (last << (4 * at));
        nv--;
      }
    }
  }
  if (st.bail) {
    p.status[s] = MRG_REDO;
    return;
  }
  p.n_out[s] = n_out;
  p.status[s] = n_out > p.out_cap ? M3TSZ_ERR_CAPACITY : 0;
}

__global__ void __launch_bounds__(128) merge_fast_kernel(const MergeParams p) {
  const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= p.n_series) return;
  const uint64_t rep0 = p.series_off[s], rep1 = p.series_off[s + 1];
  switch (rep1 - rep0) {
    case 1: merge_fast<1>(p, s, rep0); break;
    case 2: merge_fast<2>(p, s, rep0); break;
    case 3: merge_fast<3>(p, s, rep0); break;
    case 4: merge_fast<4>(p, s, rep0); break;
    default: static_assert(MRG_FAST_R == 4, "one case per replica count"); p.status[s] = MRG_REDO; break;
  }
}

// general kernel; redo_only: handle only the series the fast kernel gave up on
__global__ void __launch_bounds__(128) merge_kernel(const MergeParams p, bool redo_only) {
  const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= p.n_series) return;
  if (redo_only && p.status[s] != MRG_REDO) return;
  const uint64_t rep0 = p.series_off[s], rep1 = p.series_off[s + 1];
  int err = 0;
  uint32_t n_out = 0;
  if (rep1 - rep0 > (uint64_t)MRG_K) {
    p.n_out[s] = 0;
    p.status[s] = M3TSZ_ERR_TOO_MANY_ITERATORS;
    return;
  }
  const int n_rep = (int)(rep1 - rep0);
  Mri reps[MRG_K];
  Iters top;
  top.reset();
  Filter f;
  f.on = (p.start != 0 && p.end != 0);  // series_iterator.go:146-148
  f.start = p.start;
  f.end = p.end;
  ReplicaSet set{reps, p};
  for (int r = 0; r < n_rep; r++) {  // seriesIterator.Reset :157-168
    Mri &m = reps[r];
    m.iters.reset();
    m.slice_cur = p.replica_off[rep0 + r];
    m.slice_end = p.replica_off[rep0 + r + 1];
    m.slices_open = true;
    m.first_next = true;
    m.err = 0;
    mri_move_to_next(m, p);  // ResetSliceOfSlices :186-196
    if (!set.next(r) || !iters_push(top, f, set, r)) {
      if (m.err) err = m.err;
    }
  }
  int64_t *ts_out = p.ts_out + s * p.out_seq_stride;
  double *val_out = p.val_out + s * p.out_seq_stride;
  bool first_next = true;
  for (;;) {  // seriesIterator.Next :74-83
    if (!first_next) {
      if (err || top.n_values == 0) break;
      for (;;) {  // moveToNext :196-215
        const int64_t prev = top.earliest_at;
        int e2;
        const bool next = move_to_valid_next(top, f, set, e2);
        if (e2) {
          err = e2;
          break;
        }
        if (!next) break;
        if (top.earliest_at != prev) break;
      }
    }
    first_next = false;
    if (err || top.n_values == 0) break;
    int64_t t;
    double v;
    iters_current(top, p.strategy, set, t, v);
    if (n_out < p.out_cap) {
      ts_out[(uint64_t)n_out * p.out_pt_stride] = t;
      val_out[(uint64_t)n_out * p.out_pt_stride] = v;
    }
    n_out++;
  }
  p.n_out[s] = n_out;
  p.status[s] = err ? err : (n_out > p.out_cap ? M3TSZ_ERR_CAPACITY : 0);
}

cudaError_t launch_merge(const MergeParams &p, cudaStream_t stream) {
  if (p.n_series == 0) return cudaSuccess;
  const unsigned tb = 128;
  const uint64_t blocks = (p.n_series + tb - 1) / tb;
  if (blocks > 0x7fffffffull) return cudaErrorInvalidValue;
  const bool fast = (p.strategy == 0);
  if (fast) merge_fast_kernel<<<(unsigned)blocks, tb, 0, stream>>>(p);
  merge_kernel<<<(unsigned)blocks, tb, 0, stream>>>(p, fast);
  return cudaGetLastError();
}

}  // namespace m3tsz
