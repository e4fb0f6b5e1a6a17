"""Pins the merge oracle (iterators / multiReaderIterator / seriesIterator restatement,
oracle/m3tsz_merge_oracle.c) against the reference's own table tests:
  src/dbnode/encoding/iterators_test.go:36-130
  src/dbnode/encoding/multi_reader_iterator_test.go:52-275
  src/dbnode/encoding/series_iterator_test.go:61-207
CPU only."""
import numpy as np

import oracle_lib as O

SEC = 1_000_000_000
START = 1_600_000_020 * SEC  # "xtime.Now().Truncate(time.Minute)" stand-in
ERR_OUT_OF_ORDER = 13


def build(series):
    """series: list of replicas; replica: list of slices; slice: list of readers;
    reader: list of (value, ts) or (list, err).  Returns the flat batch arrays."""
    seqs, slice_off, replica_off, series_off = [], [0], [0], [0]
    for replicas in series:
        for slices in replicas:
            for readers in slices:
                for rd in readers:
                    err = 0
                    if isinstance(rd, tuple) and len(rd) == 2 and isinstance(rd[0], list):
                        rd, err = rd
                    seqs.append((rd, err))
                slice_off.append(len(seqs))
            replica_off.append(len(slice_off) - 1)
        series_off.append(len(replica_off) - 1)
    cap = max([len(s[0]) for s in seqs] + [1])
    n = len(seqs)
    ts = np.zeros((max(n, 1), cap), dtype=np.int64)
    val = np.zeros((max(n, 1), cap), dtype=np.float64)
    npts = np.zeros(max(n, 1), dtype=np.uint32)
    st = np.zeros(max(n, 1), dtype=np.int32)
    for i, (rd, err) in enumerate(seqs):
        npts[i] = len(rd)
        st[i] = err
        for j, (v, t) in enumerate(rd):
            ts[i, j], val[i, j] = t, v
    return ts, val, npts, st, slice_off, replica_off, series_off


def merge(series, start=0, end=0, strategy=0):
    args = build(series)
    ts_out, val_out, n_out, status = O.series_merge_batch(*args, start=start, end=end, strategy=strategy)
    return [[(float(val_out[s, i]), int(ts_out[s, i])) for i in range(n_out[s])] for s in range(len(series))], \
        [int(x) for x in status]


def multi_reader(slices, **kw):
    """One MultiReaderIterator (slices of readers) == a series with a single replica and no filter."""
    out, st = merge([[slices]], **kw)
    return out[0], st[0]


def at(k):
    return START + k * SEC


# ---------------------------------------------------------------- multi_reader_iterator_test.go
V0 = [(1.0, at(1)), (2.0, at(2)), (3.0, at(3))]
V1 = [(4.0, at(4)), (5.0, at(5)), (6.0, at(6))]


def test_mri_merges_multi():  # :52-75
    assert multi_reader([[V0, V1]]) == (V0 + V1, 0)


def test_mri_merges_empty():  # :77-101
    assert multi_reader([[[]], [[]]]) == ([], 0)
    assert multi_reader([[[], []]]) == ([], 0)


def test_mri_reads_slices_in_order():  # :103-127
    assert multi_reader([[V0], [V1]]) == (V0 + V1, 0)


def test_mri_slices_with_no_entries():  # :129-154
    assert multi_reader([[V0], [], [V1]]) == (V0 + V1, 0)


def test_mri_slices_with_empty_entries():  # :156-183
    assert multi_reader([[V0], [[]], [V1]]) == (V0 + V1, 0)


def test_mri_deduplicates_single():  # :185-204
    v = [(1.0, at(1)), (2.0, at(2)), (2.0, at(2))]
    assert multi_reader([[v]]) == (v[:2], 0)


def test_mri_deduplicates_multi():  # :206-228
    assert multi_reader([[V0, V0, V0]]) == (V0, 0)


def test_mri_error_on_out_of_order():  # :230-252
    v = [(1.0, at(1)), (3.0, at(3)), (2.0, at(2))]
    out, st = multi_reader([[v]])
    assert out == v[:2] and st == ERR_OUT_OF_ORDER


def test_mri_error_on_inner_iterator_error():  # :254-283: the reader fails after 2 datapoints
    out, st = multi_reader([[(V0[:2], 77)]])
    assert out == V0[:2] and st == 77


# ---------------------------------------------------------------- series_iterator_test.go
def test_series_merges_replicas():  # :61-104
    a = [(1.0, at(1)), (2.0, at(2)), (3.0, at(3))]
    c = [(3.0, at(3)), (4.0, at(4)), (5.0, at(5))]
    out, st = merge([[[[a]], [[a]], [[c]]]], start=START, end=START + 60 * SEC)
    assert st == [0]
    assert out[0] == [(1.0, at(1)), (2.0, at(2)), (3.0, at(3)), (4.0, at(4)), (5.0, at(5))]


def test_series_filters_to_range():  # :106-138
    v = [(0.0, at(-2)), (1.0, at(-1)), (2.0, at(0)), (3.0, at(1)), (4.0, at(60)), (5.0, at(61))]
    out, st = merge([[[[v]]]], start=START, end=START + 60 * SEC)
    assert st == [0] and out[0] == v[2:4]


def test_series_ignores_empty_replicas():  # :140-165
    v = [(1.0, at(1)), (2.0, at(2)), (3.0, at(3))]
    out, st = merge([[[[v]], [[[]]], [[v]]]], start=START, end=START + 60 * SEC)
    assert st == [0] and out[0] == v


def test_series_does_not_ignore_replicas_with_errors():  # :167-186
    out, st = merge([[[[([], 55)]]]], start=START, end=START + 60 * SEC)
    assert out[0] == [] and st == [55]


def test_series_last_failing_replica_sets_the_error():
    """series_iterator.go:157-168 (Reset): `it.err = replica.Err()` is an unconditional assignment inside the loop
    over ALL replicas -- when several replicas fail, the LAST one's error is the series' error (and the healthy
    replicas after a failing one are still pushed, although hasNext() then refuses to iterate).  Found by running
    the device merge kernels on the host against this oracle (tests/test_device_merge_on_host.py): the oracle
    used to stop at the first failing replica."""
    v = [(1.0, at(1)), (2.0, at(2))]
    out, st = merge([[[[([], 55)]], [[([], 66)]], [[v]]]], start=START, end=START + 60 * SEC)
    assert out[0] == [] and st == [66]
    out, st = merge([[[[([], 66)]], [[v]], [[([], 55)]]]], start=START, end=START + 60 * SEC)
    assert out[0] == [] and st == [55]


def test_series_error_on_out_of_order():  # :188-213
    v = [(1.0, at(1)), (3.0, at(3)), (2.0, at(2))]
    out, st = merge([[[[v]]]], start=START, end=START + 60 * SEC)
    assert out[0] == v[:2] and st == [ERR_OUT_OF_ORDER]


# ---------------------------------------------------------------- iterators_test.go
COMMON = [
    [(2.0, at(0)), (6.0, at(1)), (7.0, at(2))],
    [(1.0, at(0)), (5.0, at(1)), (9.0, at(2))],
    [(3.0, at(0)), (4.0, at(1)), (8.0, at(2))],
]


def _as_replicas(lists):
    return [[[[v]] for v in lists]]


def test_iterators_last_pushed():  # :58-66
    out, st = merge(_as_replicas(COMMON), strategy=0)
    assert st == [0] and out[0] == COMMON[2]


def test_iterators_highest_value():  # :78-91
    out, st = merge(_as_replicas(COMMON), strategy=1)
    assert st == [0] and out[0] == [COMMON[2][0], COMMON[0][1], COMMON[1][2]]


def test_iterators_lowest_value():  # :93-106
    out, st = merge(_as_replicas(COMMON), strategy=2)
    assert st == [0] and out[0] == [COMMON[1][0], COMMON[2][1], COMMON[0][2]]


def test_iterators_highest_frequency_value():  # :108-139
    v = [
        [(2.0, at(0)), (6.0, at(1)), (8.0, at(2))],
        [(2.0, at(0)), (6.0, at(1)), (9.0, at(2))],
        [(3.0, at(0)), (5.0, at(1)), (8.0, at(2)), (10.0, at(3))],
    ]
    out, st = merge(_as_replicas(v), strategy=3)
    assert st == [0] and out[0] == [v[0][0], v[1][1], v[2][2], v[2][3]]


# ---------------------------------------------------------------- extra behaviours
def test_swap_removal_changes_tie_break_order():
    """values[idx] = values[n-1] on exhaustion (iterators.go:189-195) permutes the
    'last pushed' order: after replica 0 runs out, replica 2 sits in its slot, so
    replica 1 becomes the last of the list."""
    a = [(10.0, at(1))]
    b = [(20.0, at(1)), (21.0, at(2)), (22.0, at(3))]
    c = [(30.0, at(1)), (31.0, at(2)), (32.0, at(3))]
    out, st = merge(_as_replicas([a, b, c]), strategy=0)
    assert st == [0] and out[0] == [(30.0, at(1)), (21.0, at(2)), (22.0, at(3))]


def test_batch_of_series_and_blocks():
    rng = np.random.default_rng(0)
    series = []
    for s in range(20):
        base = np.sort(rng.choice(np.arange(1, 400), size=120, replace=False))
        replicas = []
        for r in range(3):
            keep = base[rng.random(len(base)) < 0.9]
            pts = [(float(k * 10 + r), at(int(k))) for k in keep]
            blocks = [[[p for p in pts if at(0) + b * 100 * SEC <= p[1] < at(0) + (b + 1) * 100 * SEC]]
                      for b in range(4)]
            replicas.append(blocks)
        series.append(replicas)
    out, st = merge(series)
    assert st == [0] * 20
    for s in range(20):
        tss = [t for _, t in out[s]]
        assert tss == sorted(set(tss))
        union = set()
        for rep in series[s]:
            for sl in rep:
                for rd in sl:
                    union |= {t for _, t in rd}
        assert set(tss) == union
