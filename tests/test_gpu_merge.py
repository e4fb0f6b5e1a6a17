"""GPU parity of the series-merge row against the merge oracle:
the reference's own table tests (through m3tsz_merge_series_batch) and random
multi-replica / multi-block batches, plus the full fetch path
decode (GPU) -> merge (GPU) on real M3TSZ streams."""
import numpy as np
import pytest

import oracle_lib as O
from test_merge_oracle import COMMON, START, V0, V1, at, build

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
SEC = 1_000_000_000


@pytest.fixture(scope="module")
def codec():
    from m3_b200.codec import BatchCodec
    return BatchCodec(0, True)


def gpu_merge(codec, series, start=0, end=0, strategy=0, out_cap=None):
    ts, val, npts, st, slice_off, replica_off, series_off = build(series)
    if out_cap is None:
        out_cap = max(1, int(npts.sum()))
    d = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).cuda()
    r = codec.merge_series(d(ts, np.int64), d(val, np.float64), d(npts.astype(np.int32), np.int32),
                           d(st, np.int32), d(slice_off, np.int64), d(replica_off, np.int64),
                           d(series_off, np.int64), out_cap, start, end, strategy)
    # the same merge over POINT-major arrays ([cap][n_seq] in, [out_cap][n_series] out) must agree
    rp = codec.merge_series(d(ts.T, np.int64), d(val.T, np.float64), d(npts.astype(np.int32), np.int32),
                            d(st, np.int32), d(slice_off, np.int64), d(replica_off, np.int64),
                            d(series_off, np.int64), out_cap, start, end, strategy, point_major=True)
    torch.cuda.synchronize()
    assert torch.equal(r[2], rp[2]) and torch.equal(r[3], rp[3])
    n_host = r[2].cpu().numpy().view(np.uint32)
    a_ts, a_v = r[0].cpu().numpy(), r[1].cpu().numpy().view(np.uint64)
    b_ts, b_v = rp[0].cpu().numpy().T, rp[1].cpu().numpy().view(np.uint64).T
    for i in range(len(n_host)):
        k = min(int(n_host[i]), out_cap)
        assert (a_ts[i, :k] == b_ts[i, :k]).all() and (a_v[i, :k] == b_v[i, :k]).all(), ("point-major", i)
    ts_o, val_o, n_o, st_o = [x.cpu().numpy() for x in r]
    exp = O.series_merge_batch(ts, val, npts, st, slice_off, replica_off, series_off, start=start, end=end,
                               strategy=strategy, out_cap=out_cap)
    return (ts_o, val_o, n_o.view(np.uint32), st_o), exp


def assert_same(got, exp):
    ts_o, val_o, n_o, st_o = got
    ets, eval_, en, est = exp
    assert (n_o == en).all(), (n_o, en)
    assert (st_o == est).all(), (st_o, est)
    for s in range(len(en)):
        k = min(int(en[s]), ts_o.shape[1])
        assert (ts_o[s, :k] == ets[s, :k]).all(), s
        assert (val_o[s, :k].view(np.uint64) == eval_[s, :k].view(np.uint64)).all(), s


def test_reference_tables_on_gpu(codec):
    a = [(1.0, at(1)), (2.0, at(2)), (3.0, at(3))]
    c = [(3.0, at(3)), (4.0, at(4)), (5.0, at(5))]
    oo = [(1.0, at(1)), (3.0, at(3)), (2.0, at(2))]
    rng_filter = [(0.0, at(-2)), (1.0, at(-1)), (2.0, at(0)), (3.0, at(1)), (4.0, at(60)), (5.0, at(61))]
    cases = [
        ([[[[V0, V1]]]], {}),                                   # MergesMulti
        ([[[[V0], [V1]]]], {}),                                 # ReadsSlicesInOrder
        ([[[[V0], [], [V1]]]], {}),                             # SlicesWithNoEntries
        ([[[[V0], [[]], [V1]]]], {}),                           # SlicesWithEmptyEntries
        ([[[[[(1.0, at(1)), (2.0, at(2)), (2.0, at(2))]]]]], {}),  # DeduplicatesSingle
        ([[[[V0, V0, V0]]]], {}),                               # DeduplicatesMulti
        ([[[[oo]]]], {}),                                       # ErrorOnOutOfOrder (MRI level)
        ([[[[(V0[:2], 77)]]]], {}),                             # inner iterator error
        ([[[[a]], [[a]], [[c]]]], dict(start=START, end=START + 60 * SEC)),       # MergesReplicas
        ([[[[rng_filter]]]], dict(start=START, end=START + 60 * SEC)),           # FiltersToRange
        ([[[[a]], [[[]]], [[a]]]], dict(start=START, end=START + 60 * SEC)),     # IgnoresEmptyReplicas
        ([[[[([], 55)]]]], dict(start=START, end=START + 60 * SEC)),             # replica with error
        # several failing replicas: the LAST error is the series' error (series_iterator.go:157-168)
        ([[[[([], 55)]], [[([], 66)]], [[a]]]], dict(start=START, end=START + 60 * SEC)),
        ([[[[([], 66)]], [[a]], [[([], 55)]]]], dict(start=START, end=START + 60 * SEC)),
        ([[[[oo]]]], dict(start=START, end=START + 60 * SEC)),                   # out of order (series)
    ]
    for strategy in (0, 1, 2, 3):
        cases.append(([[[[v]] for v in COMMON]], dict(strategy=strategy)))
    for series, kw in cases:
        got, exp = gpu_merge(codec, series, **kw)
        assert_same(got, exp)
    # all cases as ONE batch (heterogeneous series in the same launch, no filter)
    batch = [c[0][0] for c in cases if not c[1]]
    got, exp = gpu_merge(codec, batch)
    assert_same(got, exp)


@pytest.mark.parametrize("strategy", [0, 1, 2, 3])
def test_random_replicas_and_blocks(codec, strategy):
    rng = np.random.default_rng(100 + strategy)
    series = []
    for s in range(120):
        n_rep = int(rng.integers(1, 5))
        base = np.sort(rng.choice(np.arange(1, 600), size=int(rng.integers(0, 200)), replace=False))
        replicas = []
        for r in range(n_rep):
            keep = base[rng.random(len(base)) < 0.85]
            # values: mostly agree across replicas, sometimes differ (exercises the strategies)
            pts = [(float(k) if rng.random() < 0.8 else float(k * 10 + r), at(int(k))) for k in keep]
            if rng.random() < 0.1 and len(pts) > 3:  # duplicate timestamp inside one reader
                j = int(rng.integers(1, len(pts)))
                pts.insert(j, (pts[j - 1][0] + 0.5, pts[j - 1][1]))
            n_blocks = int(rng.integers(1, 5))
            bounds = np.linspace(0, 600, n_blocks + 1)
            slices = []
            for b in range(n_blocks):
                blk = [p for p in pts if at(bounds[b]) <= p[1] < at(bounds[b + 1])]
                if rng.random() < 0.2 and len(blk) > 2:  # two unmerged readers for this block
                    slices.append([blk[::2], blk[1::2]])
                else:
                    slices.append([blk])
            if rng.random() < 0.03 and slices and slices[-1][0]:
                slices[-1][0] = (slices[-1][0], 1)  # the last reader ends with a decode error (EOF)
            replicas.append(slices)
        series.append(replicas)
    for kw in ({}, dict(start=at(100), end=at(400))):
        got, exp = gpu_merge(codec, series, strategy=strategy, **kw)
        assert_same(got, exp)
    got, exp = gpu_merge(codec, series, strategy=strategy, out_cap=16)  # capacity status
    ts_o, val_o, n_o, st_o = got
    ets, eval_, en, est = exp
    assert (n_o == en).all()
    assert ((st_o == 100) == ((est == 0) & (en > 16))).all()


@pytest.mark.parametrize("strategy", [0, 2])
def test_many_replicas_and_readers(codec, strategy):
    """up to 12 replicas per series and 12 unmerged readers per block slice (the limit of the merge:
    the reference's sort.Slice is a stable insertion sort up to 12 elements), 13 is refused"""
    rng = np.random.default_rng(300 + strategy)
    series = []
    for s in range(40):
        n_rep = int(rng.integers(5, 13))
        base = np.sort(rng.choice(np.arange(1, 400), size=int(rng.integers(20, 120)), replace=False))
        replicas = []
        for r in range(n_rep):
            keep = base[rng.random(len(base)) < 0.8]
            pts = [(float(k) if rng.random() < 0.7 else float(k * 10 + r), at(int(k))) for k in keep]
            n_readers = int(rng.integers(1, 13)) if r == 0 else 1
            replicas.append([[pts[i::n_readers] for i in range(n_readers)]])
        series.append(replicas)
    got, exp = gpu_merge(codec, series, strategy=strategy)
    assert_same(got, exp)
    # 13 replicas: refused with the documented status, like the oracle
    pts = [(1.0, at(1)), (2.0, at(2))]
    got, exp = gpu_merge(codec, [[[[pts]] for _ in range(13)]], strategy=strategy)
    assert (got[3] == exp[3]).all() and int(got[3][0]) == 14


def test_fetch_path_decode_then_merge(codec):
    """RF=3 fetch: each replica of each series = 2 blocks of real M3TSZ streams (one
    replica misses a few writes); decode all streams in one launch, merge in one launch."""
    rng = np.random.default_rng(5)
    S, P, R, B = 48, 120, 3, 2
    start = 1599955200 * SEC
    streams, truth = [], []
    for s in range(S):
        ts_full = start + (np.arange(B * P) * 60 + rng.integers(0, 30, size=B * P)) * SEC
        vals_full = np.round(100 + np.cumsum(rng.normal(size=B * P)), 2)
        truth.append((ts_full, vals_full))
        for r in range(R):
            for b in range(B):
                sl = slice(b * P, (b + 1) * P)
                keep = np.ones(P, dtype=bool)
                if r == 1:
                    keep = rng.random(P) < 0.9
                blk_start = start + b * P * 60 * SEC
                streams.append(O.encode_series(ts_full[sl][keep], vals_full[sl][keep], blk_start, O.UNIT_S, True))
    blob = b"".join(streams)
    off = np.zeros(len(streams) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(x) for x in streams])
    buf = torch.zeros(len(blob) + 16, dtype=torch.uint8, device="cuda")
    buf[: len(blob)] = torch.frombuffer(bytearray(blob), dtype=torch.uint8).cuda()
    dec = codec.decode(buf[: len(blob)], torch.from_numpy(off).cuda(), P)
    n_seq = len(streams)
    slice_off = torch.arange(n_seq + 1, dtype=torch.int64, device="cuda")          # one reader per slice
    replica_off = torch.arange(0, n_seq + 1, B, dtype=torch.int64, device="cuda")  # B slices per replica
    series_off = torch.arange(0, S * R + 1, R, dtype=torch.int64, device="cuda")   # R replicas per series
    ts_o, val_o, n_o, st_o = codec.merge_series(dec.ts, dec.values, dec.n_points, dec.status, slice_off,
                                                replica_off, series_off, B * P)
    torch.cuda.synchronize()
    assert (st_o.cpu().numpy() == 0).all()
    n_o = n_o.cpu().numpy()
    for s in range(S):
        assert n_o[s] == B * P
        assert (ts_o[s].cpu().numpy() == truth[s][0]).all()
        assert (val_o[s].cpu().numpy() == truth[s][1]).all()
