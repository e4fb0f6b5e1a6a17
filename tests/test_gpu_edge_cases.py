"""GPU parity, edge cases: batch shapes, ragged lengths, empty series, random
structured streams (units / annotations / value families) against the oracle."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
SEC = 1_000_000_000


@pytest.fixture(scope="module")
def codecs():
    from m3_b200.codec import BatchCodec
    return {True: BatchCodec(0, True), False: BatchCodec(0, False)}


def _oracle_stream(dps, start, int_opt):
    """dps: list of (ts, value, unit, annotation)."""
    e = O.Encoder(0, int_opt)
    e.reset(start)
    for t, v, u, a in dps:
        assert e.encode(t, v, u, a) == 0
    return e.stream() or b""


def _upload(streams):
    blob = b"".join(streams)
    off = np.zeros(len(streams) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(s) for s in streams])
    buf = torch.zeros(len(blob) + 16, dtype=torch.uint8, device="cuda")
    if blob:
        buf[: len(blob)] = torch.frombuffer(bytearray(blob), dtype=torch.uint8).cuda()
    return buf[: len(blob)], torch.from_numpy(off).cuda()


@pytest.mark.parametrize("n_series", [1, 2, 31, 32, 33, 127, 129])
@pytest.mark.parametrize("int_opt", [True, False])
def test_batch_shapes(codecs, n_series, int_opt):
    """Batches that do not fill a warp / a block, including a single series."""
    from m3_b200 import synth
    P = 97
    ts, vals, start = synth.gaussian_walk(n_series, P, "cuda", seed=n_series)
    vals[::3] = torch.round(vals[::3] * 10) / 10
    codec = codecs[int_opt]
    enc = codec.encode(ts, vals, start, unit=O.UNIT_S)
    torch.cuda.synchronize()
    o_out, o_len, o_st = O.encode_batch(ts.cpu().numpy(), vals.cpu().numpy(), int(start[0]), O.UNIT_S,
                                        int_opt)
    g_len = enc.out_len.cpu().numpy()
    g_out = enc.out.cpu().numpy()
    assert (enc.status.cpu().numpy() == 0).all()
    for s in range(n_series):
        assert g_len[s] == o_len[s] and (g_out[s, : g_len[s]] == o_out[s, : o_len[s]]).all(), s
    packed, offsets = codec.compact(enc, align=1)
    dec = codec.decode(packed, offsets, P)
    torch.cuda.synchronize()
    assert (dec.n_points.cpu().numpy() == P).all() and (dec.status.cpu().numpy() == 0).all()
    assert torch.equal(dec.ts, ts)
    for s in range(n_series):
        ots, ovals, n, err = O.decode_series(o_out[s, : o_len[s]].tobytes(), int_opt, cap=P)
        assert err == 0 and n == P
        assert (dec.values[s].cpu().numpy().view(np.uint64) == ovals.view(np.uint64)).all()


@pytest.mark.parametrize("int_opt", [True, False])
def test_ragged_n_points_including_empty(codecs, int_opt):
    rng = np.random.default_rng(4)
    S, P = 200, 150
    start = 1599955200 * SEC
    n_pts = rng.integers(0, P + 1, size=S).astype(np.int32)
    n_pts[:8] = [0, 1, 2, P, 0, 7, 8, 9]
    ts = start + np.cumsum(rng.integers(1, 90, size=(S, P)), axis=1) * SEC
    vals = np.round(100 + np.cumsum(rng.normal(size=(S, P)), axis=1), 3)
    codec = codecs[int_opt]
    enc = codec.encode(torch.from_numpy(ts).cuda(), torch.from_numpy(vals).cuda(),
                       torch.full((S,), start, dtype=torch.int64, device="cuda"), unit=O.UNIT_S,
                       n_points=torch.from_numpy(n_pts).cuda())
    torch.cuda.synchronize()
    assert (enc.status.cpu().numpy() == 0).all()
    g_len = enc.out_len.cpu().numpy()
    g_out = enc.out.cpu().numpy()
    streams = []
    for s in range(S):
        n = int(n_pts[s])
        exp = _oracle_stream([(int(ts[s, i]), float(vals[s, i]), O.UNIT_S, b"") for i in range(n)], start,
                             int_opt)
        assert g_len[s] == len(exp), (s, n)
        assert g_out[s, : g_len[s]].tobytes() == exp, (s, n)
        streams.append(exp)
    # decode the ragged batch: empty streams report io.EOF with zero points, like the reference
    d, off = _upload(streams)
    dec = codec.decode(d, off, P)
    torch.cuda.synchronize()
    n = dec.n_points.cpu().numpy()
    st = dec.status.cpu().numpy()
    for s in range(S):
        assert n[s] == n_pts[s]
        assert st[s] == (O.ERR_EOF if n_pts[s] == 0 else 0)
        assert (dec.ts[s, : n[s]].cpu().numpy() == ts[s, : n[s]]).all()


@pytest.mark.parametrize("int_opt", [True, False])
def test_random_structured_streams_vs_oracle(codecs, int_opt):
    """Random unit changes, annotations (short, repeated, long), value families and
    timestamp patterns: the GPU encoder (through per-datapoint units + sparse
    annotations) and decoder against the oracle, stream by stream."""
    from m3_b200 import capi
    rng = np.random.default_rng(21)
    S, P = 96, 120
    start = 1427162400 * SEC
    codec = codecs[int_opt]
    ts = np.zeros((S, P), dtype=np.int64)
    vals = np.zeros((S, P), dtype=np.float64)
    units = np.zeros((S, P), dtype=np.uint8)
    ann_entries, ann_off, blob = [], [0], bytearray()
    expected = []
    for s in range(S):
        t = start + int(rng.integers(0, 3)) * 500_000_000  # sometimes not second-aligned
        unit = int(rng.choice([1, 1, 1, 2, 3, 4]))
        dps = []
        last_ann = None
        fam = s % 6
        for i in range(P):
            un = {1: SEC, 2: 10 ** 6, 3: 10 ** 3, 4: 1}[unit]
            t += int(rng.integers(1, 3000)) * un
            if fam == 0:
                v = float(rng.normal() * 100)
            elif fam == 1:
                v = float(rng.integers(-1000, 1000))
            elif fam == 2:
                v = float(np.round(rng.normal() * 10, 2))
            elif fam == 3:
                v = float(rng.integers(0, 5)) if rng.random() < 0.8 else float(rng.normal())
            elif fam == 4:
                v = float(rng.integers(0, 3) * 1e12)
            else:
                v = float(np.round(rng.normal(), int(rng.integers(0, 7))))
            if rng.random() < 0.04:
                unit = int(rng.choice([1, 2, 3, 4]))
            a = b""
            r = rng.random()
            if r < 0.03:
                a = bytes(rng.integers(0, 256, size=int(rng.integers(1, 20)), dtype=np.uint8))
            elif r < 0.05 and last_ann:
                a = last_ann  # repeated annotation: must not be rewritten
            elif r < 0.055:
                a = bytes(rng.integers(0, 256, size=300, dtype=np.uint8))  # long
            if a:
                last_ann = a
                ann_entries.append((i, len(a), len(blob)))
                blob += a
            ts[s, i], vals[s, i], units[s, i] = t, v, unit
            dps.append((t, v, unit, a))
        ann_off.append(len(ann_entries))
        expected.append(_oracle_stream(dps, start, int_opt))
    ent = np.zeros(len(ann_entries), dtype=[("dp", "<u4"), ("len", "<u4"), ("off", "<u8")])
    for k, e in enumerate(ann_entries):
        ent[k] = e
    ann = (torch.tensor(ann_off, dtype=torch.int64, device="cuda"),
           torch.from_numpy(ent.view(np.uint8).reshape(-1, 16).copy()).cuda(),
           torch.frombuffer(bytearray(bytes(blob) + b"\0"), dtype=torch.uint8).cuda())
    stride = codec.encode_bound(P) + ((len(blob) + 16 * len(ann_entries) + 15) // 16) * 16
    enc = codec.encode(torch.from_numpy(ts).cuda(), torch.from_numpy(vals).cuda(),
                       torch.full((S,), start, dtype=torch.int64, device="cuda"), unit=O.UNIT_S,
                       units=torch.from_numpy(units).cuda(), annotations=ann, out_stride=stride)
    torch.cuda.synchronize()
    assert (enc.status.cpu().numpy() == 0).all()
    g_len = enc.out_len.cpu().numpy()
    g_out = enc.out.cpu().numpy()
    for s in range(S):
        assert g_out[s, : g_len[s]].tobytes() == expected[s], (s, s % 6)
    d, off = _upload(expected)
    dec = codec.decode(d, off, P, want_annotations=True)
    torch.cuda.synchronize()
    assert (dec.status.cpu().numpy() == 0).all() and (dec.n_points.cpu().numpy() == P).all()
    # NB: decoded timestamps are compared with the ORACLE's decode, not with the inputs:
    # after a unit change deltas that are not multiples of the new unit are truncated
    # by the reference encoder itself
    gt = dec.ts.cpu().numpy()
    gv = dec.values.cpu().numpy().view(np.uint64)
    gu = dec.unit.cpu().numpy()
    for s in range(S):
        dps, err = O.decode_all(expected[s], int_opt)
        assert err == 0
        assert (gt[s] == np.array([d_[0] for d_ in dps], dtype=np.int64)).all(), s
        assert (gv[s] == np.array([d_[1] for d_ in dps]).view(np.uint64)).all(), s
        assert gu[s] == dps[-1][2]


def test_max_points_one_and_tiny_capacity(codecs):
    from m3_b200 import synth
    ts, vals, start = synth.gaussian_walk(40, 10, "cuda", seed=2)
    codec = codecs[True]
    enc = codec.encode(ts, vals, start, unit=O.UNIT_S)
    packed, offsets = codec.compact(enc, align=4)
    dec = codec.decode(packed, offsets, 1)
    torch.cuda.synchronize()
    assert (dec.n_points.cpu().numpy() == 10).all() and (dec.status.cpu().numpy() == 100).all()
    assert torch.equal(dec.ts[:, 0], ts[:, 0])


def test_downsample_window_edges(codecs):
    """Windows that do not cover the series, a range starting mid-series, and one
    datapoint per window."""
    rng = np.random.default_rng(8)
    S, P = 64, 240
    start = 1599955200 * SEC
    ts = start + np.arange(P, dtype=np.int64)[None, :] * 60 * SEC + np.zeros((S, 1), dtype=np.int64)
    vals = 100 + np.cumsum(rng.normal(size=(S, P)), axis=1)
    vals[5, 10:20] = np.nan
    codec = codecs[True]
    enc = codec.encode(torch.from_numpy(ts).cuda(), torch.from_numpy(vals).cuda(),
                       torch.full((S,), start, dtype=torch.int64, device="cuda"), unit=O.UNIT_S)
    packed, offsets = codec.compact(enc, align=16)
    for rs, win, nw in [(start + 3600 * SEC, 300 * SEC, 12), (start - 600 * SEC, 60 * SEC, 400),
                        (start, 7 * 60 * SEC, 5)]:
        r = codec.decode_downsample(packed, offsets, rs, win, nw)
        torch.cuda.synchronize()
        assert (r.status.cpu().numpy() == 0).all()
        gs, gc = r.sum.cpu().numpy(), r.count.cpu().numpy()
        gmn, gmx = r.min.cpu().numpy(), r.max.cpu().numpy()
        for s in (0, 5, 63):
            dps, _ = O.decode_all(enc.out[s, : int(enc.out_len[s])].cpu().numpy().tobytes(), True)
            ots = np.array([d_[0] for d_ in dps], dtype=np.int64)
            ovs = np.array([d_[1] for d_ in dps], dtype=np.float64)
            es, ec, emn, emx, _ = O.downsample_series(ots, ovs, rs, win, nw)
            assert (gc[:, s] == ec).all()
            assert (gs[:, s].view(np.uint64) == es.view(np.uint64)).all()
            assert (gmn[:, s].view(np.uint64) == emn.view(np.uint64)).all()
            assert (gmx[:, s].view(np.uint64) == emx.view(np.uint64)).all()


def test_output_arrays_not_sector_aligned(codecs):
    """The decoder stores whole 32-byte sectors when it can; output arrays that are only
    8-byte aligned (a view one element into an allocation) take the row-by-row path and
    must give the same result."""
    from m3_b200 import synth
    from m3_b200.codec import DecodeResult
    S, P = 70, 64  # P % 4 == 0
    codec = codecs[True]
    ts, vals, start = synth.gaussian_walk(S, P, "cuda", seed=5)
    enc = codec.encode(ts, vals, start, unit=O.UNIT_S)
    packed, offsets = codec.compact(enc, align=64)
    ref = codec.decode(packed, offsets, P)
    raw_t = torch.zeros(S * P + 1, dtype=torch.int64, device="cuda")
    raw_v = torch.zeros(S * P + 1, dtype=torch.float64, device="cuda")
    out = DecodeResult(ts=raw_t[1:].view(S, P), values=raw_v[1:].view(S, P),
                       n_points=torch.empty(S, dtype=torch.int32, device="cuda"),
                       status=torch.empty(S, dtype=torch.int32, device="cuda"),
                       unit=torch.empty(S, dtype=torch.uint8, device="cuda"), annotations=None)
    assert out.ts.data_ptr() % 32 == 8
    codec.decode(packed, offsets, P, out=out)
    torch.cuda.synchronize()
    assert torch.equal(out.ts, ref.ts) and torch.equal(out.ts, ts)
    assert torch.equal(out.values.view(torch.int64), ref.values.view(torch.int64))
    assert bool((out.n_points == P).all()) and bool((out.status == 0).all())
    assert int(raw_t[0]) == 0 and float(raw_v[0]) == 0.0  # nothing written before the view
