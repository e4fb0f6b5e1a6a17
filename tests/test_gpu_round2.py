"""GPU parity tests of the round-2 additions, through the C ABI against the CPU oracle:
Gauge `last` in the fused downsample (lockstep / out-of-phase / out-of-order series),
per-datapoint unit + annotation events, non-CSR stream placement (offset + length),
packed encode (no slots, no compaction pass)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

SEC = 1_000_000_000
START = 1599955200 * SEC


@pytest.fixture(scope="module")
def codecs():
    from m3_b200.codec import BatchCodec
    return {True: BatchCodec(0, True), False: BatchCodec(0, False)}


def to_device_streams(streams, align=1):
    off = np.zeros(len(streams) + 1, dtype=np.int64)
    pos = 0
    starts = []
    for s in streams:
        pos = (pos + align - 1) // align * align
        starts.append(pos)
        pos += len(s)
    buf = np.zeros(pos + 16, dtype=np.uint8)
    for st, s in zip(starts, streams):
        buf[st:st + len(s)] = np.frombuffer(s, dtype=np.uint8)
    off[:-1] = starts
    off[-1] = pos
    d = torch.from_numpy(buf).cuda()
    return d[:pos], torch.from_numpy(off).cuda()


def oracle_decode(stream, int_opt):
    dps, err = O.decode_all(stream, int_opt)
    ts = np.array([d[0] for d in dps], dtype=np.int64)
    vals = np.array([d[1] for d in dps], dtype=np.float64)
    return ts, vals, err, dps


def _ds_series(rng, S, P):
    """Families for the downsample sink: lockstep minute cadence (the hot path with
    warp-uniform window advance), out-of-phase cadences, jitter, gaps wider than a window,
    out-of-order timestamps, NaN / Inf / -0, integers, repeats, timestamps before / after the range."""
    ts = np.zeros((S, P), dtype=np.int64)
    vals = np.zeros((S, P), dtype=np.float64)
    for s in range(S):
        fam = (s // 32) % 8 if s >= 64 else s % 8  # whole warps of one family + mixed warps
        walk = 100.0 + np.cumsum(rng.normal(size=P))
        if fam == 0:
            ts[s] = START + np.arange(P) * 60 * SEC
            vals[s] = walk
        elif fam == 1:  # different phase / cadence per series
            ts[s] = START + int(rng.integers(0, 300)) * SEC + np.arange(P) * int(rng.integers(7, 200)) * SEC
            vals[s] = walk
        elif fam == 2:  # jitter
            ts[s] = START + np.cumsum(rng.integers(1, 120, size=P)) * SEC
            vals[s] = np.round(walk, 2)
        elif fam == 3:  # gaps of several windows
            ts[s] = START + np.cumsum(rng.integers(1, 2000, size=P)) * SEC
            vals[s] = walk
        elif fam == 4:  # out of order
            t = START + np.arange(P) * 45 * SEC
            for _ in range(6):
                i = int(rng.integers(12, P - 2))
                j = i - int(rng.integers(1, 12))
                t[i], t[j] = t[j], t[i]
            ts[s] = t
            vals[s] = walk
        elif fam == 5:  # NaN / Inf / -0
            ts[s] = START + np.arange(P) * 60 * SEC
            v = walk.copy()
            v[rng.integers(0, P, size=P // 6)] = np.nan
            v[rng.integers(0, P, size=3)] = np.inf
            v[rng.integers(0, P, size=3)] = -np.inf
            v[rng.integers(0, P, size=8)] = -0.0
            v[rng.integers(0, P, size=8)] = 0.0
            vals[s] = v
        elif fam == 6:  # integers and repeats (int mode under int_opt)
            ts[s] = START + np.arange(P) * 60 * SEC
            vals[s] = np.repeat(np.round(walk[: (P + 3) // 4]), 4)[:P]
        else:  # starts before the range and runs past its end
            ts[s] = START - 1000 * SEC + np.arange(P) * 150 * SEC
            vals[s] = walk
    return ts, vals


@pytest.mark.parametrize("int_opt", [True, False])
@pytest.mark.parametrize("want_last", [False, True])
def test_downsample_last_vs_oracle(codecs, int_opt, want_last):
    rng = np.random.default_rng(21)
    S, P = 320, 400
    ts, vals = _ds_series(rng, S, P)
    o_out, o_len, o_st = O.encode_batch(ts, vals, START, O.UNIT_S, int_opt, n_threads=8)
    assert (o_st == 0).all()
    streams = [o_out[s, : o_len[s]].tobytes() for s in range(S)]
    for align in (1, 64):
        d, off = to_device_streams(streams, align)
        window, n_win = 300 * SEC, 90
        r = codecs[int_opt].decode_downsample(d, off, START, window, n_win, want_last=want_last)
        torch.cuda.synchronize()
        assert (r.status.cpu().numpy() == 0).all()
        assert (r.n_points.cpu().numpy() == P).all()
        gs, gc = r.sum.cpu().numpy(), r.count.cpu().numpy()
        gmn, gmx = r.min.cpu().numpy(), r.max.cpu().numpy()
        gl = r.last.cpu().numpy() if want_last else None
        for s in range(S):
            ots, ovals, err, _ = oracle_decode(streams[s], int_opt)
            assert err == 0
            es, ec, emn, emx, el = O.downsample_series(ots, ovals, START, window, n_win)
            assert (gc[:, s] == ec).all(), (s, align)
            assert (gs[:, s].view(np.uint64) == es.view(np.uint64)).all(), (s, align)
            assert (gmn[:, s].view(np.uint64) == emn.view(np.uint64)).all(), (s, align)
            assert (gmx[:, s].view(np.uint64) == emx.view(np.uint64)).all(), (s, align)
            if want_last:
                assert (gl[:, s].view(np.uint64) == el.view(np.uint64)).all(), (s, align)


def test_downsample_gauge_test_vectors(codecs):
    """The reference's own Gauge tables (aggregator/aggregation/gauge_test.go:35-57,173-…,
    TestGaugeLastOutOfOrderValues) pushed through encode -> fused decode+downsample."""
    codec = codecs[True]
    # 1..100 at increasing times, one window: last 100, count 100, sum 5050, min 1, max 100
    ts = START + np.arange(100, dtype=np.int64) * SEC
    v = np.arange(1, 101, dtype=np.float64)
    # out-of-order `last`: mid 42, pre 41, after 43, prepre 40 -> last 43
    mid = START + 60 * SEC
    ts2 = np.array([mid, mid - SEC, mid + SEC, mid - SEC], dtype=np.int64)
    v2 = np.array([42.0, 41.0, 43.0, 40.0])
    streams = [O.encode_series(ts, v, START, O.UNIT_S, True), O.encode_series(ts2, v2, START, O.UNIT_S, True)]
    d, off = to_device_streams(streams)
    r = codec.decode_downsample(d, off, START, 1000 * SEC, 2, want_last=True)
    torch.cuda.synchronize()
    assert r.count[:, 0].tolist() == [100, 0] and r.sum[0, 0].item() == 5050.0
    assert r.min[0, 0].item() == 1.0 and r.max[0, 0].item() == 100.0 and r.last[0, 0].item() == 100.0
    assert r.sum[0, 0].item() / r.count[0, 0].item() == 50.5  # Mean()
    # empty gauge: 0 / NaN / NaN / last 0
    assert r.sum[1, 0].item() == 0.0 and np.isnan(r.min[1, 0].item()) and np.isnan(r.max[1, 0].item())
    assert r.last[1, 0].item() == 0.0
    assert r.last[0, 1].item() == 43.0 and r.count[0, 1].item() == 4
    assert r.last_at[0, 1].item() == mid + SEC


def _annotated_streams(rng, S, P, int_opt):
    """Oracle-encoded streams with unit changes and annotations at random datapoints."""
    streams, expect = [], []
    units = [O.UNIT_S, O.UNIT_MS, O.UNIT_US, O.UNIT_NS]
    for s in range(S):
        e = O.Encoder(START, int_opt)
        t = START
        unit = O.UNIT_S
        rows = []
        v = 100.0
        for i in range(P):
            if s % 3 and rng.random() < 0.04:
                unit = units[int(rng.integers(0, 4))]
            t += int(rng.integers(1, 90)) * SEC
            v += float(rng.normal())
            ann = b""
            if s % 2 and rng.random() < 0.05:
                ann = bytes(rng.integers(0, 256, size=int(rng.integers(1, 40)), dtype=np.uint8))
            if i == 0 and s % 5 == 0:
                ann = b"first-annotation-%d" % s
            assert e.encode(t, v if s % 4 else round(v), unit, ann) == 0
            rows.append((t, unit, ann))
        streams.append(e.stream())
        expect.append(rows)
    return streams, expect


def _extract_bits(buf, bit_off, nbytes):
    out = bytearray()
    for k in range(nbytes):
        b = bit_off + 8 * k
        i, sh = b >> 3, b & 7
        w = (buf[i] << 8) | (buf[i + 1] if i + 1 < len(buf) else 0)
        out.append((w >> (8 - sh)) & 0xFF)
    return bytes(out)


@pytest.mark.parametrize("int_opt", [True, False])
def test_per_datapoint_unit_and_annotation_events(codecs, int_opt):
    from m3_b200 import capi
    rng = np.random.default_rng(5)
    S, P = 96, 120
    streams, _ = _annotated_streams(rng, S, P, int_opt)
    d, off = to_device_streams(streams)
    r = codecs[int_opt].decode(d, off, P + 2, want_annotations=True, want_events=S * P)
    torch.cuda.synchronize()
    n_ev = int(r.event_count.item())
    assert n_ev <= S * P
    ev = (capi.DpEvent * n_ev).from_buffer_copy(r.events.cpu().numpy().tobytes()[: 32 * n_ev])
    by_series = {}
    for e in ev:
        by_series.setdefault(int(e.series), []).append(e)
    first = r.unit_first.cpu().numpy()
    n = r.n_points.cpu().numpy()
    for s in range(S):
        dps, err = O.decode_all(streams[s], int_opt)
        assert err == 0 and n[s] == len(dps) == P
        # reconstruct Current()'s (unit, annotation) of every datapoint from the events
        unit = int(first[s])
        evs = sorted(by_series.get(s, []), key=lambda e: (e.dp_index, e.kind))
        units = {}
        anns = {}
        for e in evs:
            if e.kind == capi.EVENT_TIME_UNIT:
                units[e.dp_index] = int(e.unit)
            else:
                anns[e.dp_index] = _extract_bits(streams[s], int(e.bit_offset), int(e.length))
        for i, (t, v, u, a) in enumerate(dps):
            unit = units.get(i, unit)
            assert unit == u, (s, i)
            # the reference's Current() keeps returning the annotation only for the datapoint that
            # carried it (iterator.go:229-231 after timestamp_iterator.go:96 resets PrevAnt)
            assert anns.get(i, b"") == a, (s, i)


@pytest.mark.parametrize("int_opt", [True, False])
def test_decode_offsets_plus_lengths(codecs, int_opt):
    """Streams placed in arbitrary order with gaps (index-entry (Offset, Size) addressing)."""
    from m3_b200 import synth
    rng = np.random.default_rng(8)
    S, P = 300, 97
    ts, vals, start = synth.gaussian_walk(S, P, "cpu", seed=11)
    o_out, o_len, o_st = O.encode_batch(ts.numpy(), vals.numpy(), start.numpy(), O.UNIT_S, int_opt, n_threads=4)
    order = rng.permutation(S)
    starts = np.zeros(S, dtype=np.int64)
    pos = 0
    for s in order:
        pos += int(rng.integers(0, 40))
        starts[s] = pos
        pos += int(o_len[s])
    buf = np.full(pos + 64, 0xAB, dtype=np.uint8)  # garbage between the streams
    for s in range(S):
        buf[int(starts[s]):int(starts[s]) + int(o_len[s])] = o_out[s, : int(o_len[s])]
    d = torch.from_numpy(buf).cuda()
    r = codecs[int_opt].decode(d, torch.from_numpy(starts).cuda(), P,
                               lengths=torch.from_numpy(o_len.astype(np.int64)).cuda())
    torch.cuda.synchronize()
    assert (r.status.cpu().numpy() == 0).all() and (r.n_points.cpu().numpy() == P).all()
    assert torch.equal(r.ts.cpu(), ts)
    for s in range(0, S, 7):
        _, ovals, err, _ = oracle_decode(o_out[s, : o_len[s]].tobytes(), int_opt)
        assert (r.values[s].cpu().numpy().view(np.uint64) == ovals.view(np.uint64)).all()


def _mixed(rng, S, P):
    import test_gpu_parity as T
    return T._mixed_series(rng, S, P)


@pytest.mark.parametrize("int_opt", [True, False])
@pytest.mark.parametrize("align", [64, 16, 1])
def test_encode_packed_byte_identical(codecs, int_opt, align):
    rng = np.random.default_rng(31)
    S, P = 1500, 200  # > one batch per resident warp is not needed for correctness; ragged tail warp
    ts, vals, start = _mixed(rng, S, P)
    n_points = rng.integers(0, P + 1, size=S).astype(np.int32)
    n_points[::3] = P
    codec = codecs[int_opt]
    d_ts, d_vals = torch.from_numpy(ts).cuda(), torch.from_numpy(vals).cuda()
    d_start = torch.full((S,), start, dtype=torch.int64, device="cuda")
    r = codec.encode_packed(d_ts, d_vals, d_start, unit=O.UNIT_S, n_points=torch.from_numpy(n_points).cuda(),
                            align=align)
    torch.cuda.synchronize()
    st = r.status.cpu().numpy()
    assert (st == 0).all(), st[st != 0]
    g_len, g_off = r.out_len.cpu().numpy(), r.offsets.cpu().numpy()
    packed = r.packed.cpu().numpy()
    total = int(r.total.item())
    assert (g_off % align == 0).all()
    # streams do not overlap and fill [0, total) up to alignment padding
    iv = sorted((int(g_off[s]), int(g_off[s] + g_len[s])) for s in range(S) if g_len[s])
    for (a0, a1), (b0, b1) in zip(iv, iv[1:]):
        assert a1 <= b0
    assert iv[-1][1] <= total <= iv[-1][1] + align
    for s in range(S):
        exp = O.encode_series(ts[s, : n_points[s]], vals[s, : n_points[s]], start, O.UNIT_S, int_opt) \
            if n_points[s] else b""
        got = packed[g_off[s]: g_off[s] + g_len[s]].tobytes()
        assert got == exp, (s, s % 12, n_points[s])
    # and the decoder reads the packed layout through (offset, length)
    dec = codec.decode(r.packed, r.offsets, P, lengths=r.out_len)
    torch.cuda.synchronize()
    d_st = dec.status.cpu().numpy()
    assert (d_st[n_points > 0] == 0).all()
    assert (d_st[n_points == 0] == 1).all()  # an empty reader: io.EOF on the first read (istream.go:86-92)
    assert (dec.n_points.cpu().numpy() == n_points).all()
    got_ts = dec.ts.cpu().numpy()
    for s in range(0, S, 11):
        assert (got_ts[s, : n_points[s]] == ts[s, : n_points[s]]).all()


@pytest.mark.parametrize("int_opt", [True, False])
def test_encode_packed_point_major_inputs(codecs, int_opt):
    """packed output from point-major ([point][series]) inputs, ragged lengths: every stream equals the
    series-major segment encode (the packed placement is completion order: compare through the index)"""
    from m3_b200 import synth
    codec = codecs[int_opt]
    S, P = 40_003, 97
    ts, vals, start = synth.gaussian_walk(S, P, "cuda", seed=9)
    vals[::7] = torch.round(vals[::7] * 100) / 100  # some int-like series
    n_points = torch.randint(0, P + 1, (S,), dtype=torch.int32, device="cuda")
    n_points[:64] = P
    enc = codec.encode(ts, vals, start, unit=O.UNIT_S, n_points=n_points)
    r = codec.encode_packed(ts.t().contiguous(), vals.t().contiguous(), start, unit=O.UNIT_S, n_points=n_points,
                            point_major=True)
    torch.cuda.synchronize()
    assert torch.equal(r.status, enc.status) and torch.equal(r.out_len, enc.out_len)
    off, ln = r.offsets.cpu().numpy(), r.out_len.cpu().numpy()
    packed, slots = r.packed.cpu().numpy(), enc.out.cpu().numpy()
    for s_ in list(range(0, S, 53)) + [S - 1]:
        assert (packed[off[s_]: off[s_] + ln[s_]] == slots[s_, : ln[s_]]).all(), s_
    # and the device-side checksums of ALL streams agree
    a, st_a = codec.segment_checksums(r.packed, r.offsets, lengths=r.out_len)
    flat_off = torch.arange(S, dtype=torch.int64, device="cuda") * enc.out.shape[1]
    b, st_b = codec.segment_checksums(enc.out.view(-1), flat_off, lengths=enc.out_len)
    torch.cuda.synchronize()
    assert torch.equal(a, b)


def test_encode_packed_many_batches_and_capacity(codecs):
    """More batches than resident warps (the persistent loop reuses its slots) + overflow."""
    from m3_b200 import synth
    codec = codecs[True]
    S, P = 150_000, 48
    ts, vals, start = synth.gaussian_walk(S, P, "cuda", seed=5)
    r = codec.encode_packed(ts, vals, start, unit=O.UNIT_S)
    enc = codec.encode(ts, vals, start, unit=O.UNIT_S)
    torch.cuda.synchronize()
    assert (r.status == 0).all() and torch.equal(r.out_len, enc.out_len)
    total = int(r.total.item())
    assert total >= int(r.out_len.sum().item())
    # every stream equals its slot
    idx = torch.arange(0, S, 997, device="cuda")
    for s in idx.tolist():
        n = int(r.out_len[s].item())
        o = int(r.offsets[s].item())
        assert torch.equal(r.packed[o:o + n], enc.out[s, :n])
    dec = codec.decode(r.packed[:total + 64], r.offsets, P, lengths=r.out_len)
    ref = codec.decode(*codec.compact(enc, align=64), P)
    torch.cuda.synchronize()
    assert torch.equal(dec.ts, ts) and torch.equal(dec.values.view(torch.int64), ref.values.view(torch.int64))
    # capacity: half the space -> some series report M3TSZ_ERR_CAPACITY, the others are intact
    small = codec.encode_packed(ts, vals, start, unit=O.UNIT_S, capacity=(total // 2) & ~63)
    torch.cuda.synchronize()
    st = small.status.cpu().numpy()
    assert (st == 100).any() and (st == 0).any() and set(np.unique(st)) <= {0, 100}
    ok = np.nonzero(st == 0)[0][:50]
    for s in ok.tolist():
        n = int(small.out_len[s].item())
        o = int(small.offsets[s].item())
        assert torch.equal(small.packed[o:o + n], enc.out[s, :n])


def test_unit_change_capacity_bound(codecs):
    """a unit change on EVERY datapoint with incompressible values must fit
    m3tsz_encode_bound_units(n, 1) (83 + 80 bits per datapoint)."""
    from m3_b200 import capi
    rng = np.random.default_rng(2)
    S, P = 64, 200
    ts = START + np.cumsum(rng.integers(1, 1 << 33, size=(S, P)), axis=1).astype(np.int64)
    vals = rng.integers(0, 1 << 63, size=(S, P), dtype=np.int64).view(np.float64).copy()
    vals[~np.isfinite(vals)] = 1.5
    units = np.zeros((S, P), dtype=np.uint8)
    units[:, 0::2] = O.UNIT_NS
    units[:, 1::2] = O.UNIT_US
    codec = codecs[True]
    stride = int(capi.lib().m3tsz_encode_bound_units(P, 1))
    enc = codec.encode(torch.from_numpy(ts).cuda(), torch.from_numpy(vals).cuda(),
                       torch.full((S,), START, dtype=torch.int64, device="cuda"), unit=O.UNIT_S,
                       units=torch.from_numpy(units).cuda(), out_stride=stride)
    torch.cuda.synchronize()
    assert (enc.status.cpu().numpy() == 0).all()
    out, ln = enc.out.cpu().numpy(), enc.out_len.cpu().numpy()
    for s in range(0, S, 9):
        e = O.Encoder(START, True)
        for i in range(P):
            assert e.encode(int(ts[s, i]), float(vals[s, i]), int(units[s, i])) == 0
        assert out[s, : ln[s]].tobytes() == e.stream()


# ------------------------------------------------------------------ row N4: Prometheus epilogue
def _golden():
    import json, os
    return json.load(open(os.path.join(os.path.dirname(__file__), "golden", "m3tsz_goldens.json")))


def _pad(rows, cap, dtype):
    a = np.zeros((len(rows), cap), dtype=dtype)
    for i, r in enumerate(rows):
        a[i, : len(r)] = r
    return a


def test_prom_convert_reference_tables(codecs):
    """prom_converter_test.go tables through the device epilogue."""
    G = _golden()
    codec = codecs[True]
    T = G["prom_counter_normalization"]
    for res in sorted({c["max_resolution_ns"] for c in T["cases"]}):
        cases = [c for c in T["cases"] if c["max_resolution_ns"] == res]
        cap = 8
        ts = _pad([[d[0] for d in c["given"]] for c in cases], cap, np.int64)
        vals = _pad([[float(d[1]) for d in c["given"]] for c in cases], cap, np.float64)
        n = np.array([len(c["given"]) for c in cases], dtype=np.int32)
        hr = np.array([int(c["is_counter"] and res >= T["resolution_threshold_ns"]) for c in cases], dtype=np.uint8)
        to, vo, no, st = codec.prom_convert(torch.from_numpy(ts).cuda(), torch.from_numpy(vals).cuda(),
                                            torch.from_numpy(n).cuda(), res, torch.from_numpy(hr).cuda())
        torch.cuda.synchronize()
        assert (st.cpu().numpy() == 0).all()
        for i, c in enumerate(cases):
            k = int(no[i].item())
            got = [[int(a), float(b)] for a, b in zip(to[i, :k].tolist(), vo[i, :k].tolist())]
            assert got == [[w[0], float(w[1])] for w in c["want"]], c["name"]
    T = G["prom_value_decrease_tolerance"]
    for c in T["cases"]:
        n = len(c["given"])
        ts = np.array([[T["now_ns"] + i * T["step_ns"] for i in range(n)]], dtype=np.int64)
        vals = np.array([c["given"]], dtype=np.float64)
        to, vo, no, st = codec.prom_convert(torch.from_numpy(ts).cuda(), torch.from_numpy(vals).cuda(),
                                            torch.tensor([n], dtype=torch.int32, device="cuda"), 0, None,
                                            c["tolerance"], c["until_ns"])
        torch.cuda.synchronize()
        assert int(no[0].item()) == n and vo[0].tolist() == [float(v) for v in c["want"]], c["name"]
        assert to[0].tolist() == [t // 1_000_000 for t in ts[0].tolist()]


@pytest.mark.parametrize("mode", ["plain", "tolerance", "resets"])
def test_prom_convert_vs_oracle(codecs, mode):
    rng = np.random.default_rng(17)
    S, P = 500, 333
    ts = START + np.cumsum(rng.integers(1, 400, size=(S, P)), axis=1).astype(np.int64) * SEC + \
        rng.integers(0, 10**9, size=(S, P))
    ts.sort(axis=1)
    vals = np.cumsum(np.abs(rng.normal(size=(S, P))), axis=1)
    drop = rng.random(size=(S, P)) < 0.05
    vals[drop] *= rng.choice([0.0, 0.5, 0.99999], size=int(drop.sum()))  # counter resets / tiny dips
    vals[rng.random(size=(S, P)) < 0.01] = np.nan
    n = rng.integers(0, P + 1, size=S).astype(np.int32)
    n[::4] = P
    res = 900 * SEC
    hr = (rng.random(size=S) < 0.5).astype(np.uint8) if mode == "resets" else None
    tol, until = (1e-4, START + 40000 * SEC) if mode != "plain" else (0.0, 0)
    codec = codecs[True]
    to, vo, no, st = codec.prom_convert(torch.from_numpy(ts).cuda(), torch.from_numpy(vals).cuda(),
                                        torch.from_numpy(n).cuda(), res,
                                        None if hr is None else torch.from_numpy(hr).cuda(), tol, until)
    torch.cuda.synchronize()
    to, vo, no = to.cpu().numpy(), vo.cpu().numpy(), no.cpu().numpy()
    assert (st.cpu().numpy() == 0).all()
    for s in range(S):
        eto, evo = O.prom_convert_series(ts[s, : n[s]], vals[s, : n[s]], res, bool(hr[s]) if hr is not None else False,
                                         tol, until)
        assert no[s] == len(eto), s
        assert (to[s, : no[s]] == eto).all(), s
        assert (vo[s, : no[s]].view(np.uint64) == evo.view(np.uint64)).all(), s


# ------------------------------------------------------------------ row N3: tile aggregation
@pytest.mark.parametrize("int_opt", [True, False])
@pytest.mark.parametrize("agg", [1, 2, 3, 4, 6, 7])
def test_aggregate_tiles_byte_identical(codecs, int_opt, agg):
    """decode -> Gauge per Step -> re-encode == oracle decode -> oracle gauge -> oracle encode."""
    rng = np.random.default_rng(40 + agg)
    S, P = 192, 360
    ts, vals = _ds_series(rng, S, P)
    o_out, o_len, o_st = O.encode_batch(ts, vals, START, O.UNIT_S, int_opt, n_threads=8)
    streams = [o_out[s, : o_len[s]].tobytes() for s in range(S)]
    streams[7] = streams[7][: len(streams[7]) // 2]  # a truncated source stream
    d, off = to_device_streams(streams, 64)
    step, n_win = 600 * SEC, 50
    codec = codecs[int_opt]
    r, n_tiles = codec.aggregate_tiles(d, off, START, step, n_win, agg_type=agg)
    torch.cuda.synchronize()
    st = r.status.cpu().numpy()
    g_len, g_off, packed = r.out_len.cpu().numpy(), r.offsets.cpu().numpy(), r.packed.cpu().numpy()
    nt = n_tiles.cpu().numpy()
    for s in range(S):
        ots, ovals, err, _ = oracle_decode(streams[s], int_opt)
        if err != 0:
            assert st[s] == err and g_len[s] == 0 and nt[s] == 0, s
            continue
        assert st[s] == 0, (s, st[s])
        tts, tvals = O.aggregate_tiles_series(ots, ovals, START, step, n_win, agg)
        assert nt[s] == len(tts), s
        exp = O.encode_series(tts, tvals, START, O.UNIT_S, int_opt) if len(tts) else b""
        assert packed[g_off[s]: g_off[s] + g_len[s]].tobytes() == exp, (s, agg)


# ------------------------------------------------------------------ boundary: streaming handles + pools
@pytest.mark.parametrize("int_opt", [True, False])
def test_streaming_handles_vs_oracle(int_opt):
    """m3tsz_encoder_* / m3tsz_iter_* (through the Python mirror of the Go interfaces) against the
    oracle's Encoder / Iterator: bytes, accessors incl. the LastEncoded scaled-int quirk
    (encoder.go:305-319), per-datapoint unit + annotation out of Current()."""
    from m3_b200 import capi
    from m3_b200.encoding import Encoder, ReaderIterator
    rng = np.random.default_rng(77)
    units = [O.UNIT_S, O.UNIT_MS, O.UNIT_US, O.UNIT_NS]
    for trial in range(6):
        e, o = Encoder(START, int_opt), O.Encoder(START, int_opt)
        assert e.empty() and e.stream() is None and e.len() == 0 and e.num_encoded() == 0
        with pytest.raises(capi.M3tszError) as ei:
            e.last_encoded()
        assert ei.value.status == 3  # errNoEncodedDatapoints
        t, v, unit = START, 50.0, O.UNIT_S
        for i in range(150):
            if trial % 2 and rng.random() < 0.05:
                unit = units[int(rng.integers(0, 4))]
            t += int(rng.integers(1, 50)) * SEC
            v += float(rng.normal())
            val = [round(v, 2), float(round(v)), v, round(v, 1)][trial % 4]
            ann = bytes(rng.integers(0, 256, size=int(rng.integers(1, 30)), dtype=np.uint8)) \
                if rng.random() < 0.06 else b""
            e.encode(t, val, unit, ann)
            assert o.encode(t, val, unit, ann) == 0
            if i in (0, 1, 7, 60, 149):  # accessors mid-stream force a launch each time
                assert e.stream() == o.stream() and e.len() == o.len()
                ot, ov, oerr = o.last_encoded()
                assert oerr == 0 and e.last_encoded() == (ot, ov), (trial, i)
                assert e.last_annotation_checksum() == o.last_annotation_checksum()[0]
                assert e.num_encoded() == o.num_encoded() and not e.empty()
        head, tail = e.segment()
        raw, pos = o.raw()
        assert head == raw[:-1] and len(tail) in (2, 3) and head + tail == o.stream()
        data = e.stream()
        it = ReaderIterator(data, int_opt)
        oit = O.Iterator(data, int_opt)
        n = 0
        while oit.next():
            assert it.next()
            assert it.current_full() == oit.current(), (trial, n)
            n += 1
        assert not it.next() and it.err() == 0 and n == 150
        # truncated stream: the error becomes visible with the Next() that fails
        cut = data[: len(data) // 2]
        it.reset(cut)
        oit = O.Iterator(cut, int_opt)
        while oit.next():
            assert it.next() and it.err() == 0
            assert it.current() == oit.current()[:3]
        if not int_opt:  # int mode: DESIGN.md §6.1 (the reference may run on past a failed read)
            assert not it.next() and it.err() == oit.err() != 0
        it.close()
        assert it.err() == 10 and not it.next()
        seg = e.discard()
        assert seg == data
        with pytest.raises(capi.M3tszError) as ei:
            e.encode(t + SEC, 1.0, O.UNIT_S)
        assert ei.value.status == 2  # errEncoderClosed
        e.reset(START + 7200 * SEC)
        e.encode(START + 7201 * SEC, 1.5, O.UNIT_S)
        assert e.num_encoded() == 1


def test_last_encoded_quirk_values():
    from m3_b200.encoding import Encoder
    e = Encoder(START, True)
    e.encode(START + SEC, 12.5, O.UNIT_S)  # int mode with multiplier 1: intVal = 125
    assert e.last_encoded() == (START + SEC, 125.0)
    e.encode(START + 2 * SEC, 0.123456789012, O.UNIT_S)  # float mode: the value itself
    assert e.last_encoded() == (START + 2 * SEC, 0.123456789012)
    f = Encoder(START, False)
    f.encode(START + SEC, 12.5, O.UNIT_S)  # no int optimisation: isFloat never set -> intVal 0
    assert f.last_encoded() == (START + SEC, 0.0)


def test_pools():
    from m3_b200.encoding import EncoderPool, ReaderIteratorPool
    pool = EncoderPool(2)
    a, b, c = pool.get(START), pool.get(START), pool.get(START)  # the third is allocated on demand
    for k, e in enumerate((a, b, c)):
        e.encode(START + SEC, float(k), O.UNIT_S)
    streams = [e.discard() for e in (a, b, c)]  # Discard closes: back to the pool
    d = pool.get(START)
    assert d.empty() and d.num_encoded() == 0
    d.encode(START + SEC, 2.0, O.UNIT_S)
    assert d.stream() == streams[2]
    ip = ReaderIteratorPool(1)
    it = ip.get(streams[1])
    assert it.next() and it.current() == (START + SEC, 1.0, O.UNIT_S) and not it.next() and it.err() == 0
    it.close()
    it2 = ip.get(streams[0])
    assert it2.next() and it2.current()[1] == 0.0


# ------------------------------------------------------------------ row N2: fileset ingestion
def test_fileset_ingest_and_tool(codecs, tmp_path, capsys):
    from m3_b200 import fileset as F
    from m3_b200.tools import read_data_files
    rng = np.random.default_rng(3)
    S, P = 400, 60
    streams, _ = _annotated_streams(rng, S, P, True)
    ids = [b"id-%04d" % int(i) for i in rng.permutation(S)]
    blob = b"".join(streams)
    off = np.concatenate([[0], np.cumsum([len(s) for s in streams])])[:-1]
    F.write_fileset(str(tmp_path), "ns", 3, START, 7200 * SEC, ids, blob, off, [len(s) for s in streams])
    fs = F.read_fileset(str(tmp_path), "ns", 3, START)
    res, ck = F.decode_fileset(codecs[True], fs, P)
    torch.cuda.synchronize()
    assert (ck.cpu().numpy() == 0).all() and (res.status.cpu().numpy() == 0).all()
    by_id = dict(zip(ids, streams))
    gts, gv = res.ts.cpu().numpy(), res.values.cpu().numpy()
    for k in range(0, S, 13):
        ots, ovals, err, _ = oracle_decode(by_id[fs.ids[k]], True)
        assert (gts[k, :P] == ots).all() and (gv[k, :P].view(np.uint64) == ovals.view(np.uint64)).all()
    # a flipped data byte is caught by the DEVICE checksum against the index entry
    bad = F.FilesetData(fs.info, fs.ids, fs.tags, fs.offsets, fs.sizes, fs.data_checksums, np.array(fs.data))
    bad.data[int(fs.offsets[5]) + 3] ^= 1
    _, ck = F.decode_fileset(codecs[True], bad, P)
    torch.cuda.synchronize()
    ckn = ck.cpu().numpy()
    assert ckn[5] == 15 and (np.delete(ckn, 5) == 0).all()  # M3TSZ_ERR_CHECKSUM_MISMATCH
    # the tool: same report as the reference's read_data_files --benchmark datapoints
    rc = read_data_files.main(["-p", str(tmp_path), "-n", "ns", "-s", "3", "-b", str(START), "-B", "datapoints"])
    out = capsys.readouterr().out
    assert rc == 0 and "%d series read" % S in out and "%d datapoints decoded" % (S * P) in out
    rc = read_data_files.main(["-p", str(tmp_path / "gen"), "-n", "g", "-s", "0", "--generate", "3000", "--points",
                               "50", "-B", "datapoints"])
    out = capsys.readouterr().out
    assert rc == 0 and "150000 datapoints decoded" in out
    rc = read_data_files.main(["-p", str(tmp_path), "-n", "ns", "-s", "3", "-b", str(START), "-f", "id-0007"])
    out = capsys.readouterr().out
    assert rc == 0 and out.count("{id: id-0007") == P


# ------------------------------------------------------------------ point-major decode output
@pytest.mark.parametrize("int_opt", [True, False])
def test_point_major_decode_equals_series_major(codecs, int_opt):
    """extras.point_major: same datapoints, [point][series] layout (coalesced stores)."""
    rng = np.random.default_rng(12)
    S, P = 333, 150
    ts, vals, start = _mixed(rng, S, P)
    n_points = rng.integers(0, P + 1, size=S).astype(np.int32)
    n_points[::2] = P
    streams = []
    for s in range(S):
        streams.append(O.encode_series(ts[s, : n_points[s]], vals[s, : n_points[s]], start, O.UNIT_S, int_opt)
                       if n_points[s] else b"")
    ann_streams, _ = _annotated_streams(rng, 40, 60, int_opt)  # markers, unit changes: the slow path
    streams += ann_streams
    d, off = to_device_streams(streams, 1)
    codec = codecs[int_opt]
    for cap in (P, 64):
        a = codec.decode(d, off, cap)
        b = codec.decode(d, off, cap, point_major=True)
        b.ts.fill_(-7)
        b.values.fill_(-7.0)
        b = codec.decode(d, off, cap, point_major=True, out=b)
        torch.cuda.synchronize()
        assert torch.equal(a.n_points, b.n_points) and torch.equal(a.status, b.status) and torch.equal(a.unit, b.unit)
        n = a.n_points.cpu().numpy().view(np.uint32).astype(np.int64)
        at, av = a.ts.cpu().numpy(), a.values.cpu().numpy().view(np.uint64)
        bt, bv = b.ts.cpu().numpy().T, b.values.cpu().numpy().view(np.uint64).T
        for s in range(len(streams)):
            k = min(int(n[s]), cap)
            assert (at[s, :k] == bt[s, :k]).all() and (av[s, :k] == bv[s, :k]).all(), (s, cap)
            assert (bt[s, k:] == -7).all(), (s, cap)  # rows past the series' end are untouched


@pytest.mark.parametrize("int_opt", [True, False])
def test_point_major_encode_byte_identical(codecs, int_opt):
    """extras.point_major_input: [point][series] inputs, same bytes as the series-major path and the oracle."""
    rng = np.random.default_rng(19)
    S, P = 517, 130
    ts, vals, start = _mixed(rng, S, P)
    n_points = rng.integers(0, P + 1, size=S).astype(np.int32)
    n_points[::3] = P
    codec = codecs[int_opt]
    d_ts, d_vals = torch.from_numpy(ts).cuda(), torch.from_numpy(vals).cuda()
    d_start = torch.full((S,), start, dtype=torch.int64, device="cuda")
    d_n = torch.from_numpy(n_points).cuda()
    a = codec.encode(d_ts, d_vals, d_start, unit=O.UNIT_S, n_points=d_n)
    b = codec.encode(d_ts.t().contiguous(), d_vals.t().contiguous(), d_start, unit=O.UNIT_S, n_points=d_n,
                     point_major=True)
    torch.cuda.synchronize()
    assert torch.equal(a.status, b.status) and int((a.status != 0).sum()) == 0
    assert torch.equal(a.out_len, b.out_len)
    la, oa, ob = a.out_len.cpu().numpy(), a.out.cpu().numpy(), b.out.cpu().numpy()
    for s in range(S):
        assert (oa[s, : la[s]] == ob[s, : la[s]]).all(), s
    for s in range(0, S, 17):
        exp = O.encode_series(ts[s, : n_points[s]], vals[s, : n_points[s]], start, O.UNIT_S, int_opt) \
            if n_points[s] else b""
        assert ob[s, : la[s]].tobytes() == exp, s
    # decode(point-major) -> encode(point-major): the device-resident round trip in one layout
    stride = a.out.shape[1]
    off = torch.arange(S, dtype=torch.int64, device="cuda") * stride
    dec = codec.decode(a.out.view(-1), off, P, lengths=a.out_len, point_major=True)
    c = codec.encode(dec.ts, dec.values, d_start, unit=O.UNIT_S, n_points=dec.n_points, point_major=True)
    torch.cuda.synchronize()
    assert int((c.status != 0).sum()) == 0
    oc, lc = c.out.cpu().numpy(), c.out_len.cpu().numpy()
    # re-encoding decoded values is byte-identical where decode is lossless (always in float mode; in
    # int mode whenever the values were exactly representable -- compare against the oracle re-encode)
    for s in range(0, S, 13):
        k = int(n_points[s])
        if not k:
            assert lc[s] == 0
            continue
        dps, err = O.decode_all(oa[s, : la[s]].tobytes(), int_opt)
        exp = O.encode_series([d[0] for d in dps], [d[1] for d in dps], start, O.UNIT_S, int_opt)
        assert oc[s, : lc[s]].tobytes() == exp, s


@pytest.mark.parametrize("int_opt", [True, False])
@pytest.mark.parametrize("S,P", [(64, 1), (64, 130), (518, 7), (4098, 130), (33_000, 61)])
def test_point_major_encode_even_series_counts(codecs, int_opt, S, P):
    """Point-major inputs with an EVEN series count (16-byte aligned row pitch: the shape the tensor-copy input
    stage of a -DM3_ENC_BULK_PM=2 build describes with a tensor map; odd counts take the cp.async fills): full and
    ragged warps, point counts that end inside a tile, ragged lengths, segment and packed output, and array bases
    that are only 8-byte aligned.  Every stream equals the series-major encode (itself oracle-pinned)."""
    rng = np.random.default_rng(1000 + S + P)
    ts, vals, start = _mixed(rng, S, P)
    codec = codecs[int_opt]
    d_ts, d_vals = torch.from_numpy(ts).cuda(), torch.from_numpy(vals).cuda()
    d_start = torch.full((S,), start, dtype=torch.int64, device="cuda")
    pm_ts, pm_vals = d_ts.t().contiguous(), d_vals.t().contiguous()
    for ragged in (False, True):
        d_n = None
        if ragged:
            n_points = rng.integers(0, P + 1, size=S).astype(np.int32)
            n_points[: min(S, 96)] = P  # whole warps of full length next to ragged ones
            d_n = torch.from_numpy(n_points).cuda()
        a = codec.encode(d_ts, d_vals, d_start, unit=O.UNIT_S, n_points=d_n)
        b = codec.encode(pm_ts, pm_vals, d_start, unit=O.UNIT_S, n_points=d_n, point_major=True)
        r = codec.encode_packed(pm_ts, pm_vals, d_start, unit=O.UNIT_S, n_points=d_n, point_major=True)
        torch.cuda.synchronize()
        assert torch.equal(a.status, b.status) and torch.equal(a.out_len, b.out_len)
        assert torch.equal(a.status, r.status) and torch.equal(a.out_len, r.out_len)
        la, oa, ob = a.out_len.cpu().numpy(), a.out.cpu().numpy(), b.out.cpu().numpy()
        off, packed = r.offsets.cpu().numpy(), r.packed.cpu().numpy()
        step = max(1, S // 600)
        for s in list(range(0, S, step)) + [S - 1]:
            assert (oa[s, : la[s]] == ob[s, : la[s]]).all(), (s, ragged)
            assert (packed[off[s]: off[s] + la[s]] == oa[s, : la[s]]).all(), (s, ragged)
        # all streams through the device-side checksums
        flat = torch.arange(S, dtype=torch.int64, device="cuda") * a.out.shape[1]
        ca, _ = codec.segment_checksums(a.out.view(-1), flat, lengths=a.out_len)
        cb, _ = codec.segment_checksums(b.out.view(-1), flat, lengths=b.out_len)
        cr, _ = codec.segment_checksums(r.packed, r.offsets, lengths=r.out_len)
        torch.cuda.synchronize()
        assert torch.equal(ca, cb) and torch.equal(ca, cr), ragged
    # bases that are 8- but not 16-byte aligned (views one element into a larger allocation)
    big_t = torch.empty(P * S + 1, dtype=torch.int64, device="cuda")
    big_v = torch.empty(P * S + 1, dtype=torch.float64, device="cuda")
    u_ts, u_vals = big_t[1:].view(P, S), big_v[1:].view(P, S)
    u_ts.copy_(pm_ts)
    u_vals.copy_(pm_vals)
    assert u_ts.data_ptr() % 16 == 8
    a = codec.encode(d_ts, d_vals, d_start, unit=O.UNIT_S)
    c = codec.encode(u_ts, u_vals, d_start, unit=O.UNIT_S, point_major=True)
    torch.cuda.synchronize()
    assert torch.equal(a.out_len, c.out_len) and torch.equal(a.status, c.status)
    flat = torch.arange(S, dtype=torch.int64, device="cuda") * a.out.shape[1]
    ca, _ = codec.segment_checksums(a.out.view(-1), flat, lengths=a.out_len)
    cc, _ = codec.segment_checksums(c.out.view(-1), flat, lengths=c.out_len)
    torch.cuda.synchronize()
    assert torch.equal(ca, cc)
