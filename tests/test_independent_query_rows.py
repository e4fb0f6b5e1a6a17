"""Independent pure-Python restatements of the two small "next" rows that sit after decode -- the Gauge window
aggregation of BASELINE config 4 (src/aggregator/aggregation/gauge.go:45-106, window = Truncate(resolution),
aggregator/generic_elem.go:219-236) and the Prometheus conversion epilogue (src/query/storage/
prom_converter.go:42-120, converter.go:388-391) -- cross-checked against the C oracle on random series (the
oracle itself is pinned by the reference's gauge_test.go / prom_converter_test.go tables; these tests cover what
the tables do not enumerate: NaN / Inf inside windows, out-of-order timestamps, empty windows, points outside the
range, negative values, resets in the middle of a resolution window)."""
import math
import struct

import numpy as np
import pytest

import oracle_lib as O

SEC = 10 ** 9


def gauge_windows(ts, vals, start, window, n_windows):
    """one aggregation.Gauge per window (gauge.go:45-106)"""
    nan = struct.unpack("<d", struct.pack("<Q", 0x7FF8000000000001))[0]  # Go's math.NaN() (math/bits.go uvnan):
    # a window that only saw NaNs hands THIS payload to the tile encoder, whose XOR path keeps payload bits
    w = [dict(sum=0.0, count=0, mn=nan, mx=nan, last=0.0, last_at=None) for _ in range(n_windows)]
    for t, v in zip(ts, vals):
        i = (t - start) // window  # floor: timestamps before the range give negative windows
        if i < 0 or i >= n_windows:
            continue
        g = w[i]
        if g["last_at"] is None or t > g["last_at"]:  # updateTotals :73-81
            g["last_at"], g["last"] = t, v
        g["count"] += 1
        if v != v:
            continue
        g["sum"] += v
        if g["mx"] != g["mx"] or g["mx"] < v:
            g["mx"] = v
        if g["mn"] != g["mn"] or g["mn"] > v:
            g["mn"] = v
    return w


def _eq(a, b):
    return a == b or (a != a and b != b)


@pytest.mark.parametrize("seed", range(6))
def test_gauge_windows_match_oracle(seed):
    rng = np.random.default_rng(100 + seed)
    start, window, n_windows = 1599955200 * SEC, 300 * SEC, 40
    for _ in range(30):
        n = int(rng.integers(0, 400))
        ts = start + rng.integers(-2 * window, (n_windows + 2) * window, size=n)  # unordered, also outside the range
        if seed % 2 == 0:
            ts = np.sort(ts)
        vals = np.round(rng.normal(size=n) * 100, int(rng.integers(0, 4)))
        for special in (np.nan, np.inf, -np.inf, -0.0, 0.0):
            if n:
                vals[rng.integers(0, n, size=max(1, n // 25))] = special
        s, c, mn, mx, last = O.downsample_series(ts, vals, start, window, n_windows)
        ref = gauge_windows(ts.tolist(), vals.tolist(), start, window, n_windows)
        for i, g in enumerate(ref):
            assert c[i] == g["count"], i
            assert _eq(float(s[i]), g["sum"]), (i, s[i], g["sum"])
            assert _eq(float(mn[i]), g["mn"]) and _eq(float(mx[i]), g["mx"]), i
            assert _eq(float(last[i]), g["last"]), i
            # signed zeros are kept apart by the reference's comparisons: compare bit patterns where both are 0
            if float(mn[i]) == 0.0 and g["mn"] == 0.0:
                assert math.copysign(1, float(mn[i])) == math.copysign(1, g["mn"]), i
            if float(mx[i]) == 0.0 and g["mx"] == 0.0:
                assert math.copysign(1, float(mx[i])) == math.copysign(1, g["mx"]), i


def prom_convert(ts, vals, resolution, handle_resets, tolerance=0.0, until=0):
    """iteratorToPromResult (prom_converter.go:42-120) for one series; handle_resets = the decision the reference
    takes from the first annotation and the resolution threshold, passed in"""
    out = []
    first, cum = True, 0.0
    prev_t, prev_v = 0, 0.0
    for t, v in zip(ts, vals):
        if tolerance > 0 and t < until:
            if not first and v < prev_v and v > prev_v * (1 - tolerance):
                v = prev_v
        if handle_resets:
            if _go_div(t, resolution) != _go_div(prev_t, resolution) and not first:
                out.append((_go_div(prev_t, 10 ** 6), cum))
            if v < prev_v:
                cum += v
            else:
                cum += v - prev_v
        else:
            out.append((_go_div(t, 10 ** 6), v))
        prev_t, prev_v = t, v
        first = False
    if handle_resets and not first:  # handleResets is only ever set while looking at the first datapoint (:76-84)
        out.append((_go_div(prev_t, 10 ** 6), cum))
    return out


def _go_div(a, b):  # Go's integer division truncates toward zero
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


@pytest.mark.parametrize("seed", range(6))
def test_prom_convert_matches_oracle(seed):
    rng = np.random.default_rng(200 + seed)
    for _ in range(60):
        n = int(rng.integers(1, 300))
        t0 = 1599955200 * SEC + int(rng.integers(0, 10 ** 9))
        ts = t0 + np.cumsum(rng.integers(1, 120, size=n)) * SEC // 2
        kind = int(rng.integers(0, 3))
        if kind == 0:  # a counter with resets
            inc = np.abs(np.round(rng.normal(size=n) * 10, 1))
            vals = np.cumsum(inc)
            for r in rng.integers(1, n, size=max(1, n // 40)) if n > 1 else []:
                vals[r:] -= vals[r] - inc[r]
        elif kind == 1:  # small decreases (the tolerance case) and real ones
            vals = 1000.0 + np.cumsum(np.abs(rng.normal(size=n)))
            vals[rng.integers(0, n, size=max(1, n // 10))] *= (1 - 10.0 ** -rng.integers(2, 7))
        else:
            vals = rng.normal(size=n) * 100
            vals[rng.integers(0, n, size=max(1, n // 20))] = np.nan
        resolution = int(rng.choice([60, 300, 600, 3600])) * SEC
        handle = bool(rng.integers(0, 2))
        tol = float(rng.choice([0.0, 1e-3, 1e-6]))
        until = int(ts[min(n - 1, int(rng.integers(0, n)))]) if tol else 0
        o_ts, o_v = O.prom_convert_series(ts, vals, resolution, handle, tol, until)
        ref = prom_convert(ts.tolist(), vals.tolist(), resolution, handle, tol, until)
        assert len(ref) == len(o_ts), (kind, handle, tol)
        for (rt, rv), ot, ov in zip(ref, o_ts.tolist(), o_v.tolist()):
            assert rt == ot
            assert np.float64(rv).view(np.uint64) == np.float64(ov).view(np.uint64) or (rv != rv and ov != ov)


@pytest.mark.parametrize("agg", [O.AGG_LAST, O.AGG_MIN, O.AGG_MAX, O.AGG_MEAN, O.AGG_COUNT, O.AGG_SUM])
def test_tile_aggregation_matches_oracle_and_independent_encoder(agg):
    """row N3 (storage.TileAggregator compute): one datapoint per non-empty Step window at the window's END
    (aggregator/list.go:541-543), value = Gauge.ValueOf(type) (gauge.go:144-165); and the re-encoded stream: the
    oracle's encoder against the independent encoder of tests/indep_m3tsz.py on those tiles."""
    import indep_m3tsz as I
    rng = np.random.default_rng(300 + agg)
    start, step, n_windows = 1599955200 * SEC, 300 * SEC, 48
    for _ in range(25):
        n = int(rng.integers(0, 500))
        ts = np.sort(start + rng.integers(0, n_windows * step, size=n))
        vals = np.round(rng.normal(size=n) * 50, int(rng.integers(0, 3)))
        if n:
            vals[rng.integers(0, n, size=max(1, n // 30))] = np.nan
        t_o, v_o = O.aggregate_tiles_series(ts, vals, start, step, n_windows, agg)
        exp = []
        for i, g in enumerate(gauge_windows(ts.tolist(), vals.tolist(), start, step, n_windows)):
            if g["count"] == 0:
                continue
            value = {O.AGG_LAST: g["last"], O.AGG_MIN: g["mn"], O.AGG_MAX: g["mx"],
                     O.AGG_MEAN: (g["sum"] / float(g["count"])) if g["count"] else 0.0,
                     O.AGG_COUNT: float(g["count"]), O.AGG_SUM: g["sum"]}[agg]
            exp.append((start + (i + 1) * step, value))
        assert t_o.tolist() == [e[0] for e in exp]
        for a, (_, b) in zip(v_o.tolist(), exp):
            assert _eq(a, b)
        if exp:
            mine = I.encode(start, [(t, v, O.UNIT_S, b"") for t, v in exp], True)
            assert mine == O.encode_series(t_o, v_o, start, O.UNIT_S, True)
