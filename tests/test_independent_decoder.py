"""Cross-check of the C oracle against tests/indep_m3tsz.py, a second restatement of the reference's iterator
written separately in pure Python: the goldens and fixtures decode identically, and -- the point of the exercise --
so do ORACLE-ENCODED streams of the reference's round-trip families in BOTH modes (int-optimised encodes have no
golden bytes in the reference, SURVEY.md §8c "Gap"), with unit changes and annotations as in roundtrip_test.go."""
import base64
import json
import os
import random
import struct

import numpy as np
import pytest

import indep_m3tsz as I
import oracle_lib as O

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "m3tsz_goldens.json")))
SEC = 10 ** 9


def _bits(v):
    return struct.unpack("<Q", struct.pack("<d", v))[0]


def _same(stream, int_opt):
    got = I.decode(stream, int_opt)
    exp, err = O.decode_all(stream, int_opt)
    assert err == 0
    assert len(got) == len(exp)
    for i, ((t, vb, u, a), e) in enumerate(zip(got, exp)):
        assert t == e[0], i
        assert vb == _bits(e[1]), (i, vb, e[1])
        if len(e) > 2:
            assert u == e[2] and a == e[3], i
    return got


def test_goldens_and_fixtures():
    for s in G["streams"]:
        got = _same(bytes.fromhex(s["bytes"]) if "bytes" in s else bytes(s["raw"]), False)
        assert len(got) > 0
    n = 0
    for b64 in G["fixtures_b64"]["streams"]:
        n += len(_same(base64.b64decode(b64), True))
    assert n == sum(G["fixtures_b64"]["expected_points"])
    assert len(_same(base64.b64decode(G["regression_b64"]["stream"]), True)) == 150


def _gen(r, num_dig, num_dec):
    dig = r.getrandbits(62) % 10 ** num_dig
    return float(dig) if num_dec == 0 else float("%d.%d" % (dig, r.getrandbits(62) % 10 ** num_dec))


def _encode(dps, int_opt, with_markers):
    e = O.Encoder(1427162400 * SEC, int_opt)
    for i, (t, v) in enumerate(dps):
        unit, ann = O.UNIT_S, b""
        if with_markers:
            unit = O.UNIT_MS if i == 0 else (O.UNIT_US if i == 10 else O.UNIT_S)
            ann = b"foo" if i < 5 else (b"bar" if i < 7 else (b"long annotation " * 4 if i == 10 else b""))
        assert e.encode(t, v, unit, ann) == 0
    return e.stream()


@pytest.mark.parametrize("int_opt", [True, False])
@pytest.mark.parametrize("with_markers", [False, True])
def test_oracle_encoded_round_trip_families(int_opt, with_markers):
    r = random.Random(99 + int(int_opt) * 2 + int(with_markers))
    fams = [(12, 0), (7, 6), (0, 1), (2, 16), (5, 3), (3, 0), (18, 0), (1, 2), (0, 6)]
    for num_dig, num_dec in fams:
        for rep in range(3):
            t = 1427162462 * SEC
            dps = [(t, 1.0)]
            for _ in range(400):
                t += SEC * r.randrange(1200)
                v = _gen(r, num_dig, num_dec)
                if rep == 1 and r.random() < 0.5:
                    v = -v
                if rep == 2 and r.random() < 0.3:
                    v = dps[-1][1]  # repeats
                dps.append((t, v))
            got = _same(_encode(dps, int_opt, with_markers), int_opt)
            assert [g[0] for g in got] == [d[0] for d in dps]
            dec = [struct.unpack("<d", struct.pack("<Q", g[1]))[0] for g in got]
            if int_opt and num_dec == 0 and num_dig > 15:
                # integers beyond 2^53 in int mode: the reference subtracts and re-adds in float64
                # (encoder.go:148-176 `valDiff := enc.intVal - val`, iterator.go:163-176 `intVal += sign * float64(bits)`),
                # so the decoded series is the float64 recurrence below, not the input -- pinned here as it is
                assert dec[0] == dps[0][1]
                for i in range(1, len(dps)):
                    assert dec[i] == dec[i - 1] - (dps[i - 1][1] - dps[i][1]), i
                assert any(d != v for d, (_, v) in zip(dec, dps))
            else:
                assert dec == [v for _, v in dps]  # the families round-trip exactly


@pytest.mark.parametrize("int_opt", [True, False])
def test_oracle_encoded_gaussian_walk_and_mode_switches(int_opt):
    """the bench's data family (where accidental int-mode transitions live), int <-> float switches, special
    values, significant-bit updates and multiplier growth"""
    rng = np.random.default_rng(3)
    start = 1599955200 * SEC
    for s in range(24):
        P = 300
        ts = start + np.arange(1, P + 1, dtype=np.int64) * 60 * SEC
        if s % 4 == 0:
            vals = 100.0 + np.cumsum(rng.normal(size=P))
        elif s % 4 == 1:  # integers whose magnitude wanders over many significant-bit counts
            vals = np.round(np.cumsum(rng.normal(size=P) * 10.0 ** rng.integers(0, 9))).astype(np.float64)
        elif s % 4 == 2:  # decimals with growing precision (multiplier updates), then floats, then ints again
            vals = np.concatenate([np.round(rng.normal(size=100) * 50, 1), np.round(rng.normal(size=50) * 50, 4),
                                   rng.normal(size=50), np.round(rng.normal(size=100) * 1000)])
        else:
            vals = np.round(rng.normal(size=P) * 5, 2)
            vals[::37] = [np.nan, np.inf, -np.inf, -0.0, 2.0 ** 63, -2.0 ** 63, 1e300, 5e-324, 0.0][: len(vals[::37])]
        out, ln, st = O.encode_batch(ts[None, :], vals[None, :], start, O.UNIT_S, int_opt)
        assert st[0] == 0
        got = _same(out[0, : ln[0]].tobytes(), int_opt)
        assert [g[0] for g in got] == ts.tolist()


# ------------------------------------------------------------------------------------------------------
# the independent ENCODER: same bytes as the oracle's encoder
# ------------------------------------------------------------------------------------------------------
def _oracle_encode(start, dps, int_opt):
    e = O.Encoder(start, int_opt)
    for t, v, u, a in dps:
        assert e.encode(t, v, u, a) == 0
    return e.stream()


def test_independent_encoder_reproduces_the_golden_streams():
    assert len(G["streams"]) == 4
    for s in G["streams"]:  # encoder_test.go:207-393: the reference's own bytes, not the oracle's
        dps = [(p["ts"], float(p["value"]), p["unit"], bytes.fromhex(p["annotation"])) for p in s["datapoints"]]
        assert I.encode(s["encoder_start"], dps, s["int_optimized"]) == bytes.fromhex(s["bytes"]), s["name"]
        got = I.decode(bytes.fromhex(s["bytes"]), s["int_optimized"])
        assert [(g[0], g[1]) for g in got] == [(p["ts"], _bits(float(p["value"]))) for p in s["datapoints"]]
        assert [g[3].hex() for g in got] == s["decoded_annotations"]


@pytest.mark.parametrize("int_opt", [True, False])
@pytest.mark.parametrize("with_markers", [False, True])
def test_independent_encoder_matches_oracle_bytes_on_families(int_opt, with_markers):
    r = random.Random(7 + int(int_opt) * 2 + int(with_markers))
    start = 1427162400 * SEC
    fams = [(12, 0), (7, 6), (0, 1), (2, 16), (5, 3), (3, 0), (18, 0), (1, 2), (0, 6), (4, 4), (9, 2), (0, 0)]
    for num_dig, num_dec in fams:
        for rep in range(4):
            t = 1427162462 * SEC
            dps = []
            for i in range(300):
                v = 1.0 if i == 0 else _gen(r, num_dig, num_dec)
                if rep == 1 and r.random() < 0.5:
                    v = -v
                if rep == 2 and dps and r.random() < 0.3:
                    v = dps[-1][1]
                if rep == 3 and r.random() < 0.15:  # another family now and then: mode / multiplier / sig changes
                    v = _gen(r, *fams[r.randrange(len(fams))])
                unit, ann = O.UNIT_S, b""
                if with_markers:
                    unit = O.UNIT_MS if i == 0 else (O.UNIT_US if i == 10 else (O.UNIT_NS if 200 <= i < 210 else O.UNIT_S))
                    ann = b"foo" if i < 5 else (b"bar" if i < 7 else (b"long annotation " * 9 if i == 10 else b""))
                dps.append((t, v, unit, ann))
                t += SEC * r.randrange(1200) + (r.randrange(1000) * 1000 if with_markers and i >= 10 else 0)
            assert I.encode(start, dps, int_opt) == _oracle_encode(start, dps, int_opt), (num_dig, num_dec, rep)


@pytest.mark.parametrize("int_opt", [True, False])
def test_independent_encoder_matches_oracle_bytes_on_walks_and_specials(int_opt):
    rng = np.random.default_rng(11)
    start = 1599955200 * SEC
    for s in range(40):
        P = 400
        ts = start + np.arange(1, P + 1, dtype=np.int64) * 60 * SEC
        if s % 5 == 0:
            vals = 100.0 + np.cumsum(rng.normal(size=P))
        elif s % 5 == 1:
            vals = np.round(np.cumsum(rng.normal(size=P) * 10.0 ** rng.integers(0, 12))).astype(np.float64)
        elif s % 5 == 2:
            vals = np.concatenate([np.round(rng.normal(size=100) * 50, 1), np.round(rng.normal(size=100) * 50, 4),
                                   rng.normal(size=100), np.round(rng.normal(size=100) * 1000)])
        elif s % 5 == 3:
            vals = np.round(rng.normal(size=P) * 5, int(rng.integers(0, 7)))
            sp = [np.nan, np.inf, -np.inf, -0.0, 2.0 ** 63, -2.0 ** 63, 1e300, -1e300, 5e-324, 0.0, 9.2e18, 1e13, 1e13 - 1]
            vals[::29] = (sp * 2)[: len(vals[::29])]
        else:  # jittered timestamps: every delta-of-delta bucket
            vals = np.round(rng.normal(size=P) * 100, 2)
            ts = start + np.cumsum(rng.choice([1, 10, 60, 300, 3000, 100000], size=P)).astype(np.int64) * SEC
        dps = [(int(t), float(v), O.UNIT_S, b"") for t, v in zip(ts, vals)]
        mine = I.encode(start, dps, int_opt)
        out, ln, st = O.encode_batch(ts[None, :], vals[None, :], start, O.UNIT_S, int_opt)
        assert st[0] == 0
        assert mine == out[0, : ln[0]].tobytes(), s


def test_independent_pair_reencodes_the_production_fixtures_byte_identically():
    """encoder_benchmark_test.go:36-47: ten int-optimised production streams (ms unit, markers, int mode, repeats).
    Independent decoder -> independent encoder (the stream's own start, per-datapoint unit and annotations replayed)
    gives back the reference's bytes: the int-optimised encoder pinned against reference-produced bytes without the
    C oracle in the loop."""
    n = 0
    for b64 in G["fixtures_b64"]["streams"]:
        data = base64.b64decode(b64)
        got = I.decode(data, True)
        start = struct.unpack(">q", data[:8])[0]
        dps = [(t, struct.unpack("<d", struct.pack("<Q", vb))[0], u, a) for t, vb, u, a in got]
        assert I.encode(start, dps, True) == data
        n += len(dps)
    assert n == 7197


def test_independent_convert_to_int_float_matches_oracle():
    """m3tsz.go:78-119 restated twice (C oracle, pure Python): same (value, multiplier, isFloat) for every current
    multiplier on the reference's families, on +-ulp neighbours of k / 10^m (where Modf / Nextafter decide) and on
    random bit patterns."""
    r = random.Random(31)
    vals = []
    for num_dig, num_dec in [(0, 0), (2, 0), (10, 0), (18, 0), (0, 6), (3, 6), (7, 6), (0, 1), (1, 3), (5, 3), (2, 16),
                             (9, 2), (11, 3), (13, 0), (12, 1)]:
        for _ in range(300):
            v = _gen(r, num_dig, num_dec)
            vals += [v, -v]
    base = np.array([(r.getrandbits(r.randrange(1, 50)) + 1) / 10.0 ** r.randrange(0, 7) for _ in range(3000)])
    bits = base.view(np.int64)
    for j in (-3, -2, -1, 1, 2, 3, 9, 40):
        vals += (bits + j).view(np.float64).tolist()
    vals += base.tolist() + (-base).tolist()
    rb = np.random.default_rng(5).integers(0, 2 ** 63, size=4000, dtype=np.int64).view(np.float64)
    vals += rb.tolist() + (-rb).tolist()
    vals += [0.0, -0.0, float("inf"), float("-inf"), 2.0 ** 63, -2.0 ** 63, 1e13, 1e13 - 0.5, 9999999999999.9, 1e300, -1e300]
    n = 0
    for v in vals:
        for cur in range(7):
            ov, om, of, err = O.convert_to_int_float(v, cur)
            assert err == 0
            iv, im, isf = I.convert_to_int_float(v, cur)
            assert (im, isf) == (om, of), (v, cur)
            assert iv == ov or (iv != iv and ov != ov), (v, cur, iv, ov)
            n += 1
    assert n > 250000
