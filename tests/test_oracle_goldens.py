"""Pins the CPU oracle against every golden vector the reference's own tests hold
for the M3TSZ path (tests/golden/m3tsz_goldens.json; provenance in each entry's
"src" and in tests/golden/README.md).  CPU only."""
import base64
import json
import math
import os
import random
import struct

import numpy as np
import pytest

import oracle_lib as O

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "m3tsz_goldens.json")))
SEC = 1_000_000_000


def hb(s):
    return bytes.fromhex(s)


# ---------------------------------------------------------------- bit I/O
def test_ostream_write_bits():
    os_ = O.OStream()
    for st in G["ostream_write_bits"]["steps"]:
        os_.write_bits(st["value"], st["nbits"])
        raw, pos = os_.raw()
        assert raw == hb(st["bytes"])
        assert pos == st["pos"]


def test_istream_read_bits():
    g = G["istream_read_bits"]
    is_ = O.IStream(hb(g["bytes"]))
    got = []
    for n in g["nbits"]:
        v, err = is_.read_bits(n)
        assert err == 0
        got.append(v)
    assert got == g["expected"]
    _, err = is_.read_bits(8)
    assert err == O.ERR_EOF


def test_istream_read_byte():
    data = hb(G["istream_read_bits"]["bytes"])
    is_ = O.IStream(data)
    assert bytes(is_.read_bits(8)[0] for _ in data) == data
    assert is_.read_bits(8)[1] == O.ERR_EOF


def test_istream_peek():
    g = G["istream_peek_bits"]
    is_ = O.IStream(hb(g["bytes"]))
    for n, exp in g["cases"]:
        v, err = is_.peek_bits(n)
        assert err == 0 and v == exp
    g = G["istream_peek_error"]
    is_ = O.IStream(hb(g["bytes"]))
    v, err = is_.peek_bits(g["peek"])
    assert err == O.ERR_EOF and v == 0


def test_istream_read_after_peek():
    g = G["istream_read_after_peek"]
    is_ = O.IStream(hb(g["bytes"]))
    assert is_.peek_bits(10) == (g["peek10"], 0)
    assert is_.peek_bits(20)[1] == O.ERR_EOF
    for n, exp in g["reads"]:
        assert is_.read_bits(n) == (exp, 0)
    assert is_.read_bits(8)[1] == O.ERR_EOF


def test_istream_peek_after_read():
    g = G["istream_peek_after_read"]
    is_ = O.IStream(hb(g["bytes"]))
    for op, n, exp in g["ops"]:
        if op == "read":
            assert is_.read_bits(n) == (exp, 0)
        elif op == "peek":
            assert is_.peek_bits(n) == (exp, 0)
        else:
            assert is_.peek_bits(n)[1] == O.ERR_EOF


def test_istream_remaining_bits_in_current_byte():
    # istream_test.go:164-178
    data = bytes([0xFF, 0, 0x42])
    is_ = O.IStream(data)
    for b in data:
        for i in range(8):
            assert is_.remaining_bits_in_current_byte() == (8 - i if i > 0 else 0)
            bit, err = is_.read_bits(1)
            assert err == 0 and bit == (b >> i) & 1


# ---------------------------------------------------------------- field level (encode)
def test_write_dod_unit_unchanged():
    for c in G["write_dod_unit_unchanged"]["cases"]:
        os_ = O.OStream()
        err = O.lib().m3o_write_dod_unit_unchanged(os_.h, 0, c["delta_ns"], c["unit"])
        assert err == 0
        assert os_.raw() == (hb(c["bytes"]), c["pos"])


def test_write_dod_unit_changed():
    for c in G["write_dod_unit_changed"]["cases"]:
        os_ = O.OStream()
        O.lib().m3o_write_dod_unit_changed(os_.h, 0, c["delta_ns"])
        assert os_.raw() == (hb(c["bytes"]), c["pos"])


def test_write_xor():
    for c in G["write_xor"]["cases"]:
        os_ = O.OStream()
        O.lib().m3o_write_xor(os_.h, c["prev_xor"], c["cur_xor"])
        assert os_.raw() == (hb(c["bytes"]), c["pos"])


def test_write_annotation_and_time_unit_via_encoder():
    # encoder_test.go:125-155: NewTimestampEncoder(0, ns).writeAnnotation -> bytes after the
    # 64-bit start (start=0 -> 8 zero bytes) and before the dod.  We drive the full
    # encoder (start=0, default unit ns) and compare the annotation field bytes.
    for c in G["write_annotation"]["cases"]:
        if not c["annotation"]:
            continue
        e = O.Encoder(0, False, default_unit=O.UNIT_NS)
        assert e.encode(0, 0.0, O.UNIT_NS, hb(c["annotation"])) == 0
        raw, _ = e.raw()
        field = hb(c["bytes"])
        nbits = (len(field) - 1) * 8 + c["pos"]
        got = int.from_bytes(raw[8:8 + len(field)], "big") >> (len(field) * 8 - nbits)
        exp = int.from_bytes(field, "big") >> (len(field) * 8 - nbits)
        assert got == exp
    # encoder_test.go:172-205: time unit marker = 0x80 0x40 0x20 pos 3 for unit Second
    e = O.Encoder(1, False, default_unit=O.UNIT_S)  # start=1ns -> initial unit None
    assert e.encode(1, 0.0, O.UNIT_S) == 0
    raw, _ = e.raw()
    assert raw[8:10] == bytes([0x80, 0x40]) and raw[10] >> 5 == 0x20 >> 5
    # unit None / invalid unit: no marker is written; Encode errors on Value()
    e = O.Encoder(1, False, default_unit=O.UNIT_S)
    assert e.encode(1, 0.0, 0) != 0
    e = O.Encoder(1, False, default_unit=O.UNIT_S)
    assert e.encode(1, 0.0, 255) != 0


def test_init_time_unit():
    for c in G["init_time_unit"]["cases"]:
        assert O.lib().m3o_initial_time_unit(c["start_ns"], c["unit"]) == c["expected"]


# ---------------------------------------------------------------- full streams
@pytest.mark.parametrize("idx", range(4))
def test_golden_stream_encode(idx):
    s = G["streams"][idx]
    e = O.Encoder(s["encoder_start"], s["int_optimized"])
    assert e.stream() is None  # Stream() -> (nil,false) on empty encoder
    assert e.len() == 0 and e.empty()
    for dp in s["datapoints"]:
        assert e.encode(dp["ts"], dp["value"], dp["unit"], hb(dp["annotation"])) == 0
    assert e.stream() == hb(s["bytes"])
    assert e.len() == len(hb(s["bytes"]))
    assert e.num_encoded() == len(s["datapoints"])
    if "raw" in s:
        assert e.raw() == (hb(s["raw"]), s["raw_pos"])


@pytest.mark.parametrize("idx", range(4))
def test_golden_stream_decode(idx):
    s = G["streams"][idx]
    it = O.Iterator(hb(s["bytes"]), s["int_optimized"])
    for dp, ann in zip(s["datapoints"], s["decoded_annotations"]):
        assert it.next()
        t, v, u, a = it.current()
        assert (t, v, u, a) == (dp["ts"], float(dp["value"]), dp["unit"], hb(ann))
        assert it.err() == 0 and not it.done()
    for _ in range(2):
        assert not it.next()
        assert it.err() == 0 and it.done()


def test_iterator_read_next_timestamp():
    g = G["read_next_timestamp"]
    for c in g["cases"]:
        it = O.Iterator(hb(c["bytes"]), False)
        O.lib().m3o_iter_set_ts_state(it.h, c["unit"], c["prev_delta_ns"])
        assert O.lib().m3o_iter_read_next_timestamp(it.h) == 0
        assert O.lib().m3o_iter_prev_time_delta(it.h) == c["expected_delta_ns"]
    it = O.Iterator(hb(g["error_stream"]), False)
    assert O.lib().m3o_iter_read_first_timestamp(it.h) != 0
    assert O.lib().m3o_iter_read_next_timestamp(it.h) != 0
    assert O.lib().m3o_iter_read_next_timestamp(it.h) != 0


def test_iterator_read_next_value():
    import ctypes as C
    g = G["read_next_value"]
    for c in g["cases"]:
        it = O.Iterator(hb(c["bytes"]), False)
        O.lib().m3o_iter_set_float_state(it.h, c["prev_value"], c["prev_xor"])
        O.lib().m3o_iter_read_next_value(it.h)
        pb, px = C.c_uint64(), C.c_uint64()
        O.lib().m3o_iter_get_float_state(it.h, C.byref(pb), C.byref(px))
        assert px.value == c["expected_xor"] and pb.value == c["expected_value"]
        assert it.err() == 0
    it = O.Iterator(hb(g["error_stream"]), False)
    O.lib().m3o_iter_read_next_value(it.h)
    assert it.err() != 0
    # iterator_test.go:217-221: after the failed value read Next() is false, not done
    assert not it.next() and not it.done() and it.err() != 0


def test_iterator_read_annotation():
    import ctypes as C
    for c in G["read_annotation"]["cases"]:
        it = O.Iterator(hb(c["bytes"]), False)
        p = O.u8p()
        n = C.c_size_t()
        assert O.lib().m3o_iter_read_annotation(it.h, C.byref(p), C.byref(n)) == 0
        assert bytes(p[: n.value]) == hb(c["annotation"])


def test_iterator_read_time_unit():
    import ctypes as C
    for c in G["read_time_unit"]["cases"]:
        it = O.Iterator(hb(c["bytes"]), False)
        O.lib().m3o_iter_set_ts_state(it.h, c["unit"], 0)
        u, ch = C.c_int(), C.c_int()
        assert O.lib().m3o_iter_read_time_unit(it.h, C.byref(u), C.byref(ch)) == 0
        assert u.value == c["expected_unit"] and bool(ch.value) == c["expected_changed"]


def test_iterator_error_streams():
    for c in G["iterator_error_streams"]["cases"]:
        it = O.Iterator(hb(c["bytes"]), c["int_optimized"])
        assert not it.next()
        assert not it.done()
        assert it.err() != 0


# ---------------------------------------------------------------- int-optimized fixtures
def _fixture_streams():
    return [base64.b64decode(s) for s in G["fixtures_b64"]["streams"]]


def test_fixtures_decode_counts():
    total_dp = total_b = 0
    for data, exp in zip(_fixture_streams(), G["fixtures_b64"]["expected_points"]):
        dps, err = O.decode_all(data, True)
        assert err == 0
        assert len(dps) == exp
        ts = [d[0] for d in dps]
        assert ts == sorted(ts)
        total_dp += len(dps)
        total_b += len(data)
    assert total_dp == 7197 and total_b == 11458


def test_fixtures_reencode_byte_identical():
    """Decode -> re-encode with the stream's own start, per-dp unit and
    annotations must reproduce the fixture bytes (the only byte-level pin that
    exists for intOptimized=true encoding)."""
    for data in _fixture_streams():
        dps, err = O.decode_all(data, True)
        assert err == 0
        start = struct.unpack(">q", data[:8])[0]
        e = O.Encoder(0, True)
        e.reset(start)
        for t, v, u, a in dps:
            assert e.encode(t, v, u, a) == 0
        assert e.stream() == data


def test_regression_stream():
    data = base64.b64decode(G["regression_b64"]["stream"])
    dps, err = O.decode_all(data, True)
    assert err == 0
    assert len(dps) == 150
    assert dps[0][1] == -(2.0 ** 63)
    start = struct.unpack(">q", data[:8])[0]
    e = O.Encoder(0, True)
    e.reset(start)
    for t, v, u, a in dps:
        assert e.encode(t, v, u, a) == 0
    dps2, err2 = O.decode_all(e.stream(), True)
    assert err2 == 0 and dps2 == dps


# ---------------------------------------------------------------- DoD overflow errors
@pytest.mark.parametrize("delta_h,unit,overflow", [
    (1, O.UNIT_S, False), (25 * 24, O.UNIT_S, False), (1000 * 25 * 24, O.UNIT_S, True),
    (1, O.UNIT_MS, False), (24 * 24, O.UNIT_MS, False), (25 * 24, O.UNIT_MS, True),
    (1, O.UNIT_US, False), (25 * 24, O.UNIT_US, False),
    (1, O.UNIT_NS, False), (25 * 24, O.UNIT_NS, False),
])
def test_dod_overflow(delta_h, unit, overflow):
    # encoder_test.go:581-730
    start = 1427162400 * SEC
    delta = delta_h * 3600 * SEC
    one_unit = {O.UNIT_S: SEC, O.UNIT_MS: 10 ** 6, O.UNIT_US: 10 ** 3, O.UNIT_NS: 1}[unit]
    e = O.Encoder(start, False)
    assert e.encode(start, 1, unit) == 0
    err = e.encode(start + delta, 2, unit)
    assert (err == O.ERR_DOD_OVERFLOW) == overflow
    if not overflow:
        dps, derr = O.decode_all(e.stream(), False)
        assert derr == 0
        assert [(d[0], d[1], d[2]) for d in dps] == [(start, 1.0, unit), (start + delta, 2.0, unit)]
    e = O.Encoder(start, False)
    pts = [(start, 1), (start + delta // 2, 2), (start + delta // 2 + delta, 3),
           (start + delta // 2 + delta + one_unit, 4)]
    for i, (t, v) in enumerate(pts):
        err = e.encode(t, v, unit)
        if i == 3 and overflow:
            assert err == O.ERR_DOD_OVERFLOW
            return
        assert err == 0
    dps, derr = O.decode_all(e.stream(), False)
    assert derr == 0 and [(d[0], d[1]) for d in dps] == [(t, float(v)) for t, v in pts]


# ---------------------------------------------------------------- round trips (roundtrip_test.go)
def _gen_float_val(r, num_dig, num_dec):
    # src/dbnode/encoding/testgen/gen.go:30-44 (GenerateFloatVal): "<dig>.<dec>" parsed as float64, the decimal
    # part NOT zero-padded (dec = 5 with numDec = 3 gives x.5)
    dig = r.getrandbits(62) % 10 ** num_dig
    if num_dec == 0:
        return float(dig)
    return float("%d.%d" % (dig, r.getrandbits(62) % 10 ** num_dec))


def _roundtrip(dps, int_opt):
    start = 1427162400 * SEC
    e = O.Encoder(start, int_opt)
    units, anns = [], []
    for i, (t, v) in enumerate(dps):
        unit = O.UNIT_S
        if i == 0:
            unit = O.UNIT_MS
        elif i == 10:
            unit = O.UNIT_US
        ann = b""
        if i < 5:
            ann = b"foo"
        elif i < 7:
            ann = b"bar"
        elif i == 10:
            ann = b"long annotation long annotation long annotation long annotation"
        units.append(unit)
        anns.append(ann)
        assert e.encode(t, v, unit, ann) == 0
    out, err = O.decode_all(e.stream(), int_opt)
    assert err == 0
    assert len(out) == len(dps)
    for i, (t, v, u, a) in enumerate(out):
        exp_ann = anns[i]
        if i > 0 and anns[i - 1] == exp_ann:
            exp_ann = b""
        assert t == dps[i][0], i
        assert v == dps[i][1] or (math.isnan(v) and math.isnan(dps[i][1])), (i, v, dps[i][1])
        assert u == units[i], i
        assert a == exp_ann, i


@pytest.mark.parametrize("num_dig,num_dec,neg,mixsign", [
    (12, 0, False, False), (7, 6, False, False), (0, 1, False, False), (2, 16, False, False),
    (5, 3, True, False), (3, 0, False, True)])
@pytest.mark.parametrize("int_opt", [True, False])
def test_roundtrip_generated(num_dig, num_dec, neg, mixsign, int_opt):
    r = random.Random(num_dig * 100 + num_dec)
    for _ in range(8):
        t = 1427162462 * SEC
        end = 1427162400 * SEC + 2 * 3600 * SEC
        dps = [(t, 1.0)]
        for _i in range(1, 1000):
            t += SEC * r.randrange(1200)
            if t >= end:
                break
            v = _gen_float_val(r, num_dig, num_dec)
            if neg or (mixsign and r.random() < 0.5):
                v = -v
            dps.append((t, v))
        _roundtrip(dps, int_opt)


@pytest.mark.parametrize("int_opt", [True, False])
def test_roundtrip_mixed(int_opt):
    """roundtrip_test.go:88-95,237-262 (TestMixedRoundTrip / generateMixedDatapoints): steps of up to two hours,
    the value switches between 5-digit integers (10 %), 3.16-digit floats (18 %) and a repeat of the last one."""
    r = random.Random(4242)
    for _ in range(8):
        t = 1427162462 * SEC
        end = 1427162400 * SEC + 2 * 3600 * SEC
        v = _gen_float_val(r, 3, 16)
        dps = [(t, v)]
        for _i in range(1, 1000):
            t += SEC * r.randrange(7200)
            if r.random() < 0.1:
                v = _gen_float_val(r, 5, 0)
            elif r.random() < 0.2:
                v = _gen_float_val(r, 3, 16)
            if t >= end:
                break
            dps.append((t, v))
        _roundtrip(dps, int_opt)
    # the same value process at a one-minute cadence, so that streams are long enough to go through
    # int -> float -> int mode changes many times
    for _ in range(4):
        t = 1427162462 * SEC
        v = _gen_float_val(r, 3, 16)
        dps = [(t, v)]
        for _i in range(1, 600):
            t += 10 * SEC
            if r.random() < 0.1:
                v = _gen_float_val(r, 5, 0)
            elif r.random() < 0.2:
                v = _gen_float_val(r, 3, 16)
            dps.append((t, v))
        _roundtrip(dps, int_opt)


@pytest.mark.parametrize("int_opt", [True, False])
def test_roundtrip_overflow_and_precision(int_opt):
    g = G["overflow_values"]
    dps = [(g["start_ns"] + i * g["step_ns"], v) for i, v in enumerate(g["values"])]
    _roundtrip(dps, int_opt)
    v = G["precision_value"]["value"]
    t0 = 1600000000 * SEC
    _roundtrip([(t0 + i * 60 * SEC, v) for i in range(100)], int_opt)


# ---------------------------------------------------------------- convertToIntFloat (m3tsz_test.go)
def test_convert_to_int_float():
    # m3tsz_test.go:34-100 semantic checks
    assert O.convert_to_int_float(46.0, 0) == (46.0, 0, False, 0)
    assert O.convert_to_int_float(46.000000000000001, 0) == (46.0, 0, False, 0)
    v, m, isf, err = O.convert_to_int_float(12.5, 0)
    assert (v, m, isf, err) == (125.0, 1, False, 0)
    v, m, isf, err = O.convert_to_int_float(-12.345, 0)
    assert (v, m, isf, err) == (-12345.0, 3, False, 0)
    for bad in (float("inf"), float("-inf"), float("nan")):
        _, _, isf, err = O.convert_to_int_float(bad, 0)
        assert isf and err == 0
    assert O.convert_to_int_float(1.0, 7)[3] != 0  # errInvalidMultiplier
    r = random.Random(7)
    for _ in range(2000):
        mult = r.randrange(0, 7)
        iv = r.randrange(-10 ** 9, 10 ** 9)
        x = iv / (10.0 ** mult)
        val, m, isf, err = O.convert_to_int_float(x, 0)
        assert err == 0 and not isf
        assert O.lib().m3o_convert_from_int_float(val, m) == x


def _go_itoa_pad(dec, num_dec):
    s = str(dec)
    return s + "0" * (num_dec - len(s)) if len(s) < num_dec else s


@pytest.mark.parametrize("num_dig,num_dec,neg", [
    (0, 0, False), (1, 0, False), (2, 0, False), (10, 0, False), (18, 0, False),   # TestCountConversions :34-40
    (0, 6, False), (1, 6, False), (3, 6, False), (5, 6, False), (7, 6, False),      # TestTimerConversions :42-48
    (0, 1, False), (0, 3, False), (1, 3, False), (3, 3, False), (5, 3, False), (7, 3, False),  # SmallGauge :50-57
    (1, 0, True), (3, 0, True), (1, 2, True), (3, 2, True),                         # NegativeGauge :71-76
])
def test_convert_to_int_float_reference_int_families(num_dig, num_dec, neg):
    """m3tsz_test.go:102-136 (validateIntConversions), the reference's own property test restated family by
    family: decimal strings "<dig>.<dec>" parsed to float64 convert to the integer "<dig><dec padded>" with a
    multiplier <= numDec, starting from curMaxMult = numDec."""
    r = random.Random(1000 * num_dig + 10 * num_dec + int(neg))
    dig_mod, dec_mod = 10 ** num_dig, 10 ** num_dec
    sign = -1.0 if neg else 1.0
    for _ in range(1000):
        dig, dec = r.getrandbits(62) % dig_mod, r.getrandbits(62) % dec_mod
        if num_dec == 0:
            val = sign * float(dig)
            iv, m, isf, err = O.convert_to_int_float(val, num_dec)
            assert err == 0 and not isf and iv == val and m <= 0, (dig, dec)
        else:
            val = float("%d.%d" % (dig, dec))
            expected = int(str(dig) + _go_itoa_pad(dec, num_dec))
            iv, m, isf, err = O.convert_to_int_float(sign * val, num_dec)
            assert err == 0 and not isf and iv == sign * float(expected) and m <= num_dec, (dig, dec, iv, m)


@pytest.mark.parametrize("num_dig,num_dec", [(0, 16), (1, 16), (5, 16),   # TestPreciseGaugeConversions :59-63
                                             (9, 2), (10, 3), (11, 3)])    # TestLargeGaugeConversions :65-69
def test_convert_to_int_float_reference_float_families(num_dig, num_dec):
    """m3tsz_test.go:138-183 (testFloatConversions / validateConvertFloat): either the value stays a float
    (returned unchanged, multiplier 0) or the integer found reproduces it to within 10^-mult."""
    r = random.Random(77 * num_dig + num_dec)
    dig_mod, dec_mod = 10 ** num_dig, 10 ** num_dec
    for _ in range(1000):
        dig, dec = r.getrandbits(62) % dig_mod, r.getrandbits(62) % dec_mod
        val = float(str(dig) + "." + _go_itoa_pad(dec, num_dec))
        v, m, isf, err = O.convert_to_int_float(val, 0)
        assert err == 0
        if isf:
            assert v == val and m == 0
        else:
            assert abs(val - v / 10.0 ** m) < 1.0 / 10.0 ** m, (val, v, m)


def test_convert_to_int_float_inf_nan_and_from():
    """m3tsz_test.go:78-100: TestConvertFromIntFloat, TestInfNan (every curMaxMult), TestInvalidMult."""
    for val, mult, exp in ((1.0, 0, 1.0), (2.0, 0, 2.0), (10.0, 1, 1.0), (200.0, 2, 2.0)):
        assert O.lib().m3o_convert_from_int_float(val, mult) == exp
    for cur in (0, 3, 6):
        for bad in (float("inf"), float("-inf")):
            assert O.convert_to_int_float(bad, cur) == (bad, 0, True, 0)
        v, m, isf, err = O.convert_to_int_float(float("nan"), cur)
        assert v != v and m == 0 and isf and err == 0
    assert O.convert_to_int_float(1.0, 7)[3] != 0


def test_xxh64_known_answers():
    # XXH64 seed 0 published test values
    assert O.lib().m3o_xxh64(None, 0) == 0xEF46DB3751D8E999
    assert O.lib().m3o_xxh64(b"a", 1) == 0xD24EC4F1A98C6E5B
    assert O.lib().m3o_xxh64(b"abc", 3) == 0x44BC2CF5AD770999
    msg = b"Nobody inspects the spammish repetition"
    assert O.lib().m3o_xxh64(msg, len(msg)) == 0xFBCEA83C8A378BF1


# ---------------------------------------------------------------- encoder behaviours (encoder_test.go:412-579)
def test_encoder_reset_len_last_encoded():
    rng = np.random.default_rng(int(1427162400))
    start = 1427162400 * SEC
    for int_opt in (False, True):
        e = O.Encoder(start, int_opt)
        assert e.last_encoded()[2] != 0  # errNoEncodedDatapoints
        assert e.last_annotation_checksum()[1] != 0
        for _pass in range(8):
            n = int(rng.integers(1, 512))
            t = start
            for i in range(n):
                t += int(rng.integers(1, 10 ** 9))
                v = float(rng.normal())
                assert e.encode(t, v, O.UNIT_NS) == 0
                assert e.num_encoded() == i + 1
                lt, lv, lerr = e.last_encoded()
                assert lerr == 0 and lt == t
                if int_opt:
                    assert lv == v  # gaussian values stay in float mode
                assert e.len() == len(e.stream())
            e.reset(start)
            assert e.empty() and e.len() == 0 and e.num_encoded() == 0
        e.close()
        assert e.encode(start, 1.0, O.UNIT_S) != 0  # errEncoderClosed


def test_last_annotation_checksum_tracks_last_written():
    e = O.Encoder(1427162400 * SEC, False)
    e.encode(1427162462 * SEC, 12, O.UNIT_S, b"\x0a")
    assert e.last_annotation_checksum() == (O.lib().m3o_xxh64(b"\x0a", 1), 0)
    e.encode(1427162522 * SEC, 12, O.UNIT_S, b"")
    assert e.last_annotation_checksum() == (O.lib().m3o_xxh64(b"\x0a", 1), 0)


# ---------------------------------------------------------------- batch helpers agree with the per-dp objects
def test_batch_helpers_match_objects():
    rng = np.random.default_rng(3)
    S, P = 16, 200
    start = 1599955200 * SEC
    ts = start + np.arange(P, dtype=np.int64)[None, :] * 60 * SEC + np.zeros((S, 1), dtype=np.int64)
    vals = 100.0 + np.cumsum(rng.normal(size=(S, P)), axis=1)
    vals[:, 0] = 100.0
    vals[3] = np.round(vals[3], 2)
    vals[4] = np.round(vals[4])
    for int_opt in (True, False):
        out, out_len, status = O.encode_batch(ts, vals, start, O.UNIT_S, int_opt, n_threads=3)
        assert (status == 0).all()
        off = np.zeros(S + 1, dtype=np.uint64)
        off[1:] = np.cumsum(out_len)
        blob = np.concatenate([out[s, : out_len[s]] for s in range(S)])
        for s in range(S):
            e = O.Encoder(0, int_opt)
            e.reset(start)
            for t, v in zip(ts[s], vals[s]):
                assert e.encode(int(t), float(v), O.UNIT_S) == 0
            assert e.stream() == out[s, : out_len[s]].tobytes()
        dts, dvals, n, st = O.decode_batch(blob, off, P + 8, int_opt, n_threads=3)
        assert (st == 0).all() and (n == P).all()
        assert (dts[:, :P] == ts).all()
        assert (dvals[:, :P] == vals).all()


def test_downsample_oracle_gauge_semantics():
    ts = np.array([0, 60, 120, 299, 300, 900, 901], dtype=np.int64) * SEC
    vals = np.array([1.0, float("nan"), 3.0, -2.0, 5.0, 7.0, 8.0])
    s, c, mn, mx, last = O.downsample_series(ts, vals, 0, 300 * SEC, 4)
    assert list(c) == [4, 1, 0, 2]
    assert s[0] == 2.0 and mn[0] == -2.0 and mx[0] == 3.0 and last[0] == -2.0
    assert s[1] == 5.0 and s[2] == 0.0 and math.isnan(mn[2]) and math.isnan(mx[2])
    assert s[3] == 15.0 and last[3] == 8.0


# ------------------------------------------------------------------ Gauge tables (pins the downsample oracle)
def test_downsample_oracle_pinned_by_gauge_test_tables():
    """aggregator/aggregation/gauge_test.go tables, through m3o_downsample_series (one window
    = one Gauge) and m3o_gauge_value_of (Gauge.ValueOf)."""
    for case in G["gauge"]["cases"]:
        ts = np.array(case["times"], dtype=np.int64) * SEC + 1000 * SEC
        vals = np.array(case["values"], dtype=np.float64)
        s, c, mn, mx, last = O.downsample_series(ts, vals, 0, 10_000 * SEC, 1)
        assert last[0] == case["last"], case["name"]
        assert O.gauge_value_of(O.AGG_LAST, s[0], c[0], mn[0], mx[0], last[0]) == case["last"]
        if "count" not in case:
            continue
        assert c[0] == case["count"] and s[0] == case["sum"], case["name"]
        assert O.gauge_value_of(O.AGG_COUNT, s[0], c[0], mn[0], mx[0], last[0]) == float(case["count"])
        assert O.gauge_value_of(O.AGG_SUM, s[0], c[0], mn[0], mx[0], last[0]) == case["sum"]
        assert O.gauge_value_of(O.AGG_MEAN, s[0], c[0], mn[0], mx[0], last[0]) == case["mean"]
        for key, arr, agg in (("min", mn, O.AGG_MIN), ("max", mx, O.AGG_MAX)):
            got = O.gauge_value_of(agg, s[0], c[0], mn[0], mx[0], last[0])
            if case[key] == "NaN":
                assert math.isnan(arr[0]) and math.isnan(got)
            else:
                assert arr[0] == case[key] and got == case[key]
        # SumSq / Stdev / unsupported types report 0 without HasExpensiveAggregations (gauge_test.go:43)
        assert O.gauge_value_of(8, s[0], c[0], mn[0], mx[0], last[0]) == 0.0


# ------------------------------------------------------------------ Prometheus epilogue tables
def test_prom_convert_oracle_counter_normalization():
    T = G["prom_counter_normalization"]
    for case in T["cases"]:
        res = case["max_resolution_ns"]
        handle = case["is_counter"] and res >= T["resolution_threshold_ns"]
        ts = [d[0] for d in case["given"]]
        vals = [float(d[1]) for d in case["given"]]
        to, vo = O.prom_convert_series(ts, vals, res, handle)
        assert [[int(a), float(b)] for a, b in zip(to, vo)] == [[w[0], float(w[1])] for w in case["want"]], case["name"]


def test_prom_convert_oracle_value_decrease_tolerance():
    T = G["prom_value_decrease_tolerance"]
    for case in T["cases"]:
        n = len(case["given"])
        ts = [T["now_ns"] + i * T["step_ns"] for i in range(n)]
        to, vo = O.prom_convert_series(ts, case["given"], 0, False, case["tolerance"], case["until_ns"])
        assert list(to) == [t // 1_000_000 for t in ts], case["name"]
        assert list(vo) == [float(v) for v in case["want"]], case["name"]


def test_aggregate_tiles_oracle_shape():
    ts = np.array([0, 60, 120, 299, 300, 900, 901], dtype=np.int64) * SEC
    vals = np.array([1.0, float("nan"), 3.0, -2.0, 5.0, 7.0, 8.0])
    to, vo = O.aggregate_tiles_series(ts, vals, 0, 300 * SEC, 4, O.AGG_LAST)
    assert list(to) == [300 * SEC, 600 * SEC, 1200 * SEC] and list(vo) == [-2.0, 5.0, 8.0]
    to, vo = O.aggregate_tiles_series(ts, vals, 0, 300 * SEC, 4, O.AGG_MEAN)
    assert list(vo) == [0.5, 5.0, 7.5]
