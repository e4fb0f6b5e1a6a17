"""The encode kernel's int-likeness filter (`maybe_int`, m3_b200/csrc/m3tsz_encode.cu) claims to be a NECESSARY
condition for the reference's convertToIntFloat (m3tsz/m3tsz.go:78-119) to return an int for any current
multiplier: a value it rejects ("float for certain") must be a float for the reference.  DESIGN.md §4 has the
proof; this test hunts for counter-examples on the CPU: the filter restated in numpy float64 (same IEEE
operations: |v|, one rounded multiply, rint, one exact subtract, compares) against the oracle's classifier, on
the reference's own value families, on adversarial values a few ulps around k / 10^m, and on random doubles of
every magnitude.  (The kernel itself is compared with the oracle byte for byte in the GPU suite; this test covers
the part of the input space those batches cannot enumerate.)"""
import math
import random
import struct

import numpy as np
import pytest

import oracle_lib as O


def maybe_int(v):
    """numpy restatement of __device__ maybe_int(double) -- keep in step with m3tsz_encode.cu."""
    v = np.asarray(v, dtype=np.float64)
    with np.errstate(all="ignore"):
        a = np.abs(v)
        p = a * 1000000.0
        r = p - np.rint(p)
        far = np.abs(r) > p * 2.0 ** -49
        in_rng = (p >= 1.0) & (p < 2.0 ** 48)
        tiny_v = (p < 0.5) & (a >= 1e-300)
    return ~((in_rng & far) | tiny_v)


def _oracle_is_int_for_some_cur(v):
    for cur in range(7):
        _, _, is_float, err = O.convert_to_int_float(v, cur)
        assert err == 0
        if not is_float:
            return True
    return False


def _check(values):
    values = np.asarray(values, dtype=np.float64)
    keep = maybe_int(values)
    rejected = values[~keep]
    for v in rejected.tolist():
        assert not _oracle_is_int_for_some_cur(v), "filter rejected %r (%s) but the reference finds an int" % (
            v, struct.pack(">d", v).hex())
    return len(rejected), len(values)


def _ulp_neighbours(x, k):
    out = [x]
    lo = hi = x
    for _ in range(k):
        lo = math.nextafter(lo, -math.inf)
        hi = math.nextafter(hi, math.inf)
        out += [lo, hi]
    return out


def test_filter_never_rejects_reference_int_families():
    r = random.Random(5)
    vals = []
    for num_dig, num_dec in [(0, 0), (1, 0), (2, 0), (10, 0), (18, 0), (0, 6), (1, 6), (3, 6), (5, 6), (7, 6),
                             (0, 1), (0, 3), (1, 3), (3, 3), (5, 3), (7, 3), (12, 0), (2, 16), (5, 16), (9, 2),
                             (10, 3), (11, 3)]:
        for _ in range(1500):
            dig, dec = r.getrandbits(62) % 10 ** num_dig, r.getrandbits(62) % 10 ** num_dec
            v = float(dig) if num_dec == 0 else float("%d.%d" % (dig, dec))
            vals += [v, -v]
    rejected, total = _check(vals)
    assert total == 66000 and rejected > 3000  # the float families are rejected, none of them wrongly


def test_filter_adversarial_neighbourhood_of_decimals():
    """k / 10^m and its neighbours up to 6 ulps away, over every multiplier and magnitude the reference accepts:
    the reference's own tolerance is one ulp of val*10^m (math.Nextafter, m3tsz.go:107-114), so its int verdict
    flips somewhere inside this neighbourhood -- exactly where a too-eager filter would be wrong."""
    r = random.Random(6)
    xs = []
    for _ in range(8000):
        m = r.randrange(0, 7)
        k = r.getrandbits(r.randrange(1, 50)) + 1
        xs.append(k / 10.0 ** m)
    base = np.array(xs, dtype=np.float64).view(np.int64)  # positive doubles: +-j in the bit pattern = +-j ulps
    offs = np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 16, 20, 24, 32, 48, 64, 128, 1024], dtype=np.int64)
    near = np.concatenate([base + j for j in offs] + [base - j for j in offs[1:]]).view(np.float64)
    vals = np.concatenate([near, -near])
    rejected, total = _check(vals)
    # the filter keeps everything within its 8-ulp margin and starts rejecting beyond it; none wrongly
    assert total == 8000 * 41 * 2 and 20000 < rejected < total // 2
    kept_close = maybe_int(np.concatenate([base + j for j in range(-4, 5)]).view(np.float64))
    assert kept_close.all()


def test_filter_random_doubles_all_magnitudes():
    rng = np.random.default_rng(7)
    bits = rng.integers(0, 2 ** 63, size=120000, dtype=np.int64).astype(np.uint64)
    bits |= (rng.integers(0, 2, size=bits.size, dtype=np.uint64) << np.uint64(63))
    vals = bits.view(np.float64)
    # plus magnitudes where the filter's range tests switch: around 1e-6 (p = 1), 2^48 / 1e6, 1e-300, subnormals
    edge = []
    for c in (1e-6, 5e-7, 2.0 ** 48 / 1e6, 2.0 ** 52 / 1e6, 1e-300, 5e-324, 2.2250738585072014e-308, 2.0 ** 63, 1e300):
        edge += _ulp_neighbours(c, 8) + [-y for y in _ulp_neighbours(c, 8)]
    edge += [0.0, -0.0, math.inf, -math.inf, math.nan]
    rejected, total = _check(np.concatenate([vals, np.array(edge)]))
    assert rejected > 40000  # (random bit patterns are mostly huge or tiny: tiny ones are rejected, huge ones kept)


def test_filter_accepts_everything_it_must_keep():
    """NaN / Inf / zero / huge values are never rejected (they take the exact restatement, which classifies them
    like the reference: m3tsz_test.go:85-95)."""
    keep = maybe_int(np.array([math.nan, math.inf, -math.inf, 0.0, -0.0, 2.0 ** 60, -2.0 ** 63, 1e300, 5e-324]))
    assert keep.all()
