"""Pins the segment-checksum oracle (oracle/m3tsz_segment_oracle.c).

The reference computes ts.Segment.CalculateChecksum with m3db/stackadler32 /
hash/adler32 (src/dbnode/ts/segment.go:60-76, src/dbnode/digest/digest.go:36-38);
its own tests hold no checksum constants for stream bytes, so the oracle is pinned
by the published known answers of the algorithm (RFC 1950 Adler-32; the table of
Go's hash/adler32 tests) and by a second implementation (zlib.adler32)."""
import zlib

import numpy as np

import oracle_lib as O

# (checksum, input) -- hash/adler32 golden table (Go standard library) and the RFC example
KAT = [
    (0x00000001, b""),
    (0x00620062, b"a"),
    (0x012600c4, b"ab"),
    (0x024d0127, b"abc"),
    (0x03d8018b, b"abcd"),
    (0x05c801f0, b"abcde"),
    (0x081e0256, b"abcdef"),
    (0x0adb02bd, b"abcdefg"),
    (0x0e000325, b"abcdefgh"),
    (0x118e038e, b"abcdefghi"),
    (0x158603f8, b"abcdefghij"),
    (0x11e60398, b"Wikipedia"),
    (0x29750586, b"message digest"),
    (0x90860b20, b"abcdefghijklmnopqrstuvwxyz"),
    (0x00000001 + (0xff << 0) + ((0x100) << 16), b"\xff"),  # a = 256, b = 256
]


def test_known_answers():
    for want, data in KAT:
        assert O.adler32(data) == want, data
        assert zlib.adler32(data) == want, data  # the table itself against the second implementation


def test_against_zlib_random_lengths():
    rng = np.random.default_rng(7)
    for n in list(range(0, 70)) + [255, 256, 257, 5551, 5552, 5553, 11104, 65520, 65521, 65522, 1 << 20]:
        data = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
        assert O.adler32(data) == zlib.adler32(data), n
    # worst case for the deferred modulo: all 0xff
    for n in (5552, 5553, 100000):
        assert O.adler32(b"\xff" * n) == zlib.adler32(b"\xff" * n)


def test_batch_and_mismatch_status():
    rng = np.random.default_rng(8)
    lens = [0, 1, 15, 16, 17, 1000, 0, 4097]
    blobs = [rng.integers(0, 256, size=n, dtype=np.uint8).tobytes() for n in lens]
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    out, st = O.adler32_batch(b"".join(blobs), off)
    assert [int(x) for x in out] == [zlib.adler32(b) for b in blobs] and (st == 0).all()
    exp = out.copy()
    exp[3] ^= 1
    exp[6] = 0  # the empty stream's checksum is 1
    out2, st2 = O.adler32_batch(b"".join(blobs), off, expected=exp)
    assert (out2 == out).all()
    assert st2.tolist() == [0, 0, 0, 15, 0, 0, 15, 0]


def test_checksum_of_real_streams_is_head_then_tail():
    """Segment = head || tail: the checksum over the two parts in order equals the
    checksum of the whole stream (what the data file stores contiguously)."""
    ts = 1599955200 * 10**9 + np.arange(100) * 60 * 10**9
    vals = 100 + np.cumsum(np.random.default_rng(1).normal(size=100))
    stream = O.encode_series(ts, vals, int(ts[0]), O.UNIT_S, True)
    head, tail = stream[:-1], stream[-1:]
    whole = O.adler32(stream)
    # incremental digest: adler32(tail, seed = adler32(head))
    assert zlib.adler32(tail, zlib.adler32(head)) == whole
