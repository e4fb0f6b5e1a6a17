"""GPU parity of the segment-checksum row against the oracle
(oracle/m3tsz_segment_oracle.c, pinned by tests/test_checksum_oracle.py)."""
import zlib

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def codec():
    from m3_b200.codec import BatchCodec
    return BatchCodec(0, True)


def _upload(blobs, pad_to=1):
    padded = [b + b"\0" * ((-len(b)) % pad_to) for b in blobs]
    off = np.zeros(len(blobs) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(b) for b in padded])
    blob = b"".join(padded)
    buf = torch.zeros(len(blob) + 16, dtype=torch.uint8, device="cuda")
    if blob:
        buf[: len(blob)] = torch.frombuffer(bytearray(blob), dtype=torch.uint8).cuda()
    return buf[: len(blob)], torch.from_numpy(off).cuda(), blob, off


def test_random_streams_every_alignment_and_length(codec):
    rng = np.random.default_rng(3)
    lens = list(range(0, 80)) + [255, 256, 257, 511, 512, 513, 4095, 5552, 5553, 65521, 70001] + \
        [int(x) for x in rng.integers(0, 3000, size=200)]
    blobs = [rng.integers(0, 256, size=n, dtype=np.uint8).tobytes() for n in lens]
    blobs.append(b"\xff" * 300000)  # worst case for the accumulators
    d, off, blob, h_off = _upload(blobs)
    ck, st = codec.segment_checksums(d, off)
    torch.cuda.synchronize()
    got = ck.cpu().numpy().view(np.uint32)
    exp, _ = O.adler32_batch(blob, h_off.astype(np.uint64))
    assert (got == exp).all(), np.nonzero(got != exp)[0][:10]
    assert (st.cpu().numpy() == 0).all()
    assert [int(x) for x in got[:5]] == [zlib.adler32(b) for b in blobs[:5]]


def test_padded_starts_with_lengths_and_expected(codec):
    rng = np.random.default_rng(4)
    blobs = [rng.integers(0, 256, size=int(n), dtype=np.uint8).tobytes() for n in rng.integers(0, 2000, size=300)]
    d, off, blob, h_off = _upload(blobs, pad_to=64)
    lens = torch.tensor([len(b) for b in blobs], dtype=torch.int64, device="cuda")
    want = np.array([zlib.adler32(b) for b in blobs], dtype=np.uint32)
    exp = want.copy()
    bad = [7, 123, 299]
    exp[bad] ^= 0x10
    ck, st = codec.segment_checksums(d, off, lengths=lens,
                                     expected=torch.from_numpy(exp.view(np.int32)).cuda())
    torch.cuda.synchronize()
    assert (ck.cpu().numpy().view(np.uint32) == want).all()
    st = st.cpu().numpy()
    assert (np.nonzero(st)[0] == bad).all() and (st[bad] == 15).all()
    # (offset, size) addressing: a size that runs past the end of the buffer is an argument error
    # for that stream only; overlapping the next stream is the caller's business
    lens2 = lens.clone()
    lens2[5] = int(d.numel()) + 1
    _, st2 = codec.segment_checksums(d, off, lengths=lens2)
    st2 = st2.cpu().numpy()
    assert st2[5] == 101 and (np.delete(st2, 5) == 0).all()


def test_checksums_of_encoded_streams_match_the_oracle(codec):
    """Encode on the GPU, compact with padded starts, checksum the exact stream bytes: equals
    the oracle's Adler-32 of the oracle-encoded stream (bitstreams are identical)."""
    from m3_b200 import synth
    S, P = 512, 200
    ts, vals, start = synth.gaussian_walk(S, P, "cuda", seed=9)
    enc = codec.encode(ts, vals, start, unit=1)
    packed, offsets = codec.compact(enc, align=64)
    ck, st = codec.segment_checksums(packed, offsets, lengths=enc.out_len)
    torch.cuda.synchronize()
    got = ck.cpu().numpy().view(np.uint32)
    h_ts, h_vals = ts.cpu().numpy(), vals.cpu().numpy()
    for s in range(0, S, 37):
        stream = O.encode_series(h_ts[s], h_vals[s], int(start[s].item()), O.UNIT_S, True)
        assert got[s] == O.adler32(stream), s
    assert (st.cpu().numpy() == 0).all()
