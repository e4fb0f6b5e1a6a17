"""Parity at BASELINE.json's full sizes, against the ORACLE (not the kernels against themselves):
  * configs[2] 1M x 1440 encode, both modes: every stream's length and Adler-32 (computed on the
    device by m3tsz_checksum_batch) equal the oracle encoder's -- a checksum of checksums over
    10.5 GB of output, 4 B + 8 B per series crossing PCIe;
  * the accidental int-mode series (a Gaussian value within an ulp of a short
    decimal flips the encoder into int mode; ~50 per 1M x 1440): found on the device, each one
    compared with the oracle byte for byte and value for value, and at least one must exist;
  * configs[3] fused decode + 5-min downsample on 16k x 1440: every window of every series."""
import zlib

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


@pytest.mark.parametrize("int_opt", [True, False])
def test_encode_1m_x_1440_checksums_vs_oracle(codecs, int_opt):
    import os
    from m3_b200 import synth
    S, P, CH = 1_000_000, 1440, 125_000
    codec = codecs[int_opt]
    ts, vals, start = synth.gaussian_walk(S, P, "cuda", seed=4242)
    pk = codec.encode_packed(ts, vals, start, unit=O.UNIT_S, align=64, capacity=S * (8 * P + 256))
    torch.cuda.synchronize()
    assert int((pk.status != 0).sum()) == 0
    ck, st = codec.segment_checksums(pk.packed, pk.offsets, lengths=pk.out_len)
    torch.cuda.synchronize()
    assert int((st != 0).sum()) == 0
    g_ck = ck.cpu().numpy().view(np.uint32)
    g_len = pk.out_len.cpu().numpy()
    threads = max(1, len(os.sched_getaffinity(0)))
    stride = 64 + 20 * P
    bufs = (np.zeros((CH, stride), dtype=np.uint8), np.zeros(CH, dtype=np.uint64), np.zeros(CH, dtype=np.int32))
    s0 = int(start[0].item())
    for c0 in range(0, S, CH):
        h_ts, h_vals = ts[c0:c0 + CH].cpu().numpy(), vals[c0:c0 + CH].cpu().numpy()
        out, ln, status = O.encode_batch(h_ts, h_vals, s0, O.UNIT_S, int_opt, n_threads=threads, bufs=bufs)
        assert (status == 0).all()
        assert (ln.astype(np.int64) == g_len[c0:c0 + CH]).all(), c0
        o_ck = np.fromiter((zlib.adler32(out[i, : ln[i]]) for i in range(CH)), dtype=np.uint32, count=CH)
        bad = np.nonzero(o_ck != g_ck[c0:c0 + CH])[0]
        assert len(bad) == 0, (c0, bad[:5])


def test_accidental_int_series_vs_oracle(codecs):
    """int-optimised mode over 400k x 1440 Gaussian series: the series whose decoded values differ
    from the input are exactly the ones the reference encoder moves into int mode; compare every one
    of them (plus a random sample of ordinary ones) with the oracle, bytes and values."""
    from m3_b200 import synth
    S, P = 400_000, 1440
    codec = codecs[True]
    ts, vals, start = synth.gaussian_walk(S, P, "cuda", seed=31337)
    enc = codec.encode(ts, vals, start, unit=O.UNIT_S, out_stride=(64 + 9 * P + 63) // 64 * 64)
    assert int((enc.status != 0).sum()) == 0
    off = torch.arange(S, dtype=torch.int64, device="cuda") * enc.out.shape[1]
    dec = codec.decode(enc.out.view(-1), off, P, lengths=enc.out_len)
    torch.cuda.synchronize()
    assert int((dec.status != 0).sum()) == 0 and torch.equal(dec.ts, ts)
    differs = (dec.values.view(torch.int64) != vals.view(torch.int64)).any(dim=1)
    idx = torch.nonzero(differs).flatten().cpu().numpy()
    assert len(idx) >= 1, "no accidental int-mode series in 400k x 1440 (expected ~20)"
    assert len(idx) < 400
    extra = np.random.default_rng(5).integers(0, S, size=32)
    idx = np.unique(np.concatenate([idx, extra]))
    t_idx = torch.from_numpy(idx).cuda()
    h_ts, h_vals = ts[t_idx].cpu().numpy(), vals[t_idx].cpu().numpy()
    o_out, o_len, o_st = O.encode_batch(h_ts, h_vals, int(start[0].item()), O.UNIT_S, True, n_threads=8)
    g_out, g_len = enc.out[t_idx].cpu().numpy(), enc.out_len[t_idx].cpu().numpy()
    g_val = dec.values[t_idx].cpu().numpy().view(np.uint64)
    n_int_mode = 0
    for k in range(len(idx)):
        assert o_st[k] == 0 and g_len[k] == o_len[k], idx[k]
        assert (g_out[k, : g_len[k]] == o_out[k, : o_len[k]]).all(), idx[k]
        dps, err = O.decode_all(o_out[k, : o_len[k]].tobytes(), True)
        assert err == 0 and len(dps) == P
        ovals = np.array([d[1] for d in dps], dtype=np.float64).view(np.uint64)
        assert (g_val[k] == ovals).all(), idx[k]
        n_int_mode += int((ovals != h_vals[k].view(np.uint64)).any())
    assert n_int_mode >= 1


@pytest.mark.parametrize("int_opt", [True, False])
def test_downsample_16k_x_1440_every_window(codecs, int_opt):
    from m3_b200 import synth
    S, P = 16_384, 1440
    codec = codecs[int_opt]
    ts, vals, start = synth.gaussian_walk(S, P, "cuda", seed=99)
    enc = codec.encode(ts, vals, start, unit=O.UNIT_S)
    packed, offsets = codec.compact(enc, align=64)
    s0 = int(start[0].item())
    n_win = 288
    r = codec.decode_downsample(packed, offsets, s0, 300 * SEC, n_win, want_last=True)
    plain = codec.decode_downsample(packed, offsets, s0, 300 * SEC, n_win)
    torch.cuda.synchronize()
    assert int((r.status != 0).sum()) == 0
    for a, b in ((r.sum, plain.sum), (r.count, plain.count), (r.min, plain.min), (r.max, plain.max)):
        assert torch.equal(a.view(torch.int64), b.view(torch.int64))
    # oracle: decode its own streams, then the Gauge per window
    total = int(offsets[-1].item())
    h_blob = packed[:total].cpu().numpy()
    h_off = offsets.cpu().numpy().astype(np.uint64)
    o_ts, o_vals, o_n, o_st = O.decode_batch(h_blob, h_off, P, int_opt, n_threads=8)
    assert (o_st == 0).all() and (o_n == P).all()
    g = [x.cpu().numpy() for x in (r.sum, r.count, r.min, r.max, r.last)]
    for s in range(S):
        es, ec, emn, emx, el = O.downsample_series(o_ts[s], o_vals[s], s0, 300 * SEC, n_win)
        assert (g[1][:, s] == ec).all(), s
        for got, exp in ((g[0], es), (g[2], emn), (g[3], emx), (g[4], el)):
            assert (got[:, s].view(np.uint64) == exp.view(np.uint64)).all(), s
