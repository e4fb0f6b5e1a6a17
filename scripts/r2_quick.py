"""Round-2 kernel timings (not the bench contract; see bench.py): slots vs packed encode,
CSR vs (offset, length) decode, fused downsample without / with `last`."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from m3_b200 import synth
from m3_b200.codec import BatchCodec

S = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
modes = [True, False] if len(sys.argv) <= 2 else [sys.argv[2] == "1"]
P = 1440


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for int_opt in modes:
    codec = BatchCodec(0, int_opt)
    ts, vals, start = synth.gaussian_walk(S, P, "cuda", seed=1)
    dp = S * P
    enc = codec.encode(ts, vals, start, unit=1)
    torch.cuda.synchronize()
    t_enc = timeit(lambda: codec.encode(ts, vals, start, unit=1, out=enc))
    packed, offsets = codec.compact(enc, align=64)
    t_cmp = timeit(lambda: codec.compact(enc, align=64, capacity=packed.numel()), n=3)
    total = int(offsets[-1].item())
    bc = total / dp
    pk = codec.encode_packed(ts, vals, start, unit=1, align=64, capacity=total + 4096)
    torch.cuda.synchronize()
    ok_pk = bool((pk.status == 0).all()) and torch.equal(pk.out_len, enc.out_len)
    t_pk = timeit(lambda: codec.encode_packed(ts, vals, start, unit=1, align=64, out=pk))
    del enc
    dec = codec.decode(packed, offsets, P)
    t_dec = timeit(lambda: codec.decode(packed, offsets, P, out=dec))
    ok = torch.equal(dec.ts, ts) and torch.equal(dec.values.view(torch.int64), vals.view(torch.int64))
    dec2 = codec.decode(pk.packed, pk.offsets, P, lengths=pk.out_len, out=dec)
    t_dec2 = timeit(lambda: codec.decode(pk.packed, pk.offsets, P, lengths=pk.out_len, out=dec))
    ok2 = torch.equal(dec2.ts, ts) and torch.equal(dec2.values.view(torch.int64), vals.view(torch.int64))
    del dec2
    pm = codec.decode(packed, offsets, P, point_major=True)
    t_pm = timeit(lambda: codec.decode(packed, offsets, P, point_major=True, out=pm))
    ok_pm = torch.equal(pm.ts.t(), dec.ts) and torch.equal(pm.values.t().contiguous().view(torch.int64), dec.values.view(torch.int64))
    del dec, pm
    s0 = int(start[0].item())
    ds = codec.decode_downsample(packed, offsets, s0, 300 * 10**9, 288)
    t_ds = timeit(lambda: codec.decode_downsample(packed, offsets, s0, 300 * 10**9, 288, out=ds))
    dsl = codec.decode_downsample(packed, offsets, s0, 300 * 10**9, 288, want_last=True)
    t_dsl = timeit(lambda: codec.decode_downsample(packed, offsets, s0, 300 * 10**9, 288, out=dsl, want_last=True))
    ok3 = torch.equal(ds.sum, dsl.sum) and torch.equal(ds.count, dsl.count) and bool((ds.count == 5).all())
    last_ok = torch.equal(dsl.last, vals[:, 4::5].t().contiguous())
    print(f"int_opt={int_opt} S={S} B/dp={bc:.3f} roundtrip_ok={ok} packed_ok={ok_pk and ok2} ds_ok={ok3} last_ok={last_ok}")
    print(f"  encode        {t_enc:7.3f} ms  {dp/t_enc/1e6:6.1f} Gdp/s  {(16+bc)*dp/t_enc/1e6:5.0f} GB/s")
    print(f"  compact       {t_cmp:7.3f} ms")
    print(f"  encode_packed {t_pk:7.3f} ms  {dp/t_pk/1e6:6.1f} Gdp/s  {(16+bc)*dp/t_pk/1e6:5.0f} GB/s")
    print(f"  decode        {t_dec:7.3f} ms  {dp/t_dec/1e6:6.1f} Gdp/s  {(16+bc)*dp/t_dec/1e6:5.0f} GB/s")
    print(f"  decode(off,len){t_dec2:6.3f} ms  {dp/t_dec2/1e6:6.1f} Gdp/s")
    print(f"  decode(point-major){t_pm:6.3f} ms  {dp/t_pm/1e6:6.1f} Gdp/s  {(16+bc)*dp/t_pm/1e6:5.0f} GB/s  same_as_series_major={ok_pm}")
    print(f"  dec+ds        {t_ds:7.3f} ms  {dp/t_ds/1e6:6.1f} Gdp/s  {(6.4+bc)*dp/t_ds/1e6:5.0f} GB/s")
    print(f"  dec+ds+last   {t_dsl:7.3f} ms  {dp/t_dsl/1e6:6.1f} Gdp/s  {(9.6+bc)*dp/t_dsl/1e6:5.0f} GB/s")
    del ds, dsl, pk, packed
    torch.cuda.empty_cache()
