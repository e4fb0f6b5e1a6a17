import os
"""Rough first timing of the kernels (not the bench contract; see bench.py)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from m3_b200 import synth
from m3_b200.codec import BatchCodec

S = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
P = 1440
for int_opt in (True, False):
    codec = BatchCodec(0, int_opt)
    ts, vals, start = synth.gaussian_walk(S, P, "cuda", seed=1)
    enc = codec.encode(ts, vals, start, unit=1)
    torch.cuda.synchronize()
    def timeit(fn, n=5):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    t_enc = timeit(lambda: codec.encode(ts, vals, start, unit=1, out=enc))
    packed, offsets = codec.compact(enc, align=int(os.environ.get('M3_ALIGN', '64')))
    total = int(offsets[-1].item())
    dec = codec.decode(packed, offsets, P)
    t_dec = timeit(lambda: codec.decode(packed, offsets, P, out=dec))
    ds = codec.decode_downsample(packed, offsets, int(start[0].item()), 300 * 10**9, 288)
    t_ds = timeit(lambda: codec.decode_downsample(packed, offsets, int(start[0].item()), 300 * 10**9, 288, out=ds))
    ok = torch.equal(dec.ts, ts) and torch.equal(dec.values.view(torch.int64), vals.view(torch.int64))
    dp = S * P
    bc = total / dp
    print(f"int_opt={int_opt} S={S} B/dp={bc:.3f} roundtrip_ok={ok}")
    print(f"  encode {t_enc:.3f} ms  {dp/t_enc/1e6:.1f} Gdp/s  {(16+bc)*dp/t_enc/1e6:.0f} GB/s")
    print(f"  decode {t_dec:.3f} ms  {dp/t_dec/1e6:.1f} Gdp/s  {(16+bc)*dp/t_dec/1e6:.0f} GB/s")
    print(f"  dec+ds {t_ds:.3f} ms  {dp/t_ds/1e6:.1f} Gdp/s  {(6.4+bc)*dp/t_ds/1e6:.0f} GB/s")
