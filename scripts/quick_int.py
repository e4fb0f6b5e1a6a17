"""Throughput on non-float data families (int counters, 2-decimal gauges, repeats)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from m3_b200 import synth
from m3_b200.codec import BatchCodec
S = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
P = 1440
codec = BatchCodec(0, True)
ts, walk, start = synth.gaussian_walk(S, P, "cuda", seed=1)
jit = ts + torch.randint(0, 5, ts.shape, device="cuda", dtype=torch.int64) * 1_000_000_000
fams = {
    "gaussian float": (ts, walk),
    "int counter (round)": (ts, torch.round(walk * 10)),
    "2-decimal gauge": (ts, torch.round(walk * 100) / 100),
    "repeats (8x)": (ts, torch.round(walk[:, ::8]).repeat_interleave(8, dim=1)[:, :P].contiguous()),
    "float + jittered ts": (jit, walk),
}
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for name, (t, v) in fams.items():
    enc = codec.encode(t, v, start, unit=1)
    assert int((enc.status != 0).sum()) == 0
    te = timeit(lambda: codec.encode(t, v, start, unit=1, out=enc))
    packed, offsets = codec.compact(enc, align=16)
    dec = codec.decode(packed, offsets, P)
    td = timeit(lambda: codec.decode(packed, offsets, P, out=dec))
    ok = torch.equal(dec.ts, t) and int((dec.status != 0).sum()) == 0
    bpd = float(enc.out_len.sum().item()) / (S * P)
    print(f"{name:22s} B/dp={bpd:5.2f}  encode {S*P/te/1e6:6.1f} Gdp/s  decode {S*P/td/1e6:6.1f} Gdp/s  ts_ok={ok}")
