"""Timing of m3tsz_aggregate_tiles_batch at 1 M x 1440 -> 288 five-minute tiles (agg = last)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from m3_b200 import synth
from m3_b200.codec import BatchCodec
S, P = 1_000_000, 1440
codec = BatchCodec(0, True)
ts, vals, start = synth.gaussian_walk(S, P, "cuda", seed=1)
s0 = int(start[0].item())
enc = codec.encode(ts, vals, start, unit=1, out_stride=((64 + 9 * P) + 63) // 64 * 64)
del ts, vals
packed, offsets = codec.compact(enc, align=64)
del enc
r = codec.aggregate_tiles(packed, offsets, s0, 300 * 10**9, 288)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    codec.aggregate_tiles(packed, offsets, s0, 300 * 10**9, 288, out=r[0])
e1.record(); torch.cuda.synchronize()
print("aggregate_tiles %.3f ms" % (e0.elapsed_time(e1) / 5))
