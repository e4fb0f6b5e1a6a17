"""Minimal driver for ncu: one encode, one compact, a few decodes (100k x 1440)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from m3_b200 import synth
from m3_b200.codec import BatchCodec
S = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
P = 1440
codec = BatchCodec(0, True)
ts, vals, start = synth.gaussian_walk(S, P, "cuda", seed=1)
enc = codec.encode(ts, vals, start, unit=1)
packed, offsets = codec.compact(enc, align=64)
dec = codec.decode(packed, offsets, P)
for _ in range(2):
    codec.decode(packed, offsets, P, out=dec)
    codec.encode(ts, vals, start, unit=1, out=enc)
ds = codec.decode_downsample(packed, offsets, int(start[0].item()), 300 * 10**9, 288)
torch.cuda.synchronize()
print("done")
