"""Timing of the point-major encode input stage against the series-major one (1 M x 1440)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from m3_b200 import synth
from m3_b200.codec import BatchCodec
S, P = 1_000_000, 1440
codec = BatchCodec(0, True)
ts, vals, start = synth.gaussian_walk(S, P, "cuda", seed=1)
a, b = ts.t().contiguous(), vals.t().contiguous()
stride = ((64 + 9 * P) + 63) // 64 * 64
o1 = codec.encode(ts, vals, start, unit=1, out_stride=stride)
o2 = codec.encode(a, b, start, unit=1, out_stride=stride, point_major=True)
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("encode series-major %.3f ms, point-major %.3f ms, same=%s" % (
    t(lambda: codec.encode(ts, vals, start, unit=1, out=o1)),
    t(lambda: codec.encode(a, b, start, unit=1, out=o2, point_major=True)),
    torch.equal(o1.out_len, o2.out_len) and torch.equal(o1.out, o2.out)))
