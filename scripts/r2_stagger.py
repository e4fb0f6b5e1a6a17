"""packed-encode timing only (env M3TSZ_ENC_STAGGER_NS_PER_DP is read by the library at first use)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from m3_b200 import synth
from m3_b200.codec import BatchCodec
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
P = 1440
codec = BatchCodec(0, True)
ts, vals, start = synth.gaussian_walk(S, P, "cuda", seed=1)
pk = codec.encode_packed(ts, vals, start, unit=1, align=64, capacity=S * P * 8)
torch.cuda.synchronize()
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
t = timeit(lambda: codec.encode_packed(ts, vals, start, unit=1, align=64, out=pk))
print("stagger_ns_per_dp=%s encode_packed %.3f ms" % (os.environ.get("M3TSZ_ENC_STAGGER_NS_PER_DP", "default"), t))
