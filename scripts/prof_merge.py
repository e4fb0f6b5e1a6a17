"""ncu target: series merge (row N1) of 100k series x 3 replicas x 1440 points (RF=3 fetch shape).
ncu -k regex:merge_fast_kernel -s 2 -c 1 ..."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from m3_b200 import synth
from m3_b200.codec import BatchCodec
Sm, P = 300_000, 1440
codec = BatchCodec(0, True)
ts, vals, start = synth.gaussian_walk(Sm, P, "cuda", seed=1)
n = torch.full((Sm,), P, dtype=torch.int32, device="cuda")
st = torch.zeros(Sm, dtype=torch.int32, device="cuda")
ar = torch.arange(Sm + 1, dtype=torch.int64, device="cuda")
ser = torch.arange(0, Sm + 1, 3, dtype=torch.int64, device="cuda")
rep_ts = ts[::3].repeat_interleave(3, dim=0).contiguous()  # replicas share timestamps
for _ in range(3):
    out = codec.merge_series(rep_ts, vals, n, st, ar, ar, ser, P)
torch.cuda.synchronize()
print("done", int(out[2].sum().item()))
