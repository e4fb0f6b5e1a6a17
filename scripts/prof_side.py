"""ncu targets for the side kernels (three launches each; capture the third):
  python scripts/prof_side.py merge_pm   # merge_fast_kernel, point-major in/out, 100k series x RF 3 x 1440
  python scripts/prof_side.py prom       # prom_simple_kernel over 300k x 1440
  python scripts/prof_side.py prom_reset # prom_general_kernel (counter-reset normalisation)"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from m3_b200 import synth
from m3_b200.codec import BatchCodec
what = sys.argv[1]
P = 1440
codec = BatchCodec(0, True)
if what == "merge_pm":
    Sm = 300_000
    ts, vals, start = synth.gaussian_walk(Sm, P, "cuda", seed=1)
    n = torch.full((Sm,), P, dtype=torch.int32, device="cuda")
    st = torch.zeros(Sm, dtype=torch.int32, device="cuda")
    ar = torch.arange(Sm + 1, dtype=torch.int64, device="cuda")
    ser = torch.arange(0, Sm + 1, 3, dtype=torch.int64, device="cuda")
    rep_ts = ts[::3].repeat_interleave(3, dim=0)  # replicas share timestamps
    pm_ts, pm_v = rep_ts.t().contiguous(), vals.t().contiguous()
    del ts, vals, rep_ts
    for _ in range(3):
        out = codec.merge_series(pm_ts, pm_v, n, st, ar, ar, ser, P, point_major=True)
    torch.cuda.synchronize()
    print("done", int(out[2].sum().item()))
else:
    S = 300_000
    ts, vals, start = synth.gaussian_walk(S, P, "cuda", seed=1)
    n = torch.full((S,), P, dtype=torch.int32, device="cuda")
    hr = torch.ones(S, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        if what == "prom":
            out = codec.prom_convert(ts, vals, n)
        else:
            out = codec.prom_convert(ts, vals, n, 300 * 10**9, hr)
    torch.cuda.synchronize()
    print("done", what)
