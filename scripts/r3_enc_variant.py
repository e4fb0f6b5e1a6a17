"""Point-major encode (segment and packed output) of 1 M x 1440 with the library named by M3TSZ_B200_LIB:
timing (CUDA events) + byte equality with the series-major input stage of the same library."""
import json, os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from m3_b200 import synth, capi
from m3_b200.codec import BatchCodec
S, P = int(os.environ.get("S", 1_000_000)), 1440
codec = BatchCodec(0, True)
ts, vals, start = synth.gaussian_walk(S, P, "cuda", seed=1)
a, b = ts.t().contiguous(), vals.t().contiguous()
stride = ((64 + 9 * P) + 63) // 64 * 64
o1 = codec.encode(ts, vals, start, unit=1, out_stride=stride)
del ts, vals
o2 = codec.encode(a, b, start, unit=1, out_stride=stride, point_major=True)
torch.cuda.synchronize()
same = bool(torch.equal(o1.out_len, o2.out_len) and torch.equal(o1.status, o2.status))
if same:
    flat = torch.arange(S, dtype=torch.int64, device="cuda") * stride
    c1, _ = codec.segment_checksums(o1.out.view(-1), flat, lengths=o1.out_len)
    c2, _ = codec.segment_checksums(o2.out.view(-1), flat, lengths=o2.out_len)
    torch.cuda.synchronize()
    same = bool(torch.equal(c1, c2)) and int((o2.status != 0).sum()) == 0
del o1
def t(fn, n=8):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = t(lambda: codec.encode(a, b, start, unit=1, out=o2, point_major=True))
del o2
r = codec.encode_packed(a, b, start, unit=1, point_major=True, capacity=S * stride)
ms_packed = t(lambda: codec.encode_packed(a, b, start, unit=1, point_major=True, out=r), n=4)
print(json.dumps({"lib": os.path.basename(capi.LIB_PATH), "series": S, "encode_pm_ms": round(ms, 4),
                  "encode_packed_pm_ms": round(ms_packed, 4), "same_as_series_major": same}))
