"""ncu target for round 2: three launches of ONE kernel variant at S x 1440 (int-optimised):
  python scripts/prof_r2.py <dec_sm|dec_pm|ds|dsl|enc_sm|enc_pm|enc_packed|tiles> [S]
(capture the third launch: ncu -k regex:<decode_kernel|encode_kernel> -s 2 -c 1 ...)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from m3_b200 import synth
from m3_b200.codec import BatchCodec
what = sys.argv[1]
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
P = 1440
codec = BatchCodec(0, True)
ts, vals, start = synth.gaussian_walk(S, P, "cuda", seed=1)
s0 = int(start[0].item())
stride = ((64 + 9 * P) + 63) // 64 * 64
if what.startswith("enc"):
    if what == "enc_pm":
        a, b = ts.t().contiguous(), vals.t().contiguous()
        del ts, vals
        out = codec.encode(a, b, start, unit=1, out_stride=stride, point_major=True)
        for _ in range(2):
            codec.encode(a, b, start, unit=1, out=out, point_major=True)
    elif what == "enc_sm":
        out = codec.encode(ts, vals, start, unit=1, out_stride=stride)
        for _ in range(2):
            codec.encode(ts, vals, start, unit=1, out=out)
    else:
        out = codec.encode_packed(ts, vals, start, unit=1, align=64, capacity=S * (8 * P + 256))
        for _ in range(2):
            codec.encode_packed(ts, vals, start, unit=1, align=64, out=out)
else:
    # one encode (kernel name encode_kernel: not matched by the decode regex), then three of the target
    enc = codec.encode(ts, vals, start, unit=1, out_stride=stride)
    del ts, vals
    off = torch.arange(S, dtype=torch.int64, device="cuda") * stride
    flat = enc.out.view(-1)
    if what in ("ds", "dsl", "tiles"):
        packed, offsets = codec.compact(enc, align=64)
        del enc, flat
    for i in range(3):
        if what == "dec_sm":
            r = codec.decode(flat, off, P, lengths=enc.out_len, out=(r if i else None))
        elif what == "dec_pm":
            r = codec.decode(flat, off, P, lengths=enc.out_len, point_major=True, out=(r if i else None))
        elif what == "ds":
            r = codec.decode_downsample(packed, offsets, s0, 300 * 10**9, 288, out=(r if i else None))
        elif what == "dsl":
            r = codec.decode_downsample(packed, offsets, s0, 300 * 10**9, 288, want_last=True, out=(r if i else None))
        elif what == "tiles":
            r = codec.aggregate_tiles(packed, offsets, s0, 300 * 10**9, 288, out=(r[0] if i else None))
torch.cuda.synchronize()
print("done", what)
