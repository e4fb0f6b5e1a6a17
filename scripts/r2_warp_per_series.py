"""Times the warp-per-series decode PROTOTYPE (scripts/proto/decode_warp_per_series.cu) next to the
product's lane-per-series kernel on the same float-mode streams, and checks its output against it.
Run under gpurun; with `ncu` around it this is the capture committed as
profiles/r02_warp_per_series_proto.ncu_summary.txt."""
import ctypes as C
import os
import subprocess
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from m3_b200 import synth
from m3_b200.codec import BatchCodec

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "..", "m3_b200", "variants", "proto_wps.so")
if not os.path.exists(SO):
    subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
                           "-Xcompiler", "-fPIC", "-rdc=true", "-shared", "-o", SO,
                           os.path.join(HERE, "proto", "decode_warp_per_series.cu"), os.path.join(HERE, "proto", "launch.cu")])
lib = C.CDLL(SO)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
P = 1440
codec = BatchCodec(0, False)
ts, vals, start = synth.gaussian_walk(S, P, "cuda", seed=1)
enc = codec.encode(ts, vals, start, unit=1)
packed, offsets = codec.compact(enc, align=64)
del enc
dec = codec.decode(packed, offsets, P)
o_t = torch.empty((S, P), dtype=torch.int64, device="cuda")
o_v = torch.empty((S, P), dtype=torch.float64, device="cuda")
lib.launch_decode_warp_per_series.argtypes = [C.c_void_p] * 2 + [C.c_uint64, C.c_uint32] + [C.c_void_p] * 3


def proto():
    lib.launch_decode_warp_per_series(packed.data_ptr(), offsets.data_ptr(), S, P, o_t.data_ptr(), o_v.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream)


def timeit(fn, n=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


t_proto = timeit(proto)
ok = torch.equal(o_t, dec.ts) and torch.equal(o_v.view(torch.int64), dec.values.view(torch.int64))
t_lane = timeit(lambda: codec.decode(packed, offsets, P, out=dec))
dp = S * P
print(f"S={S} float mode  warp-per-series prototype: {t_proto:.3f} ms = {dp/t_proto/1e6:.1f} G dp/s (output == product: {ok})")
print(f"               lane-per-series product   : {t_lane:.3f} ms = {dp/t_lane/1e6:.1f} G dp/s  ({t_proto/t_lane:.1f}x)")
