mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2j_gpus.txt
(timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_round2.py tests/test_gpu_merge.py -q -m gpu -k "allgather or point_major or merge" 2>&1 | tail -15) > gpurun_out/r2j_tests.log 2>&1
(timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3) > gpurun_out/r2j_bench_2gpu.json 2> gpurun_out/r2j_bench_2gpu.err
(timeout 300 python scripts/prof_r2.py enc_pm 2>&1; timeout 200 python - <<'PY'
import sys, os, torch
sys.path.insert(0, os.getcwd())
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
    torch.equal(o1.out_len, o2.out_len)))
PY
) > gpurun_out/r2j_enc_pm.log 2>&1
tail -8 gpurun_out/r2j_tests.log; tail -3 gpurun_out/r2j_enc_pm.log; tail -c 2500 gpurun_out/r2j_bench_2gpu.json; tail -5 gpurun_out/r2j_bench_2gpu.err
