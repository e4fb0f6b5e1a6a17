mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -8) > gpurun_out/r2v_tests.log 2>&1
(timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3) > gpurun_out/r2v_bench_2gpu.json 2> gpurun_out/r2v_bench_2gpu.err
tail -5 gpurun_out/r2v_tests.log; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2v_bench_2gpu.json").read().strip().splitlines()[-1])
print(json.dumps(d.get("fetch_allgather"), indent=1)); print(d["value"], d["ms_per_step"], d["e2e"]["value"])
PY
tail -5 gpurun_out/r2v_bench_2gpu.err
