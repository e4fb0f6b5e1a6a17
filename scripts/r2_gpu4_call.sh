mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2y_gpus.txt
(timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 5 --warmup 3) > gpurun_out/r2y_bench_4gpu.json 2> gpurun_out/r2y_bench_4gpu.err
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 4 --steps 1 --warmup 0) > gpurun_out/r2y_ref_4gpu.json 2> gpurun_out/r2y_ref_4gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2y_bench_4gpu.json").read().strip().splitlines()[-1])
print(json.dumps(d.get("fetch_allgather"), indent=1)); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["series_per_step"], d["e2e_fetch"]["input_dps"], d["cpu_baseline"]["cores"])
PY
tail -3 gpurun_out/r2y_bench_4gpu.err; tail -c 600 gpurun_out/r2y_ref_4gpu.json
