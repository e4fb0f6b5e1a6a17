mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_round2.py -q -m gpu -k "point_major" 2>&1 | tail -5) > gpurun_out/r2g_tests.log 2>&1
(timeout 300 python scripts/r2_quick.py 1000000 1) > gpurun_out/r2g_quick.log 2>&1
(timeout 1500 python bench.py --steps 5 --warmup 3) > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1
tail -5 gpurun_out/r2g_tests.log; cat gpurun_out/r2g_quick.log; tail -3 gpurun_out/r2g_smoke.log; tail -c 7000 gpurun_out/r2g_bench.json; tail -8 gpurun_out/r2g_bench.err
