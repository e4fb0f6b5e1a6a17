mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -12) > gpurun_out/r2f_tests.log 2>&1
(timeout 900 python bench.py --steps 20 --warmup 5) > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
(timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')") > gpurun_out/r2f_smoke.log 2>&1
(timeout 600 python bench.py --impl reference --steps 1 --warmup 0) > gpurun_out/r2f_ref.json 2> gpurun_out/r2f_ref.err
tail -4 gpurun_out/r2f_tests.log; tail -c 600 gpurun_out/r2f_bench.json; tail -3 gpurun_out/r2f_bench.err; tail -1 gpurun_out/r2f_smoke.log; tail -c 400 gpurun_out/r2f_ref.json
