mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -60) > gpurun_out/r2e_tests.log 2>&1
(timeout 600 python scripts/r2_warp_per_series.py 200000) > gpurun_out/r2e_wps.log 2>&1
(timeout 900 python bench.py --steps 5 --warmup 3) > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
(timeout 600 python bench.py --impl reference --steps 2 --warmup 1) > gpurun_out/r2e_bench_ref.json 2> gpurun_out/r2e_bench_ref.err
ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 3 -c 1 -o gpurun_out/r2e_ds_full python scripts/prof_decode.py 1000000 > gpurun_out/r2e_prof.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:decode_warp_per_series -s 1 -c 1 -o gpurun_out/r2e_wps_full python scripts/r2_warp_per_series.py 200000 >> gpurun_out/r2e_prof.log 2>&1
tail -25 gpurun_out/r2e_tests.log; cat gpurun_out/r2e_wps.log; tail -c 3000 gpurun_out/r2e_bench.json; tail -5 gpurun_out/r2e_bench.err; tail -c 1500 gpurun_out/r2e_bench_ref.json; tail -5 gpurun_out/r2e_bench_ref.err
