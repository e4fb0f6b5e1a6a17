mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -20) > gpurun_out/r2i_tests.log 2>&1
(timeout 1500 python bench.py --steps 20 --warmup 5) > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2i_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r2i_bench_under_ncu.log 2>&1
python -m m3_b200.tools.read_data_files -p /tmp/m3fs -n bench -s 0 --generate 100000 -B datapoints > gpurun_out/r2i_read_data_files.log 2>&1
python -m m3_b200.tools.read_data_files -p /tmp/m3fs -n bench -s 0 -b 1599955200000000000 -B datapoints >> gpurun_out/r2i_read_data_files.log 2>&1
tail -8 gpurun_out/r2i_tests.log; tail -c 5000 gpurun_out/r2i_bench.json; tail -5 gpurun_out/r2i_bench.err; cat gpurun_out/r2i_read_data_files.log
