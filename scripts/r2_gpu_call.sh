mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -40) > gpurun_out/r2b_tests.log 2>&1
python scripts/r2_debug.py > gpurun_out/r2b_debug.log 2>&1
export M3TSZ_B200_LIB=$PWD/m3_b200/variants/dec4.so
for ns in 0 250 500 800 1200; do M3TSZ_ENC_STAGGER_NS_PER_DP=$ns timeout 200 python scripts/r2_stagger.py 1000000; done > gpurun_out/r2b_stagger.log 2>&1
(timeout 300 python scripts/r2_quick.py 1000000 1) > gpurun_out/r2b_quick_dec4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 1 -c 1 -o gpurun_out/r2b_decode_full python scripts/prof_decode.py 1000000 > gpurun_out/r2b_prof.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 3 -c 1 -o gpurun_out/r2b_ds_full python scripts/prof_decode.py 1000000 >> gpurun_out/r2b_prof.log 2>&1
M3TSZ_B200_LIB=$PWD/m3_b200/variants/ds5.so timeout 300 python scripts/r2_quick.py 1000000 1 > gpurun_out/r2b_quick_ds5.log 2>&1
tail -25 gpurun_out/r2b_tests.log; cat gpurun_out/r2b_debug.log gpurun_out/r2b_stagger.log gpurun_out/r2b_quick_*.log; tail -3 gpurun_out/r2b_prof.log
