mkdir -p gpurun_out
(timeout 300 python scripts/r2_enc_pm.py) > gpurun_out/r2h_encpm_unroll.log 2>&1
(M3TSZ_B200_LIB=$PWD/m3_b200/variants/nounroll.so timeout 300 python scripts/r2_enc_pm.py) > gpurun_out/r2h_encpm_nounroll.log 2>&1
(timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity_at_size.py -q -m gpu -k "point_major or packed or at_size or encode" 2>&1 | tail -3) > gpurun_out/r2h_tests.log 2>&1
tail -1 gpurun_out/r2h_encpm_unroll.log; tail -1 gpurun_out/r2h_encpm_nounroll.log; tail -2 gpurun_out/r2h_tests.log
