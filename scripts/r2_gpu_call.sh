mkdir -p gpurun_out
(timeout 300 python scripts/r2_quick.py 1000000 1) > gpurun_out/r2m_quick_base.log 2>&1
for v in dsm0 dsb4; do (M3TSZ_B200_LIB=$PWD/m3_b200/variants/$v.so timeout 300 python scripts/r2_quick.py 1000000 1) > gpurun_out/r2m_quick_$v.log 2>&1; done
(timeout 300 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity_at_size.py -q -m gpu -k "downsample or tiles or gauge" 2>&1 | tail -4) > gpurun_out/r2m_tests.log 2>&1
for f in gpurun_out/r2m_quick_*.log; do echo "== $f"; grep -E "dec\+ds" $f; done; tail -3 gpurun_out/r2m_tests.log
