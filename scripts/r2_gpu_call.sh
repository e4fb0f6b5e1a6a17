mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -60) > gpurun_out/r2d_tests.log 2>&1
(timeout 300 python scripts/r2_quick.py 1000000 1) > gpurun_out/r2d_quick_base.log 2>&1
for v in cg pf cgpf ds5; do (M3TSZ_B200_LIB=$PWD/m3_b200/variants/$v.so timeout 300 python scripts/r2_quick.py 1000000 1) > gpurun_out/r2d_quick_$v.log 2>&1; done
tail -30 gpurun_out/r2d_tests.log; for f in gpurun_out/r2d_quick_*.log; do echo "== $f"; cat $f; done
