mkdir -p gpurun_out
(timeout 300 python scripts/r2_quick.py 1000000 1 2>&1 | grep -E "encode|decode|dec\+ds") > gpurun_out/r2u_quick_base.log 2>&1
(M3TSZ_B200_LIB=$PWD/m3_b200/variants/r64all.so timeout 300 python scripts/r2_quick.py 1000000 1 2>&1 | grep -E "decode|dec\+ds") > gpurun_out/r2u_quick_r64all.log 2>&1
(timeout 300 python scripts/r2_enc_pm.py) > gpurun_out/r2u_encpm_base.log 2>&1
for v in encw16 encw32 encb5; do (M3TSZ_B200_LIB=$PWD/m3_b200/variants/$v.so timeout 300 python scripts/r2_enc_pm.py) > gpurun_out/r2u_encpm_$v.log 2>&1; done
for f in gpurun_out/r2u_*.log; do echo == $f; tail -8 $f; done
