mkdir -p gpurun_out
(timeout 300 python scripts/r2_enc_pm.py) > gpurun_out/r2x_encpm_base.log 2>&1
(M3TSZ_B200_LIB=$PWD/m3_b200/variants/encpm6.so timeout 300 python scripts/r2_enc_pm.py) > gpurun_out/r2x_encpm_6.log 2>&1
(M3TSZ_ENC_CARVEOUT_KB=228 M3TSZ_B200_LIB=$PWD/m3_b200/variants/encpm6.so timeout 300 python scripts/r2_enc_pm.py) > gpurun_out/r2x_encpm_6_co228.log 2>&1
for f in gpurun_out/r2x_*.log; do echo == $f; tail -1 $f; done
