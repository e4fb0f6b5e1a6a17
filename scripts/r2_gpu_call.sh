mkdir -p gpurun_out
(timeout 300 python scripts/r2_enc_pm.py) > gpurun_out/r2w_encpm.log 2>&1
(timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -4) > gpurun_out/r2w_tests.log 2>&1
(timeout 300 python scripts/r2_quick.py 1000000 1 2>&1 | grep -E "encode|decode|dec\+ds") > gpurun_out/r2w_quick.log 2>&1
tail -2 gpurun_out/r2w_encpm.log; tail -3 gpurun_out/r2w_tests.log; cat gpurun_out/r2w_quick.log
