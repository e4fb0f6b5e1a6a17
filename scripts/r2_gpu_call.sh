mkdir -p gpurun_out
(timeout 300 python scripts/r2_tiles_time.py) > gpurun_out/r2g_tiles4.log 2>&1
(M3TSZ_B200_LIB=$PWD/m3_b200/variants/tiles2.so timeout 300 python scripts/r2_tiles_time.py) > gpurun_out/r2g_tiles2.log 2>&1
tail -1 gpurun_out/r2g_tiles4.log; tail -1 gpurun_out/r2g_tiles2.log
