mkdir -p gpurun_out
(timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_round2.py -q -m gpu -x -k "not at_size and not large and not fileset" 2>&1 | tail -25) > gpurun_out/r2s_memcheck.log 2>&1
echo "exit: $?" >> gpurun_out/r2s_memcheck.log
tail -12 gpurun_out/r2s_memcheck.log
