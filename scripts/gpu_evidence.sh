#!/bin/bash
# Run on the GPU box (under gpurun): evidence for profiles/ (never a bench value).
#   1. launch list of the bench command (ncu, cold-cache, serialised: compare SHARES)
#   2. ncu --set full capture of the decode and encode kernels at the bench workload
R=${1:-r01}
S=${2:-1000000}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${R}_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras --series $S \
    > gpurun_out/${R}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 1 -c 1 \
    -o gpurun_out/${R}_decode_full python scripts/prof_decode.py $S > gpurun_out/${R}_prof.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:encode_kernel -s 1 -c 1 \
    -o gpurun_out/${R}_encode_full python scripts/prof_decode.py $S >> gpurun_out/${R}_prof.log 2>&1
tail -2 gpurun_out/${R}_prof.log
