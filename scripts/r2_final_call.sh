# Round-2 evidence run on one B200 (under gpurun): full GPU suite, bench, smoke, launch list, ncu captures.
#   bash scripts/r2_final_call.sh <tag> [kernels to capture, default: dec_pm ds enc_pm dec_sm]
R=${1:-r2z}; shift
KS=${@:-dec_pm ds enc_pm dec_sm}
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -12) > gpurun_out/${R}_tests.log 2>&1
(timeout 900 python bench.py --steps 20 --warmup 5) > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
(timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')") > gpurun_out/${R}_smoke.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${R}_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras \
    > gpurun_out/${R}_bench_under_ncu.log 2>&1
for w in $KS; do
  k=decode_kernel; case $w in enc*) k=encode_kernel;; esac
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 \
      -o gpurun_out/${R}_$w python scripts/prof_r2.py $w > gpurun_out/${R}_prof_$w.log 2>&1
done
tail -5 gpurun_out/${R}_tests.log; tail -c 1500 gpurun_out/${R}_bench.json; tail -2 gpurun_out/${R}_smoke.log; ls -la gpurun_out/${R}_*.ncu-rep
