# Round-2 last GPU call: (1) point-major encode with the cp.async fills (default build) against the TMA tensor-copy
# input stage (variants tma4 / tma8: -DM3_ENC_BULK_PM=2, 4 / 8 rows per tile);
# (2) the full GPU suite with the fastest correct library.  Budget: ~8 minutes of box time.
R=${1:-r3a}
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo $(( $(date +%s) - T0 )); }
(timeout -s KILL 200 python -c "import torch; torch.zeros(1).cuda(); print(torch.cuda.get_device_name(0))") 2>&1 | tail -1; echo "[t=$(el)s]"
for v in default tma8 tma4; do
  L=$PWD/m3_b200/variants/$v.so; [ $v = default ] && L=$PWD/m3_b200/libm3tsz_b200.so
  (M3TSZ_B200_LIB=$L timeout -s KILL 80 python scripts/r3_enc_variant.py) > gpurun_out/${R}_enc_$v.log 2>&1
  echo "$v: $(tail -1 gpurun_out/${R}_enc_$v.log | cut -c1-300) [t=$(el)s]"
done
BEST=$(python - <<PY
import json
best, bms = "default", None
for v in ("default", "tma8", "tma4", "bulk4"):
    try:
        d = json.loads(open("gpurun_out/${R}_enc_%s.log" % v).read().strip().splitlines()[-1])
    except Exception:
        continue
    if not d.get("same_as_series_major"):
        continue
    # a variant must win by more than noise to replace the default
    ms = d["encode_pm_ms"] * (1.0 if v == "default" else 1.01)
    if bms is None or ms < bms:
        best, bms = v, ms
print(best)
PY
)
echo "best=$BEST [t=$(el)s]"
L=$PWD/m3_b200/variants/$BEST.so; [ $BEST = default ] && L=$PWD/m3_b200/libm3tsz_b200.so
echo $BEST > gpurun_out/${R}_best.txt
LEFT=$(( 505 - $(el) )); [ $LEFT -lt 60 ] && LEFT=60
(M3TSZ_B200_LIB=$L timeout -s KILL $LEFT python -m pytest tests -q -m gpu -x --durations=8 2>&1 | tail -25) > gpurun_out/${R}_tests_$BEST.log 2>&1
tail -4 gpurun_out/${R}_tests_$BEST.log; echo "[t=$(el)s]"
if [ $BEST != default ] && [ $(el) -lt 420 ]; then
  (timeout -s KILL 90 python -m pytest tests/test_gpu_round2.py -q -m gpu -x -k "point_major or packed" 2>&1 | tail -5) > gpurun_out/${R}_tests_default_pm.log 2>&1
  tail -2 gpurun_out/${R}_tests_default_pm.log; echo "[t=$(el)s]"
fi
if [ $(el) -lt 360 ]; then
  (M3TSZ_B200_LIB=$L timeout -s KILL 150 python bench.py --steps 10 --warmup 3 --no-extras) > gpurun_out/${R}_bench_$BEST.json 2> gpurun_out/${R}_bench_$BEST.err
  tail -c 600 gpurun_out/${R}_bench_$BEST.json; echo "[t=$(el)s]"
fi
