# ncu captures of the side kernels (packed encode, tile re-encode, point-major merge, Prometheus epilogue)
mkdir -p gpurun_out
cap() { # tag kernel-regex script args...
  t=$1; k=$2; shift 2
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/r2p_$t "$@" > gpurun_out/r2p_prof_$t.log 2>&1
}
cap enc_packed encode_kernel python scripts/prof_r2.py enc_packed
cap tiles_enc encode_kernel python scripts/prof_r2.py tiles
cap merge_pm merge_fast_kernel python scripts/prof_side.py merge_pm
cap prom prom_simple_kernel python scripts/prof_side.py prom
cap prom_reset prom_general_kernel python scripts/prof_side.py prom_reset
ls -la gpurun_out/r2p_*.ncu-rep; tail -2 gpurun_out/r2p_prof_*.log
