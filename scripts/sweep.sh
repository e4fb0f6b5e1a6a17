#!/bin/bash
# runs quick_gpu.py for every tuning build under m3_b200/variants
for f in m3_b200/variants/*.so; do echo "== $f"; M3TSZ_B200_LIB=$PWD/$f python scripts/quick_gpu.py ${1:-100000} 2>&1 | grep -E "encode|decode|dec\+ds" ; done
