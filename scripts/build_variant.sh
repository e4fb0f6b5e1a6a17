#!/bin/bash
# builds a tuning variant of the library: scripts/build_variant.sh NAME "-DFLAG=.. -DFLAG2=.."
set -e
NAME=$1; FLAGS=$2
cd "$(dirname "$0")/../m3_b200/csrc"
D=/tmp/m3var_$NAME; rm -rf $D; mkdir -p $D
ARCH="-gencode arch=compute_100a,code=sm_100a"
for f in m3tsz_decode m3tsz_encode m3tsz_merge m3tsz_checksum m3tsz_query m3tsz_stream m3tsz_collective m3tsz_capi; do
  if [ -f $f.cu ]; then /usr/local/cuda/bin/nvcc -O3 -std=c++17 $ARCH -lineinfo -fmad=false -Xcompiler -fPIC $FLAGS -c $f.cu -o $D/$f.o & fi
done
wait
/usr/local/cuda/bin/nvcc $ARCH -shared -o ../variants/$NAME.so $D/*.o -lcudart -ldl
echo built ../variants/$NAME.so
