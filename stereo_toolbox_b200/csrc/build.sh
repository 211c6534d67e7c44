#!/bin/bash
# Build libstb200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall"
OUT=../libstb200.so
mkdir -p build
objs=""
pids=""
for f in *.cu; do
  o=build/${f%.cu}.o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ common.cuh -nt "$o" ] || [ ../../include/stb200.h -nt "$o" ] || \
     { [ -f umma.cuh ] && [ umma.cuh -nt "$o" ]; }; then
    $NVCC $FLAGS ${NVCC_EXTRA} -c "$f" -o "$o" &
    pids="$pids $!"
  fi
  objs="$objs $o"
done
for p in $pids; do wait $p; done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o $OUT $objs
echo "built $(realpath $OUT)"
