#!/bin/bash
# dev tool: build hx_fused variants (macros) into build_variants/ ; run with "run" on the GPU box
set -e
cd "$(dirname "$0")/.."
SRC=sqcircuit_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Iinclude -I$SRC"
if [ "$1" = "build" ]; then
  shift
  for obj in dense small ops lowblock; do nvcc $FLAGS -c $SRC/$obj.cu -o build_variants/$obj.o & done; wait
  i=0
  for v in "$@"; do
    nvcc $FLAGS $v -Xptxas -v -c $SRC/hx_fused.cu -o build_variants/hx_$i.o 2> build_variants/hx_$i.log
    nvcc -shared -o build_variants/lib_$i.so build_variants/{dense,small,ops,lowblock}.o build_variants/hx_$i.o -lcudart
    echo "$i: $v  spills(KS=13..): $(grep -c 'spill' build_variants/hx_$i.log)" | tee -a build_variants/index.txt
    i=$((i+1))
  done
else
  for f in build_variants/lib_*.so; do
    cp $f $SRC/libsqcircuit_b200.so
    echo "== $f"; python tools/microbench.py 2048 hx
  done
fi
