#!/bin/bash
# dev tool: build dense.cu variants (macros) into build_variants/ ; "run <microbench args>" on the GPU box
set -e
cd "$(dirname "$0")/.."
SRC=sqcircuit_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Iinclude -I$SRC"
if [ "$1" = "build" ]; then
  shift
  for obj in hx_fused small ops lowblock; do nvcc $FLAGS -c $SRC/$obj.cu -o build_variants/$obj.o & done; wait
  i=0
  for v in "$@"; do
    nvcc $FLAGS $v -c $SRC/dense.cu -o build_variants/dense_$i.o
    nvcc -shared -o build_variants/libd_$i.so build_variants/{hx_fused,small,ops,lowblock}.o build_variants/dense_$i.o -lcudart 2>/dev/null
    echo "$i: $v"
    i=$((i+1))
  done
else
  shift
  for f in build_variants/libd_*.so; do
    cp $f $SRC/libsqcircuit_b200.so
    echo "== $f"; python tools/microbench.py "$@"
  done
fi
