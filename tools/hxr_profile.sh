#!/bin/bash
# dev tool: phase cycle counts of the real H.X kernel.  "build" here (CPU box), "run" on the GPU box:
# swaps the profiling build of hx_real.cu into a scratch copy of the library and runs the microbenchmark.
set -e
cd "$(dirname "$0")/.."
SRC=sqcircuit_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
if [ "$1" = "build" ]; then
  mkdir -p build_variants
  nvcc $FLAGS -DHXR_PROFILE -c $SRC/hx_real.cu -o build_variants/hx_real_prof.o
  objs=""
  for f in dense small small_real ops hx_fused realrep lowblock; do objs="$objs $SRC/_obj/$f.o"; done
  nvcc --shared -gencode arch=compute_100a,code=sm_100a -o build_variants/lib_hxr_prof.so $objs build_variants/hx_real_prof.o
else
  cp $SRC/libsqcircuit_b200.so /tmp/lib_keep.so
  cp build_variants/lib_hxr_prof.so $SRC/libsqcircuit_b200.so
  python tools/microbench.py ${2:-1024} hxr || true
  cp /tmp/lib_keep.so $SRC/libsqcircuit_b200.so
fi
