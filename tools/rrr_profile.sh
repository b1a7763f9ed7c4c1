#!/bin/bash
# dev tool: phase cycle counts of sqc_gen_rr_real.  "build" here (CPU box), "run" on the GPU box: swaps a profiling
# build of small_real.cu into a scratch copy of the library and runs the microbenchmark (B = 1 problem wave).
set -e
cd "$(dirname "$0")/.."
SRC=sqcircuit_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
if [ "$1" = "build" ]; then
  mkdir -p build_variants
  nvcc $FLAGS -DRRR_PROFILE -c $SRC/small_real.cu -o build_variants/small_real_prof.o
  objs=""
  for f in dense small ops hx_fused hx_real hx_generic realrep lowblock buildinfo; do objs="$objs $SRC/_obj/$f.o"; done
  nvcc --shared -gencode arch=compute_100a,code=sm_100a -o build_variants/lib_rrr_prof.so $objs build_variants/small_real_prof.o
else
  cp $SRC/libsqcircuit_b200.so /tmp/lib_keep.so
  cp build_variants/lib_rrr_prof.so $SRC/libsqcircuit_b200.so
  python tools/microbench.py 64 rrr1 || true
  cp /tmp/lib_keep.so $SRC/libsqcircuit_b200.so
fi
