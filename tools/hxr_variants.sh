#!/bin/bash
# dev tool: build variants of the real H.X kernel ("build", here) and time them with the microbenchmark ("run", on
# the GPU box): each variant library is swapped in for the product one for the duration of its run.
set -e
cd "$(dirname "$0")/.."
SRC=sqcircuit_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
VARIANTS=("base:" "lag4:-DHXR_LAG_SEP=4" "lag1:-DHXR_LAG_SEP=1" "nopipe:-DHXR_PIPE=0")
if [ "$1" = "build" ]; then
  mkdir -p build_variants
  for v in "${VARIANTS[@]}"; do
    tag=${v%%:*}; defs=${v#*:}
    nvcc $FLAGS $defs -Xptxas -v -c $SRC/hx_real.cu -o build_variants/hx_real_$tag.o 2>&1 | grep -A2 "hx_real_kernelILi13ELb1" | grep "spill" | sed "s/^/$tag: /"
    objs=""
    for f in $(ls $SRC/_obj/*.o | grep -v hx_real.o); do objs="$objs $f"; done
    nvcc --shared -gencode arch=compute_100a,code=sm_100a -o build_variants/lib_hxr_$tag.so $objs build_variants/hx_real_$tag.o
  done
else
  cp $SRC/libsqcircuit_b200.so /tmp/lib_keep.so
  for v in "${VARIANTS[@]}"; do
    tag=${v%%:*}
    cp build_variants/lib_hxr_$tag.so $SRC/libsqcircuit_b200.so
    echo "== $tag"
    python tools/microbench.py ${2:-1024} hxr 2>&1 | grep "separable" || true
  done
  cp /tmp/lib_keep.so $SRC/libsqcircuit_b200.so
fi
