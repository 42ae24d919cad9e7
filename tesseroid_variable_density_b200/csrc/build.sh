#!/bin/bash
# Builds libtvd.so in-tree for sm_100a (cross-compiles without a GPU).
#   -fmad=false : no implicit FMA contraction (see tvd_internal.cuh); explicit fma() is kept.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -Xptxas -v"
mkdir -p _obj
pids=()
for f in tvd_preprocess tvd_far tvd_near tvd_sort tvd_abi; do
  stale=0
  for dep in $f.cu tvd_internal.cuh tvd_ddmath.cuh tvd_crtrig.cuh tvd_sincos_table.inc ../../include/tvd.h build.sh; do
    if [ ! -f _obj/$f.o ] || [ $dep -nt _obj/$f.o ]; then stale=1; fi
  done
  if [ $stale = 1 ]; then
    ( $NVCC $FLAGS -c $f.cu -o _obj/$f.o > _obj/$f.log 2>&1 || { cat _obj/$f.log; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o libtvd.so _obj/tvd_preprocess.o _obj/tvd_far.o _obj/tvd_near.o _obj/tvd_sort.o _obj/tvd_abi.o -lcudart
echo "built $(pwd)/libtvd.so"
