#!/bin/bash
# Builds libtvd.so in-tree for sm_100a (cross-compiles without a GPU).
#   -fmad=false : no implicit FMA contraction (see tvd_internal.cuh); explicit fma() is kept.
# Development knobs: TVD_EXTRA (extra nvcc flags, e.g. -DNW_INLINE_TRIG), TVD_OUT (output file),
# TVD_OBJ (object directory) build a variant next to the product library.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OBJ=${TVD_OBJ:-_obj}
OUT=${TVD_OUT:-libtvd.so}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -Xptxas -v -Wno-deprecated-gpu-targets $TVD_EXTRA"
mkdir -p $OBJ
pids=()
for f in tvd_preprocess tvd_far tvd_near tvd_sort tvd_abi; do
  stale=0
  for dep in $f.cu tvd_internal.cuh tvd_ddmath.cuh tvd_crtrig.cuh tvd_sincos_table.inc tvd_exp_table.inc ../../include/tvd.h build.sh; do
    if [ ! -f $OBJ/$f.o ] || [ $dep -nt $OBJ/$f.o ]; then stale=1; fi
  done
  if [ $stale = 1 ]; then
    ( $NVCC $FLAGS -c $f.cu -o $OBJ/$f.o > $OBJ/$f.log 2>&1 || { cat $OBJ/$f.log; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o $OUT $OBJ/tvd_preprocess.o $OBJ/tvd_far.o $OBJ/tvd_near.o $OBJ/tvd_sort.o $OBJ/tvd_abi.o -lcudart
echo "built $(cd "$(dirname "$OUT")" && pwd)/$(basename "$OUT")"
