#!/bin/bash
# Tuning helper: builds variants of the D3Q19-BGK step kernels with different compile-time knobs and links each into
# xlb_b200/variants/libxlb_b200_<tag>.so (selected at run time with XLB_B200_LIB=...).  Not part of the product build.
set -e
cd "$(dirname "$0")"
BUILD=${BUILD:-/tmp/xlb_b200_build}; mkdir -p $BUILD ../variants
FLAGS="-std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC"
OTHERS="$BUILD/api.o $BUILD/error.o $BUILD/ops.o $BUILD/masker.o $BUILD/halo.o $BUILD/step_inst_d3q27_bgk.o $BUILD/step_inst_d3q27_kbc.o $BUILD/step_inst_d2q9_bgk.o $BUILD/step_inst_d2q9_kbc.o"
build_one() {  # tag, defines...
  tag=$1; shift
  nvcc $FLAGS "$@" -Xptxas -v -c step_inst_d3q19_bgk.cu -o $BUILD/var_$tag.o 2> $BUILD/var_$tag.log
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libxlb_b200_$tag.so $BUILD/var_$tag.o $OTHERS
  echo "$tag: $(grep -A2 'EffLi1EEE' $BUILD/var_$tag.log | grep -E 'Used|spill' | tr '\n' ' ')"
}
for spec in "$@"; do
  tag=${spec%%:*}; defs=${spec#*:}
  build_one $tag $(echo $defs | tr ',' ' ') &
done
wait
