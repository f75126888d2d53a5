#!/bin/bash
# A/B builds of the FP64 GEMM ring: tools/gemm_selftest_<tag> and tools/gemm_stress_<tag> with other stage / prefetch counts.
# usage: tools/gemm_variants.sh "<tag> <nvcc -D flags>" ...
set -e
cd "$(dirname "$0")"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17"
OTHERS=$(ls ../autodiff_b200/build/*.o | grep -v gemm_f64.o)
for spec in "$@"; do
  set -- $spec; tag=$1; shift
  nvcc $FLAGS "$@" -c ../autodiff_b200/csrc/gemm_f64.cu -o /tmp/gemm_f64_$tag.o
  nvcc $FLAGS -o gemm_selftest_$tag gemm_selftest.cu /tmp/gemm_f64_$tag.o $OTHERS
  nvcc $FLAGS -o gemm_stress_$tag gemm_stress.cu /tmp/gemm_f64_$tag.o $OTHERS
  echo built $tag
done
