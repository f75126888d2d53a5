#!/bin/bash
# Builds the standalone GEMM self-test against the SAME objects libadb200.so is linked from (autodiff_b200/build/*.o,
# produced by autodiff_b200/build.py), so a new translation unit can never be missing here.
set -e
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o gemm_selftest gemm_selftest.cu ../autodiff_b200/build/*.o
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o conv_selftest conv_selftest.cu ../autodiff_b200/build/*.o
