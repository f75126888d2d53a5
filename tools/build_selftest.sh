#!/bin/bash
# Builds the standalone GEMM self-test (links the same translation units as libadb200.so).
set -e
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o gemm_selftest gemm_selftest.cu \
    ../autodiff_b200/csrc/gemm_f64.cu ../autodiff_b200/csrc/gemm_tf32x3.cu ../autodiff_b200/csrc/adb_core.cu ../autodiff_b200/csrc/ewise.cu ../autodiff_b200/csrc/ewise_catalog.cu ../autodiff_b200/csrc/reduce.cu ../autodiff_b200/csrc/softmax_ce_ring.cu
