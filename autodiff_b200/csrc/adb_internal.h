// Internal helpers shared by the .cu translation units of libadb200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/adb200.h"

namespace adb {
int set_error(const char* fmt, ...);   // stores a thread-local message, returns 1
void count_launch(int n = 1);
int gemm_f64(const adb_gemm_desc* d, cudaStream_t st, int force_path);
}  // namespace adb

#define ADB_CUDA_OK(expr)                                                                 \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) return adb::set_error("%s failed: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)
