// Internal helpers shared by the .cu translation units of libadb200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/adb200.h"

namespace adb {
int set_error(const char* fmt, ...);   // stores a thread-local message, returns 1
void count_launch(int n = 1);
int gemm_f64(const adb_gemm_desc* d, cudaStream_t st, int force_path);
int gemm_tf32x3(const adb_gemm_desc* d, cudaStream_t st);
// implicit-GEMM convolution (gemm_f64.cu): the im2col matrix of `img` (N,H,W,C) is read through an im2col-mode tensor map
bool conv_image_ok(const adb_tensor* img, int kh, int kw, int pad_h, int pad_w);
int conv_gemm_cols(const adb_tensor* img, int kh, int kw, int pad_h, int pad_w, const double* B, long long b_sk, long long b_sn, long long Nn,
                   double* out, long long c_sm, long long c_sn, cudaStream_t st);
int conv_gemm_wgrad(const adb_tensor* img, int kh, int kw, const double* D, long long d_sp, long long d_so, long long O,
                    double* out, long long c_sm, long long c_sn, cudaStream_t st);
extern int g_precision;   // 0 = FP64 DMMA, 1 = 3xTF32 tcgen05
int ewise_run(const adb_ew_instr* code, int n_instr, int n_in, const adb_tensor* ins, int n_out, const adb_tensor* outs,
              cudaStream_t st);
struct ReduceSpec {           // dims innermost-first, already collapsed (see reduce.cu)
    int n_ops, n_o, n_r, group, square;
    int post_sqrt;            // single scalar output only: store sqrt(sum) (FrobeniusNorm of one tensor, reference ops.py:369)
    long long odim[4], rdim[4], out_s[4];
    long long op_os[ADB_RED_MAX_OPS][4], op_rs[ADB_RED_MAX_OPS][4];
    const double* op[ADB_RED_MAX_OPS];
    double* out;
};
int reduce_run(const ReduceSpec& S, cudaStream_t st);
int softmax_ce_run(const adb_tensor* labels, const adb_tensor* logits, const adb_tensor* out, int* flag, cudaStream_t st);
int softmax_ce_ring_try(const adb_tensor* labels, const adb_tensor* logits, const adb_tensor* out, int* flag, cudaStream_t st);
int sce_exp_tab();         // 1 = the table-based exp in the log-sum-exp kernels (default), 0 = the degree-13 Taylor variant
int softmax_ce_ring_narrow_try(const adb_tensor* labels, const adb_tensor* logits, const adb_tensor* out, int* flag, cudaStream_t st);
}  // namespace adb

#define ADB_CUDA_OK(expr)                                                                 \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) return adb::set_error("%s failed: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)
