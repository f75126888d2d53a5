// Convolution support = im2col (+ its adjoint col2im) feeding the einsum/GEMM path.
// The reference only has a commented-out design (reshape.py:137-160, high_level_ops.py:152-165): overlapping
// windows via as_strided, then Einsum.  Semantics decoded from its dead test (tests/test_convolution.py:17-33):
// NHWC input, VALID, stride 1, cross-correlation, filters last.
//
//   im2col : cols[n,oh,ow, a,b,c] = x[n, oh+a, ow+b, c]            (pure gather = strided copy, HBM-bound)
//   col2im : dx[n,y,x,c] = sum_{a,b valid} dcols[n, y-a, x-b, a,b,c] (gather-sum, reads dcols once, HBM-bound)
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "adb_internal.h"

namespace adb {

__global__ void __launch_bounds__(256) col2im_kernel(const double* __restrict__ cols, double* __restrict__ dx, int N, int H, int W, int C,
                                                     int kh, int kw, long long x_sn, long long x_sh, long long x_sw, long long x_sc,
                                                     long long cols_row_stride) {
    const int OH = H - kh + 1, OW = W - kw + 1;
    const long long total = (long long)N * H * W * C;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
        const int c = (int)(i % C);
        long long r = i / C;
        const int x = (int)(r % W); r /= W;
        const int y = (int)(r % H);
        const int n = (int)(r / H);
        double acc = 0.0;
        const int a_lo = y - (OH - 1) > 0 ? y - (OH - 1) : 0, a_hi = y < kh - 1 ? y : kh - 1;
        const int b_lo = x - (OW - 1) > 0 ? x - (OW - 1) : 0, b_hi = x < kw - 1 ? x : kw - 1;
        for (int a = a_lo; a <= a_hi; ++a) {
            const long long rowbase = ((long long)n * OH + (y - a)) * OW;
            for (int b = b_lo; b <= b_hi; ++b)
                acc += __ldg(cols + (rowbase + (x - b)) * cols_row_stride + ((long long)a * kw + b) * C + c);
        }
        dx[n * x_sn + y * x_sh + x * x_sw + c * x_sc] = acc;
    }
}

// Channel-quad variant (C % 4 == 0, 32-byte aligned rows): one thread gathers 4 consecutive channels of one pixel with
// 256-bit loads; the lanes of a warp cover consecutive (pixel, channel-quad) pairs, i.e. whole 128-byte channel runs.
__global__ void __launch_bounds__(256) col2im_quad_kernel(const double* __restrict__ cols, double* __restrict__ dx, int N, int H, int W, int CQ,
                                                          int kh, int kw, long long x_sn, long long x_sh, long long x_sw,
                                                          long long cols_row_stride) {
    const int OH = H - kh + 1, OW = W - kw + 1;
    const long long total = (long long)N * H * W * CQ;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int cq = (int)(i % CQ);
    long long r = i / CQ;
    const int x = (int)(r % W); r /= W;
    const int y = (int)(r % H);
    const int n = (int)(r / H);
    const int C = CQ * 4;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    const int a_lo = y - (OH - 1) > 0 ? y - (OH - 1) : 0, a_hi = y < kh - 1 ? y : kh - 1;
    const int b_lo = x - (OW - 1) > 0 ? x - (OW - 1) : 0, b_hi = x < kw - 1 ? x : kw - 1;
    for (int a = a_lo; a <= a_hi; ++a) {
        const long long rowbase = ((long long)n * OH + (y - a)) * OW;
        for (int b = b_lo; b <= b_hi; ++b) {
            const double* p = cols + (rowbase + (x - b)) * cols_row_stride + ((long long)a * kw + b) * C + cq * 4;
            double v0, v1, v2, v3;
            asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v0), "=d"(v1), "=d"(v2), "=d"(v3) : "l"(p));
            a0 += v0; a1 += v1; a2 += v2; a3 += v3;
        }
    }
    double* q = dx + n * x_sn + y * x_sh + x * x_sw + cq * 4;
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(q), "d"(a0), "d"(a1), "d"(a2), "d"(a3) : "memory");
}

}  // namespace adb

extern "C" {

// x: (N,H,W,C) view, cols: (N*OH*OW, kh*kw*C) view with unit column stride.
int adb_im2col(const adb_tensor* x, int kh, int kw, const adb_tensor* cols, void* stream) {
    if (x->rank != 4 || cols->rank != 2) return adb::set_error("im2col: x must be 4-D (NHWC) and cols 2-D");
    const long long N = x->shape[0], H = x->shape[1], W = x->shape[2], C = x->shape[3];
    if (kh < 1 || kw < 1 || kh > H || kw > W) return adb::set_error("im2col: bad window %dx%d for %lldx%lld", kh, kw, H, W);
    const long long OH = H - kh + 1, OW = W - kw + 1;
    if (cols->shape[0] != N * OH * OW || cols->shape[1] != (long long)kh * kw * C) return adb::set_error("im2col: cols has the wrong shape");
    adb_tensor in, out;
    in.ptr = x->ptr; in.rank = 6;
    out.ptr = cols->ptr; out.rank = 6;
    const long long shp[6] = {N, OH, OW, kh, kw, C};
    const long long ist[6] = {x->stride[0], x->stride[1], x->stride[2], x->stride[1], x->stride[2], x->stride[3]};
    const long long rs = cols->stride[0], cs = cols->stride[1];
    const long long ost[6] = {OH * OW * rs, OW * rs, rs, (long long)kw * C * cs, C * cs, cs};
    for (int k = 0; k < 6; ++k) { in.shape[k] = out.shape[k] = shp[k]; in.stride[k] = ist[k]; out.stride[k] = ost[k]; }
    adb_ew_instr code[2] = {{ADB_EW_LOAD | ADB_EW_SRC_IN, 0, 0.0}, {ADB_EW_STORE_OUT, 0, 0.0}};
    return adb::ewise_run(code, 2, 1, &in, 1, &out, reinterpret_cast<cudaStream_t>(stream));
}

int adb_col2im(const adb_tensor* cols, int kh, int kw, const adb_tensor* dx, void* stream) {
    if (dx->rank != 4 || cols->rank != 2) return adb::set_error("col2im: dx must be 4-D (NHWC) and cols 2-D");
    const long long N = dx->shape[0], H = dx->shape[1], W = dx->shape[2], C = dx->shape[3];
    if (kh < 1 || kw < 1 || kh > H || kw > W) return adb::set_error("col2im: bad window %dx%d for %lldx%lld", kh, kw, H, W);
    const long long OH = H - kh + 1, OW = W - kw + 1;
    if (cols->shape[0] != N * OH * OW || cols->shape[1] != (long long)kh * kw * C) return adb::set_error("col2im: cols has the wrong shape");
    if (cols->stride[1] != 1) return adb::set_error("col2im: cols must have unit column stride");
    if (N * H * W * C == 0) return 0;
    if (H > 0x7fffffff / 4 || W > 0x7fffffff / 4) return adb::set_error("col2im: image too large");
    const bool quad = C % 4 == 0 && dx->stride[3] == 1 && cols->stride[0] % 4 == 0 && dx->stride[0] % 4 == 0 && dx->stride[1] % 4 == 0 &&
                      dx->stride[2] % 4 == 0 && ((reinterpret_cast<uintptr_t>(cols->ptr) | reinterpret_cast<uintptr_t>(dx->ptr)) & 31) == 0 &&
                      N * H * W * (C / 4) < (1LL << 31) * 256;
    if (quad) {
        const long long threads = N * H * W * (C / 4);
        adb::col2im_quad_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
            cols->ptr, dx->ptr, (int)N, (int)H, (int)W, (int)(C / 4), kh, kw, dx->stride[0], dx->stride[1], dx->stride[2], cols->stride[0]);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return adb::set_error("col2im launch failed: %s", cudaGetErrorString(e));
        adb::count_launch();
        return 0;
    }
    long long blocks = (N * H * W * C + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    adb::col2im_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        cols->ptr, dx->ptr, (int)N, (int)H, (int)W, (int)C, kh, kw, dx->stride[0], dx->stride[1], dx->stride[2], dx->stride[3], cols->stride[0]);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return adb::set_error("col2im launch failed: %s", cudaGetErrorString(e));
    adb::count_launch();
    return 0;
}

}  // extern "C"
