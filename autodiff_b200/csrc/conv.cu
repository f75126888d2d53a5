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

// ---- implicit-GEMM convolution: the three contractions around Im2Col / Col2Im without the materialised matrix ----------
// mode 0  out (P, O)          = Im2Col(x) @ w              x: (N,H,W,C), w: (kh*kw*C, O), P = N*OH*OW        [forward]
// mode 1  out (kh*kw*C, O)    = Im2Col(x)^T @ d            d: (P, O)                                            [filter gradient]
// mode 2  out (N,H,W,C)       = Col2Im(d @ w^T)            x := d seen as (N,OH,OW,O), w: (kh*kw*C, O)          [data gradient]
static int conv_gemm_check(int mode, const adb_tensor* x, int kh, int kw, const adb_tensor* w, const adb_tensor* out, bool quiet) {
    auto fail = [&](const char* why) { return quiet ? 1 : adb::set_error("conv_gemm: %s", why); };
    if (mode < 0 || mode > 2) return fail("mode must be 0, 1 or 2");
    if (x->rank != 4 || w->rank != 2) return fail("x must be 4-D (NHWC) and w / d 2-D");
    if (kh < 1 || kw < 1) return fail("bad window");
    const long long N = x->shape[0], H = x->shape[1], W = x->shape[2], C = x->shape[3];
    auto matrix_ok = [&](const adb_tensor* t) {
        if ((reinterpret_cast<uintptr_t>(t->ptr) & 15) != 0) return false;
        if (t->stride[1] == 1) return t->stride[0] > 0 && (t->stride[0] & 1) == 0;
        return t->stride[0] == 1 && t->stride[1] > 0 && (t->stride[1] & 1) == 0;
    };
    if (mode == 2) {
        // x is dOut (N,OH,OW,O); out is the image gradient (N, OH+kh-1, OW+kw-1, C)
        if (out->rank != 4) return fail("mode 2: out must be 4-D");
        const long long Hi = H + kh - 1, Wi = W + kw - 1, Ci = out->shape[3];
        if (out->shape[0] != N || out->shape[1] != Hi || out->shape[2] != Wi) return fail("mode 2: out has the wrong shape");
        if (w->shape[0] != (long long)kh * kw * Ci || w->shape[1] != C) return fail("mode 2: w has the wrong shape");
        if (out->stride[3] != 1 || out->stride[1] != Wi * out->stride[2] || out->stride[0] != Hi * out->stride[1] || out->stride[2] < Ci)
            return fail("mode 2: out must be a dense NHWC array");
        if (!adb::conv_image_ok(x, kh, kw, kh - 1, kw - 1)) return fail("mode 2: output gradient not describable by an im2col tensor map");
        if (Ci > 1024) return fail("mode 2: too many channels");
        return 0;
    }
    const long long OH = H - kh + 1, OW = W - kw + 1;
    if (OH < 1 || OW < 1) return fail("window larger than the image");
    if (!adb::conv_image_ok(x, kh, kw, 0, 0)) return fail("image not describable by an im2col tensor map (C % 16, alignment)");
    if (out->rank != 2) return fail("out must be 2-D");
    if (mode == 0) {
        if (w->shape[0] != (long long)kh * kw * C || out->shape[0] != N * OH * OW || out->shape[1] != w->shape[1]) return fail("mode 0: shapes do not match");
        if (!matrix_ok(w)) return fail("mode 0: w not describable by a tensor map");
    } else {
        if (w->shape[0] != N * OH * OW || out->shape[0] != (long long)kh * kw * C || out->shape[1] != w->shape[1]) return fail("mode 1: shapes do not match");
        if (!matrix_ok(w)) return fail("mode 1: d not describable by a tensor map");
        if ((long long)kh * kw * C * w->shape[1] > (1LL << 24)) return fail("mode 1: filter too large for the split-K workspace");
    }
    return 0;
}

int adb_conv_gemm_ok(int mode, const adb_tensor* x, int kh, int kw, const adb_tensor* w, const adb_tensor* out) {
    return conv_gemm_check(mode, x, kh, kw, w, out, true) == 0 ? 1 : 0;
}

int adb_conv_gemm(int mode, const adb_tensor* x, int kh, int kw, const adb_tensor* w, const adb_tensor* out, void* stream) {
    if (int rc = conv_gemm_check(mode, x, kh, kw, w, out, false)) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (mode == 0)
        return adb::conv_gemm_cols(x, kh, kw, 0, 0, w->ptr, w->stride[0], w->stride[1], w->shape[1], out->ptr, out->stride[0], out->stride[1], st);
    if (mode == 1)
        return adb::conv_gemm_wgrad(x, kh, kw, w->ptr, w->stride[0], w->stride[1], w->shape[1], out->ptr, out->stride[0], out->stride[1], st);
    // mode 2: dx[n,y,x,c] = sum_{a,b,o} d[n, y-a, x-b, o] * w[(a,b,c), o]  =  a FULL convolution of d (zero padding k-1) with the
    // flipped filter  wf[(a',b',o), c] = w[(kh-1-a', kw-1-b', c), o]  (built here: kh*kw*O*C elements, a strided copy)
    const long long O = x->shape[3], Ci = out->shape[3];
    double* wf = nullptr;
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&wf), (size_t)kh * kw * O * Ci * sizeof(double), st);
    if (e != cudaSuccess) return adb::set_error("conv_gemm: flipped-filter allocation failed: %s", cudaGetErrorString(e));
    adb_tensor in, o4;
    in.rank = o4.rank = 4;
    const long long r = w->stride[0], c = w->stride[1];
    in.ptr = w->ptr + ((long long)(kh - 1) * kw + (kw - 1)) * Ci * r;
    const long long shp[4] = {kh, kw, O, Ci};
    const long long ist[4] = {-(long long)kw * Ci * r, -Ci * r, c, r};
    const long long ost[4] = {(long long)kw * O * Ci, O * Ci, Ci, 1};
    for (int k = 0; k < 4; ++k) { in.shape[k] = o4.shape[k] = shp[k]; in.stride[k] = ist[k]; o4.stride[k] = ost[k]; }
    o4.ptr = wf;
    adb_ew_instr code[2] = {{ADB_EW_LOAD | ADB_EW_SRC_IN, 0, 0.0}, {ADB_EW_STORE_OUT, 0, 0.0}};
    int rc = adb::ewise_run(code, 2, 1, &in, 1, &o4, st);
    if (rc == 0)
        rc = adb::conv_gemm_cols(x, kh, kw, kh - 1, kw - 1, wf, Ci, 1, Ci, out->ptr, out->stride[2], 1, st);
    cudaFreeAsync(wf, st);
    return rc;
}

}  // extern "C"
