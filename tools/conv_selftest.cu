// Self-test of the implicit-GEMM convolution entry point (adb_conv_gemm, modes 0/1/2) against naive device kernels,
// plus timings at the BASELINE configs[3] sizes.   ./conv_selftest [time]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include "../include/adb200.h"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2);} } while (0)

__global__ void fill_rand(double* p, size_t n, unsigned long long seed) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        unsigned long long x = (i + 1) * 0x9E3779B97F4A7C15ull + seed;
        x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
        p[i] = ((double)(x >> 11) / 9007199254740992.0) * 2.0 - 1.0;
    }
}
// x (N,H,W,C) dense, w (kh*kw*C, O) dense, d (P,O) dense
__global__ void ref_fwd(const double* x, const double* w, double* out, int N, int H, int W, int C, int kh, int kw, int O) {
    const int OH = H - kh + 1, OW = W - kw + 1;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)N * OH * OW * O) return;
    const int o = i % O; long long p = i / O;
    const int ow = p % OW; const int oh = (p / OW) % OH; const int n = p / ((long long)OW * OH);
    double acc = 0;
    for (int a = 0; a < kh; ++a) for (int b = 0; b < kw; ++b) for (int c = 0; c < C; ++c)
        acc += x[(((long long)n * H + oh + a) * W + ow + b) * C + c] * w[((long long)(a * kw + b) * C + c) * O + o];
    out[i] = acc;
}
__global__ void ref_wgrad(const double* x, const double* d, double* out, int N, int H, int W, int C, int kh, int kw, int O) {
    const int OH = H - kh + 1, OW = W - kw + 1;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)kh * kw * C * O) return;
    const int o = i % O; long long k = i / O;
    const int c = k % C; const int b = (k / C) % kw; const int a = k / ((long long)C * kw);
    double acc = 0;
    for (int n = 0; n < N; ++n) for (int oh = 0; oh < OH; ++oh) for (int ow = 0; ow < OW; ++ow)
        acc += x[(((long long)n * H + oh + a) * W + ow + b) * C + c] * d[(((long long)n * OH + oh) * OW + ow) * O + o];
    out[i] = acc;
}
__global__ void ref_dgrad(const double* d, const double* w, double* out, int N, int H, int W, int C, int kh, int kw, int O) {
    const int OH = H - kh + 1, OW = W - kw + 1;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)N * H * W * C) return;
    const int c = i % C; long long r = i / C;
    const int xx = r % W; const int y = (r / W) % H; const int n = r / ((long long)W * H);
    double acc = 0;
    for (int a = 0; a < kh; ++a) for (int b = 0; b < kw; ++b) {
        const int oh = y - a, ow = xx - b;
        if (oh < 0 || oh >= OH || ow < 0 || ow >= OW) continue;
        for (int o = 0; o < O; ++o) acc += d[(((long long)n * OH + oh) * OW + ow) * O + o] * w[((long long)(a * kw + b) * C + c) * O + o];
    }
    out[i] = acc;
}
__global__ void max_diff(const double* a, const double* b, size_t n, double* res) {
    double md = 0, mr = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double d = fabs(a[i] - b[i]); if (!(d <= md)) md = d; mr = fmax(mr, fabs(b[i]));
    }
    unsigned long long* r = reinterpret_cast<unsigned long long*>(res);
    atomicMax(r, (unsigned long long)__double_as_longlong(md != md ? 1e300 : md)); atomicMax(r + 1, (unsigned long long)__double_as_longlong(mr));
}
static adb_tensor T(double* p, std::vector<long long> shp) {
    adb_tensor t; memset(&t, 0, sizeof t); t.ptr = p; t.rank = (int)shp.size();
    long long acc = 1;
    for (int k = t.rank - 1; k >= 0; --k) { t.shape[k] = shp[k]; t.stride[k] = acc; acc *= shp[k]; }
    return t;
}
static int run(int N, int H, int W, int C, int kh, int kw, int O, bool check, bool time_it) {
    const long long OH = H - kh + 1, OW = W - kw + 1, P = (long long)N * OH * OW, K = (long long)kh * kw * C;
    double *x, *w, *d, *o0, *o1, *o2, *r, *res;
    CK(cudaMalloc(&x, (size_t)N * H * W * C * 8)); CK(cudaMalloc(&w, K * O * 8)); CK(cudaMalloc(&d, P * O * 8));
    CK(cudaMalloc(&o0, P * O * 8)); CK(cudaMalloc(&o1, K * O * 8)); CK(cudaMalloc(&o2, (size_t)N * H * W * C * 8));
    size_t rmax = (size_t)N * H * W * C; if ((size_t)(P * O) > rmax) rmax = P * O; if ((size_t)(K * O) > rmax) rmax = K * O;
    CK(cudaMalloc(&r, rmax * 8)); CK(cudaMalloc(&res, 16));
    fill_rand<<<1024, 256>>>(x, (size_t)N * H * W * C, 1); fill_rand<<<64, 256>>>(w, K * O, 2); fill_rand<<<1024, 256>>>(d, P * O, 3);
    adb_tensor tx = T(x, {N, H, W, C}), tw = T(w, {K, O}), td = T(d, {P, O}), td4 = T(d, {N, OH, OW, O});
    adb_tensor t0 = T(o0, {P, O}), t1 = T(o1, {K, O}), t2 = T(o2, {N, H, W, C});
    int bad = 0;
    struct M { int mode; const adb_tensor *a, *b, *o; double* out; size_t n; double flops; const char* name; } ms[3] = {
        {0, &tx, &tw, &t0, o0, (size_t)(P * O), 2.0 * P * K * O, "forward"},
        {1, &tx, &td, &t1, o1, (size_t)(K * O), 2.0 * P * K * O, "filter gradient"},
        {2, &td4, &tw, &t2, o2, (size_t)N * H * W * C, 2.0 * N * H * W * (double)kh * kw * O * C, "data gradient"}};
    for (auto& m : ms) {
        if (m.mode == 2 && O % 16) { printf("skip data gradient (O %% 16)\n"); continue; }
        if (!adb_conv_gemm_ok(m.mode, m.a, kh, kw, m.b, m.o)) { printf("FAIL %s: adb_conv_gemm_ok says no\n", m.name); ++bad; continue; }
        CK(cudaMemset(m.out, 0xff, m.n * 8));
        if (adb_conv_gemm(m.mode, m.a, kh, kw, m.b, m.o, 0)) { printf("FAIL %s: %s\n", m.name, adb_last_error()); ++bad; continue; }
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("FAIL %s kernel error: %s\n", m.name, cudaGetErrorString(e)); exit(3); }
        if (check) {
            const int thr = 256; const long long nb = ((long long)m.n + thr - 1) / thr;
            if (m.mode == 0) ref_fwd<<<(unsigned)nb, thr>>>(x, w, r, N, H, W, C, kh, kw, O);
            if (m.mode == 1) ref_wgrad<<<(unsigned)nb, thr>>>(x, d, r, N, H, W, C, kh, kw, O);
            if (m.mode == 2) ref_dgrad<<<(unsigned)nb, thr>>>(d, w, r, N, H, W, C, kh, kw, O);
            CK(cudaMemset(res, 0, 16));
            max_diff<<<256, 256>>>(m.out, r, m.n, res);
            double h[2]; CK(cudaMemcpy(h, res, 16, cudaMemcpyDeviceToHost));
            const double rel = h[0] / (h[1] > 0 ? h[1] : 1);
            const bool f = !(rel < 1e-12);
            bad += f;
            printf("%s %-16s N=%d %dx%dx%d k=%dx%d O=%d  relerr=%.3e\n", f ? "FAIL" : "ok  ", m.name, N, H, W, C, kh, kw, O, rel);
        }
        if (time_it) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            float best = 1e30f;
            for (int rep = 0; rep < 5; ++rep) {
                cudaEventRecord(e0); adb_conv_gemm(m.mode, m.a, kh, kw, m.b, m.o, 0); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
                float t; cudaEventElapsedTime(&t, e0, e1); if (t < best) best = t;
            }
            printf("     time %-16s N=%d %dx%dx%d k=%dx%d O=%d : %.3f ms  %.2f TFLOP/s\n", m.name, N, H, W, C, kh, kw, O, best, m.flops / best * 1e-9);
        }
    }
    cudaFree(x); cudaFree(w); cudaFree(d); cudaFree(o0); cudaFree(o1); cudaFree(o2); cudaFree(r); cudaFree(res);
    return bad;
}
int main(int argc, char** argv) {
    if (adb_init(0)) { printf("adb_init failed: %s\n", adb_last_error()); return 1; }
    int bad = 0;
    bad += run(2, 9, 11, 16, 3, 3, 32, true, false);
    bad += run(3, 12, 10, 16, 5, 5, 32, true, false);
    bad += run(5, 20, 17, 32, 5, 4, 16, true, false);     // two k-tiles per tap, 16 output channels
    bad += run(4, 14, 14, 48, 3, 5, 40, true, false);     // O > 32: 128x128 tiles (data gradient skipped: 40 % 16)
    bad += run(2, 16, 16, 16, 1, 1, 16, true, false);     // 1x1 window
    bad += run(7, 30, 30, 16, 5, 5, 48, true, false);
    bad += run(16, 60, 60, 16, 5, 5, 32, true, false);    // configs[3] second layer at batch 16
    printf("conv correctness: %d failing cases\n", bad);
    if (argc > 1 && !strcmp(argv[1], "time")) run(256, 60, 60, 16, 5, 5, 32, false, true);
    return bad ? 1 : 0;
}
