// Reproducer for the data-parallel parity failure of round 1: the FP64 GEMM on one stream while an
// HBM-bound streaming kernel (what the side-stream Adam update is) runs on another.  Every repetition is compared
// BITWISE with a quiet run of the same GEMM; any mismatch is a race inside the GEMM kernel.
//   ./gemm_stress <path> [M N K] [reps] [hammer: 0|1]        path = adb_gemm_f64_path selector (0 auto, 3 small, 4 big, ...)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

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

// one-shot streaming kernel: 4 arrays read, 3 written (the shape of the fused Adam update)
__global__ void hammer(const double4* a, const double4* b, const double4* c, const double4* d, double4* x, double4* y, double4* z, size_t n4) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    double4 va = a[i], vb = b[i], vc = c[i], vd = d[i];
    x[i] = make_double4(va.x + vb.x, va.y + vb.y, va.z + vb.z, va.w + vb.w);
    y[i] = make_double4(vc.x * 0.9, vc.y * 0.9, vc.z * 0.9, vc.w * 0.9);
    z[i] = make_double4(vd.x + va.x, vd.y + va.y, vd.z + va.z, vd.w + va.w);
}

__global__ void hammer_ro(const double4* a, const double4* b, const double4* c, const double4* d, double4* x, size_t n4) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    double4 va = a[i], vb = b[i], vc = c[i], vd = d[i];
    if (va.x + vb.x + vc.x + vd.x == 123.456) x[i] = va;      // never true: loads only
}
__global__ void hammer_wo(double4* x, double4* y, double4* z, size_t n4) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    x[i] = make_double4(1, 2, 3, 4); y[i] = make_double4(1, 2, 3, 4); z[i] = make_double4(1, 2, 3, 4);
}
__global__ void hammer_spin(double* x, int iters) {           // no memory traffic: many short CTAs that only occupy issue slots
    double v = threadIdx.x;
    for (int i = 0; i < iters; ++i) v = fma(v, 1.0000001, 1e-9);
    if (v == 123.456) x[0] = v;
}

__global__ void count_diff(const unsigned long long* a, const unsigned long long* b, size_t n, unsigned long long* out /*[0]=count,[1]=first index+1*/) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride)
        if (a[i] != b[i]) {
            atomicAdd(out, 1ull);
            atomicMin(out + 1, (unsigned long long)i);
        }
}

int main(int argc, char** argv) {
    int path = argc > 1 ? atoi(argv[1]) : 0;
    long long M = argc > 4 ? atoll(argv[2]) : 2048, N = argc > 4 ? atoll(argv[3]) : 4096, K = argc > 4 ? atoll(argv[4]) : 4096;
    int reps = argc > 5 ? atoi(argv[5]) : 40;
    int do_hammer = argc > 6 ? atoi(argv[6]) : 1;
    int layout = argc > 7 ? atoi(argv[7]) : 0;          // 0: A,B K-contiguous (dX form); 1: both MN-contiguous (dW form)
    if (adb_init(0)) { printf("adb_init: %s\n", adb_last_error()); return 2; }
    double *A, *B, *C0, *C;   // C: 4 result buffers, one per queued GEMM
    CK(cudaMalloc(&A, (size_t)M * K * 8)); CK(cudaMalloc(&B, (size_t)N * K * 8));
    CK(cudaMalloc(&C0, (size_t)M * N * 8)); CK(cudaMalloc(&C, (size_t)4 * M * N * 8));
    fill_rand<<<1024, 256>>>(A, (size_t)M * K, 1); fill_rand<<<1024, 256>>>(B, (size_t)N * K, 2);
    const size_t HN = (size_t)4097 * 4096;               // one MLP weight matrix
    double* h[7];
    for (int i = 0; i < 7; ++i) { CK(cudaMalloc(&h[i], HN * 8)); fill_rand<<<1024, 256>>>(h[i], HN, 10 + i); }
    unsigned long long* stat; CK(cudaMalloc(&stat, 16));
    CK(cudaDeviceSynchronize());
    cudaStream_t s1, s2; CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    adb_gemm_desc d; memset(&d, 0, sizeof d);
    d.M = M; d.N = N; d.K = K; d.batch = 1;
    if (layout == 0) { d.A = A; d.a_sm = K; d.a_sk = 1; d.B = B; d.b_sn = K; d.b_sk = 1; }
    else { d.A = A; d.a_sm = 1; d.a_sk = M; d.B = B; d.b_sn = 1; d.b_sk = N; }
    d.c_sm = N; d.c_sn = 1;
    d.C = C0;
    if (adb_gemm_f64_path(&d, s1, path)) { printf("gemm: %s\n", adb_last_error()); return 2; }
    CK(cudaDeviceSynchronize());
    long long bad_runs = 0, bad_elems = 0;
    for (int r = 0; r < reps; ++r) {
        CK(cudaMemsetAsync(C, 0xff, (size_t)4 * M * N * 8, s1));
        CK(cudaStreamSynchronize(s1));
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        // queue several GEMMs back to back with hammers interleaved on the other stream (deep queue, like a training step)
        CK(cudaEventRecord(e0, s1));
        for (int rr = 0; rr < 4; ++rr) {
            const unsigned hg = (unsigned)((HN / 4 + 255) / 256);
            if (do_hammer == 2) hammer_ro<<<hg, 256, 0, s2>>>((double4*)h[0], (double4*)h[1], (double4*)h[2], (double4*)h[3], (double4*)h[4], HN / 4);
            else if (do_hammer == 3) hammer_wo<<<hg, 256, 0, s2>>>((double4*)h[4], (double4*)h[5], (double4*)h[6], HN / 4);
            else if (do_hammer == 4) hammer_spin<<<hg / 8, 256, 0, s2>>>(h[4], 2000);
            else if (do_hammer == 1) hammer<<<(unsigned)((HN / 4 + 255) / 256), 256, 0, s2>>>((double4*)h[0], (double4*)h[1], (double4*)h[2], (double4*)h[3], (double4*)h[4], (double4*)h[5], (double4*)h[6], HN / 4);
            d.C = C + (size_t)rr * M * N;
            if (adb_gemm_f64_path(&d, s1, path)) { printf("gemm: %s\n", adb_last_error()); return 2; }
        }
        CK(cudaEventRecord(e1, s1));
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        unsigned long long init[2] = {0ull, ~0ull};
        CK(cudaMemcpy(stat, init, 16, cudaMemcpyHostToDevice));
        for (int rr = 0; rr < 4; ++rr) count_diff<<<1024, 256>>>((unsigned long long*)(C + (size_t)rr * M * N), (unsigned long long*)C0, (size_t)M * N, stat);
        unsigned long long got[2];
        CK(cudaMemcpy(got, stat, 16, cudaMemcpyDeviceToHost));
        if (got[0]) {
            ++bad_runs; bad_elems += got[0];
            printf("  rep %d: %llu elements differ, first at row %llu col %llu   (4 GEMMs %.3f ms)\n", r, got[0], got[1] / N, got[1] % N, ms);
        } else if (r == 0) printf("  rep 0 clean (4 GEMMs %.3f ms = %.2f TFLOP/s)\n", ms, 4 * 2.0 * M * N * K / (ms * 1e-3) / 1e12);
    }
    printf("path %d layout %d M %lld N %lld K %lld hammer %d: %lld of %d repetitions differ (%lld elements)\n", path, layout, M, N, K, do_hammer, bad_runs, reps, bad_elems);
    return 0;
}
