// FP64 issue-rate probes for B200 (sm_100a): DMMA.8x8x4 vs DFMA, plus a double2 copy.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_fp64 probe_fp64.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dmma(double* out, int iters, double a0, double b0) {
    double c0[NACC], c1[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c0[i] = 0.0; c1[i] = 0.0; }
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dfma(double* out, int iters, double a0, double b0) {
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) c[i] = fma(a, c[i], b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_copy(const double2* __restrict__ in, double2* __restrict__ out, size_t n2) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n2; i += stride) out[i] = in[i];
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s sms %d clock %d kHz smem/blk optin %zu\n", p.name, p.multiProcessorCount, p.clockRate, p.sharedMemPerBlockOptin);
    double* out; CK(cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double)));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int ctas_per_sm = 1; ctas_per_sm <= 2; ++ctas_per_sm)
    for (int threads = 128; threads <= 1024; threads *= 2) {
        int grid = 148 * ctas_per_sm;
        if (threads * ctas_per_sm > 2048) continue;
        // DMMA 16 independent chains
        k_dmma<16><<<grid, threads>>>(out, 100, 1.0, 1.0);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        k_dmma<16><<<grid, threads>>>(out, iters, 1.0, 1.0);
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flop = (double)grid * (threads / 32) * (double)iters * 16 * 512.0;
        printf("DMMA884 nacc16 ctas/sm %d threads %4d : %8.3f ms  %7.2f TFLOP/s\n", ctas_per_sm, threads, ms, flop / ms * 1e-9);
        k_dfma<16><<<grid, threads>>>(out, 100, 1.0, 1.0);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        k_dfma<16><<<grid, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, e0, e1);
        flop = (double)grid * threads * (double)iters * 16 * 2.0;
        printf("DFMA    nacc16 ctas/sm %d threads %4d : %8.3f ms  %7.2f TFLOP/s\n", ctas_per_sm, threads, ms, flop / ms * 1e-9);
    }
    // fewer chains: latency probe (1 warp per SMSP, 1..8 chains)
    {
        int grid = 148, threads = 128;
        cudaEventRecord(e0); k_dmma<1><<<grid, threads>>>(out, iters, 1.0, 1.0); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("DMMA884 dependent chain (1 warp/SMSP): %.3f ms => %.1f ns per DMMA\n", ms, ms * 1e6 / iters);
        cudaEventRecord(e0); k_dmma<4><<<grid, threads>>>(out, iters, 1.0, 1.0); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, e0, e1);
        printf("DMMA884 4 chains (1 warp/SMSP): %.3f ms => %.1f ns per DMMA\n", ms, ms * 1e6 / iters / 4);
    }
    // copy bandwidth
    {
        size_t n = (size_t)1 << 28;  // 2 GiB of doubles per buffer
        double *a, *b; CK(cudaMalloc(&a, n * 8)); CK(cudaMalloc(&b, n * 8));
        CK(cudaMemset(a, 1, n * 8));
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            k_copy<<<148 * 16, 512>>>((const double2*)a, (double2*)b, n / 2);
            cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("copy double2 2GiB->2GiB: %.3f ms %.1f GB/s\n", ms, 2.0 * n * 8 / ms * 1e-6);
        }
    }
    return 0;
}
