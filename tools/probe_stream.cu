// HBM streaming probes for B200 (sm_100a): which kernel shape reaches the measured copy peak?
// copy / add (2 in, 1 out) / read-only sum / write-only fill, swept over vectors-per-thread (U), CTAs per SM,
// 128- vs 256-bit accesses, cache hints, persistent vs one-shot grids, plus a register-free TMA bulk copy ring.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o probe_stream probe_stream.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

struct __align__(32) d4 { double x, y, z, w; };

__device__ __forceinline__ double2 ld128(const double2* p) { return __ldg(p); }
__device__ __forceinline__ double2 ld128_na(const double2* p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ d4 ld256(const d4* p) {
    d4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st256(d4* p, const d4& v) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}

// MODE: 0 copy, 1 add (2 in), 2 read-only sum, 3 write-only
// HINT: 0 = __ldg + plain store, 1 = L1::no_allocate load + st.cs, 2 = __ldg + st.cs
template <int U, int MODE, int HINT, int MINB>
__global__ void __launch_bounds__(256, MINB) k128(const double2* __restrict__ a, const double2* __restrict__ b, double2* __restrict__ o, size_t n2, double* sink) {
    // tile = 256 threads * U vectors; within a tile vector j of thread t is at j*256 + t  (coalesced 4 KiB per j)
    const size_t tile = (size_t)256 * U;
    const size_t ntiles = n2 / tile;
    double acc = 0.0;
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const size_t base = t * tile + threadIdx.x;
        double2 x[U], y[U];
        if (MODE != 3) {
#pragma unroll
            for (int j = 0; j < U; ++j) x[j] = HINT == 1 ? ld128_na(a + base + j * 256) : ld128(a + base + j * 256);
        }
        if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < U; ++j) y[j] = HINT == 1 ? ld128_na(b + base + j * 256) : ld128(b + base + j * 256);
#pragma unroll
            for (int j = 0; j < U; ++j) { x[j].x += y[j].x; x[j].y += y[j].y; }
        }
        if (MODE == 2) {
#pragma unroll
            for (int j = 0; j < U; ++j) acc += x[j].x + x[j].y;
        } else {
            if (MODE == 3) {
#pragma unroll
                for (int j = 0; j < U; ++j) x[j] = make_double2(1.0, 2.0);
            }
#pragma unroll
            for (int j = 0; j < U; ++j) {
                if (HINT == 0) o[base + j * 256] = x[j];
                else __stcs(o + base + j * 256, x[j]);
            }
        }
    }
    if (MODE == 2 && acc == 123.456) *sink = acc;
}

template <int U, int MODE, int MINB>
__global__ void __launch_bounds__(256, MINB) k256(const d4* __restrict__ a, const d4* __restrict__ b, d4* __restrict__ o, size_t n4, double* sink) {
    const size_t tile = (size_t)256 * U;
    const size_t ntiles = n4 / tile;
    double acc = 0.0;
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const size_t base = t * tile + threadIdx.x;
        d4 x[U], y[U];
        if (MODE != 3) {
#pragma unroll
            for (int j = 0; j < U; ++j) x[j] = ld256(a + base + j * 256);
        }
        if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < U; ++j) y[j] = ld256(b + base + j * 256);
#pragma unroll
            for (int j = 0; j < U; ++j) { x[j].x += y[j].x; x[j].y += y[j].y; x[j].z += y[j].z; x[j].w += y[j].w; }
        }
        if (MODE == 2) {
#pragma unroll
            for (int j = 0; j < U; ++j) acc += (x[j].x + x[j].y) + (x[j].z + x[j].w);
        } else {
            if (MODE == 3) {
#pragma unroll
                for (int j = 0; j < U; ++j) { x[j].x = 1; x[j].y = 2; x[j].z = 3; x[j].w = 4; }
            }
#pragma unroll
            for (int j = 0; j < U; ++j) st256(o + base + j * 256, x[j]);
        }
    }
    if (MODE == 2 && acc == 123.456) *sink = acc;
}

// ---- register-free copy: per-CTA ring of cp.async.bulk (global->shared) / cp.async.bulk (shared->global)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" :: "r"(smem_u32(b)), "r"(parity) : "memory");
}
template <int STAGES, int CHUNK>   // CHUNK bytes per stage
__global__ void __launch_bounds__(32, 1) k_bulk(const char* __restrict__ a, char* __restrict__ o, size_t bytes) {
    extern __shared__ __align__(128) char smem[];
    __shared__ uint64_t full[STAGES];
    const size_t nchunks = bytes / CHUNK;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // prologue
        size_t c = blockIdx.x;
        int issued = 0;
        for (; issued < STAGES && c < nchunks; ++issued, c += gridDim.x) {
            mbar_expect(&full[issued], CHUNK);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(smem_u32(smem + (size_t)issued * CHUNK)), "l"(a + c * CHUNK), "r"(CHUNK), "r"(smem_u32(&full[issued])) : "memory");
        }
        size_t next = c;
        int s = 0; uint32_t ph = 0;
        for (size_t cc = blockIdx.x; cc < nchunks; cc += gridDim.x) {
            mbar_wait(&full[s], ph);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(o + cc * CHUNK), "r"(smem_u32(smem + (size_t)s * CHUNK)), "r"(CHUNK) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (next < nchunks) {
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem of stage s has been read out
                mbar_expect(&full[s], CHUNK);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(smem_u32(smem + (size_t)s * CHUNK)), "l"(a + next * CHUNK), "r"(CHUNK), "r"(smem_u32(&full[s])) : "memory");
                next += gridDim.x;
            }
            if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

static cudaEvent_t e0, e1;
template <class F>
static double time_ms(F f, int reps = 10) {
    for (int i = 0; i < 3; ++i) f();
    CK(cudaDeviceSynchronize());
    float best = 1e9f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0));
        for (int i = 0; i < reps; ++i) f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms / reps < best) best = ms / reps;
    }
    CK(cudaGetLastError());
    return best;
}

static const char* MODES[] = {"copy", "add", "sum", "fill"};
static const double TRAFFIC[] = {2, 3, 1, 1};
static size_t N;   // doubles per array
static double *A, *B, *O, *SINK;

template <int U, int MODE, int HINT, int MINB>
static void run128(int ctas_per_sm) {
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k128<U, MODE, HINT, MINB>, 256, 0));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k128<U, MODE, HINT, MINB>));
    const int grid = ctas_per_sm > 0 ? 148 * (ctas_per_sm < occ ? ctas_per_sm : occ) : (int)(N / 2 / (256 * U));
    double ms = time_ms([&] { k128<U, MODE, HINT, MINB><<<grid, 256>>>((const double2*)A, (const double2*)B, (double2*)O, N / 2, SINK); });
    printf("k128 %-4s U=%d hint=%d minb=%d regs=%3d occ=%d grid=%7d : %.4f ms  %7.1f GB/s\n", MODES[MODE], U, HINT, MINB, fa.numRegs, occ, grid, ms,
           TRAFFIC[MODE] * N * 8 / ms / 1e6);
}
template <int U, int MODE>
static void run128_occ(int occ_target) {
    // one-shot grid; occupancy limited to occ_target CTAs/SM through a dynamic shared memory request
    const int smem = occ_target >= 8 ? 0 : (227 * 1024 / occ_target - 1024 - 1024);
    CK(cudaFuncSetAttribute(k128<U, MODE, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k128<U, MODE, 0, 1>, 256, smem));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k128<U, MODE, 0, 1>));
    const int grid = (int)(N / 2 / (256 * U));
    double ms = time_ms([&] { k128<U, MODE, 0, 1><<<grid, 256, smem>>>((const double2*)A, (const double2*)B, (double2*)O, N / 2, SINK); });
    printf("oneshot %-4s U=%d regs=%3d occ=%d grid=%7d : %.4f ms  %7.1f GB/s\n", MODES[MODE], U, fa.numRegs, occ, grid, ms,
           TRAFFIC[MODE] * N * 8 / ms / 1e6);
}
template <int U, int MODE, int MINB>
static void run256(int ctas_per_sm) {
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k256<U, MODE, MINB>, 256, 0));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k256<U, MODE, MINB>));
    const int grid = ctas_per_sm > 0 ? 148 * (ctas_per_sm < occ ? ctas_per_sm : occ) : (int)(N / 4 / (256 * U));
    double ms = time_ms([&] { k256<U, MODE, MINB><<<grid, 256>>>((const d4*)A, (const d4*)B, (d4*)O, N / 4, SINK); });
    printf("k256 %-4s U=%d        minb=%d regs=%3d occ=%d grid=%7d : %.4f ms  %7.1f GB/s\n", MODES[MODE], U, MINB, fa.numRegs, occ, grid, ms,
           TRAFFIC[MODE] * N * 8 / ms / 1e6);
}
template <int STAGES, int CHUNK>
static void runbulk(int ctas_per_sm) {
    CK(cudaFuncSetAttribute(k_bulk<STAGES, CHUNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * CHUNK));
    const int grid = 148 * ctas_per_sm;
    double ms = time_ms([&] { k_bulk<STAGES, CHUNK><<<grid, 32, STAGES * CHUNK>>>((const char*)A, (char*)O, N * 8); });
    printf("bulk copy stages=%d chunk=%d ctas/sm=%d : %.4f ms  %7.1f GB/s\n", STAGES, CHUNK, ctas_per_sm, ms, 2.0 * N * 8 / ms / 1e6);
}

int main(int argc, char** argv) {
    N = (size_t)8192 * 4096;
    if (argc > 1) N = (size_t)atoll(argv[1]);
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s sms %d, arrays of %zu doubles (%.0f MiB)\n", p.name, p.multiProcessorCount, N, N * 8 / 1048576.0);
    CK(cudaMalloc(&A, N * 8)); CK(cudaMalloc(&B, N * 8)); CK(cudaMalloc(&O, N * 8)); CK(cudaMalloc(&SINK, 8));
    CK(cudaMemset(A, 0, N * 8)); CK(cudaMemset(B, 0, N * 8)); CK(cudaMemset(O, 0, N * 8));
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double ms = time_ms([&] { CK(cudaMemcpyAsync(O, A, N * 8, cudaMemcpyDeviceToDevice)); });
    printf("cudaMemcpy D2D : %.4f ms  %7.1f GB/s\n", ms, 2.0 * N * 8 / ms / 1e6);
    ms = time_ms([&] { CK(cudaMemsetAsync(O, 0, N * 8)); });
    printf("cudaMemset     : %.4f ms  %7.1f GB/s\n", ms, 1.0 * N * 8 / ms / 1e6);

    if (argc > 2) {
#define OCC(MODE) for (int oc : {1, 2, 3, 4, 6, 8}) { run128_occ<1, MODE>(oc); run128_occ<2, MODE>(oc); run128_occ<4, MODE>(oc); run128_occ<8, MODE>(oc); }
        OCC(0) OCC(1) OCC(2) OCC(3)
        return 0;
    }
#define SWEEP128(MODE)                                                        \
    run128<1, MODE, 0, 8>(0); run128<2, MODE, 0, 8>(0); run128<4, MODE, 0, 6>(0); \
    run128<2, MODE, 0, 8>(8); run128<4, MODE, 0, 6>(6); run128<4, MODE, 0, 4>(4); run128<4, MODE, 0, 3>(3); \
    run128<8, MODE, 0, 4>(4); run128<8, MODE, 0, 3>(3); run128<8, MODE, 0, 2>(2); \
    run128<4, MODE, 1, 6>(6); run128<4, MODE, 2, 6>(6); run128<8, MODE, 1, 3>(3); run128<4, MODE, 1, 6>(0);
    SWEEP128(0) SWEEP128(1) SWEEP128(2) SWEEP128(3)
#define SWEEP256(MODE)                                                        \
    run256<1, MODE, 8>(0); run256<2, MODE, 6>(0); run256<1, MODE, 8>(8); run256<2, MODE, 6>(6); run256<2, MODE, 4>(4); \
    run256<4, MODE, 4>(4); run256<4, MODE, 3>(3); run256<4, MODE, 2>(2);
    SWEEP256(0) SWEEP256(1) SWEEP256(2) SWEEP256(3)
    runbulk<4, 16384>(1); runbulk<4, 16384>(2); runbulk<8, 8192>(2); runbulk<8, 16384>(1); runbulk<4, 32768>(1); runbulk<6, 32768>(1);
    runbulk<4, 8192>(4); runbulk<8, 4096>(4); runbulk<3, 16384>(4);
    return 0;
}
