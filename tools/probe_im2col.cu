// Probe of TMA im2col mode with fp64 data on sm_100a: what lands in shared memory for given start coordinates / offsets.
//   tensor x: (N, H, W, C) NHWC fp64, value = n*1e6 + h*1e4 + w*1e2 + c   (so every element names itself)
//   map: rank 4 dims {C, W, H, N}, lower corner {lw, lh}, upper corner {uw, uh}, channelsPerPixel = 16, pixelsPerColumn = P
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o probe_im2col probe_im2col.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(2);} } while (0)

__global__ void probe(const __grid_constant__ CUtensorMap tm, double* out, int P, int c0, int w0, int h0, int n0, int offw, int offh) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    for (int i = threadIdx.x; i < P * 16; i += blockDim.x) reinterpret_cast<double*>(smem)[i] = -7.0;
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(P * 128) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
            ::"r"(dst), "l"(&tm), "r"(b), "r"(c0), "r"(w0), "r"(h0), "r"(n0), "h"((unsigned short)offw), "h"((unsigned short)offh)
            : "memory");
    }
    uint32_t ok = 0;
    int spins = 0;
    while (!ok && spins < 2000000) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(b) : "memory");
        ++spins;
    }
    __syncthreads();
    if (threadIdx.x == 0 && !ok) out[P * 16] = -1.0;      // timed out
    for (int i = threadIdx.x; i < P * 16; i += blockDim.x) {
        const int row = i / 16, chunk = (i % 16) / 2, e = i % 2;           // undo the 128B swizzle: 16-byte chunk j of row r sits at chunk j ^ (r & 7)
        out[i] = reinterpret_cast<double*>(smem)[row * 16 + ((chunk ^ (row & 7)) * 2) + e];
    }
}

int main(int argc, char** argv) {
    const int N = 3, H = 10, W = 12, C = 32;
    int kh = 5, kw = 5;
    std::vector<double> hx((size_t)N * H * W * C);
    for (int n = 0; n < N; ++n) for (int h = 0; h < H; ++h) for (int w = 0; w < W; ++w) for (int c = 0; c < C; ++c)
        hx[(((size_t)n * H + h) * W + w) * C + c] = n * 1e6 + h * 1e4 + w * 1e2 + c;
    double* dx; CK(cudaMalloc(&dx, hx.size() * 8)); CK(cudaMemcpy(dx, hx.data(), hx.size() * 8, cudaMemcpyHostToDevice));
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q));
    typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*, const int*, cuuint32_t,
                            cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    Enc enc = (Enc)fn;
    struct Case { const char* name; int lw, lh, uw, uh, P, c0, w0, h0, n0, offw, offh; };
    Case cases[] = {
        {"valid conv: start (w=0,h=0,n=0), tap (0,0)", 0, 0, -(kw - 1), -(kh - 1), 32, 0, 0, 0, 0, 0, 0},
        {"valid conv: start (w=3,h=1,n=0), tap (b=2,a=1), channels 16..31", 0, 0, -(kw - 1), -(kh - 1), 32, 16, 3, 1, 0, 2, 1},
        {"valid conv: start near the end of image 0 (w=6,h=5): wraps into image 1", 0, 0, -(kw - 1), -(kh - 1), 32, 0, 6, 5, 0, 4, 4},
        {"valid conv: last image, runs past the end (n=2,h=5,w=0)", 0, 0, -(kw - 1), -(kh - 1), 32, 0, 0, 5, 2, 0, 0},
        {"full conv (padding 4): start (w=-4,h=-4), tap (0,0)", -(kw - 1), -(kh - 1), 0, 0, 32, 0, -(kw - 1), -(kh - 1), 0, 0, 0},
        {"full conv (padding 4): start (w=-4,h=-4), tap (4,4)", -(kw - 1), -(kh - 1), 0, 0, 32, 0, -(kw - 1), -(kh - 1), 0, 4, 4},
        {"full conv: start (w=5,h=2,n=1), tap (1,3)", -(kw - 1), -(kh - 1), 0, 0, 32, 0, 5, 2, 1, 1, 3},
        {"16-pixel column, valid conv, start (w=7,h=0)", 0, 0, -(kw - 1), -(kh - 1), 16, 0, 7, 0, 0, 1, 1},
    };
    double* dout; CK(cudaMalloc(&dout, (1024 * 16 + 1) * 8));
    for (auto& cs : cases) {
        CUtensorMap tm;
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)C * 8, (cuuint64_t)W * C * 8, (cuuint64_t)H * W * C * 8};
        int lo[2] = {cs.lw, cs.lh}, up[2] = {cs.uw, cs.uh};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, dx, dims, strides, lo, up, 16, (cuuint32_t)cs.P, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("== %s : encode rc=%d\n", cs.name, (int)r);
        if (r != CUDA_SUCCESS) continue;
        CK(cudaMemset(dout, 0, (1024 * 16 + 1) * 8));
        probe<<<1, 128, cs.P * 128 + 1024>>>(tm, dout, cs.P, cs.c0, cs.w0, cs.h0, cs.n0, cs.offw, cs.offh);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("   kernel error: %s\n", cudaGetErrorString(e)); return 3; }
        std::vector<double> ho(cs.P * 16 + 1);
        CK(cudaMemcpy(ho.data(), dout, ho.size() * 8, cudaMemcpyDeviceToHost));
        if (ho[cs.P * 16] == -1.0) printf("   TIMED OUT waiting for the mbarrier\n");
        for (int p = 0; p < cs.P; ++p) {
            const double v = ho[p * 16];
            bool consistent = true;
            for (int c = 1; c < 16; ++c) consistent &= (ho[p * 16 + c] == v + c) || (v == 0 && ho[p * 16 + c] == 0) || (v == -7.0 && ho[p * 16 + c] == -7.0);
            if (v == -7.0) printf("   row %2d: untouched\n", p);
            else if (v == 0 && ho[p * 16 + 1] == 0) printf("   row %2d: zero fill\n", p);
            else { long long iv = (long long)v; printf("   row %2d: n=%lld h=%lld w=%lld c=%lld%s\n", p, iv / 1000000, iv / 10000 % 100, iv / 100 % 100, iv % 100, consistent ? "" : "  (row not contiguous!)"); }
        }
    }
    return 0;
}
