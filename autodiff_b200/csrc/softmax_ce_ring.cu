// SoftmaxCEWithLogits (reference autodiff/core/ops.py:243-255) for wide rows, decoupled-load variant.
//
// Why a second kernel: `softmax_ce_wide_kernel` (reduce.cu) keeps a row in registers, so a CTA alternates between a
// load phase and an FP64 phase (one exp per element: ~0.6 of the HBM time at 8192x4096).  The two resident CTAs of an
// SM fall into lock-step - both load, then both compute - and HBM idles during every compute phase (0.67 of the roof).
// Here the loads never stop:
//   * one persistent CTA per SM made of GROUPS independent groups of 128 threads; the rows of the CTA go to the groups
//     round-robin, and every group owns SPG slots of a shared-memory ring (192 KB in total: 3 x 64 KB at 4096 columns);
//   * a row reaches its slot by two 1-D bulk copies (`cp.async.bulk.shared::cluster.global.mbarrier::complete_tx`), logits
//     and labels.  The group drains the slot into registers - the labels are folded into sum(l) and sum(l*z) on the way
//     and never kept - and its thread 0 immediately re-arms the slot with the group's next row, so a slot is idle for
//     ~0.3 us of the ~1.5 us a row costs at the HBM rate and 2-3 rows per SM are in flight at all times;
//   * the dependent chain of a row (max -> exp -> sum -> log) then runs out of registers while the next row streams in,
//     and the groups sit at different points of it, which keeps the FP64 pipe busy.
// Per row and group: two 128-thread named barriers (slot drained; warp results exchanged).  Arithmetic is the same online
// log-sum-exp as the wide kernel: each warp reduces to (max, sum exp(z - max), sum l, sum l*z), the four warp results are
// rescaled to the row max with one exp each;  out = -(sum l*z - lse * sum l).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "adb_internal.h"
#include "adb_math.cuh"

namespace adb {
namespace {

constexpr int SCE_RING_BYTES = 192 * 1024;

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void bar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
template <int GT>
__device__ __forceinline__ void group_sync(int g) {
    asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(GT) : "memory");
}

// GT threads per group, P = 16-byte pairs per thread.  HALVES = 1: a slot holds a whole row (columns <= GT * 2 * P).
// HALVES = 2 (4096 < columns <= 8192): a row travels as two half-rows through the same slot, first GT * 2 * P columns, then
// the rest; the group keeps a running (max, sum exp) per warp across the halves - online log-sum-exp, one extra exp per
// lane and half - and exchanges / finalises once per row.  GROUPS * SPG slots of 2 * half_bytes each.
template <int P, int GROUPS, int SPG, int GT, int HALVES>
__global__ void __launch_bounds__(GROUPS * GT, 1)
softmax_ce_ring_kernel(const double* __restrict__ labels, long long l_sr, const double* __restrict__ logits, long long z_sr,
                       double* __restrict__ out, long long o_s, long long rows, unsigned cols, unsigned half_bytes,
                       int* __restrict__ flag, int exp_tab) {
    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ __align__(8) unsigned long long bars[GROUPS * SPG];
    constexpr int NW = GT / 32;
    constexpr unsigned HALF_COLS = GT * 2 * P;         // columns of the first half when HALVES == 2
    __shared__ double xch[GROUPS][2][4][NW];           // [group][row parity][quantity][warp]
    const uint32_t slot_bytes = 2u * half_bytes;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < GROUPS * SPG; ++s) bar_init(s_u32(bars + s), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    // rows of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...; the j-th of them belongs to group j % GROUPS.  A group walks
    // its rows as UNITS (row k / HALVES, half k % HALVES); unit k uses the group's slot k % SPG
    const long long first = blockIdx.x, step = gridDim.x;
    const long long mine = first < rows ? (rows - first + step - 1) / step : 0;
    const int g = threadIdx.x / GT, t = threadIdx.x % GT, w = t >> 5, lane = t & 31;
    const long long gmine = (mine > g ? (mine - g + GROUPS - 1) / GROUPS : 0) * HALVES;
    unsigned char* const gring = ring + (size_t)g * SPG * slot_bytes;
    const uint32_t gbar = s_u32(bars + g * SPG);
    // bytes of each half: whole 16-byte pairs (the tail pair may reach into the row padding)
    const unsigned cols0 = HALVES == 2 ? HALF_COLS : cols, cols1 = HALVES == 2 ? cols - HALF_COLS : 0u;
    const unsigned bytes0 = (cols0 + 1) / 2 * 16, bytes1 = (cols1 + 1) / 2 * 16;

    auto arm = [&](long long k) {                      // thread 0 of the group: start the copies of the group's k-th unit
        const int s = (int)(k % SPG);
        const int h = (int)(k % HALVES);
        const long long r = first + (g + (k / HALVES) * GROUPS) * step;
        const uint32_t dst = s_u32(gring) + s * slot_bytes;
        const unsigned nb = h == 0 ? bytes0 : bytes1;
        bar_expect_tx(gbar + 8 * s, 2 * nb);
        bulk_load(dst, logits + r * z_sr + h * HALF_COLS, nb, gbar + 8 * s);
        bulk_load(dst + half_bytes, labels + r * l_sr + h * HALF_COLS, nb, gbar + 8 * s);
    };
    if (t == 0) {
        for (long long k = 0; k < SPG && k < gmine; ++k) arm(k);
    }
    double mw = -INFINITY, sew = 0.0, lsum = 0.0, dot = 0.0;      // running state of the current row (per lane; mw per warp)
    for (long long k = 0; k < gmine; ++k) {
        const int s = (int)(k % SPG);
        const int h = (int)(k % HALVES);
        const unsigned ncols = h == 0 ? cols0 : cols1;            // columns of this unit
        bar_wait(gbar + 8 * s, (uint32_t)(k / SPG) & 1u);
        const double2* zs = reinterpret_cast<const double2*>(gring + (size_t)s * slot_bytes);
        const double2* ls = reinterpret_cast<const double2*>(gring + (size_t)s * slot_bytes + half_bytes);
        double z[2 * P];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const unsigned c = 2u * (p * GT + t);
            double2 zv = make_double2(-INFINITY, -INFINITY), lv = make_double2(0.0, 0.0);
            if (c < ncols) { zv = zs[p * GT + t]; lv = ls[p * GT + t]; }
            if (c + 1 >= ncols) { zv.y = -INFINITY; lv.y = 0.0; }     // odd width: the pair's second half is row padding
            z[2 * p] = zv.x; z[2 * p + 1] = zv.y;
            if (c < ncols) dot += lv.x * zv.x;
            if (c + 1 < ncols) dot += lv.y * zv.y;
            lsum += lv.x;
            lsum += lv.y;
        }
        group_sync<GT>(g);                             // the slot is drained: everything the unit needs is in registers
        if (t == 0 && k + SPG < gmine) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            arm(k + SPG);
        }
        double m = mw;
#pragma unroll
        for (int e = 0; e < 2 * P; ++e) m = fmax(m, z[e]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
        // fmax skips NaN, exp_nonpos does not: a NaN logit makes the sum NaN without a per-element test (np.max propagates NaN).
        // A warp that has seen nothing but -inf / padding subtracts 0 instead of its max (-inf - -inf would be NaN): sum = 0.
        const double mm = m == -INFINITY ? 0.0 : m;
        if (HALVES > 1) sew *= (mw == -INFINITY) ? 0.0 : exp_nonpos(mw - m);   // earlier halves, rescaled to the new warp max
        double se = 0.0;
        if (exp_tab) {                                 // (uniform branch; straight-line: the 2P evaluations interleave; padding: exp(-inf) = 0)
#pragma unroll
            for (int e = 0; e < 2 * P; ++e) se += exp_nonpos_tab(z[e] - mm);
        } else {
#pragma unroll
            for (int e = 0; e < 2 * P; ++e) se += exp_nonpos(z[e] - mm);
        }
        sew += se;
        mw = m;
        if (h != HALVES - 1) continue;                 // the row's second half follows in the next unit
        const long long krow = k / HALVES;
        se = sew;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            se += __shfl_xor_sync(0xffffffffu, se, off);
            lsum += __shfl_xor_sync(0xffffffffu, lsum, off);
            dot += __shfl_xor_sync(0xffffffffu, dot, off);
        }
        double (*x)[NW] = xch[g][krow & 1];
        if (lane == 0) { x[0][w] = m; x[1][w] = se; x[2][w] = lsum; x[3][w] = dot; }
        mw = -INFINITY; sew = 0.0; lsum = 0.0; dot = 0.0;
        group_sync<GT>(g);
        if (w == 0) {
            double M = x[0][lane % NW], S = x[1][lane % NW], L = x[2][lane % NW], D = x[3][lane % NW];
            double Mx = M;
#pragma unroll
            for (int off = 1; off < NW; off <<= 1) Mx = fmax(Mx, __shfl_xor_sync(0xffffffffu, Mx, off));
            S = (M == -INFINITY) ? S : S * exp_nonpos(M - Mx);      // (S is 0, or NaN when the warp saw a NaN, for M = -inf)
#pragma unroll
            for (int off = 1; off < NW; off <<= 1) {
                S += __shfl_xor_sync(0xffffffffu, S, off);
                L += __shfl_xor_sync(0xffffffffu, L, off);
                D += __shfl_xor_sync(0xffffffffu, D, off);
            }
            if (lane == 0) {
                if (Mx == -INFINITY) S = NAN;          // a row of -inf: numpy computes exp(-inf - -inf) = nan
                const double lse = Mx + log(S);
                out[(first + (g + krow * GROUPS) * step) * o_s] = -(D - lse * L);
                if (!(fabs(L - 1.0) <= 1e-8 + 1e-5)) atomicOr(flag, 1);
            }
        }
    }
}

template <int P, int GROUPS, int SPG, int GT, int HALVES>
cudaError_t launch_ring(const adb_tensor* labels, const adb_tensor* logits, const adb_tensor* out, int* flag, cudaStream_t st,
                        unsigned grid, unsigned row_bytes) {
    static bool attr_set = false;                      // one host thread per device
    const int smem = GROUPS * SPG * 2 * (int)row_bytes;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(softmax_ce_ring_kernel<P, GROUPS, SPG, GT, HALVES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SCE_RING_BYTES);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    softmax_ce_ring_kernel<P, GROUPS, SPG, GT, HALVES><<<grid, GROUPS * GT, smem, st>>>(
        static_cast<const double*>(labels->ptr), labels->stride[0], static_cast<const double*>(logits->ptr), logits->stride[0],
        static_cast<double*>(out->ptr), out->stride[0], logits->shape[0], (unsigned)logits->shape[1], row_bytes, flag, sce_exp_tab());
    return cudaGetLastError();
}

// ----------------------------------------------------------------------------- narrow rows (128 < cols <= 512)
// The same decoupling for rows of 1-4 KB: one persistent CTA per SM, every WARP an independent consumer that owns SPW slots of
// the ring, a slot = one row (logits + labels, two bulk copies on the slot's mbarrier).  The warp drains a slot into registers
// (K quads per lane), its lane 0 re-arms the slot with the warp's next row, then the dependent chain (max -> exp -> sum -> log)
// runs out of registers while WARPS x SPW rows per SM keep streaming in.  The register-only warp kernel (reduce.cu) alternates
// load and FP64 phases with 16 warps per SM: 0.71 of the HBM roof at 512 columns.  Arithmetic: the warp kernel's.
template <int K>
__global__ void __launch_bounds__(512, 1)
softmax_ce_ring_narrow_kernel(const double* __restrict__ labels, long long l_sr, const double* __restrict__ logits, long long z_sr,
                              double* __restrict__ out, long long o_s, long long rows, unsigned cols, unsigned half_bytes, unsigned copy_bytes,
                              int spw, int* __restrict__ flag, int exp_tab) {
    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ __align__(8) unsigned long long bars[64];
    const int warps = blockDim.x >> 5, w = threadIdx.x >> 5;
    const unsigned lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < warps * spw; ++s) bar_init(s_u32(bars + s), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const long long first = (long long)blockIdx.x * warps + w, step = (long long)gridDim.x * warps;
    const long long mine = first < rows ? (rows - first + step - 1) / step : 0;
    const uint32_t slot_bytes = 2u * half_bytes;
    unsigned char* const wring = ring + (size_t)w * spw * slot_bytes;
    const uint32_t wbar = s_u32(bars + w * spw);
    auto arm = [&](long long k) {                      // lane 0: start the copies of the warp's k-th row
        const int s = (int)(k % spw);
        const long long r = first + k * step;
        const uint32_t dst = s_u32(wring) + s * slot_bytes;
        bar_expect_tx(wbar + 8 * s, 2 * copy_bytes);
        bulk_load(dst, logits + r * z_sr, copy_bytes, wbar + 8 * s);
        bulk_load(dst + half_bytes, labels + r * l_sr, copy_bytes, wbar + 8 * s);
    };
    if (lane == 0)
        for (long long k = 0; k < spw && k < mine; ++k) arm(k);
    for (long long k = 0; k < mine; ++k) {
        const int s = (int)(k % spw);
        bar_wait(wbar + 8 * s, (uint32_t)(k / spw) & 1u);
        const double2* zs = reinterpret_cast<const double2*>(wring + (size_t)s * slot_bytes);
        const double2* ls = reinterpret_cast<const double2*>(wring + (size_t)s * slot_bytes + half_bytes);
        double z[4 * K];
        double lsum = 0.0, dot = 0.0;
#pragma unroll
        for (int q = 0; q < 2 * K; ++q) {              // 16-byte pairs, lane-interleaved: conflict-free LDS.128
            const unsigned c = 2u * (q * 32u + lane);
            double2 zv = make_double2(-INFINITY, -INFINITY), lv = make_double2(0.0, 0.0);
            if (c < cols) { zv = zs[q * 32 + lane]; lv = ls[q * 32 + lane]; }
            if (c + 1 >= cols) { zv.y = -INFINITY; lv.y = 0.0; }      // odd width: the pair's second half is row padding
            z[2 * q] = zv.x; z[2 * q + 1] = zv.y;
            if (c < cols) dot += lv.x * zv.x;
            if (c + 1 < cols) dot += lv.y * zv.y;
            lsum += lv.x;
            lsum += lv.y;
        }
        __syncwarp();                                  // the slot is drained
        if (lane == 0 && k + spw < mine) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            arm(k + spw);
        }
        double m = -INFINITY;
#pragma unroll
        for (int e = 0; e < 4 * K; ++e) m = fmax(m, z[e]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
        const double mm = m == -INFINITY ? 0.0 : m;    // (fmax skips NaN, exp_nonpos does not: a NaN logit makes `se` NaN)
        double se = 0.0;
        if (exp_tab) {
#pragma unroll
            for (int e = 0; e < 4 * K; ++e) se += exp_nonpos_tab(z[e] - mm);
        } else {
#pragma unroll
            for (int e = 0; e < 4 * K; ++e) se += exp_nonpos(z[e] - mm);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            se += __shfl_xor_sync(0xffffffffu, se, off);
            lsum += __shfl_xor_sync(0xffffffffu, lsum, off);
            dot += __shfl_xor_sync(0xffffffffu, dot, off);
        }
        if (lane == 0) {
            if (m == -INFINITY) se = NAN;              // a row of -inf: numpy computes exp(-inf - -inf) = nan
            const double lse = m + log(se);
            out[(first + k * step) * o_s] = -(dot - lse * lsum);
            if (!(fabs(lsum - 1.0) <= 1e-8 + 1e-5)) atomicOr(flag, 1);
        }
    }
}

template <int K>
cudaError_t launch_ring_narrow(const adb_tensor* labels, const adb_tensor* logits, const adb_tensor* out, int* flag, cudaStream_t st, unsigned grid) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(softmax_ce_ring_narrow_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, SCE_RING_BYTES);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const unsigned cols = (unsigned)logits->shape[1];
    const unsigned copy_bytes = (cols + 1) / 2 * 16, half_bytes = (copy_bytes + 127) / 128 * 128;
    // as many (warp, slot) pairs as fit the 192 KB ring: 12 warps x 2 slots at 512 columns, 16 x 3 at 256, 16 x 4 below
    int warps = 16, spw = (int)(SCE_RING_BYTES / (16 * 2 * half_bytes));
    if (spw < 2) { warps = 12; spw = (int)(SCE_RING_BYTES / (12 * 2 * half_bytes)); }
    if (spw > 4) spw = 4;
    if (spw < 1) return cudaErrorInvalidValue;
    const int smem = warps * spw * 2 * (int)half_bytes;
    softmax_ce_ring_narrow_kernel<K><<<grid, warps * 32, smem, st>>>(
        static_cast<const double*>(labels->ptr), labels->stride[0], static_cast<const double*>(logits->ptr), logits->stride[0],
        static_cast<double*>(out->ptr), out->stride[0], logits->shape[0], cols, half_bytes, copy_bytes, spw, flag, sce_exp_tab());
    return cudaGetLastError();
}

}  // namespace

// Narrow rows through the ring: 0 = launched, 1 = error, -1 = not applicable.  Same preconditions as softmax_ce_ring_try.
int softmax_ce_ring_narrow_try(const adb_tensor* labels, const adb_tensor* logits, const adb_tensor* out, int* flag, cudaStream_t st) {
    const long long rows = logits->shape[0], cols = logits->shape[1];
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    static int lo = -1;                            // ADB_SCE_NARROW_RING=0 disables; =N sets the smallest width (A/B measurements)
    if (lo < 0) { const char* e = getenv("ADB_SCE_NARROW_RING"); lo = e ? atoi(e) : 257; if (e && lo <= 0) lo = 1 << 30; }
    // (measured, profiles/r02_softmax_ce_warp.log: ring / register-only warp kernel = 0.76 / 0.70 of the HBM roof at 512 columns,
    // 0.64 / 0.55 at 384, but 0.65 / 0.73 at 256 and 0.52 / 0.59 at 200: 2 KB copies no longer amortise the slot round trip)
    if (cols < lo || cols > 512 || rows < 64LL * sms) return -1;      // a persistent grid needs many rows per warp
    const unsigned grid = (unsigned)sms;
    cudaError_t e;
    if (cols <= 128) e = launch_ring_narrow<1>(labels, logits, out, flag, st, grid);
    else if (cols <= 256) e = launch_ring_narrow<2>(labels, logits, out, flag, st, grid);
    else e = launch_ring_narrow<4>(labels, logits, out, flag, st, grid);
    if (e != cudaSuccess) return set_error("softmax_ce narrow ring launch failed: %s", cudaGetErrorString(e));
    count_launch();
    return 0;
}

// Returns 0 = launched, 1 = error (set_error called), -1 = not applicable (the caller uses the register-resident kernels).
// Preconditions checked by the caller: 2-D, unit column stride, 32-byte aligned bases, row strides multiples of 4.
int softmax_ce_ring_try(const adb_tensor* labels, const adb_tensor* logits, const adb_tensor* out, int* flag, cudaStream_t st) {
    const long long rows = logits->shape[0], cols = logits->shape[1];
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    // Measured per width (tools/bench_sce.py, profiles/r01_softmax_ce_ab.log): the ring wins where a row is too big for two
    // resident register-kernel CTAs to overlap (2048 < cols <= 4096: 0.83-0.87 vs 0.60-0.75 of the HBM roof); up to 2048 columns
    // the register-resident kernel is as fast or faster (0.87 at 2048); beyond 4096 a whole row no longer fits a slot three
    // times, so up to 8192 columns a row goes through as two half-rows.
    // A persistent grid also needs several rows per SM.
    if (cols <= 2048 || cols > 8192 || rows < 4LL * sms) return -1;
    const unsigned grid = (unsigned)(rows < sms ? rows : sms);
    cudaError_t e;                                                    // slots of 2 * half_bytes fill the 192 KB ring
    if (cols <= 3072) e = launch_ring<12, 4, 1, 128, 1>(labels, logits, out, flag, st, grid, (unsigned)((cols + 1) / 2 * 16));   // 4 x 48 KB
    else if (cols <= 4096) e = launch_ring<16, 3, 1, 128, 1>(labels, logits, out, flag, st, grid, (unsigned)((cols + 1) / 2 * 16));   // 3 x 64 KB
    else if (cols <= 6144) e = launch_ring<12, 4, 1, 128, 2>(labels, logits, out, flag, st, grid, 3072u * 8u);     // 4 x 48 KB, a row = two half-rows
    else e = launch_ring<16, 3, 1, 128, 2>(labels, logits, out, flag, st, grid, 4096u * 8u);          // 3 x 64 KB, a row = two half-rows
    if (e != cudaSuccess) return set_error("softmax_ce ring launch failed: %s", cudaGetErrorString(e));
    count_launch();
    return 0;
}

}  // namespace adb
