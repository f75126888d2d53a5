// FP64 strided-batched GEMM for sm_100a — the contraction kernel behind Einsum.forward
// (replaces np.einsum at reference autodiff/core/ops.py:180 for true contractions).
//
//   C[b][m][n] = sum_k A[b][m][k] * B[b][k][n]        (all strides in elements, any of NN/NT/TN/TT)
//
// Fast path (gemm_tma_kernel): TMA (cp.async.bulk.tensor, 128B swizzle) -> 6-stage shared-memory ring
// guarded by mbarriers -> 8 warps issuing DMMA.8x8x4 (mma.sync.m8n8k4.f64) on 64x32 warp tiles ->
// direct register epilogue.  Lane 0 of warp 0 doubles as the TMA producer (4 k-tiles ahead).  tcgen05 has no f64 kind
// (ptxas rejects .kind::f64), so on B200 the FP64 tensor path IS the warp-level DMMA; measured issue
// peak 37.0 TFLOP/s (profiles/r01_probe_fp64_issue_rates.log).
// Stream-K (gemm_sk_kernel): the same loop for the full waves of tiles, plus CTAs that share the ragged last wave along K with a
// last-arriver fix-up in k order - one launch, no wave quantisation, bitwise repeatable (profiles/r02_gemm_streamk.md).
// Implicit-GEMM convolution (CONV = 1, conv_wgrad_kernel): the A operand is an NHWC image read through im2col-mode tensor maps.
//
// Shared-memory fragment addressing is conflict-free under the TMA 128B swizzle for both operand
// layouts (see DESIGN.md "GEMM fragment maps"):
//   K-contiguous operand ("KC"): tile = [128 rows][16 k] (one TMA box), one LDS.128 yields the k-pair
//        (8h+2q, 8h+2q+1) used by two consecutive k4-steps; MMA row g <-> tile row (g>>1)|((g&1)<<2).
//   MN-contiguous operand ("MC"): tile = 8 blocks of [16 k][16 mn] (eight TMA boxes), one LDS.128 yields
//        two interleaved 8-wide MMA tiles (columns 2g and 2g+1) for one k4-step.
//   Both use the k-slot map  q -> k = 8h + 2q + p  so A and B fragments always agree.
//
// Fallback (gemm_generic_kernel): any stride / any alignment, 64x64 DFMA register tiling.  Used only
// when an operand cannot be described by a TMA tensor map (unaligned base or pitch).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "adb_internal.h"

namespace adb {

constexpr int BK = 16;
// Tile configuration: 8 warps as 2 (M) x 4 (N); warp tile = (BM/2) x (BN/4) = MT x NT fragments of 8x8.
//   Big   128x128, 6 stages, 1 CTA/SM  : large contractions (C2a/C2b/C3)
//   Small  64x64,  6 stages, 2 CTAs/SM : problems with fewer than one wave of 128x128 tiles (char-rnn steps)
template <int BM_, int BN_, int STAGES_, int MINB_, int WM_ = 2, int WN_ = 4, int PREFETCH_ = STAGES_ - 2>
struct TileCfg {
    static constexpr int BM = BM_, BN = BN_, STAGES = STAGES_, MINB = MINB_;
    static constexpr int WM = WM_, WN = WN_;                  // warp grid (WM x WN = 8 consumer warps)
    static constexpr int MT = BM_ / (8 * WM_), NT = BN_ / (8 * WN_);   // 8-wide MMA fragments per warp
    static_assert(WM_ * WN_ == 8 && MT % 2 == 0 && NT % 2 == 0, "warp grid must use 8 warps and even fragment counts");
    static constexpr int A_TILE_BYTES = BM_ * BK * 8;
    static constexpr int B_TILE_BYTES = BN_ * BK * 8;
    static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
    // k-tiles in flight ahead of the consumers.  The slot of k-tile t is released when a warp has passed the full-barrier wait of
    // t+1 (see the main loop), so the slot the producer refills at step kt (tile kt + PREFETCH, last used by kt + PREFETCH - STAGES)
    // is free once every warp has started step kt + PREFETCH - STAGES + 1: STAGES - PREFETCH - 1 steps of slack between warps.
    static constexpr int PREFETCH = PREFETCH_;
    static constexpr int SMEM = STAGES_ * STAGE_BYTES + 1024 /*align slack*/ + 2 * STAGES_ * 8;
};
#ifndef ADB_BIG_STAGES          // (compile-time knobs for the A/B builds of tools/gemm_variants.sh)
#define ADB_BIG_STAGES 6
#endif
#ifndef ADB_BIG_PREFETCH
#define ADB_BIG_PREFETCH (ADB_BIG_STAGES - 2)
#endif
#ifndef ADB_SMALL_PREFETCH
#define ADB_SMALL_PREFETCH 4
#endif
using CfgBig = TileCfg<128, 128, ADB_BIG_STAGES, 1, 2, 4, ADB_BIG_PREFETCH>;
using CfgSmall = TileCfg<64, 64, 6, 2, 2, 4, ADB_SMALL_PREFETCH>;
// Tall-and-skinny problems (im2col convolutions: M = batch*OH*OW, N = output channels <= 32): all 8 warps stack along M
// and share the whole 32-column B tile, so no MMA work is spent on padding columns.
using CfgSkinny = TileCfg<256, 32, 6, 1, 8, 1>;
// Few rows AND few columns with a long K (weight gradients of the first conv layer: 75 x 16 x 921600): 128x32 tiles,
// two CTAs per SM, split-K supplies the parallelism.
using CfgSkinnyS = TileCfg<128, 32, 6, 2, 8, 1>;
// Implicit-GEMM convolution onto <= 16 output channels (the data gradient of a conv layer with 16 input channels).
using CfgSkinny16 = TileCfg<256, 16, 6, 1, 8, 1>;
constexpr int CONSUMER_WARPS = 8;
constexpr int GEMM_THREADS = CONSUMER_WARPS * 32;   // 9 warps would round to 12 for register allocation (168 regs)

// ----------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// TMA im2col mode (tensor (C, W, H, N)): `pixels` consecutive positions of the bounding box starting at (w, h, n), each read at
// (w + off_w, h + off_h), 16 channels from c on - one 128-byte row per position (semantics pinned by tools/probe_im2col.cu,
// profiles/r02_probe_im2col.log: wraps W -> H -> N inside the box, zero-fills outside the tensor).
__device__ __forceinline__ void tma_load_im2col(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w, int h, int n, int off_w, int off_h) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"((unsigned short)off_w), "h"((unsigned short)off_h)
        : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}
__device__ __forceinline__ double2 lds128(const unsigned char* p) { return *reinterpret_cast<const double2*>(p); }

struct GemmParams {
    int M, N, K;
    int tiles_m, tiles_n;
    int kt_per_split;            // k-tiles each blockIdx.z handles (split-K); == all k-tiles when gridDim.z == 1
    int blocked;                 // bit 0 / 1: the MN-contiguous A / B operand has a BLOCKED tensor map (one TMA box per tile)
    double* C;
    long long c_sm, c_sn, c_sb, c_ss;   // c_ss: stride between split-K partial results
    // implicit-GEMM convolution (CONV kernels): operand A is never materialised, tmA is an im2col-mode tensor map of the image
    int cv_wb, cv_hb;            // positions per row / rows per image of the bounding box the windows slide over
    int cv_lw, cv_lh;            // its lower corner (0 for a VALID convolution, -(k-1) for the data gradient's full convolution)
    int cv_kw, cv_cpt;           // taps per filter row; k-tiles (16 channels) per tap
    // stream-K kernel (gemm_sk_kernel): tiles [0, dp_tiles) are taken whole, one per CTA; the k-tiles ("units") of the remaining
    // tiles are dealt out evenly and contiguously to the last sk_ctas CTAs of the grid
    int total_kt, batch;
    long long tiles_total, dp_tiles, sk_units;
    int sk_ctas;
    double* sk_ws;               // [2 * sk_ctas] partial tiles, thread-major (see gemm_sk_kernel)
    unsigned* sk_count;          // [tiles_total - dp_tiles] arrivals per split tile (zeroed by the host)
};

// One k-tile (BK = 16) of a warp tile: fragment loads (conflict-free under the 128B swizzle, see the header) + MT x NT x 4 DMMAs.
template <class Cfg, bool A_KC, bool B_KC>
__device__ __forceinline__ void ktile_mma(const unsigned char* sA, const unsigned char* sB, int g, int q, int rmap,
                                          double (&acc)[Cfg::MT][Cfg::NT][2]) {
    constexpr int MT = Cfg::MT, NT = Cfg::NT;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        double a[2][MT], b[2][NT];
        if (A_KC) {
            const int ch = ((4 * h + q) ^ rmap) << 4;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const double2 v = lds128(sA + mt * 1024 + ch);
                a[0][mt] = v.x;
                a[1][mt] = v.y;
            }
        } else {
#pragma unroll
            for (int pp = 0; pp < 2; ++pp) {
                const int kr = 8 * h + 2 * q + pp;
                const int off = kr * 128 + ((g ^ (kr & 7)) << 4);
#pragma unroll
                for (int blk = 0; blk < MT / 2; ++blk) {
                    const double2 v = lds128(sA + blk * 2048 + off);
                    a[pp][2 * blk] = v.x;
                    a[pp][2 * blk + 1] = v.y;
                }
            }
        }
        if (B_KC) {
            const int ch = ((4 * h + q) ^ rmap) << 4;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const double2 v = lds128(sB + nt * 1024 + ch);
                b[0][nt] = v.x;
                b[1][nt] = v.y;
            }
        } else {
#pragma unroll
            for (int pp = 0; pp < 2; ++pp) {
                const int kr = 8 * h + 2 * q + pp;
                const int off = kr * 128 + ((g ^ (kr & 7)) << 4);
#pragma unroll
                for (int blk = 0; blk < NT / 2; ++blk) {
                    const double2 v = lds128(sB + blk * 2048 + off);
                    b[pp][2 * blk] = v.x;
                    b[pp][2 * blk + 1] = v.y;
                }
            }
        }
#pragma unroll
        for (int pp = 0; pp < 2; ++pp)
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[pp][mt], b[pp][nt]);
    }
}

// Register epilogue: accumulators -> C (general strides; 16B vector stores when the n-stride is 1).
template <class Cfg, bool A_KC, bool B_KC>
__device__ __forceinline__ void store_acc(const double (&acc)[Cfg::MT][Cfg::NT][2], double* __restrict__ Cb, long long c_sm, long long c_sn,
                                          int m0, int n0, int M, int N, int warp_m, int warp_n, int g, int q, int rmap) {
    constexpr int BM = Cfg::BM, BN = Cfg::BN, MT = Cfg::MT, NT = Cfg::NT, WM = Cfg::WM, WN = Cfg::WN;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const int row = m0 + warp_m * (BM / WM) + (A_KC ? mt * 8 + rmap : (mt >> 1) * 16 + 2 * g + (mt & 1));
        if (row >= M) continue;
        double* crow = Cb + (long long)row * c_sm;
        if (B_KC) {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int col = n0 + warp_n * (BN / WN) + nt * 8 + q + 4 * e;
                    if (col < N) crow[(long long)col * c_sn] = acc[mt][nt][e];
                }
            }
        } else {
#pragma unroll
            for (int blk = 0; blk < NT / 2; ++blk) {
                // thread owns 4 consecutive columns: 4q + {0,1,2,3} = (e=0,j=0),(e=0,j=1),(e=1,j=0),(e=1,j=1)
                const int col = n0 + warp_n * (BN / WN) + blk * 16 + 4 * q;
                const double v0 = acc[mt][2 * blk][0], v1 = acc[mt][2 * blk + 1][0];
                const double v2 = acc[mt][2 * blk][1], v3 = acc[mt][2 * blk + 1][1];
                double* cp = crow + (long long)col * c_sn;
                if (c_sn == 1 && col + 3 < N && ((reinterpret_cast<uintptr_t>(cp) & 15) == 0)) {
                    reinterpret_cast<double2*>(cp)[0] = make_double2(v0, v1);
                    reinterpret_cast<double2*>(cp)[1] = make_double2(v2, v3);
                } else {
                    if (col < N) cp[0] = v0;
                    if (col + 1 < N) cp[c_sn] = v1;
                    if (col + 2 < N) cp[2 * c_sn] = v2;
                    if (col + 3 < N) cp[3 * c_sn] = v3;
                }
            }
        }
    }
}

// ----------------------------------------------------------------------------- TMA + DMMA kernel
// CONV = 1: A is the im2col matrix of an NHWC image, read straight from the image by im2col-mode TMA:
//   A_KC  (forward / data gradient): rows = window positions, k = (tap, channel): one BM-position column per k-tile;
//   !A_KC (weight gradient):         k = window positions, m = (tap, channel): BM/16 blocks of [16 positions][16 channels].
template <class Cfg, bool A_KC, bool B_KC, int CONV = 0>
__global__ void __launch_bounds__(GEMM_THREADS, Cfg::MINB)
gemm_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    constexpr int BM = Cfg::BM, BN = Cfg::BN, STAGES = Cfg::STAGES, MT = Cfg::MT, NT = Cfg::NT;
    constexpr int A_TILE_BYTES = Cfg::A_TILE_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES, PREFETCH = Cfg::PREFETCH;
    extern __shared__ unsigned char smem_raw[];
    // keep the pointer derived from the __shared__ array so loads compile to LDS (not generic LD)
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    const uint32_t full0 = smem_u32(bars);
    const uint32_t empty0 = smem_u32(bars + STAGES);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // grouped rasterisation: 8 tile-rows share their B column panels in L2
    const int GROUP = 8;
    const int tile = blockIdx.x;
    const int group_size = GROUP * p.tiles_n;
    const int first_m = (tile / group_size) * GROUP;
    const int gm = min(p.tiles_m - first_m, GROUP);
    const int tm = first_m + (tile % group_size) % gm;
    const int tn = (tile % group_size) / gm;
    const int m0 = tm * BM, n0 = tn * BN;
    const int batch = blockIdx.y;
    const int total_kt = (p.K + BK - 1) / BK;
    const int kt_first = blockIdx.z * p.kt_per_split;            // split-K: this CTA's slice of the k-tiles
    const int num_kt = min(p.kt_per_split, total_kt - kt_first);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    // ===== TMA producer: lane 0 of warp 0 keeps PREFETCH k-tiles in flight (issued between its own MMAs) =====
    int cv_n = 0, cv_h = 0, cv_w = 0, cv_blocks = BM / 16;
    if (CONV && threadIdx.x == 0) {
        if (A_KC) {                                    // first window position of this row tile
            const int per_img = p.cv_hb * p.cv_wb;
            cv_n = m0 / per_img;
            const int r = m0 - cv_n * per_img;
            cv_h = r / p.cv_wb;
            cv_w = r - cv_h * p.cv_wb;
        } else {                                       // (tap, channel-chunk) blocks of this tile that exist
            const int left = (p.M - m0 + 15) / 16;
            cv_blocks = left < BM / 16 ? left : BM / 16;
        }
    }
    auto issue_tile = [&](int kt) {
        const int s = kt % STAGES;
        const uint32_t ph = (kt / STAGES) & 1;
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        const uint32_t fb = full0 + 8 * s;
        mbar_expect_tx(fb, (CONV && !A_KC) ? cv_blocks * 2048 + Cfg::B_TILE_BYTES : STAGE_BYTES);
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t sb = sa + A_TILE_BYTES;
        const int k0 = (kt_first + kt) * BK;
        if (CONV && A_KC) {
            const int kk = kt_first + kt;
            const int tap = kk / p.cv_cpt, c0 = (kk - tap * p.cv_cpt) * 16;
            const int ta = tap / p.cv_kw, tb = tap - ta * p.cv_kw;
            tma_load_im2col(sa, &tmA, fb, c0, cv_w + p.cv_lw, cv_h + p.cv_lh, cv_n, tb, ta);
        } else if (CONV) {
            const int per_img = p.cv_hb * p.cv_wb;
            const int n = k0 / per_img;
            const int r = k0 - n * per_img;
            const int h = r / p.cv_wb, w = r - h * p.cv_wb;
            for (int b = 0; b < cv_blocks; ++b) {
                const int blk = m0 / 16 + b;
                const int tap = blk / p.cv_cpt, c0 = (blk - tap * p.cv_cpt) * 16;
                const int ta = tap / p.cv_kw, tb = tap - ta * p.cv_kw;
                tma_load_im2col(sa + b * 2048, &tmA, fb, c0, w + p.cv_lw, h + p.cv_lh, n, tb, ta);
            }
        } else if (A_KC) {
            tma_load_3d(sa, &tmA, fb, k0, m0, batch);
        } else if (p.blocked & 1) {
            tma_load_4d(sa, &tmA, fb, 0, k0, m0 / 16, batch);          // the BM/16 blocks of [16 k][16 m] in ONE box
        } else {
#pragma unroll
            for (int b = 0; b < BM / 16; ++b) tma_load_3d(sa + b * 2048, &tmA, fb, m0 + 16 * b, k0, batch);
        }
        if (B_KC) {
            tma_load_3d(sb, &tmB, fb, k0, n0, batch);
        } else if (p.blocked & 2) {
            tma_load_4d(sb, &tmB, fb, 0, k0, n0 / 16, batch);
        } else {
#pragma unroll
            for (int b = 0; b < BN / 16; ++b) tma_load_3d(sb + b * 2048, &tmB, fb, n0 + 16 * b, k0, batch);
        }
    };
    const bool is_producer = threadIdx.x == 0;
    if (is_producer) {
        const int npre = num_kt < PREFETCH ? num_kt : PREFETCH;
        for (int kt = 0; kt < npre; ++kt) issue_tile(kt);
    }

    // ===== consumers: WM (M) x WN (N) warps, warp tile (BM/WM) x (BN/WN) =====
    constexpr int WM = Cfg::WM, WN = Cfg::WN;
    const int warp_m = warp % WM;
    const int warp_n = warp / WM;
    const int g = lane >> 2;
    const int q = lane & 3;
    const int rmap = (g >> 1) | ((g & 1) << 2);

    double acc[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // per-thread byte offsets inside a stage
    const int a_off = A_KC ? (warp_m * (BM / WM) + rmap) * 128 : warp_m * (MT / 2) * 2048;
    const int b_off = A_TILE_BYTES + (B_KC ? (warp_n * (BN / WN) + rmap) * 128 : warp_n * (NT / 2) * 2048);

    for (int kt = 0; kt < num_kt; ++kt) {
        if (is_producer && kt + PREFETCH < num_kt) issue_tile(kt + PREFETCH);
        __syncwarp();
        const int s = kt % STAGES;
        const uint32_t ph = (kt / STAGES) & 1;
        mbar_wait(full0 + 8 * s, ph);
        // Release of the PREVIOUS k-tile's slot.  A slot may go back to the TMA producer only once every LDS that reads it has
        // COMPLETED, not merely issued.  Arriving at the end of the iteration is not enough: ptxas hoists the arrive above the
        // last MMAs, right behind the LDS issue, and under heavy L1/LSU traffic from a kernel that shares the SM (the side-stream
        // Adam update; tools/gemm_stress.cu reproduces it with any HBM-bound kernel) a queued LDS was overtaken by the refill
        // of its slot - one warp then multiplied a row fragment of the NEXT k-tile (the round-1 data-parallel parity failure).
        // Here every MMA of iteration kt-1 has issued (the spin loop above is a basic-block boundary, and an MMA issues only
        // when its LDS operands have landed), so the arrive can be neither early nor hoisted.
        if (kt > 0) {
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * ((kt - 1) % STAGES));
        }
        const unsigned char* st = smem + s * STAGE_BYTES;
        ktile_mma<Cfg, A_KC, B_KC>(st + a_off, st + b_off, g, q, rmap, acc);
    }

    // ===== epilogue: registers -> C (general strides; 16B vector stores when the n-stride is 1) =====
    double* __restrict__ Cb = p.C + (long long)batch * p.c_sb + (long long)blockIdx.z * p.c_ss;
    store_acc<Cfg, A_KC, B_KC>(acc, Cb, p.c_sm, p.c_sn, m0, n0, p.M, p.N, warp_m, warp_n, g, q, rmap);
}

// ----------------------------------------------------------------------------- stream-K kernel
// ONE launch without wave quantisation.  Tiles in raster order; grid = dp_tiles + sk_ctas CTAs:
//   * CTA x < dp_tiles computes tile x whole, with the loop of gemm_tma_kernel (dp_tiles is a multiple of the SM slots: these are
//     the full waves, and the hardware's dynamic CTA scheduler keeps dealing them out as SMs free up);
//   * the k-tiles ("units") of the remaining tiles - the wave that would fill only part of the machine - are dealt out evenly and
//     contiguously to the last sk_ctas CTAs (<= one per SM slot, so they all run at once when the last full wave drains).  A CTA's
//     share is a run of SEGMENTS (tile, [lo, hi)): at most one that starts inside a tile, whole tiles, at most one that ends inside
//     a tile.  Its smem ring, mbarriers and producer run across segment boundaries, so the loads of the next segment are in flight
//     while the accumulators of this one are stored.
// A partial segment parks its accumulators thread-major in its own slot of the workspace (slot 2c for a segment of stream-K CTA c
// that starts inside a tile, 2c + 1 for one that starts a tile), then bumps the tile's arrival counter; the CTA that arrives last
// re-adds ALL parts of the tile in k order - the order, and therefore the rounding, does not depend on which CTA arrives last -
// and stores the tile.  Nobody waits for anybody, so the kernel cannot deadlock whatever else shares the GPU.
// (A fully persistent variant - whole tiles round-robin inside the same CTAs - measured 2 % slower per tile than relaunching CTAs:
// profiles/r02_gemm_streamk.md.)
struct SkCursor {         // (32-bit: the host keeps tiles_total and sk_units below 2^31)
    int u, ue;            // units of the stream-K region still to take
    int tile;             // current segment
    int lo, hi;
};
__device__ __forceinline__ int sk_unit_begin(const GemmParams& p, int cta) {
    return (int)((long long)cta * p.sk_units / p.sk_ctas);
}
__device__ __forceinline__ bool sk_advance(SkCursor& c, const GemmParams& p) {
    if (c.u >= c.ue) return false;
    const int t = c.u / p.total_kt;
    c.lo = c.u - t * p.total_kt;
    const int left = c.ue - c.u;
    c.hi = (left < p.total_kt - c.lo) ? c.lo + left : p.total_kt;
    c.tile = (int)p.dp_tiles + t;
    c.u += c.hi - c.lo;
    return true;
}
// stream-K CTA whose share holds unit u
__device__ __forceinline__ int sk_owner(const GemmParams& p, int u) {
    int c = (int)((long long)u * p.sk_ctas / p.sk_units);
    if (c >= p.sk_ctas) c = p.sk_ctas - 1;
    while (sk_unit_begin(p, c) > u) --c;
    while (sk_unit_begin(p, c + 1) <= u) ++c;
    return c;
}
template <int BM, int BN>
__device__ __forceinline__ void sk_tile_origin(const GemmParams& p, int tile, int& m0, int& n0, int& batch) {
    const int GROUP = 8;
    const int tiles2d = p.tiles_m * p.tiles_n;
    batch = tile / tiles2d;
    const int t2 = tile - batch * tiles2d;
    const int group_size = GROUP * p.tiles_n;
    const int first_m = (t2 / group_size) * GROUP;
    const int gm = min(p.tiles_m - first_m, GROUP);
    m0 = (first_m + (t2 % group_size) % gm) * BM;
    n0 = ((t2 % group_size) / gm) * BN;
}

template <class Cfg, bool A_KC, bool B_KC>
__global__ void __launch_bounds__(GEMM_THREADS, Cfg::MINB)
gemm_sk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    constexpr int BM = Cfg::BM, BN = Cfg::BN, STAGES = Cfg::STAGES, MT = Cfg::MT, NT = Cfg::NT;
    constexpr int A_TILE_BYTES = Cfg::A_TILE_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES, PREFETCH = Cfg::PREFETCH;
    constexpr int ACC2 = MT * NT;                        // double2 values per thread
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    const uint32_t full0 = smem_u32(bars);
    const uint32_t empty0 = smem_u32(bars + STAGES);
    volatile int* last_flag = reinterpret_cast<volatile int*>(bars + 2 * STAGES);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    constexpr int WM = Cfg::WM, WN = Cfg::WN;
    const int warp_m = warp % WM;
    const int warp_n = warp / WM;
    const int g = lane >> 2;
    const int q = lane & 3;
    const int rmap = (g >> 1) | ((g & 1) << 2);
    const int a_off = A_KC ? (warp_m * (BM / WM) + rmap) * 128 : warp_m * (MT / 2) * 2048;
    const int b_off = A_TILE_BYTES + (B_KC ? (warp_n * (BN / WN) + rmap) * 128 : warp_n * (NT / 2) * 2048);
    const bool is_producer = threadIdx.x == 0;

    // one k-tile of tile (m0, n0, batch) into ring slot (it % STAGES)
    auto issue = [&](int it, int kt, int m0, int n0, int batch) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        const uint32_t fb = full0 + 8 * s;
        mbar_expect_tx(fb, STAGE_BYTES);
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t sb = sa + A_TILE_BYTES;
        const int k0 = kt * BK;
        if (A_KC) {
            tma_load_3d(sa, &tmA, fb, k0, m0, batch);
        } else if (p.blocked & 1) {
            tma_load_4d(sa, &tmA, fb, 0, k0, m0 / 16, batch);
        } else {
#pragma unroll
            for (int b = 0; b < BM / 16; ++b) tma_load_3d(sa + b * 2048, &tmA, fb, m0 + 16 * b, k0, batch);
        }
        if (B_KC) {
            tma_load_3d(sb, &tmB, fb, k0, n0, batch);
        } else if (p.blocked & 2) {
            tma_load_4d(sb, &tmB, fb, 0, k0, n0 / 16, batch);
        } else {
#pragma unroll
            for (int b = 0; b < BN / 16; ++b) tma_load_3d(sb + b * 2048, &tmB, fb, n0 + 16 * b, k0, batch);
        }
    };

    double acc[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    if ((long long)blockIdx.x < p.dp_tiles) {
        // ===== a whole tile: the loop of gemm_tma_kernel =====
        int m0, n0, batch;
        sk_tile_origin<BM, BN>(p, (int)blockIdx.x, m0, n0, batch);
        const int num_kt = p.total_kt;
        if (is_producer) {
            const int npre = num_kt < PREFETCH ? num_kt : PREFETCH;
            for (int kt = 0; kt < npre; ++kt) issue(kt, kt, m0, n0, batch);
        }
        for (int kt = 0; kt < num_kt; ++kt) {
            if (is_producer && kt + PREFETCH < num_kt) issue(kt + PREFETCH, kt + PREFETCH, m0, n0, batch);
            __syncwarp();
            const int s = kt % STAGES;
            mbar_wait(full0 + 8 * s, (kt / STAGES) & 1);
            if (kt > 0) {                               // release of the PREVIOUS k-tile's slot: see gemm_tma_kernel
                __syncwarp();
                if (lane == 0) mbar_arrive(empty0 + 8 * ((kt - 1) % STAGES));
            }
            const unsigned char* st = smem + s * STAGE_BYTES;
            ktile_mma<Cfg, A_KC, B_KC>(st + a_off, st + b_off, g, q, rmap, acc);
        }
        store_acc<Cfg, A_KC, B_KC>(acc, p.C + (long long)batch * p.c_sb, p.c_sm, p.c_sn, m0, n0, p.M, p.N, warp_m, warp_n, g, q, rmap);
        return;
    }

    // ===== a share of the stream-K region =====
    const int cta = (int)(blockIdx.x - p.dp_tiles);
    SkCursor pc, cc;                                    // producer (PREFETCH k-tiles ahead) and consumer cursors
    cc.u = pc.u = sk_unit_begin(p, cta);
    cc.ue = pc.ue = sk_unit_begin(p, cta + 1);
    pc.tile = pc.lo = pc.hi = 0;
    int p_kt = 0, p_m0 = 0, p_n0 = 0, p_batch = 0, p_it = 0;
    bool p_more = true;
    auto issue_next = [&]() {
        if (p_kt >= pc.hi) {
            if (!sk_advance(pc, p)) { p_more = false; return; }
            p_kt = pc.lo;
            sk_tile_origin<BM, BN>(p, pc.tile, p_m0, p_n0, p_batch);
        }
        issue(p_it, p_kt, p_m0, p_n0, p_batch);
        ++p_kt;
        ++p_it;
    };
    if (is_producer)
        for (int k = 0; k < PREFETCH && p_more; ++k) issue_next();

    int it = 0;
    bool first = true;
    while (sk_advance(cc, p)) {
        if (!first) {
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        }
        first = false;
        for (int kt = cc.lo; kt < cc.hi; ++kt, ++it) {
            if (is_producer && p_more) issue_next();
            __syncwarp();
            const int s = it % STAGES;
            mbar_wait(full0 + 8 * s, (it / STAGES) & 1);
            if (it > 0) {
                __syncwarp();
                if (lane == 0) mbar_arrive(empty0 + 8 * ((it - 1) % STAGES));
            }
            const unsigned char* st = smem + s * STAGE_BYTES;
            ktile_mma<Cfg, A_KC, B_KC>(st + a_off, st + b_off, g, q, rmap, acc);
        }

        if (cc.lo > 0 || cc.hi < p.total_kt) {
            // ---- partial segment: park the accumulators, count the arrival; the last arriver rebuilds the tile in k order
            double2* slot = reinterpret_cast<double2*>(p.sk_ws) + ((long long)(2 * cta + (cc.lo > 0 ? 0 : 1)) * ACC2) * GEMM_THREADS + threadIdx.x;
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) slot[(i * NT + j) * GEMM_THREADS] = make_double2(acc[i][j][0], acc[i][j][1]);
            __threadfence();
            __syncthreads();
            const int t = cc.tile - (int)p.dp_tiles;
            const int u0 = t * p.total_kt;
            const int c_first = sk_owner(p, u0), c_last = sk_owner(p, u0 + p.total_kt - 1);
            if (threadIdx.x == 0) {
                const unsigned old = atomicAdd(p.sk_count + t, 1u);
                *last_flag = (old == (unsigned)(c_last - c_first)) ? 1 : 0;
            }
            __syncthreads();
            const bool last = *last_flag != 0;
            __syncthreads();                            // (the flag is rewritten by the next partial segment)
            if (!last) continue;
            __threadfence();
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
            for (int c = c_first; c <= c_last; ++c) {
                const double2* part = reinterpret_cast<const double2*>(p.sk_ws) + ((long long)(2 * c + (c == c_first ? 1 : 0)) * ACC2) * GEMM_THREADS + threadIdx.x;
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        const double2 v = __ldcg(part + (i * NT + j) * GEMM_THREADS);
                        acc[i][j][0] += v.x;
                        acc[i][j][1] += v.y;
                    }
            }
        }
        int m0, n0, batch;
        sk_tile_origin<BM, BN>(p, cc.tile, m0, n0, batch);
        store_acc<Cfg, A_KC, B_KC>(acc, p.C + (long long)batch * p.c_sb, p.c_sm, p.c_sn, m0, n0, p.M, p.N, warp_m, warp_n, g, q, rmap);
    }
}

// ----------------------------------------------------------------------------- filter gradient of a convolution
// dW[(a,b,c), o] = sum_{n,oh,ow} x[n, oh+a, ow+b, c] * d[n, oh, ow, o]     (== Einsum('pk,po->ko', Im2Col(x), d))
// As a GEMM over im2col blocks this is bound by the TMA unit, not the tensor pipe: every 128-byte image row (16 channels of one
// pixel) feeds only 16 x 32 products, and each filter tap loads it again (measured 2.43 ms at configs[3], 8.5 TFLOP/s).  Here a
// CTA owns one filter ROW a (and 16 channels x 32 outputs): one image row segment in shared memory serves all kw taps of that
// filter row as kw overlapping views shifted by one 128-byte row each, so per stage the kernel loads SEG + kw - 1 image rows
// and 2 x SEG gradient rows with three plain tiled TMA boxes and issues 2 kw x 2 x SEG/4 DMMAs per warp pair.
//   warps 0-7: consumers; pair (w >> 1) takes every 4th stage (a 4-way split of the position sum whose partial results
//              are separate slabs of the workspace, summed afterwards in a fixed order), w & 1 picks 16 of the 32 outputs
//   warp 8   : TMA producer
// Fragment addressing as in gemm_tma_kernel's MN-contiguous operands: one LDS.128 = two interleaved 8-wide tiles of one k-row.
struct WgradParams {
    int kw, C, O;
    int cblocks, opairs;          // C / 16, ceil(O / 32)
    int OH, nseg, SEG;            // output rows per image, segments per output row, positions per segment (multiple of 8)
    int xrows;                    // image rows per stage: roundup(SEG + kw - 1, 8)
    int chunks, chunks_per_split; // N * OH * nseg
    double* ws;                   // [gridDim.y * 4][kh * kw * C][Np]
    long long slab;               // kh * kw * C * Np
    int Np;
};
constexpr int WG_STAGES = 8;

template <int KW>
__global__ void __launch_bounds__(288, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmD, const WgradParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int x_bytes = p.xrows * 128, d_bytes = p.SEG * 128, stage_bytes = x_bytes + 2 * d_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_STAGES * stage_bytes);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + WG_STAGES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // task = (filter row a, channel block cb, output pair op); blockIdx.y = split of the chunk range
    int t = blockIdx.x;
    const int op = t % p.opairs; t /= p.opairs;
    const int cb = t % p.cblocks;
    const int a = t / p.cblocks;
    const int j0 = blockIdx.y * p.chunks_per_split;
    const int j1 = min(j0 + p.chunks_per_split, p.chunks);
    const int nch = max(j1 - j0, 0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < WG_STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 2);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (warp == 8) {
        if (lane == 0) {
            for (int i = 0; i < nch; ++i) {
                const int s = i % WG_STAGES;
                mbar_wait(empty0 + 8 * s, ((i / WG_STAGES) & 1) ^ 1);
                const uint32_t fb = full0 + 8 * s;
                mbar_expect_tx(fb, stage_bytes);
                const int j = j0 + i;
                const int seg = j % p.nseg;
                const int row = j / p.nseg;               // n * OH + oh
                const int n = row / p.OH, oh = row - n * p.OH;
                const uint32_t sx = smem_u32(smem + s * stage_bytes);
                tma_load_4d(sx, &tmX, fb, cb * 16, seg * p.SEG, oh + a, n);
                tma_load_4d(sx + x_bytes, &tmD, fb, op * 32, seg * p.SEG, oh, n);
                tma_load_4d(sx + x_bytes + d_bytes, &tmD, fb, op * 32 + 16, seg * p.SEG, oh, n);
            }
        }
        return;
    }
    const int ob = warp & 1, wk = warp >> 1;
    const int g = lane >> 2, q = lane & 3;
    double acc[2 * KW][2][2];
#pragma unroll
    for (int i = 0; i < 2 * KW; ++i) acc[i][0][0] = acc[i][0][1] = acc[i][1][0] = acc[i][1][1] = 0.0;
    const int k8s = p.SEG >> 3;
    int prev = -1;
    for (int i = wk; i < nch; i += 4) {
        const int s = i % WG_STAGES;
        mbar_wait(full0 + 8 * s, (i / WG_STAGES) & 1);
        // the previous stage of this warp goes back to the producer only now: every MMA that consumed it has issued, so every
        // LDS that read it has completed (see the main loop of gemm_tma_kernel)
        if (prev >= 0) {
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * prev);
        }
        prev = s;
        const unsigned char* sx = smem + s * stage_bytes;
        const unsigned char* sd = sx + x_bytes + ob * d_bytes;
        for (int h = 0; h < k8s; ++h) {
#pragma unroll
            for (int pp = 0; pp < 2; ++pp) {
                const int kr = 8 * h + 2 * q + pp;
                const double2 bv = lds128(sd + kr * 128 + ((g ^ (kr & 7)) << 4));
                double av[2 * KW];
#pragma unroll
                for (int b = 0; b < KW; ++b) {
                    const int xr = kr + b;
                    const double2 v = lds128(sx + xr * 128 + ((g ^ (xr & 7)) << 4));
                    av[2 * b] = v.x;
                    av[2 * b + 1] = v.y;
                }
#pragma unroll
                for (int m = 0; m < 2 * KW; ++m) {
                    dmma884(acc[m][0][0], acc[m][0][1], av[m], bv.x);
                    dmma884(acc[m][1][0], acc[m][1][1], av[m], bv.y);
                }
            }
        }
    }
    // partial result of this warp: slab (split, wk), rows (a, b, cb*16 + 2g + e), columns op*32 + ob*16 + 4q + {0,1,2,3}
    double* slab = p.ws + ((long long)blockIdx.y * 4 + wk) * p.slab;
    const int col = op * 32 + ob * 16 + 4 * q;
#pragma unroll
    for (int m = 0; m < 2 * KW; ++m) {
        const int b = m >> 1;
        const long long row = ((long long)(a * p.kw + b)) * p.C + cb * 16 + 2 * g + (m & 1);
        double* cp = slab + row * p.Np + col;
        const double v0 = acc[m][0][0], v1 = acc[m][1][0], v2 = acc[m][0][1], v3 = acc[m][1][1];
        if (col + 3 < p.Np) {
            reinterpret_cast<double2*>(cp)[0] = make_double2(v0, v1);
            reinterpret_cast<double2*>(cp)[1] = make_double2(v2, v3);
        } else {
            if (col < p.Np) cp[0] = v0;
            if (col + 1 < p.Np) cp[1] = v1;
            if (col + 2 < p.Np) cp[2] = v2;
        }
    }
}

// ----------------------------------------------------------------------------- generic fallback
// 64x64 tile, 256 threads, 4x4 per thread, BK=16; arbitrary element strides.
__global__ void __launch_bounds__(256)
gemm_generic_kernel(int M, int N, int K,
                    const double* __restrict__ A, long long a_sm, long long a_sk, long long a_sb,
                    const double* __restrict__ B, long long b_sk, long long b_sn, long long b_sb,
                    double* __restrict__ C, long long c_sm, long long c_sn, long long c_sb) {
    __shared__ double As[16][64 + 1];
    __shared__ double Bs[16][64 + 1];
    const int batch = blockIdx.z;
    A += (long long)batch * a_sb;
    B += (long long)batch * b_sb;
    C += (long long)batch * c_sb;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int i = threadIdx.x; i < 64 * 16; i += 256) {
            // choose the faster-varying index to follow the contiguous direction where possible
            int mm, kk;
            if (a_sk == 1) { kk = i & 15; mm = i >> 4; } else { mm = i & 63; kk = i >> 6; }
            const int gmm = m0 + mm, gk = k0 + kk;
            As[kk][mm] = (gmm < M && gk < K) ? A[(long long)gmm * a_sm + (long long)gk * a_sk] : 0.0;
            int nn, kb;
            if (b_sk == 1) { kb = i & 15; nn = i >> 4; } else { nn = i & 63; kb = i >> 6; }
            const int gn = n0 + nn, gkb = k0 + kb;
            Bs[kb][nn] = (gn < N && gkb < K) ? B[(long long)gkb * b_sk + (long long)gn * b_sn] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            double av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = As[kk][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = m0 + ty + 16 * i;
        if (row >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = n0 + tx + 16 * j;
            if (col < N) C[(long long)row * c_sm + (long long)col * c_sn] = acc[i][j];
        }
    }
}

// ----------------------------------------------------------------------------- host side
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

static bool load_encode() {
    if (g_encode) return true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        cudaGetLastError();
        return false;
    }
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    return true;
}

// One operand of a batched 2-D problem, as TMA sees it: `inner` is the stride-1 extent.
static bool make_tmap(CUtensorMap* map, const double* base, long long inner, long long outer, long long batch,
                      long long outer_stride, long long batch_stride, int box_inner, int box_outer) {
    cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)outer, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)outer_stride * 8ull, (cuuint64_t)batch_stride * 8ull};
    cuuint32_t box[3] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer, 1u};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// An MN-contiguous operand [K rows][MN columns] seen as 4-D: (16 columns, K, MN/16 column blocks, batch), so that ONE box of
// (16, BK, tile/16, 1) lands in shared memory as the tile/16 blocks of [16 k][16 mn] the fragment loads expect - instead of
// tile/16 separate 2 KB boxes.  (ncu, 4096x4096x8192: with 16 boxes per k-tile the weight-gradient form ran its tensor pipe at
// 90.6 % with 1.47 long-scoreboard stall cycles per issue, against 98.1 % / 0.15 for two K-contiguous operands: the TMA unit's
// per-instruction cost.)  The last column block may reach past MN: the caller guarantees those elements exist (MN % 16 == 0 or
// the row pitch covers them); what they hold only reaches output rows / columns that are never stored.
static bool make_tmap_blocked(CUtensorMap* map, const double* base, long long mn, long long K, long long batch,
                              long long k_stride, long long batch_stride, int box_blocks) {
    cuuint64_t dims[4] = {16u, (cuuint64_t)K, (cuuint64_t)((mn + 15) / 16), (cuuint64_t)batch};
    cuuint64_t strides[3] = {(cuuint64_t)k_stride * 8ull, 128ull, (cuuint64_t)batch_stride * 8ull};
    cuuint32_t box[4] = {16u, (cuuint32_t)BK, (cuuint32_t)box_blocks, 1u};
    cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}
static bool blocked_ok(long long mn, long long k_stride) {
    static int pref = -1;                          // ADB_GEMM_BLOCKED=0: keep the one-box-per-block maps (A/B measurements)
    if (pref < 0) { const char* e = getenv("ADB_GEMM_BLOCKED"); pref = (e && e[0] == '0') ? 0 : 1; }
    return pref == 1 && (mn % 16 == 0 || k_stride >= (mn + 15) / 16 * 16);
}

static bool tma_ok(const double* base, long long s_other, long long s_batch, long long batch) {
    if (reinterpret_cast<uintptr_t>(base) & 15) return false;
    if (s_other <= 0 || (s_other & 1)) return false;
    if (batch > 1 && (s_batch <= 0 || (s_batch & 1))) return false;
    return true;
}

template <class Cfg, bool A_KC, bool B_KC, int CONV = 0>
static int launch_tma(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int batch, int splits, cudaStream_t st) {
    static bool attr_set = false;
    auto kern = gemm_tma_kernel<Cfg, A_KC, B_KC, CONV>;
    constexpr int GEMM_SMEM = Cfg::SMEM;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
        if (e != cudaSuccess) return set_error("gemm: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    dim3 grid((unsigned)(p.tiles_m * p.tiles_n), (unsigned)batch, (unsigned)splits);
    kern<<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(ta, tb, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("gemm_tma launch failed: %s", cudaGetErrorString(e));
    count_launch();
    return 0;
}

static int sm_count() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}

// The stream-K schedule of a problem whose tile grid (tiles_m, tiles_n) is already set: full waves taken whole, the k-tiles of the
// remaining tiles dealt out to <= `slots` CTAs in shares of >= min_units (see gemm_sk_kernel).  Pure host arithmetic
// (adb_gemm_sk_plan exposes it to the CPU tests).
static int sk_schedule(GemmParams& p, int batch, long long slots, int min_units) {
    p.total_kt = (p.K + BK - 1) / BK;
    p.batch = batch;
    p.tiles_total = (long long)p.tiles_m * p.tiles_n * batch;
    p.dp_tiles = p.tiles_total / slots * slots;            // the full waves
    const long long rem = p.tiles_total - p.dp_tiles;
    p.sk_units = rem * p.total_kt;
    if (p.tiles_total >= (1LL << 31) - slots || p.sk_units >= (1LL << 31)) return set_error("gemm: problem too large for the stream-K schedule");
    p.sk_ctas = 0;
    p.sk_ws = nullptr; p.sk_count = nullptr;
    if (rem > 0) {
        long long c = p.sk_units / min_units;
        if (c < rem) c = rem;
        if (c > slots) c = slots;
        p.sk_ctas = (int)c;
    }
    return 0;
}

// Stream-K launch: fills in the schedule (see gemm_sk_kernel), allocates the partial-tile workspace when some tile
// is split, launches ONE kernel.  min_units: a CTA's share of the stream-K region is never shorter than this many k-tiles.
template <class Cfg, bool A_KC, bool B_KC>
static int launch_sk(const CUtensorMap& ta, const CUtensorMap& tb, GemmParams p, int batch, cudaStream_t st) {
    static bool attr_set = false;
    auto kern = gemm_sk_kernel<Cfg, A_KC, B_KC>;
    constexpr int SMEM = Cfg::SMEM;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) return set_error("gemm: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    static int min_units = -1;         // ADB_GEMM_SK_MIN_UNITS (experiments)
    if (min_units < 0) { const char* e = getenv("ADB_GEMM_SK_MIN_UNITS"); min_units = e ? atoi(e) : 8; if (min_units < 1) min_units = 1; }
    const long long slots = (long long)sm_count() * Cfg::MINB;
    if (int rc = sk_schedule(p, batch, slots, min_units)) return rc;
    const long long rem = p.tiles_total - p.dp_tiles;
    const long long grid = p.dp_tiles + p.sk_ctas;
    void* ws = nullptr;
    if (rem > 0 && p.sk_ctas != rem) {
        const size_t tile_bytes = (size_t)Cfg::BM * Cfg::BN * sizeof(double);
        const size_t ws_bytes = 2 * (size_t)p.sk_ctas * tile_bytes, cnt_bytes = ((size_t)rem * sizeof(unsigned) + 255) & ~(size_t)255;
        cudaError_t e = cudaMallocAsync(&ws, ws_bytes + cnt_bytes, st);
        if (e != cudaSuccess) return set_error("gemm: stream-K workspace allocation failed: %s", cudaGetErrorString(e));
        p.sk_ws = reinterpret_cast<double*>(ws);
        p.sk_count = reinterpret_cast<unsigned*>(reinterpret_cast<unsigned char*>(ws) + ws_bytes);
        e = cudaMemsetAsync(p.sk_count, 0, cnt_bytes, st);
        if (e != cudaSuccess) { cudaFreeAsync(ws, st); return set_error("gemm: stream-K counter reset failed: %s", cudaGetErrorString(e)); }
    }
    kern<<<(unsigned)grid, GEMM_THREADS, SMEM, st>>>(ta, tb, p);
    cudaError_t e = cudaGetLastError();
    if (ws) cudaFreeAsync(ws, st);
    if (e != cudaSuccess) return set_error("gemm_sk launch failed: %s", cudaGetErrorString(e));
    count_launch();
    return 0;
}

int gemm_f64(const adb_gemm_desc* d, cudaStream_t st, int force_path) {
    const long long M = d->M, N = d->N, K = d->K, batch = d->batch;
    if (M <= 0 || N <= 0 || batch <= 0) return 0;
    if (M > 0x7fffffff || N > 0x7fffffff || K > 0x7fffffff) return set_error("gemm: extent too large");
    if (K <= 0) return set_error("gemm: K must be positive (caller zero-fills)");

    // strides of size-1 dims are meaningless; normalise them so layout detection and TMA encoding are clean
    adb_gemm_desc dd = *d;
    if (K == 1) { dd.a_sk = 1; dd.b_sk = 1; if (dd.a_sm == 1 && M > 1) dd.a_sk = 0; if (dd.b_sn == 1 && N > 1) dd.b_sk = 0; }
    if (M == 1 && dd.a_sk == 1) dd.a_sm = (K + 1) & ~1LL;
    if (N == 1 && dd.b_sk == 1) dd.b_sn = (K + 1) & ~1LL;
    if (M == 1 && dd.a_sk != 1) dd.a_sm = 1;
    if (N == 1 && dd.b_sk != 1) dd.b_sn = 1;
    // Skinny in M (few rows, many columns): compute C^T = B^T A^T so the long extent is tiled along M
    if (M <= 32 && N >= 256) {
        adb_gemm_desc t = dd;
        dd.M = t.N; dd.N = t.M;
        dd.A = t.B; dd.a_sm = t.b_sn; dd.a_sk = t.b_sk; dd.a_sb = t.b_sb;
        dd.B = t.A; dd.b_sn = t.a_sm; dd.b_sk = t.a_sk; dd.b_sb = t.a_sb;
        dd.c_sm = t.c_sn; dd.c_sn = t.c_sm;
        return gemm_f64(&dd, st, force_path);
    }
    d = &dd;
    // Operands TMA cannot describe (broadcast / odd pitch / unaligned base / neither stride is 1) are repacked into a
    // dense K-contiguous workspace when the problem is large enough for the DMMA path to pay for the extra pass.
    double* ws_a = nullptr;
    double* ws_b = nullptr;
    auto describable = [&](const double* p, long long s_row, long long s_k, long long s_b) {
        const bool kc = s_k == 1, mc = !kc && s_row == 1;
        return (kc || mc) && tma_ok(p, kc ? s_row : s_k, s_b, batch);
    };
    const bool big_enough = (double)M * (double)N * (double)K * (double)batch >= 16.0 * 1024 * 1024 && batch <= 65535;
    if (force_path != 2 && big_enough && load_encode()) {
        const long long Kp = (K + 1) & ~1LL;
        auto repack = [&](const double* src, long long rows, long long s_row, long long s_k, long long s_b, double** ws) -> int {
            cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(ws), (size_t)batch * rows * Kp * sizeof(double), st);
            if (e != cudaSuccess) return set_error("gemm: workspace allocation failed: %s", cudaGetErrorString(e));
            adb_tensor in, out;
            in.ptr = const_cast<double*>(src); in.rank = 3;
            in.shape[0] = batch; in.shape[1] = rows; in.shape[2] = K;
            in.stride[0] = s_b; in.stride[1] = s_row; in.stride[2] = s_k;
            out.ptr = *ws; out.rank = 3;
            out.shape[0] = batch; out.shape[1] = rows; out.shape[2] = K;
            out.stride[0] = rows * Kp; out.stride[1] = Kp; out.stride[2] = 1;
            adb_ew_instr code[2] = {{ADB_EW_LOAD | ADB_EW_SRC_IN, 0, 0.0}, {ADB_EW_STORE_OUT, 0, 0.0}};
            return ewise_run(code, 2, 1, &in, 1, &out, st);
        };
        if (!describable(dd.A, dd.a_sm, dd.a_sk, dd.a_sb)) {
            if (int rc = repack(dd.A, M, dd.a_sm, dd.a_sk, dd.a_sb, &ws_a)) return rc;
            dd.A = ws_a; dd.a_sm = Kp; dd.a_sk = 1; dd.a_sb = M * Kp;
        }
        if (!describable(dd.B, dd.b_sn, dd.b_sk, dd.b_sb)) {
            if (int rc = repack(dd.B, N, dd.b_sn, dd.b_sk, dd.b_sb, &ws_b)) { if (ws_a) cudaFreeAsync(ws_a, st); return rc; }
            dd.B = ws_b; dd.b_sn = Kp; dd.b_sk = 1; dd.b_sb = N * Kp;
        }
    }
    struct WsGuard {
        double *a, *b; cudaStream_t st;
        ~WsGuard() { if (a) cudaFreeAsync(a, st); if (b) cudaFreeAsync(b, st); }
    } guard{ws_a, ws_b, st};
    const bool a_kc = d->a_sk == 1;
    const bool a_mc = !a_kc && d->a_sm == 1;
    const bool b_kc = d->b_sk == 1;
    const bool b_mc = !b_kc && d->b_sn == 1;
    bool use_tma = force_path != 2 && (a_kc || a_mc) && (b_kc || b_mc) && batch <= 65535 && load_encode();
    if (use_tma)
        use_tma = tma_ok(d->A, a_kc ? d->a_sm : d->a_sk, d->a_sb, batch) && tma_ok(d->B, b_kc ? d->b_sn : d->b_sk, d->b_sb, batch);
    if ((force_path == 1 || force_path >= 3) && !use_tma) return set_error("gemm: TMA path forced but operands are not TMA-describable");

    if (use_tma) {
        CUtensorMap ta, tb;
        bool ok;
        const long long a_bs = batch > 1 ? d->a_sb : 2, b_bs = batch > 1 ? d->b_sb : 2;
        // Tile choice: 128x128 tiles unless wave quantisation on 148 SMs would idle > 5% of the machine, then 64x64
        // (2 CTAs/SM, 4x finer granularity).  Split-K when even the small tiles cannot fill the SMs and K is long
        // (weight-gradient GEMMs of conv / rnn layers: tiny M x N, huge K).
        const long long big_tiles = ((M + 127) / 128) * ((N + 127) / 128) * batch;
        const double waste_big = (double)((big_tiles + 147) / 148 * 148) / (double)big_tiles;
        static int tile_pref = -1;     // ADB_GEMM_TILE=small|big overrides the heuristic (experiments)
        if (tile_pref < 0) {
            const char* e = getenv("ADB_GEMM_TILE");
            tile_pref = !e ? 0 : (!strcmp(e, "small") ? 1 : (!strcmp(e, "big") ? 2 : 0));
        }
        const bool skinny_s = force_path == 6 || (force_path == 0 && tile_pref == 0 && N <= 32 && M <= 128 && K >= 4096);
        const bool skinny = !skinny_s && (force_path == 5 || (force_path == 0 && tile_pref == 0 && N <= 32 && M >= 256));
        // short K (<= 8 k-tiles): a CTA is mostly prologue + epilogue, so two 64x64 CTAs per SM that overlap each other beat
        // one 128x128 CTA (conv dX 802816x400x32: 1.38 -> 1.05 ms)
        const bool short_k = (K + BK - 1) / BK <= 8;
        // Stream-K kernel (one launch, no wave quantisation): forced by path 7 (128x128 tiles) / 8 (64x64 tiles);
        // chosen automatically for 128x128-tile problems whose k-loop is long enough to hide a tile's epilogue.
        static int sk_pref = -1;       // ADB_GEMM_SK=0: the one-tile-per-CTA kernels (A/B measurements)
        if (sk_pref < 0) { const char* e = getenv("ADB_GEMM_SK"); sk_pref = (e && e[0] == '0') ? 0 : 1; }
        const double fill_big = (double)M * (double)N / ((double)((M + 127) / 128 * 128) * (double)((N + 127) / 128 * 128));
        // (measured, profiles/r02_gemm_streamk.md: with at least one full wave it never loses; below one wave every tile is split and
        // the parked partial tiles cost more than the 64x64-tile kernel's idle SMs unless a CTA's share is a few dozen k-tiles)
        const bool sk_auto = sk_pref && force_path == 0 && tile_pref == 0 && !skinny && !skinny_s && !short_k && batch <= 65535 &&
                             fill_big >= 0.75 && (big_tiles >= 148 || big_tiles * ((K + BK - 1) / BK) >= 148 * 48);
        const bool sk_big = force_path == 7 || sk_auto, sk_small = force_path == 8;
        const bool small = !sk_big && !skinny && !skinny_s && (sk_small || force_path == 3 || (force_path != 4 && (tile_pref == 1 || (tile_pref == 0 && (waste_big > 1.05 || short_k)))));
        const int bm = skinny_s ? CfgSkinnyS::BM : skinny ? CfgSkinny::BM : small ? CfgSmall::BM : CfgBig::BM;
        const int bn = skinny_s ? CfgSkinnyS::BN : skinny ? CfgSkinny::BN : small ? CfgSmall::BN : CfgBig::BN;
        const long long tiles = ((M + bm - 1) / bm) * ((N + bn - 1) / bn) * batch;
        const long long slots = (small || skinny_s) ? 296 : 148;
        const int total_kt = (int)((K + BK - 1) / BK);
        int splits = 1;
        if (!sk_big && !sk_small && tiles * 2 <= slots && total_kt >= 32) {
            long long s1 = slots / tiles;
            if (s1 > total_kt / 16) s1 = total_kt / 16;
            if (s1 > 128) s1 = 128;
            if (s1 > 1) splits = (int)s1;
            if (small && splits > 1) {
                // 64x64 tiles, a few hundred CTAs: pick the split count by a cost model fitted to measurements
                // (profiles/r01_gemm_small_splitk.log; unit = one k-tile step of one CTA that shares its SM, 0.52 us):
                //   T(s) = [CTAs on the busiest SM] * [k-tiles per CTA] * (1.23 if the CTA is alone on its SM: one 64x64 CTA does
                //          not saturate the DMMA pipe) + 12 (prologue + epilogue) + 15 if s > 1 (the partial-sum reduction launch).
                // 512x1024x1024 (128 tiles): s = 1, 47 us instead of 51 with s = 2;  256x1024x1024 (64 tiles): s = 4, 31 us.
                double best = 1e30;
                for (long long sc = 1; sc <= s1; ++sc) {
                    const long long per_sm = (tiles * sc + 147) / 148;
                    const double kts = (double)((total_kt + sc - 1) / sc);
                    const double t = (double)per_sm * kts * (per_sm >= 2 ? 1.0 : 1.23) + 12.0 + (sc > 1 ? 15.0 : 0.0);
                    if (t < best - 1e-9) { best = t; splits = (int)sc; }
                }
            }
        }
        int kt_per_split = (total_kt + splits - 1) / splits;
        splits = (total_kt + kt_per_split - 1) / kt_per_split;
        double* ws_c = nullptr;
        const long long Np = (N + 1) & ~1LL;
        if (splits > 1) {
            cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&ws_c), (size_t)splits * batch * M * Np * sizeof(double), st);
            if (e != cudaSuccess) return set_error("gemm: split-K workspace allocation failed: %s", cudaGetErrorString(e));
        }
        int blocked = 0;
        if (a_kc) ok = make_tmap(&ta, d->A, K, M, batch, d->a_sm, a_bs, BK, bm);
        else if (blocked_ok(M, d->a_sk) && make_tmap_blocked(&ta, d->A, M, K, batch, d->a_sk, a_bs, bm / 16)) { ok = true; blocked |= 1; }
        else      ok = make_tmap(&ta, d->A, M, K, batch, d->a_sk, a_bs, 16, BK);
        if (ok) {
            if (b_kc) ok = make_tmap(&tb, d->B, K, N, batch, d->b_sn, b_bs, BK, bn);
            else if (blocked_ok(N, d->b_sk) && make_tmap_blocked(&tb, d->B, N, K, batch, d->b_sk, b_bs, bn / 16)) blocked |= 2;
            else      ok = make_tmap(&tb, d->B, N, K, batch, d->b_sk, b_bs, 16, BK);
        }
        if (ok) {
            GemmParams p;
            memset(&p, 0, sizeof p);
            p.M = (int)M; p.N = (int)N; p.K = (int)K;
            p.tiles_m = (int)((M + bm - 1) / bm);
            p.tiles_n = (int)((N + bn - 1) / bn);
            p.kt_per_split = kt_per_split; p.blocked = blocked;
            if (splits > 1) { p.C = ws_c; p.c_sm = Np; p.c_sn = 1; p.c_sb = M * Np; p.c_ss = (long long)batch * M * Np; }
            else { p.C = d->C; p.c_sm = d->c_sm; p.c_sn = d->c_sn; p.c_sb = d->c_sb; p.c_ss = 0; }
            p.cv_wb = p.cv_hb = p.cv_lw = p.cv_lh = p.cv_kw = p.cv_cpt = 0;
            if (sk_big || sk_small) {
                int rc;
                if (sk_big) {
                    if (a_kc && b_kc) rc = launch_sk<CfgBig, true, true>(ta, tb, p, (int)batch, st);
                    else if (a_kc && !b_kc) rc = launch_sk<CfgBig, true, false>(ta, tb, p, (int)batch, st);
                    else if (!a_kc && b_kc) rc = launch_sk<CfgBig, false, true>(ta, tb, p, (int)batch, st);
                    else rc = launch_sk<CfgBig, false, false>(ta, tb, p, (int)batch, st);
                } else {
                    if (a_kc && b_kc) rc = launch_sk<CfgSmall, true, true>(ta, tb, p, (int)batch, st);
                    else if (a_kc && !b_kc) rc = launch_sk<CfgSmall, true, false>(ta, tb, p, (int)batch, st);
                    else if (!a_kc && b_kc) rc = launch_sk<CfgSmall, false, true>(ta, tb, p, (int)batch, st);
                    else rc = launch_sk<CfgSmall, false, false>(ta, tb, p, (int)batch, st);
                }
                return rc;
            }
            auto dispatch = [&](const GemmParams& pp, int sp) -> int {
                if (skinny_s) {
                    if (a_kc && b_kc) return launch_tma<CfgSkinnyS, true, true>(ta, tb, pp, (int)batch, sp, st);
                    else if (a_kc && !b_kc) return launch_tma<CfgSkinnyS, true, false>(ta, tb, pp, (int)batch, sp, st);
                    else if (!a_kc && b_kc) return launch_tma<CfgSkinnyS, false, true>(ta, tb, pp, (int)batch, sp, st);
                    else return launch_tma<CfgSkinnyS, false, false>(ta, tb, pp, (int)batch, sp, st);
                } else if (skinny) {
                    if (a_kc && b_kc) return launch_tma<CfgSkinny, true, true>(ta, tb, pp, (int)batch, sp, st);
                    else if (a_kc && !b_kc) return launch_tma<CfgSkinny, true, false>(ta, tb, pp, (int)batch, sp, st);
                    else if (!a_kc && b_kc) return launch_tma<CfgSkinny, false, true>(ta, tb, pp, (int)batch, sp, st);
                    else return launch_tma<CfgSkinny, false, false>(ta, tb, pp, (int)batch, sp, st);
                } else if (small) {
                    if (a_kc && b_kc) return launch_tma<CfgSmall, true, true>(ta, tb, pp, (int)batch, sp, st);
                    else if (a_kc && !b_kc) return launch_tma<CfgSmall, true, false>(ta, tb, pp, (int)batch, sp, st);
                    else if (!a_kc && b_kc) return launch_tma<CfgSmall, false, true>(ta, tb, pp, (int)batch, sp, st);
                    else return launch_tma<CfgSmall, false, false>(ta, tb, pp, (int)batch, sp, st);
                } else {
                    if (a_kc && b_kc) return launch_tma<CfgBig, true, true>(ta, tb, pp, (int)batch, sp, st);
                    else if (a_kc && !b_kc) return launch_tma<CfgBig, true, false>(ta, tb, pp, (int)batch, sp, st);
                    else if (!a_kc && b_kc) return launch_tma<CfgBig, false, true>(ta, tb, pp, (int)batch, sp, st);
                    else return launch_tma<CfgBig, false, false>(ta, tb, pp, (int)batch, sp, st);
                }
            };
            int rc = dispatch(p, splits);
            if (rc == 0 && splits > 1) {
                // C[b][m][n] = sum_z ws[z][b][m][n]  (deterministic order, column-mode reduction kernel)
                ReduceSpec S;
                memset(&S, 0, sizeof S);
                S.n_ops = 1; S.n_o = 3; S.n_r = 1; S.group = 1;
                S.odim[0] = N; S.odim[1] = M; S.odim[2] = batch;
                S.out_s[0] = d->c_sn; S.out_s[1] = d->c_sm; S.out_s[2] = d->c_sb;
                S.op_os[0][0] = 1; S.op_os[0][1] = Np; S.op_os[0][2] = M * Np;
                S.rdim[0] = splits; S.op_rs[0][0] = (long long)batch * M * Np;
                S.op[0] = ws_c; S.out = d->C;
                rc = reduce_run(S, st);
            }
            if (ws_c) cudaFreeAsync(ws_c, st);
            return rc;
        }
        if (ws_c) cudaFreeAsync(ws_c, st);
        if (force_path == 1 || force_path >= 3) return set_error("gemm: cuTensorMapEncodeTiled rejected the operands");
    }
    // generic path
    for (long long b0 = 0; b0 < batch; b0 += 65535) {
        const long long nb = (batch - b0 < 65535) ? batch - b0 : 65535;
        dim3 grid((unsigned)((N + 63) / 64), (unsigned)((M + 63) / 64), (unsigned)nb);
        gemm_generic_kernel<<<grid, 256, 0, st>>>((int)M, (int)N, (int)K,
                                                   d->A + b0 * d->a_sb, d->a_sm, d->a_sk, d->a_sb,
                                                   d->B + b0 * d->b_sb, d->b_sk, d->b_sn, d->b_sb,
                                                   d->C + b0 * d->c_sb, d->c_sm, d->c_sn, d->c_sb);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return set_error("gemm_generic launch failed: %s", cudaGetErrorString(e));
        count_launch();
    }
    return 0;
}

// ----------------------------------------------------------------------------- implicit-GEMM convolution
// The im2col matrix of an NHWC image is never written: the GEMM's A tiles come straight from the image through an
// im2col-mode tensor map.  At configs[3] the materialised matrix of the second
// layer is 2.57 GB, written once and read twice per step, and its gradient another 2.57 GB written and read by col2im.
static PFN_cuTensorMapEncodeIm2col_v12000 g_encode_im2col = nullptr;

static bool load_encode_im2col() {
    if (g_encode_im2col) return true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        cudaGetLastError();
        return false;
    }
    g_encode_im2col = reinterpret_cast<PFN_cuTensorMapEncodeIm2col_v12000>(fn);
    return true;
}

bool conv_image_ok(const adb_tensor* img, int kh, int kw, int pad_h, int pad_w) {
    if (img->rank != 4) return false;
    const long long N = img->shape[0], H = img->shape[1], W = img->shape[2], Cc = img->shape[3];
    if (N < 1 || H < 1 || W < 1 || Cc < 16 || Cc % 16) return false;
    if (img->stride[3] != 1) return false;
    for (int k = 0; k < 3; ++k) if (img->stride[k] <= 0 || (img->stride[k] & 1)) return false;
    if (reinterpret_cast<uintptr_t>(img->ptr) & 15) return false;
    if (kh < 1 || kw < 1 || pad_h < 0 || pad_w < 0 || pad_h > 127 || pad_w > 127 || kh - 1 - pad_h > 128 || kw - 1 - pad_w > 128) return false;
    const long long hb = H + 2 * pad_h - kh + 1, wb = W + 2 * pad_w - kw + 1;
    if (hb < 1 || wb < 1 || N * hb * wb >= (1LL << 31) - 1024 || (long long)kh * kw * Cc >= (1LL << 31)) return false;
    return load_encode() && load_encode_im2col();
}

static bool make_tmap_im2col(CUtensorMap* map, const adb_tensor* img, int kh, int kw, int pad_h, int pad_w, int pixels) {
    cuuint64_t dims[4] = {(cuuint64_t)img->shape[3], (cuuint64_t)img->shape[2], (cuuint64_t)img->shape[1], (cuuint64_t)img->shape[0]};
    cuuint64_t strides[3] = {(cuuint64_t)img->stride[2] * 8ull, (cuuint64_t)img->stride[1] * 8ull, (cuuint64_t)img->stride[0] * 8ull};
    int lo[2] = {-pad_w, -pad_h}, up[2] = {pad_w - (kw - 1), pad_h - (kh - 1)};
    cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    CUresult r = g_encode_im2col(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, img->ptr, dims, strides, lo, up, 16u, (cuuint32_t)pixels, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// The B operand (K x Nn, strides b_sk / b_sn) of a CONV launch: K-contiguous or N-contiguous, as in gemm_f64.
static bool make_b_map(CUtensorMap* tb, const double* B, long long K, long long Nn, long long b_sk, long long b_sn, int bn, bool* b_kc, int* blocked) {
    *b_kc = b_sk == 1;
    if (!*b_kc && b_sn != 1) return false;
    if (!tma_ok(B, *b_kc ? b_sn : b_sk, 2, 1)) return false;
    if (*b_kc) return make_tmap(tb, B, K, Nn, 1, b_sn, 2, BK, bn);
    if (blocked_ok(Nn, b_sk) && make_tmap_blocked(tb, B, Nn, K, 1, b_sk, 2, bn / 16)) { *blocked |= 2; return true; }
    return make_tmap(tb, B, Nn, K, 1, b_sk, 2, 16, BK);
}

// out[(n,hp,wp), j] = sum_{a,b,c} img[n, hp - pad_h + a, wp - pad_w + b, c] * B[(a*kw + b)*C + c, j]
//   pad = 0:      forward convolution            == Einsum('pk,ko->po', Im2Col(img), B)
//   pad = k - 1:  data gradient (img = dOut, B = flipped filter)  == Col2Im(Einsum('po,ko->pk', dOut, W))
int conv_gemm_cols(const adb_tensor* img, int kh, int kw, int pad_h, int pad_w, const double* B, long long b_sk, long long b_sn, long long Nn,
                   double* out, long long c_sm, long long c_sn, cudaStream_t st) {
    if (!conv_image_ok(img, kh, kw, pad_h, pad_w)) return set_error("conv_gemm: image not describable by an im2col tensor map");
    const long long Cc = img->shape[3], hb = img->shape[1] + 2 * pad_h - kh + 1, wb = img->shape[2] + 2 * pad_w - kw + 1;
    const long long M = img->shape[0] * hb * wb, K = (long long)kh * kw * Cc;
    if (Nn < 1) return 0;
    const int cfg = Nn <= 16 ? 0 : (Nn <= 32 ? 1 : 2);
    const int bm = cfg == 2 ? CfgBig::BM : CfgSkinny::BM, bn = cfg == 0 ? CfgSkinny16::BN : (cfg == 1 ? CfgSkinny::BN : CfgBig::BN);
    CUtensorMap ta, tb;
    bool b_kc = false;
    int blocked = 0;
    if (!make_tmap_im2col(&ta, img, kh, kw, pad_h, pad_w, bm)) return set_error("conv_gemm: cuTensorMapEncodeIm2col rejected the image");
    if (!make_b_map(&tb, B, K, Nn, b_sk, b_sn, bn, &b_kc, &blocked)) return set_error("conv_gemm: filter matrix not describable by a tensor map");
    GemmParams p;
    memset(&p, 0, sizeof p);
    p.M = (int)M; p.N = (int)Nn; p.K = (int)K;
    p.tiles_m = (int)((M + bm - 1) / bm); p.tiles_n = (int)((Nn + bn - 1) / bn);
    p.kt_per_split = (int)(K / BK); p.blocked = blocked;
    p.C = out; p.c_sm = c_sm; p.c_sn = c_sn;
    p.cv_wb = (int)wb; p.cv_hb = (int)hb; p.cv_lw = -pad_w; p.cv_lh = -pad_h; p.cv_kw = kw; p.cv_cpt = (int)(Cc / 16);
    if (cfg == 0) return b_kc ? launch_tma<CfgSkinny16, true, true, 1>(ta, tb, p, 1, 1, st) : launch_tma<CfgSkinny16, true, false, 1>(ta, tb, p, 1, 1, st);
    if (cfg == 1) return b_kc ? launch_tma<CfgSkinny, true, true, 1>(ta, tb, p, 1, 1, st) : launch_tma<CfgSkinny, true, false, 1>(ta, tb, p, 1, 1, st);
    return b_kc ? launch_tma<CfgBig, true, true, 1>(ta, tb, p, 1, 1, st) : launch_tma<CfgBig, true, false, 1>(ta, tb, p, 1, 1, st);
}

// 4-D tiled map of an NHWC array: dims (C, W, H, N), box (16 channels, rows pixels, 1, 1).
static bool make_tmap_nhwc_rows(CUtensorMap* map, const double* base, long long Cn, long long Wn, long long Hn, long long Nn,
                                long long s_w, long long s_h, long long s_n, int rows) {
    cuuint64_t dims[4] = {(cuuint64_t)Cn, (cuuint64_t)Wn, (cuuint64_t)Hn, (cuuint64_t)Nn};
    cuuint64_t strides[3] = {(cuuint64_t)s_w * 8ull, (cuuint64_t)s_h * 8ull, (cuuint64_t)s_n * 8ull};
    cuuint32_t box[4] = {16u, (cuuint32_t)rows, 1u, 1u};
    cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <int KW>
static int launch_wgrad(const CUtensorMap& tx, const CUtensorMap& td, const WgradParams& p, int tasks, int splits, int smem, cudaStream_t st) {
    static bool attr_set = false;
    auto kern = conv_wgrad_kernel<KW>;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return set_error("conv_wgrad: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    kern<<<dim3((unsigned)tasks, (unsigned)splits), 288, smem, st>>>(tx, td, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("conv_wgrad launch failed: %s", cudaGetErrorString(e));
    count_launch();
    return 0;
}

static int conv_wgrad_rows(const adb_tensor* img, int kh, int kw, const double* D, long long d_sp, long long O,
                           double* out, long long c_sm, long long c_sn, cudaStream_t st) {
    const long long Nn = img->shape[0], H = img->shape[1], W = img->shape[2], Cc = img->shape[3];
    const long long OH = H - kh + 1, OW = W - kw + 1, M = (long long)kh * kw * Cc;
    WgradParams p;
    memset(&p, 0, sizeof p);
    const long long owp = (OW + 7) / 8 * 8;
    p.nseg = owp <= 64 ? 1 : (int)((OW + 63) / 64);
    p.SEG = (int)(((OW + p.nseg - 1) / p.nseg + 7) / 8 * 8);
    p.xrows = (p.SEG + kw - 1 + 7) / 8 * 8;
    p.kw = kw; p.C = (int)Cc; p.O = (int)O; p.cblocks = (int)(Cc / 16); p.opairs = (int)((O + 31) / 32); p.OH = (int)OH;
    const long long chunks = Nn * OH * p.nseg;
    if (chunks >= (1LL << 31)) return set_error("conv_wgrad: too many positions");
    p.chunks = (int)chunks;
    const int tasks = kh * p.cblocks * p.opairs;
    int splits = tasks >= 148 ? 1 : 148 / tasks;
    if (splits > chunks / 8) splits = (int)(chunks / 8);
    if (splits < 1) splits = 1;
    p.chunks_per_split = (int)((chunks + splits - 1) / splits);
    splits = (int)((chunks + p.chunks_per_split - 1) / p.chunks_per_split);
    p.Np = (int)((O + 3) & ~3LL);
    p.slab = M * p.Np;
    const int stage_bytes = (p.xrows + 2 * p.SEG) * 128;
    const int smem = WG_STAGES * stage_bytes + 1024 + 2 * WG_STAGES * 8;
    if (smem > 227 * 1024) return set_error("conv_wgrad: stage does not fit shared memory");
    CUtensorMap tx, td;
    if (!make_tmap_nhwc_rows(&tx, img->ptr, Cc, W, H, Nn, img->stride[2], img->stride[1], img->stride[0], p.xrows) ||
        !make_tmap_nhwc_rows(&td, D, O, OW, OH, Nn, d_sp, OW * d_sp, OH * OW * d_sp, p.SEG))
        return set_error("conv_wgrad: cuTensorMapEncodeTiled rejected the operands");
    const int nslabs = splits * 4;
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&p.ws), (size_t)nslabs * p.slab * sizeof(double), st);
    if (e != cudaSuccess) return set_error("conv_wgrad: workspace allocation failed: %s", cudaGetErrorString(e));
    int rc;
    switch (kw) {
        case 1: rc = launch_wgrad<1>(tx, td, p, tasks, splits, smem, st); break;
        case 2: rc = launch_wgrad<2>(tx, td, p, tasks, splits, smem, st); break;
        case 3: rc = launch_wgrad<3>(tx, td, p, tasks, splits, smem, st); break;
        case 4: rc = launch_wgrad<4>(tx, td, p, tasks, splits, smem, st); break;
        case 5: rc = launch_wgrad<5>(tx, td, p, tasks, splits, smem, st); break;
        case 6: rc = launch_wgrad<6>(tx, td, p, tasks, splits, smem, st); break;
        default: rc = launch_wgrad<7>(tx, td, p, tasks, splits, smem, st); break;
    }
    if (rc == 0) {
        ReduceSpec S;
        memset(&S, 0, sizeof S);
        S.n_ops = 1; S.n_o = 2; S.n_r = 1; S.group = 1;
        S.odim[0] = O; S.odim[1] = M;
        S.out_s[0] = c_sn; S.out_s[1] = c_sm;
        S.op_os[0][0] = 1; S.op_os[0][1] = p.Np;
        S.rdim[0] = nslabs; S.op_rs[0][0] = p.slab;
        S.op[0] = p.ws; S.out = out;
        rc = reduce_run(S, st);
    }
    cudaFreeAsync(p.ws, st);
    return rc;
}

// out[(a*kw + b)*C + c, o] = sum_{(n,hp,wp)} img[n, hp + a, wp + b, c] * D[(n,hp,wp), o]     == Einsum('pk,po->ko', Im2Col(img), D)
int conv_gemm_wgrad(const adb_tensor* img, int kh, int kw, const double* D, long long d_sp, long long d_so, long long O,
                    double* out, long long c_sm, long long c_sn, cudaStream_t st) {
    if (!conv_image_ok(img, kh, kw, 0, 0)) return set_error("conv_gemm: image not describable by an im2col tensor map");
    const long long Cc = img->shape[3], hb = img->shape[1] - kh + 1, wb = img->shape[2] - kw + 1;
    const long long P = img->shape[0] * hb * wb, M = (long long)kh * kw * Cc;
    if (O < 1) return 0;
    {
        static int pref = -1;          // ADB_CONV_WGRAD=blocks: the im2col-block GEMM below instead of conv_wgrad_kernel (A/B measurements)
        if (pref < 0) { const char* e = getenv("ADB_CONV_WGRAD"); pref = (e && !strcmp(e, "blocks")) ? 0 : 1; }
        if (pref && kw <= 7 && d_so == 1 && d_sp >= O && (d_sp & 1) == 0 && (reinterpret_cast<uintptr_t>(D) & 15) == 0 && wb <= 4096)
            return conv_wgrad_rows(img, kh, kw, D, d_sp, O, out, c_sm, c_sn, st);
    }
    const bool skinny = O <= 32;
    const int bm = 128, bn = skinny ? CfgSkinnyS::BN : CfgBig::BN;
    CUtensorMap ta, tb;
    bool b_kc = false;
    int blocked = 0;
    if (!make_tmap_im2col(&ta, img, kh, kw, 0, 0, BK)) return set_error("conv_gemm: cuTensorMapEncodeIm2col rejected the image");
    if (!make_b_map(&tb, D, P, O, d_sp, d_so, bn, &b_kc, &blocked)) return set_error("conv_gemm: output gradient not describable by a tensor map");
    const long long tiles = ((M + bm - 1) / bm) * ((O + bn - 1) / bn), slots = skinny ? 296 : 148;
    const int total_kt = (int)((P + BK - 1) / BK);
    long long splits = slots / tiles;
    if (splits > total_kt / 16) splits = total_kt / 16;
    if (splits > 128) splits = 128;
    if (splits < 1) splits = 1;
    const int kt_per_split = (int)((total_kt + splits - 1) / splits);
    splits = (total_kt + kt_per_split - 1) / kt_per_split;
    const long long Np = (O + 1) & ~1LL;
    double* ws = nullptr;
    if (splits > 1) {
        cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&ws), (size_t)splits * M * Np * sizeof(double), st);
        if (e != cudaSuccess) return set_error("conv_gemm: split-K workspace allocation failed: %s", cudaGetErrorString(e));
    }
    GemmParams p;
    memset(&p, 0, sizeof p);
    p.M = (int)M; p.N = (int)O; p.K = (int)P;
    p.tiles_m = (int)((M + bm - 1) / bm); p.tiles_n = (int)((O + bn - 1) / bn);
    p.kt_per_split = kt_per_split; p.blocked = blocked;
    if (splits > 1) { p.C = ws; p.c_sm = Np; p.c_sn = 1; p.c_ss = M * Np; }
    else { p.C = out; p.c_sm = c_sm; p.c_sn = c_sn; }
    p.cv_wb = (int)wb; p.cv_hb = (int)hb; p.cv_kw = kw; p.cv_cpt = (int)(Cc / 16);
    int rc;
    if (skinny) rc = b_kc ? launch_tma<CfgSkinnyS, false, true, 1>(ta, tb, p, 1, (int)splits, st) : launch_tma<CfgSkinnyS, false, false, 1>(ta, tb, p, 1, (int)splits, st);
    else rc = b_kc ? launch_tma<CfgBig, false, true, 1>(ta, tb, p, 1, (int)splits, st) : launch_tma<CfgBig, false, false, 1>(ta, tb, p, 1, (int)splits, st);
    if (rc == 0 && splits > 1) {
        ReduceSpec S;
        memset(&S, 0, sizeof S);
        S.n_ops = 1; S.n_o = 2; S.n_r = 1; S.group = 1;
        S.odim[0] = O; S.odim[1] = M;
        S.out_s[0] = c_sn; S.out_s[1] = c_sm;
        S.op_os[0][0] = 1; S.op_os[0][1] = Np;
        S.rdim[0] = splits; S.op_rs[0][0] = M * Np;
        S.op[0] = ws; S.out = out;
        rc = reduce_run(S, st);
    }
    if (ws) cudaFreeAsync(ws, st);
    return rc;
}

}  // namespace adb

extern "C" int adb_gemm_sk_plan(long long M, long long N, long long K, long long batch, int bm, int bn, long long slots, int min_units, long long* out) {
    if (M <= 0 || N <= 0 || K <= 0 || batch <= 0 || bm <= 0 || bn <= 0 || slots <= 0 || min_units <= 0 || !out) return adb::set_error("adb_gemm_sk_plan: bad argument");
    if (M > 0x7fffffff || N > 0x7fffffff || K > 0x7fffffff || batch > 0x7fffffff) return adb::set_error("adb_gemm_sk_plan: extent too large");
    adb::GemmParams p;
    memset(&p, 0, sizeof p);
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.tiles_m = (int)((M + bm - 1) / bm); p.tiles_n = (int)((N + bn - 1) / bn);
    if (int rc = adb::sk_schedule(p, (int)batch, slots, min_units)) return rc;
    out[0] = p.total_kt; out[1] = p.tiles_total; out[2] = p.dp_tiles; out[3] = p.sk_units; out[4] = p.sk_ctas; out[5] = p.dp_tiles + p.sk_ctas;
    return 0;
}
extern "C" int adb_gemm_f64(const adb_gemm_desc* desc, void* stream) {
    return adb::gemm_f64(desc, reinterpret_cast<cudaStream_t>(stream), 0);
}
extern "C" int adb_gemm_f64_path(const adb_gemm_desc* desc, void* stream, int path) {
    return adb::gemm_f64(desc, reinterpret_cast<cudaStream_t>(stream), path);
}
