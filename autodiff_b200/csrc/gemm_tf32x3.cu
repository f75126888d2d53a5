// 3xTF32 strided-batched GEMM on the 5th-generation tensor cores (tcgen05 + TMEM) — the "TF32/FP32 variant" of the
// contraction path (BASELINE.json north_star; parity bar 1e-4).  Same contract as adb_gemm_f64: fp64 operands
// and fp64 result with arbitrary element strides; only the arithmetic inside differs.
//
// A single TF32 pass cannot meet 1e-4 (measured 3e-4 at K = 1024..8192, SURVEY.md hard part 3), so every fp64
// operand element a is split  a ~= hi + lo  with hi = tf32(a), lo = tf32(a - hi), and the product is
//   A.B ~= hi_A.hi_B + lo_A.hi_B + hi_A.lo_B          (three tcgen05.mma per k-step, fp32 accumulation in TMEM)
// which gives ~1e-7 relative error.
//
//   split_tf32_kernel / split_tf32_t_kernel : fp64 (any strides) -> two K-major fp32 matrices (hi, lo) per operand, zero-padded
//                       pitch; the _t variant transposes 32 x 32 tiles through shared memory for MN-contiguous operands
//   gemm_tf32x3_kernel<MT>: TMA (128B swizzle) -> smem ring -> tcgen05.mma.kind::tf32 (M=128, N=128, K=8) issued by one
//                       thread, MT accumulators of 128 TMEM columns each, TWO such sets -> tcgen05.ld -> fp64 -> C.  MT = 2
//                       stacks two 128-row tiles on the same B tiles (256x128 per CTA, 2 stages of 96 KiB).  K runs in
//                       chains of <= 2048 (fp32 accumulation error grows with the chain): chain c + 1 fills one TMEM set
//                       while the four epilogue warps add chain c to C in fp64.
//                       Measured at 8192^3 (profiles/r02_ncu_summary.md): kernel 4.86 ms = 678 TFLOP/s of raw tf32 MMA = 0.90 of
//                       the cuBLAS TF32 burst rate of this pool (753; profiles/tf32_peak.json), 5.30 ms with the split passes
//                       = 207 TFLOP/s effective (round 1, chains as separate CTAs + a reduction kernel, column-major tile
//                       order: 6.2 ms, 11.5 GB of DRAM reads per launch against 5.9 GB now)
//   gemm_tf32x3_2cta_kernel: the same with two CTAs of a cluster on a 256 x 256 tile (tcgen05.mma.cta_group::2, see below):
//                       8192^3 4.23 ms kernel / 4.69 ms with the split passes = 235 TFLOP/s effective
//   warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4..7 = epilogue (one TMEM lane quadrant each)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "adb_internal.h"

namespace adb {

namespace tf32 {

constexpr int BM = 128, BN = 128, BK = 32;       // one tcgen05.mma: M = 128, N = 128; BK * 4 B = 128 B = one swizzle row
constexpr int TILE_BYTES = BM * BK * 4;          // 16 KiB: 128 rows of one operand matrix for one k-block
constexpr int THREADS = 256;
// MT = number of 128-row MMA tiles a CTA stacks along M.  The kernel is L2-bandwidth-bound (4 operand tiles per 3 MMAs):
// with MT = 2 the B tiles of a stage feed two accumulators, 96 KiB per 256x128x32 instead of 64 KiB per 128x128x32.
template <int MT>
struct Cfg {
    static constexpr int STAGES = MT == 1 ? 3 : 2;
    static constexpr int STAGE_BYTES = (2 * MT + 2) * TILE_BYTES;     // A_hi, A_lo (MT tiles each), B_hi, B_lo
    static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 128;
    static constexpr int ACC_COLS = 128 * MT;        // one accumulator set (MT tiles of 128 columns)
    static constexpr int TMEM_COLS = 2 * ACC_COLS;   // two sets: the epilogue drains chunk c while the MMAs fill chunk c + 1
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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

// Shared-memory matrix descriptor (tcgen05 / "UMMA"): K-major operand, 128-byte swizzle, rows of 128 B,
// 8-row groups 1024 B apart.  bits [0,14) addr>>4, [16,30) LBO>>4 (unused here), [32,46) SBO>>4, [46,48) version=1,
// [61,64) layout (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                               // LBO = 1 (canonical value for swizzled K-major; unused)
    d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24.
__device__ __forceinline__ uint32_t make_idesc() {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

struct Params {
    int M, N, K;
    int tiles_m, tiles_n;
    int kb_per_chunk, _pad;          // k-blocks per accumulation chain (K is chunked so fp32 chains stay short, see gemm_tf32x3)
    double* C;
    long long c_sm, c_sn, c_sb;
};

template <int MT>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                   const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl, const Params p) {
    constexpr int STAGES = Cfg<MT>::STAGES, STAGE_BYTES = Cfg<MT>::STAGE_BYTES, TMEM_COLS = Cfg<MT>::TMEM_COLS, ACC_COLS = Cfg<MT>::ACC_COLS;
    constexpr int A_BYTES = MT * TILE_BYTES;          // one A matrix (hi or lo) of a stage
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    const uint32_t full0 = smem_u32(bars);
    const uint32_t empty0 = smem_u32(bars + STAGES);
    const uint32_t tmem_full0 = smem_u32(bars + 2 * STAGES);          // [2]: accumulator set complete (MMA -> epilogue)
    const uint32_t tmem_empty0 = smem_u32(bars + 2 * STAGES + 2);     // [2]: accumulator set drained (epilogue -> MMA)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // grouped rasterisation (as gemm_f64.cu): the CTAs of a wave cover 8 tile-rows x ~18 tile-columns, so a wave reads 8 A panels
    // and ~18 B panels instead of every A panel (ncu, 8192^3 with the column-major order: 11.5 GB of DRAM reads for 1 GB of operands)
    const int GROUP = 8;
    const int tile = blockIdx.x;
    const int group_size = GROUP * p.tiles_n;
    const int first_m = (tile / group_size) * GROUP;
    const int gm = min(p.tiles_m - first_m, GROUP);
    const int tm = first_m + (tile % group_size) % gm, tn = (tile % group_size) / gm;
    const int m0 = tm * BM * MT, n0 = tn * BN;
    const int batch = blockIdx.y;
    const int total_kb = (p.K + BK - 1) / BK;
    const int chunks = (total_kb + p.kb_per_chunk - 1) / p.kb_per_chunk;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tmem_full0 + 8 * s, 1);
            mbar_init(tmem_empty0 + 8 * s, 4);        // lane 0 of each of the four epilogue warps
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {   // one warp owns TMEM allocation (and deallocation at the end)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer: all k-blocks of the tile, chunk after chunk =====
        if (lane == 0) {
            for (int kb = 0; kb < total_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(empty0 + 8 * s, ph ^ 1);
                const uint32_t fb = full0 + 8 * s;
                mbar_expect_tx(fb, STAGE_BYTES);
                const uint32_t st = smem_u32(smem + s * STAGE_BYTES);
                const int k0 = kb * BK;
                tma_load_3d(st, &tmAh, fb, k0, m0, batch);                       // box = (BK, 128 * MT rows)
                tma_load_3d(st + A_BYTES, &tmAl, fb, k0, m0, batch);
                tma_load_3d(st + 2 * A_BYTES, &tmBh, fb, k0, n0, batch);
                tma_load_3d(st + 2 * A_BYTES + TILE_BYTES, &tmBl, fb, k0, n0, batch);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one thread drives the tensor core for the whole CTA.  Chunk c accumulates into TMEM set c & 1. =====
        if (lane == 0) {
            const uint32_t idesc = make_idesc();
            int kb = 0;
            for (int c = 0; c < chunks; ++c) {
                const int set = c & 1;
                if (c >= 2) {                                            // the epilogue has drained chunk c - 2 out of this set
                    mbar_wait(tmem_empty0 + 8 * set, ((c >> 1) - 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                const int kb_end = min(kb + p.kb_per_chunk, total_kb);
                for (bool first = true; kb < kb_end; ++kb, first = false) {
                    const int s = kb % STAGES;
                    const uint32_t ph = (kb / STAGES) & 1;
                    mbar_wait(full0 + 8 * s, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = smem_u32(smem + s * STAGE_BYTES);
                    const uint64_t dBh = make_desc(st + 2 * A_BYTES), dBl = make_desc(st + 2 * A_BYTES + TILE_BYTES);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {       // accumulator mt = TMEM columns [128 * mt, +128) of the set, A rows [128 * mt, +128)
                        const uint64_t dAh = make_desc(st + mt * TILE_BYTES), dAl = make_desc(st + A_BYTES + mt * TILE_BYTES);
                        const uint32_t acc = tmem_base + (uint32_t)(set * ACC_COLS + mt * BN);
#pragma unroll
                        for (int k = 0; k < BK / 8; ++k) {
                            const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);      // 32 bytes per K = 8 step inside the swizzled row
                            umma_tf32(acc, dAl + adv, dBh + adv, idesc, (!first || k > 0) ? 1u : 0u);   // small terms first
                            umma_tf32(acc, dAh + adv, dBl + adv, idesc, 1u);
                            umma_tf32(acc, dAh + adv, dBh + adv, idesc, 1u);
                        }
                    }
                    umma_commit(empty0 + 8 * s);           // frees the smem slot once these MMAs have read it
                }
                umma_commit(tmem_full0 + 8 * set);         // chunk complete
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> fp64 -> C.  Warp w may only touch TMEM lanes [32*(w%4), +32).
        // Chunk 0 stores, every later chunk adds to what this same thread stored (fp64, chunk order: the sum the separate
        // partial-tile reduction of round 1 produced, bit for bit) - the tile stays in L2 between chunks. =====
        const int quad = warp & 3;
        for (int c = 0; c < chunks; ++c) {
            const int set = c & 1;
            mbar_wait(tmem_full0 + 8 * set, (c >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int mc = 0; mc < MT * (BN / 32); ++mc) {
                const int mt = mc / (BN / 32), cc = mc % (BN / 32);
                const int row = m0 + mt * BM + quad * 32 + lane;
                double* crow = p.C + (long long)batch * p.c_sb + (long long)row * p.c_sm;
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(set * ACC_COLS + mt * BN + cc * 32);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                      "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                      "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                      "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (row < p.M) {
                    const int col0 = n0 + cc * 32;
                    if (p.c_sn == 1 && col0 + 31 < p.N && ((reinterpret_cast<uintptr_t>(crow + col0) & 15) == 0)) {
                        if (c == 0) {
#pragma unroll
                            for (int j = 0; j < 32; j += 2)
                                *reinterpret_cast<double2*>(crow + col0 + j) = make_double2((double)__uint_as_float(v[j]), (double)__uint_as_float(v[j + 1]));
                        } else {
                            // all 16 loads first: interleaved with the stores they would be serialised (the compiler must assume
                            // the stores alias the next load: ncu showed one exposed L2 round trip per DADD)
                            double2 o[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) o[j] = *reinterpret_cast<const double2*>(crow + col0 + 2 * j);
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                o[j].x += (double)__uint_as_float(v[2 * j]);
                                o[j].y += (double)__uint_as_float(v[2 * j + 1]);
                            }
#pragma unroll
                            for (int j = 0; j < 16; ++j) *reinterpret_cast<double2*>(crow + col0 + 2 * j) = o[j];
                        }
                    } else {
                        double o[32];                              // (loads first, as above)
#pragma unroll
                        for (int j = 0; j < 32; ++j) o[j] = (c > 0 && col0 + j < p.N) ? crow[(long long)(col0 + j) * p.c_sn] : 0.0;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < p.N) crow[(long long)(col0 + j) * p.c_sn] = o[j] + (double)__uint_as_float(v[j]);
                    }
                }
            }
            // this set may be overwritten by chunk c + 2
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmem_empty0 + 8 * set) : "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// ----------------------------------------------------------------------------- CTA-pair variant (cta_group::2)
// Two CTAs of a cluster (one TPC) compute a 256 x 256 tile together: each loads ITS 128 rows of A (hi, lo) and ITS 128 of the 256
// B rows (hi, lo) - 64 KiB per k-block and CTA instead of 96 KiB for the same 6.3 MFLOP, so three stages fit and the operand bytes
// per flop drop by a third (the single-CTA kernel is bound by operand delivery).  All eight TMA loads of a stage
// signal the LEADER's full barrier (`cp.async.bulk.tensor...cta_group::2`, barrier address with the peer bit cleared); the leader's
// MMA thread issues `tcgen05.mma.cta_group::2` (M = 256, N = 256, K = 8: each CTA's tensor core produces its 128 rows from its own
// A tile and both CTAs' B halves) and `tcgen05.commit...multicast::cluster` releases the smem slot / hands over the accumulator
// set in BOTH CTAs.  Each CTA's epilogue warps drain their own 128 TMEM lanes x 256 columns and arrive on the leader's
// tmem_empty barrier.  Chains of <= 2048 and the fp64 accumulation in C as in the single-CTA kernel.
struct Cfg2 {
    static constexpr int STAGES = 3;
    static constexpr int STAGE_BYTES = 4 * TILE_BYTES;            // A_hi, A_lo, B_hi, B_lo: 128 rows each
    static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 128;
    static constexpr int BN2 = 256;                               // columns of the pair's tile = columns of one accumulator set
    static constexpr int TMEM_COLS = 2 * BN2;
};
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;                   // shared::cluster address of the same offset in CTA rank 0

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((unsigned short)3) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm_tf32x3_2cta_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                        const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl, const Params p) {
    constexpr int STAGES = Cfg2::STAGES, STAGE_BYTES = Cfg2::STAGE_BYTES, BN2 = Cfg2::BN2, TMEM_COLS = Cfg2::TMEM_COLS;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    const uint32_t full0 = smem_u32(bars);                            // used in the leader only
    const uint32_t empty0 = smem_u32(bars + STAGES);
    const uint32_t tmem_full0 = smem_u32(bars + 2 * STAGES);
    const uint32_t tmem_empty0 = smem_u32(bars + 2 * STAGES + 2);     // used in the leader only
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int GROUP = 8;                                              // grouped rasterisation over the pair tiles
    const int tile = blockIdx.x >> 1;
    const int group_size = GROUP * p.tiles_n;
    const int first_m = (tile / group_size) * GROUP;
    const int gm = min(p.tiles_m - first_m, GROUP);
    const int tm = first_m + (tile % group_size) % gm, tn = (tile % group_size) / gm;
    const int m0 = tm * 256 + (int)rank * BM;                         // this CTA's rows of A and C
    const int n0 = tn * BN2;
    const int nb0 = n0 + (int)rank * BN;                              // this CTA's half of the B rows
    const int batch = blockIdx.y;
    const int total_kb = (p.K + BK - 1) / BK;
    const int chunks = (total_kb + p.kb_per_chunk - 1) / p.kb_per_chunk;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, 2);              // the leader's arrive.expect_tx + the peer's arrive
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tmem_full0 + 8 * s, 1);
            mbar_init(tmem_empty0 + 8 * s, 8);        // lane 0 of the four epilogue warps of BOTH CTAs
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                               // both CTAs' barriers initialised and TMEM allocated before anyone signals
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer (both CTAs) =====
        if (lane == 0) {
            for (int kb = 0; kb < total_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(empty0 + 8 * s, ph ^ 1);
                const uint32_t fb = (full0 + 8 * s) & PEER_BIT_MASK;      // the leader's barrier
                if (leader) mbar_expect_tx(full0 + 8 * s, 2 * STAGE_BYTES);
                const uint32_t st = smem_u32(smem + s * STAGE_BYTES);
                const int k0 = kb * BK;
                tma_load_3d_2sm(st, &tmAh, fb, k0, m0, batch);
                tma_load_3d_2sm(st + TILE_BYTES, &tmAl, fb, k0, m0, batch);
                tma_load_3d_2sm(st + 2 * TILE_BYTES, &tmBh, fb, k0, nb0, batch);
                tma_load_3d_2sm(st + 3 * TILE_BYTES, &tmBl, fb, k0, nb0, batch);
                if (!leader) mbar_arrive_cluster(fb);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the leader's thread drives both tensor cores =====
        if (leader && lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN2 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            int kb = 0;
            for (int c = 0; c < chunks; ++c) {
                const int set = c & 1;
                if (c >= 2) {
                    mbar_wait(tmem_empty0 + 8 * set, ((c >> 1) - 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                const int kb_end = min(kb + p.kb_per_chunk, total_kb);
                const uint32_t acc = tmem_base + (uint32_t)(set * BN2);
                for (bool first = true; kb < kb_end; ++kb, first = false) {
                    const int s = kb % STAGES;
                    const uint32_t ph = (kb / STAGES) & 1;
                    mbar_wait(full0 + 8 * s, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = smem_u32(smem + s * STAGE_BYTES);
                    const uint64_t dAh = make_desc(st), dAl = make_desc(st + TILE_BYTES);
                    const uint64_t dBh = make_desc(st + 2 * TILE_BYTES), dBl = make_desc(st + 3 * TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {
                        const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);
                        umma_tf32_2sm(acc, dAl + adv, dBh + adv, idesc, (!first || k > 0) ? 1u : 0u);   // small terms first
                        umma_tf32_2sm(acc, dAh + adv, dBl + adv, idesc, 1u);
                        umma_tf32_2sm(acc, dAh + adv, dBh + adv, idesc, 1u);
                    }
                    umma_commit_2sm(empty0 + 8 * s);
                }
                umma_commit_2sm(tmem_full0 + 8 * set);
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue (both CTAs): own 128 TMEM lanes x 256 columns -> fp64 -> C =====
        const int quad = warp & 3;
        const int row = m0 + quad * 32 + lane;
        double* crow = p.C + (long long)batch * p.c_sb + (long long)row * p.c_sm;
        for (int c = 0; c < chunks; ++c) {
            const int set = c & 1;
            mbar_wait(tmem_full0 + 8 * set, (c >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int cc = 0; cc < BN2 / 32; ++cc) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(set * BN2 + cc * 32);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                      "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                      "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                      "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (row < p.M) {
                    const int col0 = n0 + cc * 32;
                    if (p.c_sn == 1 && col0 + 31 < p.N && ((reinterpret_cast<uintptr_t>(crow + col0) & 15) == 0)) {
                        if (c == 0) {
#pragma unroll
                            for (int j = 0; j < 32; j += 2)
                                *reinterpret_cast<double2*>(crow + col0 + j) = make_double2((double)__uint_as_float(v[j]), (double)__uint_as_float(v[j + 1]));
                        } else {
                            double2 o[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) o[j] = *reinterpret_cast<const double2*>(crow + col0 + 2 * j);
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                o[j].x += (double)__uint_as_float(v[2 * j]);
                                o[j].y += (double)__uint_as_float(v[2 * j + 1]);
                            }
#pragma unroll
                            for (int j = 0; j < 16; ++j) *reinterpret_cast<double2*>(crow + col0 + 2 * j) = o[j];
                        }
                    } else {
                        double o[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) o[j] = (c > 0 && col0 + j < p.N) ? crow[(long long)(col0 + j) * p.c_sn] : 0.0;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < p.N) crow[(long long)(col0 + j) * p.c_sn] = o[j] + (double)__uint_as_float(v[j]);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster((tmem_empty0 + 8 * set) & PEER_BIT_MASK);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                               // the peer's smem / barriers / TMEM are in use until both CTAs are here
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// fp64 (rows x K, MN-contiguous: s_row == 1) -> hi/lo fp32, K-major: a 32 x 32 tile goes through shared memory so that both the
// fp64 reads (along rows) and the fp32 writes (along k) are coalesced (the element-per-thread kernel below reads such an operand
// with a stride of a whole row: 8 useful bytes per 32-byte sector).
__global__ void __launch_bounds__(256)
split_tf32_t_kernel(const double* __restrict__ src, long long s_k, long long s_b, float* __restrict__ hi, float* __restrict__ lo,
                    long long rows, long long K, long long Kp) {
    __shared__ double t[32][33];
    const long long b = blockIdx.z;
    const long long r0 = (long long)blockIdx.x * 32, k0 = (long long)blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    src += b * s_b;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const long long k = k0 + ty + 8 * j, r = r0 + tx;
        t[ty + 8 * j][tx] = (k < K && r < rows) ? __ldg(src + k * s_k + r) : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const long long r = r0 + ty + 8 * j, k = k0 + tx;
        if (r < rows && k < Kp) {
            const double v = t[tx][ty + 8 * j];
            uint32_t hb, lb;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"((float)v));
            const float h = __uint_as_float(hb);
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"((float)(v - (double)h)));
            const long long o = (b * rows + r) * Kp + k;
            hi[o] = h;
            lo[o] = __uint_as_float(lb);
        }
    }
}

// fp64 (rows x K, any strides) -> hi/lo fp32, K-major with pitch Kp; one thread per output element, k fastest.
__global__ void __launch_bounds__(256)
split_tf32_kernel(const double* __restrict__ src, long long s_row, long long s_k, long long s_b, float* __restrict__ hi, float* __restrict__ lo,
                  long long rows, long long K, long long Kp, long long batch) {
    const long long total = batch * rows * Kp;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
        const long long k = i % Kp;
        const long long r = (i / Kp) % rows;
        const long long b = i / (Kp * rows);
        float h = 0.f, l = 0.f;
        if (k < K) {
            const double v = __ldg(src + b * s_b + r * s_row + k * s_k);
            uint32_t hb, lb;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"((float)v));
            h = __uint_as_float(hb);
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"((float)(v - (double)h)));
            l = __uint_as_float(lb);
        }
        hi[i] = h;
        lo[i] = l;
    }
}

static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static bool load_encode() {
    if (g_encode) return true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        cudaGetLastError();
        return false;
    }
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    return true;
}
static bool make_tmap_f32(CUtensorMap* map, const float* base, long long K, long long rows, long long batch, long long Kp, int box_rows) {
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)Kp * 4ull, (cuuint64_t)Kp * (cuuint64_t)rows * 4ull};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1u};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    return g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tf32

int gemm_tf32x3(const adb_gemm_desc* d, cudaStream_t st) {
    using namespace tf32;
    const long long M = d->M, N = d->N, K = d->K, batch = d->batch;
    if (M <= 0 || N <= 0 || batch <= 0) return 0;
    if (K <= 0) return set_error("gemm_tf32x3: K must be positive");
    if (M > 0x7fffffff || N > 0x7fffffff || K > 0x7fffffff || batch > 65535) return set_error("gemm_tf32x3: extent too large");
    if (!load_encode()) return set_error("gemm_tf32x3: cuTensorMapEncodeTiled unavailable");
    const long long Kp = (K + 3) & ~3LL;          // 16-byte row pitch for TMA
    float* ws = nullptr;
    const size_t a_elems = (size_t)batch * M * Kp, b_elems = (size_t)batch * N * Kp;
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&ws), (2 * a_elems + 2 * b_elems) * sizeof(float), st);
    if (e != cudaSuccess) return set_error("gemm_tf32x3: workspace allocation failed: %s", cudaGetErrorString(e));
    float *a_hi = ws, *a_lo = ws + a_elems, *b_hi = ws + 2 * a_elems, *b_lo = ws + 2 * a_elems + b_elems;
    auto grid_for = [](size_t n) { size_t g = (n + 255) / 256; return (unsigned)(g > 148 * 32 ? 148 * 32 : g); };
    auto split = [&](const double* src, long long s_row, long long s_k, long long s_b, float* hi, float* lo, long long rows, size_t elems) {
        if (s_row == 1 && s_k != 1 && rows >= 32 && K >= 32 && (rows + 31) / 32 < (1LL << 31) && (Kp + 31) / 32 <= 65535)
            split_tf32_t_kernel<<<dim3((unsigned)((rows + 31) / 32), (unsigned)((Kp + 31) / 32), (unsigned)batch), 256, 0, st>>>(src, s_k, s_b, hi, lo, rows, K, Kp);
        else
            split_tf32_kernel<<<grid_for(elems), 256, 0, st>>>(src, s_row, s_k, s_b, hi, lo, rows, K, Kp, batch);
    };
    split(d->A, d->a_sm, d->a_sk, d->a_sb, a_hi, a_lo, M, a_elems);
    split(d->B, d->b_sn, d->b_sk, d->b_sb, b_hi, b_lo, N, b_elems);
    count_launch(2);
    // 256-row CTA tiles (two accumulators sharing the B tiles) when there are enough of them to fill the GPU; ADB_TF32_MT=1|2 overrides
    static int mt_pref = -1;
    if (mt_pref < 0) { const char* ev = getenv("ADB_TF32_MT"); mt_pref = ev ? atoi(ev) : 0; }
    const long long tiles256 = ((M + 255) / 256) * ((N + BN - 1) / BN) * batch;
    const int MT = mt_pref == 1 ? 1 : (mt_pref == 2 ? 2 : (tiles256 >= 148 ? 2 : 1));
    // CTA-pair kernel (256 x 256 tile per cluster of two) when there are enough pair tiles to fill the 74 TPCs; ADB_TF32_2CTA=0 never, =2 always
    const char* pair_env = getenv("ADB_TF32_2CTA");
    const int pair_pref = pair_env ? atoi(pair_env) : 1;
    const long long pair_tiles = ((M + 255) / 256) * ((N + 255) / 256);
    const bool use_pair = pair_pref == 2 || (pair_pref == 1 && mt_pref == 0 && pair_tiles * batch >= 74);
    if (use_pair) {
        CUtensorMap pAh, pAl, pBh, pBl;
        if (!make_tmap_f32(&pAh, a_hi, K, M, batch, Kp, BM) || !make_tmap_f32(&pAl, a_lo, K, M, batch, Kp, BM) ||
            !make_tmap_f32(&pBh, b_hi, K, N, batch, Kp, BN) || !make_tmap_f32(&pBl, b_lo, K, N, batch, Kp, BN)) {
            cudaFreeAsync(ws, st);
            return set_error("gemm_tf32x3: cuTensorMapEncodeTiled rejected the operands");
        }
        static bool pair_attr = false;
        if (!pair_attr) {
            e = cudaFuncSetAttribute(gemm_tf32x3_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2::SMEM);
            if (e != cudaSuccess) { cudaFreeAsync(ws, st); return set_error("gemm_tf32x3: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); }
            pair_attr = true;
        }
        const int total_kb2 = (int)((K + BK - 1) / BK);
        const int chunks2 = (total_kb2 + 2048 / BK - 1) / (2048 / BK);
        Params q;
        q.M = (int)M; q.N = (int)N; q.K = (int)K;
        q.tiles_m = (int)((M + 255) / 256); q.tiles_n = (int)((N + 255) / 256);
        q.kb_per_chunk = (total_kb2 + chunks2 - 1) / chunks2; q._pad = 0;
        q.C = d->C; q.c_sm = d->c_sm; q.c_sn = d->c_sn; q.c_sb = d->c_sb;
        gemm_tf32x3_2cta_kernel<<<dim3((unsigned)(2 * pair_tiles), (unsigned)batch, 1u), THREADS, Cfg2::SMEM, st>>>(pAh, pAl, pBh, pBl, q);
        e = cudaGetLastError();
        cudaFreeAsync(ws, st);
        if (e != cudaSuccess) return set_error("gemm_tf32x3 (CTA pair) launch failed: %s", cudaGetErrorString(e));
        count_launch();
        return 0;
    }
    CUtensorMap tAh, tAl, tBh, tBl;
    if (!make_tmap_f32(&tAh, a_hi, K, M, batch, Kp, BM * MT) || !make_tmap_f32(&tAl, a_lo, K, M, batch, Kp, BM * MT) ||
        !make_tmap_f32(&tBh, b_hi, K, N, batch, Kp, BN) || !make_tmap_f32(&tBl, b_lo, K, N, batch, Kp, BN)) {
        cudaFreeAsync(ws, st);
        return set_error("gemm_tf32x3: cuTensorMapEncodeTiled rejected the operands");
    }
    static bool attr_set = false;
    if (!attr_set) {
        e = cudaFuncSetAttribute(gemm_tf32x3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<1>::SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tf32x3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<2>::SMEM);
        if (e != cudaSuccess) { cudaFreeAsync(ws, st); return set_error("gemm_tf32x3: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); }
        attr_set = true;
    }
    // The tensor core accumulates in fp32 with truncation, so the error of one accumulation chain grows ~linearly
    // with its length (measured 1.4e-5 at K = 2048, 3.3e-5 at 4096).  K is therefore processed in chunks of <= 2048: each chunk is
    // its own fp32 chain in TMEM, and the chunks are added in fp64 by the epilogue warps (chunk c + 1 accumulates in the second
    // TMEM set while chunk c is drained) - round 1 wrote every chunk's tile to a workspace and summed them in a second kernel.
    const int total_kb = (int)((K + BK - 1) / BK);
    const int max_kb = 2048 / BK;
    const int chunks = (total_kb + max_kb - 1) / max_kb;
    Params p;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.tiles_m = (int)((M + BM * MT - 1) / (BM * MT)); p.tiles_n = (int)((N + BN - 1) / BN);
    p.kb_per_chunk = (total_kb + chunks - 1) / chunks; p._pad = 0;
    p.C = d->C; p.c_sm = d->c_sm; p.c_sn = d->c_sn; p.c_sb = d->c_sb;
    dim3 grid((unsigned)(p.tiles_m * p.tiles_n), (unsigned)batch, 1u);
    if (MT == 2) gemm_tf32x3_kernel<2><<<grid, THREADS, Cfg<2>::SMEM, st>>>(tAh, tAl, tBh, tBl, p);
    else gemm_tf32x3_kernel<1><<<grid, THREADS, Cfg<1>::SMEM, st>>>(tAh, tAl, tBh, tBl, p);
    e = cudaGetLastError();
    cudaFreeAsync(ws, st);
    if (e != cudaSuccess) return set_error("gemm_tf32x3 launch failed: %s", cudaGetErrorString(e));
    count_launch();
    return 0;
}

}  // namespace adb

extern "C" int adb_gemm_tf32x3(const adb_gemm_desc* desc, void* stream) {
    return adb::gemm_tf32x3(desc, reinterpret_cast<cudaStream_t>(stream));
}
