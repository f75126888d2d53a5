// Sum-of-products reductions for sm_100a:  out[o] = sum_r prod_t op_t[o, r]   (1..4 operands, fp64).
// Replaces np.einsum (reference autodiff/core/ops.py:180) for every form that is NOT a true GEMM —
// axis sums ('ab->b', 'ijkl->ij', 'i->'), GEMV-like contractions ('ijkl,jl->ik', 'ijkl,ik->jl'), the
// Softmax denominator ('bi,o->bo') — and np.sum(keepdims) at reshape.py:14 and np.sum(np.square) at
// ops.py:369.  All of these are HBM-bound: each operand element is read once.
//
// One kernel template, parameterised by G = lanes cooperating on one output element:
//   G = 32 / 8 : the innermost REDUCED index is memory-contiguous -> lanes stride along it (coalesced),
//                partial sums combined with warp shuffles ("row" reductions).
//   G = 1      : the innermost OUTPUT index is memory-contiguous -> adjacent threads own adjacent outputs
//                (coalesced) and walk the reduced indices sequentially ("column" reductions).
// The reduced index space can be split across gridDim.y; partials go to a workspace and a second tiny
// kernel adds them in a fixed order (deterministic, no atomics).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>

#include "adb_internal.h"
#include "adb_math.cuh"

namespace adb {

constexpr int RD = 4;  // max collapsed dims on each side

struct RedParams {
    int n_ops;
    int splits;
    long long O;                 // number of outputs
    long long odim[RD], rdim[RD];
    long long op_os[ADB_RED_MAX_OPS][RD], op_rs[ADB_RED_MAX_OPS][RD];
    long long out_s[RD];
    const double* op[ADB_RED_MAX_OPS];
    double* out;
    double* partial;             // [splits][O] when splits > 1
    long long router;            // rdim[1]*rdim[2]*rdim[3]
    long long nblk0;             // inner blocks per outer index
    long long blk0;              // inner block length
    long long steps;             // router * nblk0
    long long steps_per_split;
    int square;                  // 1: single operand, accumulate x*x (FrobeniusNorm)
};

template <int G, int NOPS>
__global__ void __launch_bounds__(256) reduce_kernel(const __grid_constant__ RedParams P) {
    constexpr int GROUPS = 256 / G;
    const int lane = threadIdx.x % G;
    const long long gstep = (long long)gridDim.x * GROUPS;
    const int z = blockIdx.y;
    const long long s_lo = (long long)z * P.steps_per_split;
    long long s_hi = s_lo + P.steps_per_split;
    if (s_hi > P.steps) s_hi = P.steps;

    for (long long o = (long long)blockIdx.x * GROUPS + threadIdx.x / G; o < P.O; o += gstep) {
        long long rem = o;
        long long base[NOPS];
#pragma unroll
        for (int t = 0; t < NOPS; ++t) base[t] = 0;
        long long out_off = 0;
#pragma unroll
        for (int k = 0; k < RD; ++k) {
            const long long d = P.odim[k];
            long long i = 0;
            if (d > 1) { const long long qn = rem / d; i = rem - qn * d; rem = qn; }
#pragma unroll
            for (int t = 0; t < NOPS; ++t) base[t] += i * P.op_os[t][k];
            out_off += i * P.out_s[k];
        }
        double acc = 0.0;
        // position of the first step of this split in (outer multi-index, inner block)
        long long s = s_lo;
        long long outer = s / P.nblk0;
        long long blk = s - outer * P.nblk0;
        long long ri[RD];
        {
            long long r2 = outer;
#pragma unroll
            for (int k = 1; k < RD; ++k) {
                const long long d = P.rdim[k];
                long long i = 0;
                if (d > 1) { const long long qn = r2 / d; i = r2 - qn * d; r2 = qn; }
                ri[k] = i;
            }
        }
        for (; s < s_hi; ++s) {
            long long roff[NOPS];
#pragma unroll
            for (int t = 0; t < NOPS; ++t) {
                roff[t] = base[t];
#pragma unroll
                for (int k = 1; k < RD; ++k) roff[t] += ri[k] * P.op_rs[t][k];
            }
            const long long c_lo = blk * P.blk0;
            long long c_hi = c_lo + P.blk0;
            if (c_hi > P.rdim[0]) c_hi = P.rdim[0];
            long long c = c_lo + lane;
            // 4-way unrolled so several loads are in flight per lane
            for (; c + 3 * G < c_hi; c += 4 * G) {
                double pr[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const long long cc = c + u * G;
                    double v = __ldg(P.op[0] + roff[0] + cc * P.op_rs[0][0]);
                    if (NOPS == 1 && P.square) v = v * v;
#pragma unroll
                    for (int t = 1; t < NOPS; ++t) v *= __ldg(P.op[t] + roff[t] + cc * P.op_rs[t][0]);
                    pr[u] = v;
                }
                acc += pr[0]; acc += pr[1]; acc += pr[2]; acc += pr[3];
            }
            for (; c < c_hi; c += G) {
                double v = __ldg(P.op[0] + roff[0] + c * P.op_rs[0][0]);
                if (NOPS == 1 && P.square) v = v * v;
#pragma unroll
                for (int t = 1; t < NOPS; ++t) v *= __ldg(P.op[t] + roff[t] + c * P.op_rs[t][0]);
                acc += v;
            }
            // advance (blk, ri[1..]) odometer
            if (++blk == P.nblk0) {
                blk = 0;
#pragma unroll
                for (int k = 1; k < RD; ++k) {
                    if (++ri[k] < P.rdim[k]) break;
                    ri[k] = 0;
                }
            }
        }
        if (G > 1) {
#pragma unroll
            for (int off = G / 2; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off, G);
        }
        if (lane == 0) {
            if (P.splits == 1) P.out[out_off] = acc;
            else P.partial[(long long)z * P.O + o] = acc;
        }
    }
}

__global__ void __launch_bounds__(256) reduce_finalize_kernel(const __grid_constant__ RedParams P) {
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < P.O; o += step) {
        long long rem = o, out_off = 0;
#pragma unroll
        for (int k = 0; k < RD; ++k) {
            const long long d = P.odim[k];
            long long i = 0;
            if (d > 1) { const long long qn = rem / d; i = rem - qn * d; rem = qn; }
            out_off += i * P.out_s[k];
        }
        double acc = 0.0;
        for (int z = 0; z < P.splits; ++z) acc += P.partial[(long long)z * P.O + o];
        P.out[out_off] = acc;
    }
}

// ----------------------------------------------------------------------------- fast path (<= 3 output dims, <= 2 summed dims)
// Two streaming kernels, both built from what tools/probe_stream.cu measured on B200: 256-bit loads, four independent
// loads per operand in flight per lane, one-shot grids, multiply-shift index decoding.
//   red_row_kernel<G>  : the fastest-varying index of the big operand is SUMMED.  G lanes (a warp or a whole CTA)
//                        stride along the flattened summed index in quads; shuffle (+ shared memory) combine.
//   red_col_kernel<WR> : the fastest-varying index is an OUTPUT index.  A lane owns one output quad; the WR warps of a
//                        CTA take interleaved slices of the summed index and combine through shared memory.
// Few outputs => the summed range is sliced over gridDim.y; partials go to a workspace and the LAST slice to arrive
// (self-resetting arrival counter per blockIdx.x) re-adds them in slice order: one launch, deterministic result.
struct FastDiv32 {
    unsigned mul, shr, d, _pad;
};
static FastDiv32 make_fd(unsigned d) {
    FastDiv32 f;
    f.d = d; f._pad = 0;
    if (d <= 1) { f.mul = 0; f.shr = 0; return f; }
    unsigned lg = 0;
    while ((1ull << lg) < d) ++lg;
    const unsigned p = 31 + lg;
    f.mul = (unsigned)(((1ull << p) + d - 1) / d);
    f.shr = p - 32;
    return f;
}
__device__ __forceinline__ unsigned fd_div(unsigned n, const FastDiv32& f) { return f.d == 1 ? n : (__umulhi(n, f.mul) >> f.shr); }

struct Red3Params {
    int n_ops, square, splits, fin_vec;
    unsigned NG;          // row: outputs; col: output quads (groups of work that own outputs)
    unsigned O0;          // extent of the innermost output dim (elements)
    unsigned R, chunk;    // flattened summed range per output (row: in units of VEC elements) and its per-split length
    unsigned vecmask;     // bit t: operand t has unit stride along the vector dimension (else it is broadcast along it)
    unsigned out_vec;     // col: the output can be stored as one 256-bit quad
    FastDiv32 odiv[3];    // row: output dims; col: [quads of dim 0, dim 1, dim 2]
    FastDiv32 r0div;      // row: quads (VEC = 4) or elements of summed dim 0; col: elements of summed dim 0
    long long out_s[3];
    long long op_os[ADB_RED_MAX_OPS][3];
    long long op_r0s[ADB_RED_MAX_OPS], op_r1s[ADB_RED_MAX_OPS];
    const double* op[ADB_RED_MAX_OPS];
    double* out;
    double* partial;      // [splits][NG * fin_vec]
    int* counter;         // one arrival counter per blockIdx.x (self-resetting); the last slice to arrive adds the partials
    // row kernels: an operand that does not depend on the summed indices ('bi,o->bo': the Softmax denominator's [1];
    // 'ab,a->a') is factored out of the sum and multiplied in once per output - one load per output instead of one per step
    const double* post;
    long long post_os[3];
    int post_sqrt;        // row kernels, one output: store sqrt(sum)
};

// Arrival of this CTA's slice: returns true in every thread of the LAST slice to arrive for blockIdx.x, after which
// the partials of all slices are visible.  The counter is reset so the buffer can be reused by the next launch.
__device__ __forceinline__ bool red_last_slice(const Red3Params& P) {
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int ticket = atomicAdd(P.counter + blockIdx.x, 1);
        s_last = ticket == P.splits - 1;
        if (s_last) P.counter[blockIdx.x] = 0;
    }
    __syncthreads();
    if (s_last) __threadfence();
    return s_last != 0;
}

struct Quad {
    double v[4];
};
__device__ __forceinline__ Quad ldq(const double* p) {
    Quad r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3]) : "l"(p));
    return r;
}

constexpr int RED_U = 4;   // independent loads per operand in flight per lane

// One step of the summed index: fetch every operand's VEC elements (no use of the values here, so that the loads of
// several steps issue back to back), then fold them into a product.
template <int NOPS, int VEC, bool COL>
__device__ __forceinline__ void red_fetch(const Red3Params& P, const double* const (&base)[NOPS], unsigned rho, unsigned nvalid,
                                          double (&x)[NOPS][VEC]) {
    const unsigned r1 = fd_div(rho, P.r0div);
    const unsigned r0 = (rho - r1 * P.r0div.d) * (COL ? 1u : (unsigned)VEC);
#pragma unroll
    for (int t = 0; t < NOPS; ++t) {
        const double* p = base[t] + (long long)r0 * P.op_r0s[t] + (long long)r1 * P.op_r1s[t];
        if (VEC == 4) {
            if (!((P.vecmask >> t) & 1u)) { const double sc = __ldg(p); _Pragma("unroll") for (int e = 0; e < VEC; ++e) x[t][e] = sc; }
            else if (!COL || nvalid == 4) { const Quad q = ldq(p); x[t][0] = q.v[0]; x[t][1 % VEC] = q.v[1]; x[t][2 % VEC] = q.v[2]; x[t][3 % VEC] = q.v[3]; }
            else { _Pragma("unroll") for (int e = 0; e < VEC; ++e) x[t][e] = e < (int)nvalid ? __ldg(p + e) : 0.0; }
        } else {
            x[t][0] = __ldg(p);
        }
    }
}
template <int NOPS, int VEC>
__device__ __forceinline__ void red_fold(const Red3Params& P, const double (&x)[NOPS][VEC], double (&acc)[VEC]) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
        double v = (NOPS == 1 && P.square) ? x[0][e] * x[0][e] : x[0][e];
#pragma unroll
        for (int t = 1; t < NOPS; ++t) v *= x[t][e];
        acc[e] += v;
    }
}
// Walks rho = first, first + stride, ... < hi : full trips of RED_U steps without bounds checks (straight-line code,
// all loads of a trip in flight together), then the remainder one step at a time.
template <int NOPS, int VEC, bool COL>
__device__ __forceinline__ void red_walk(const Red3Params& P, const double* const (&base)[NOPS], unsigned first, unsigned stride, unsigned hi,
                                         unsigned nvalid, double (&acc)[VEC]) {
    unsigned r = first;
    for (; r + (RED_U - 1) * stride < hi; r += RED_U * stride) {
        double x[RED_U][NOPS][VEC];
#pragma unroll
        for (int u = 0; u < RED_U; ++u) red_fetch<NOPS, VEC, COL>(P, base, r + u * stride, nvalid, x[u]);
#pragma unroll
        for (int u = 0; u < RED_U; ++u) red_fold<NOPS, VEC>(P, x[u], acc);
    }
    for (; r < hi; r += stride) {
        double x[NOPS][VEC];
        red_fetch<NOPS, VEC, COL>(P, base, r, nvalid, x);
        red_fold<NOPS, VEC>(P, x, acc);
    }
}

// What one thread per output does with the finished sum: the factored-out operand (its offset is decoded here, off the hot
// path: keeping a running offset in the main loop's prologue cost the 'ijkl,jl->ik' kernel 19 %) or the root.
__device__ __noinline__ double red_post(const Red3Params& P, unsigned o, double v) {
    if (P.post) {
        long long off = 0;
        unsigned rem = o;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const unsigned qn = fd_div(rem, P.odiv[k]);
            off += (long long)(rem - qn * P.odiv[k].d) * P.post_os[k];
            rem = qn;
        }
        return v * __ldg(P.post + off);
    }
    return P.post_sqrt ? sqrt(v) : v;
}

template <int G, int NOPS, int VEC>
__global__ void __launch_bounds__(256) red_row_kernel(const __grid_constant__ Red3Params P) {
    constexpr int GROUPS = 256 / G;
    const unsigned lane = threadIdx.x % G;
    const unsigned o = blockIdx.x * GROUPS + threadIdx.x / G;
    const bool live = o < P.NG;
    const unsigned z = blockIdx.y;
    const unsigned lo = z * P.chunk;
    const unsigned hi = (lo + P.chunk < P.R) ? lo + P.chunk : P.R;

    unsigned rem = live ? o : 0;
    long long out_off = 0;
    const double* base[NOPS];
#pragma unroll
    for (int t = 0; t < NOPS; ++t) base[t] = P.op[t];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const unsigned qn = fd_div(rem, P.odiv[k]);
        const unsigned i = rem - qn * P.odiv[k].d;
        rem = qn;
        out_off += (long long)i * P.out_s[k];
#pragma unroll
        for (int t = 0; t < NOPS; ++t) base[t] += (long long)i * P.op_os[t][k];
    }
    double acc[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e] = 0.0;
    if (live) red_walk<NOPS, VEC, false>(P, base, lo + lane, G, hi, VEC, acc);
    double a = acc[0];
#pragma unroll
    for (int e = 1; e < VEC; ++e) a += acc[e];
#pragma unroll
    for (int off = (G > 32 ? 16 : G / 2); off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off, G > 32 ? 32 : G);
    if (G == 256) {
        __shared__ double warp_part[8];
        if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = a;
        __syncthreads();
        if (threadIdx.x == 0) {
            a = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) a += warp_part[w];
        }
    }
    if (P.splits == 1) {
        if (lane == 0 && live) P.out[out_off] = red_post(P, o, a);
        return;
    }
    if (lane == 0 && live) P.partial[(long long)z * P.NG + o] = a;
    if (!red_last_slice(P)) return;
    // last slice: every group re-adds its output's partials in slice order (fixed order => deterministic)
    if (!live) return;
    double t = 0.0;
    for (int zz = lane; zz < P.splits; zz += G) t += __ldcg(P.partial + (long long)zz * P.NG + o);
#pragma unroll
    for (int off = (G > 32 ? 16 : G / 2); off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off, G > 32 ? 32 : G);
    if (G == 256) {
        __shared__ double warp_fin[8];
        if ((threadIdx.x & 31) == 0) warp_fin[threadIdx.x >> 5] = t;
        __syncthreads();
        if (threadIdx.x == 0) {
            t = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) t += warp_fin[w];
        }
    }
    if (lane == 0) P.out[out_off] = red_post(P, o, t);
}

template <int WR, int NOPS, int VEC>
__global__ void __launch_bounds__(256, (VEC == 4 && NOPS == 2) ? 3 : 1) red_col_kernel(const __grid_constant__ Red3Params P) {
    constexpr int OQ_PER_CTA = 256 / WR;
    const unsigned lane = threadIdx.x % OQ_PER_CTA;
    const unsigned w = threadIdx.x / OQ_PER_CTA;       // slice of the summed index (0 when WR == 1)
    const unsigned oq = blockIdx.x * OQ_PER_CTA + lane;
    const bool live = oq < P.NG;
    const unsigned z = blockIdx.y;
    const unsigned lo = z * P.chunk;
    const unsigned hi = (lo + P.chunk < P.R) ? lo + P.chunk : P.R;

    unsigned rem = live ? oq : 0;
    long long out_off = 0;
    const double* base[NOPS];
    unsigned o0 = 0;
#pragma unroll
    for (int t = 0; t < NOPS; ++t) base[t] = P.op[t];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const unsigned qn = fd_div(rem, P.odiv[k]);
        unsigned i = rem - qn * P.odiv[k].d;
        rem = qn;
        if (k == 0) { i *= VEC; o0 = i; }
        out_off += (long long)i * P.out_s[k];
#pragma unroll
        for (int t = 0; t < NOPS; ++t) base[t] += (long long)i * P.op_os[t][k];
    }
    const unsigned nvalid = live ? ((P.O0 - o0 < (unsigned)VEC) ? P.O0 - o0 : (unsigned)VEC) : 0u;
    double acc[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e] = 0.0;
    if (live) red_walk<NOPS, VEC, true>(P, base, lo + w, WR, hi, nvalid, acc);
    if (WR > 1) {
        __shared__ double part[WR][OQ_PER_CTA][VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) part[w][lane][e] = acc[e];
        __syncthreads();
        if (w == 0) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                double a = 0.0;
#pragma unroll
                for (int ww = 0; ww < WR; ++ww) a += part[ww][lane][e];
                acc[e] = a;
            }
        }
    }
    const bool owner = live && w == 0;
    if (P.splits > 1) {
        if (owner) {
            double* q = P.partial + ((long long)z * P.NG + oq) * VEC;
#pragma unroll
            for (int e = 0; e < VEC; ++e) q[e] = acc[e];
        }
        if (!red_last_slice(P)) return;
        // last slice: the WR warps take interleaved slices of the partials, then combine in warp order (fixed order)
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] = 0.0;
        if (live) {
#pragma unroll 4
            for (int zz = w; zz < P.splits; zz += WR) {
                const double* q = P.partial + ((long long)zz * P.NG + oq) * VEC;
#pragma unroll
                for (int e = 0; e < VEC; ++e) acc[e] += __ldcg(q + e);
            }
        }
        if (WR > 1) {
            __shared__ double fin[WR][OQ_PER_CTA][VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) fin[w][lane][e] = acc[e];
            __syncthreads();
            if (w == 0) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    double a = 0.0;
#pragma unroll
                    for (int ww = 0; ww < WR; ++ww) a += fin[ww][lane][e];
                    acc[e] = a;
                }
            }
        }
    }
    if (!owner) return;
    double* q = P.out + out_off;
    if (VEC == 4 && nvalid == 4 && P.out_vec) {
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(q), "d"(acc[0]), "d"(acc[1 % VEC]), "d"(acc[2 % VEC]), "d"(acc[3 % VEC]) : "memory");
    } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e)
            if (e < (int)nvalid) q[e * P.out_s[0]] = acc[e];
    }
}

typedef void (*Red3Kernel)(const Red3Params);
template <int NOPS>
static Red3Kernel red3_kernel(bool row, int width, int vec) {
    if (row) {
        if (width == 256) return vec == 4 ? red_row_kernel<256, NOPS, 4> : red_row_kernel<256, NOPS, 1>;
        return vec == 4 ? red_row_kernel<32, NOPS, 4> : red_row_kernel<32, NOPS, 1>;
    }
    if (width == 8) return vec == 4 ? red_col_kernel<8, NOPS, 4> : red_col_kernel<8, NOPS, 1>;
    return vec == 4 ? red_col_kernel<1, NOPS, 4> : red_col_kernel<1, NOPS, 1>;
}

// CTA slots of the device for a kernel (SMs x resident CTAs per SM), cached per kernel.
static long long red3_slots(Red3Kernel k) {
    static std::map<Red3Kernel, long long> cache;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(k);
    if (it != cache.end()) return it->second;
    int occ = 0, dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, 256, 0) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 2; }
    const long long v = (long long)sms * occ;
    cache[k] = v;
    return v;
}

// Partial-sum workspace: one grow-only buffer per stream (launches on one stream are ordered, so consecutive
// reductions can reuse it without stream-ordered malloc/free calls between the kernels).  Growing it is stream-ordered
// (the old buffer is released behind the kernels already queued, nothing blocks).  While the stream is being CAPTURED
// into a CUDA graph the shared buffer is never used: a graph must not bake a pointer that a later growth would free, so
// the partials come from cudaMallocAsync / cudaFreeAsync, which turn into memory nodes private to that graph
// (*transient = the caller frees the buffer behind its kernel).
static double* red3_workspace(cudaStream_t st, size_t bytes, bool* transient) {
    struct Ws { double* p = nullptr; size_t n = 0; };
    static std::map<cudaStream_t, Ws> ws;
    static std::mutex mu;
    *transient = false;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); cs = cudaStreamCaptureStatusNone; }
    if (cs != cudaStreamCaptureStatusNone) {
        double* p = nullptr;
        if (cudaMallocAsync(reinterpret_cast<void**>(&p), bytes, st) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        *transient = true;
        return p;
    }
    std::lock_guard<std::mutex> lk(mu);
    Ws& w = ws[st];
    if (w.n < bytes) {
        if (w.p) { cudaFreeAsync(w.p, st); w.p = nullptr; w.n = 0; }
        size_t want = bytes < (size_t(8) << 20) ? (size_t(8) << 20) : bytes + bytes / 2;
        if (cudaMallocAsync(reinterpret_cast<void**>(&w.p), want, st) != cudaSuccess) { cudaGetLastError(); w.p = nullptr; return nullptr; }
        w.n = want;
    }
    return w.p;
}

// Self-resetting arrival counters: a ring, so that launches in flight on different streams never share a segment.
static int* red3_counters(long long n) {
    constexpr long long RING = 1 << 20;
    static int* ring = nullptr;
    static long long head = 0;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (n > RING) return nullptr;
    if (!ring) {
        if (cudaMalloc(reinterpret_cast<void**>(&ring), RING * sizeof(int)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        cudaMemset(ring, 0, RING * sizeof(int));
    }
    if (head + n > RING) head = 0;
    int* p = ring + head;
    head += n;
    return p;
}

static bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31) == 0; }

// Returns 0 = ran, -1 = not applicable (caller falls back to the generic kernel), >0 = error.
static int reduce3_try(const ReduceSpec& S, cudaStream_t st) {
    if (S.n_o > 3 || S.n_r > 2 || S.n_r < 1) return -1;
    long long O = 1, R = 1;
    for (int k = 0; k < S.n_o; ++k) O *= S.odim[k];
    for (int k = 0; k < S.n_r; ++k) R *= S.rdim[k];
    if (O <= 0) return 0;
    if (O >= (1LL << 31) || R >= (1LL << 31) || R <= 0) return -1;
    const bool row = S.group > 1;
    if (S.post_sqrt && (!row || O != 1)) return -1;   // (the generic path below takes the root in a second launch)
    if (row && R < 128) return -1;                    // short rows, many outputs: the sub-warp groups of reduce_kernel fit better
    const long long R0 = S.rdim[0], O0 = S.n_o > 0 ? S.odim[0] : 1;

    Red3Params P;
    memset(&P, 0, sizeof P);
    P.square = S.square;
    P.post_sqrt = S.post_sqrt;
    int nt = 0;
    for (int t = 0; t < S.n_ops; ++t) {
        const long long r0s = S.op_rs[t][0], r1s = S.n_r > 1 ? S.op_rs[t][1] : 0;
        if (row && !P.post && r0s == 0 && r1s == 0 && S.n_ops - t + nt > 1 && !S.square) {
            // independent of the summed indices: sum_r x[o,r] * s[o] = s[o] * sum_r x[o,r]  (1 ulp-level re-association)
            P.post = S.op[t];
            for (int k = 0; k < 3; ++k) P.post_os[k] = k < S.n_o ? S.op_os[t][k] : 0;
            continue;
        }
        P.op[nt] = S.op[t];
        P.op_r0s[nt] = r0s;
        P.op_r1s[nt] = r1s;
        for (int k = 0; k < 3; ++k) P.op_os[nt][k] = k < S.n_o ? S.op_os[t][k] : 0;
        ++nt;
    }
    P.n_ops = nt;
    const int n_ops = nt;
    for (int k = 0; k < 3; ++k) P.out_s[k] = k < S.n_o ? S.out_s[k] : 0;
    P.out = S.out;
    P.O0 = (unsigned)O0;

    // 256-bit path: every operand is unit-stride or broadcast along the vector dimension, and quad-aligned elsewhere
    int vec = 4;
    unsigned vecmask = 0;
    for (int t = 0; t < n_ops && vec == 4; ++t) {
        const long long vs = row ? P.op_r0s[t] : P.op_os[t][0];
        if (vs == 0) continue;
        if (vs != 1 || !aligned32(P.op[t])) { vec = 1; break; }
        vecmask |= 1u << t;
        if (row) { if (P.op_r1s[t] % 4) vec = 1; }
        else { if (P.op_r0s[t] % 4 || P.op_r1s[t] % 4) vec = 1; }
        for (int k = row ? 0 : 1; k < 3; ++k) if (P.op_os[t][k] % 4) vec = 1;
    }
    if (row && (R0 % 4)) vec = 1;
    if (vecmask == 0) vec = 1;
    P.vecmask = vecmask;

    long long NG, Rl;        // groups that own outputs, flattened summed length in loop units
    int width;
    if (row) {
        NG = O;
        Rl = vec == 4 ? R / 4 : R;
        P.r0div = make_fd((unsigned)(vec == 4 ? R0 / 4 : R0));
        for (int k = 0; k < 3; ++k) P.odiv[k] = make_fd(k < S.n_o ? (unsigned)S.odim[k] : 1u);
        width = (O < 148LL * 16) ? 256 : 32;          // few outputs: a whole CTA per output
        P.fin_vec = 1;
    } else {
        const long long Q0 = (O0 + vec - 1) / vec;
        NG = Q0;
        for (int k = 1; k < S.n_o; ++k) NG *= S.odim[k];
        Rl = R;
        P.r0div = make_fd((unsigned)R0);
        P.odiv[0] = make_fd((unsigned)Q0);
        for (int k = 1; k < 3; ++k) P.odiv[k] = make_fd(k < S.n_o ? (unsigned)S.odim[k] : 1u);
        width = (R >= 64) ? 8 : 1;                    // warps of a CTA share the summed range when it is long enough
        P.fin_vec = vec;
        bool ov = P.out_s[0] == 1 && aligned32(P.out);
        for (int k = 1; k < 3; ++k) if (P.out_s[k] % 4) ov = false;
        P.out_vec = ov ? 1u : 0u;
    }
    P.NG = (unsigned)NG;
    Red3Kernel kernel = nullptr;
    switch (P.n_ops) {
        case 1: kernel = red3_kernel<1>(row, width, vec); break;
        case 2: kernel = red3_kernel<2>(row, width, vec); break;
        case 3: kernel = red3_kernel<3>(row, width, vec); break;
        default: kernel = red3_kernel<4>(row, width, vec); break;
    }
    const long long groups_per_cta = 256 / width;
    const long long blocks_x = (NG + groups_per_cta - 1) / groups_per_cta;
    // Slicing the summed range: below a few waves of CTAs, slice the summed range (one-shot grids of many short CTAs stream
    // best on B200, tools/probe_stream.cu).  Every slice keeps >= 2 trips.
    const long long per_trip = (long long)width * RED_U;
    const long long slots = red3_slots(kernel);
    const long long max_splits = Rl / (per_trip * 2) > 0 ? Rl / (per_trip * 2) : 1;
    long long splits = 1;
    // Row kernels: 4 waves (measured, profiles/r01_reduce_waves.log: 'ijkl,jl->ik' 0.322 -> 0.303 ms, 'ab->a' and 'ab->' a little
    // better too); column kernels: 8 (fewer, longer CTAs help the 256 MiB 'ab->b' but cost 'ijkl,ik->jl' on 2 GiB 10-29 %).
    // ADB_RED_WAVES overrides both (experiments).
    static long long waves_env = -1;
    if (waves_env < 0) { const char* e = getenv("ADB_RED_WAVES"); waves_env = e ? atoll(e) : 0; if (waves_env < 0) waves_env = 0; }
    const long long waves = waves_env > 0 ? waves_env : (row ? 4 : 8);
    if (blocks_x < waves * slots) splits = (waves * slots + blocks_x - 1) / blocks_x;
    if (splits > max_splits) splits = max_splits;
    if (splits > 65535) splits = 65535;
    if (splits < 1) splits = 1;
    long long chunk = (Rl + splits - 1) / splits;
    chunk = (chunk + per_trip - 1) / per_trip * per_trip;
    splits = (Rl + chunk - 1) / chunk;
    if (splits > 1) {
        P.counter = red3_counters(blocks_x);
        if (!P.counter) { splits = 1; chunk = Rl; }
    }
    P.R = (unsigned)Rl;
    P.chunk = (unsigned)chunk;
    P.splits = (int)splits;
    bool transient_ws = false;
    if (splits > 1) {
        P.partial = red3_workspace(st, (size_t)splits * NG * P.fin_vec * sizeof(double), &transient_ws);
        if (!P.partial) return set_error("reduce: workspace allocation failed");
    }
    dim3 grid((unsigned)blocks_x, (unsigned)splits, 1);
    kernel<<<grid, 256, 0, st>>>(P);
    cudaError_t e = cudaGetLastError();
    if (transient_ws) cudaFreeAsync(P.partial, st);
    if (e != cudaSuccess) return set_error("reduce launch failed: %s", cudaGetErrorString(e));
    count_launch();
    return 0;
}

__global__ void sqrt_scalar_kernel(double* p) { *p = sqrt(*p); }

template <int G>
static void launch_reduce(const RedParams& P, dim3 grid, cudaStream_t st) {
    switch (P.n_ops) {
        case 1: reduce_kernel<G, 1><<<grid, 256, 0, st>>>(P); break;
        case 2: reduce_kernel<G, 2><<<grid, 256, 0, st>>>(P); break;
        case 3: reduce_kernel<G, 3><<<grid, 256, 0, st>>>(P); break;
        default: reduce_kernel<G, 4><<<grid, 256, 0, st>>>(P); break;
    }
}

// dims are given innermost-first and already collapsed by the planner.
int reduce_run(const ReduceSpec& S, cudaStream_t st) {
    if (S.n_ops < 1 || S.n_ops > ADB_RED_MAX_OPS) return set_error("reduce: %d operands (max %d)", S.n_ops, ADB_RED_MAX_OPS);
    {
        const int rc = reduce3_try(S, st);
        if (rc >= 0) return rc;
    }
    RedParams P;
    P.n_ops = S.n_ops;
    P.square = S.square;
    P.O = 1;
    long long R = 1;
    for (int k = 0; k < RD; ++k) {
        P.odim[k] = k < S.n_o ? S.odim[k] : 1;
        P.rdim[k] = k < S.n_r ? S.rdim[k] : 1;
        P.out_s[k] = k < S.n_o ? S.out_s[k] : 0;
        P.O *= P.odim[k];
        R *= P.rdim[k];
        for (int t = 0; t < S.n_ops; ++t) {
            P.op_os[t][k] = k < S.n_o ? S.op_os[t][k] : 0;
            P.op_rs[t][k] = k < S.n_r ? S.op_rs[t][k] : 0;
        }
    }
    for (int t = 0; t < S.n_ops; ++t) P.op[t] = S.op[t];
    P.out = S.out;
    if (P.O == 0) return 0;
    const int G = S.group;
    if (G != 1 && G != 8 && G != 32) return set_error("reduce: bad group %d", G);
    P.router = P.rdim[1] * P.rdim[2] * P.rdim[3];
    P.blk0 = (long long)G * 16;                       // inner block: 16 elements per lane
    if (G == 1) P.blk0 = 64;
    P.nblk0 = (P.rdim[0] + P.blk0 - 1) / P.blk0;
    if (P.nblk0 < 1) P.nblk0 = 1;
    P.steps = P.router * P.nblk0;

    const int sms = 148;
    const long long groups_per_block = 256 / G;
    long long blocks_x = (P.O + groups_per_block - 1) / groups_per_block;
    const long long target_blocks = (long long)sms * 8;
    long long splits = 1;
    if (blocks_x < target_blocks && R >= 4096) {
        splits = target_blocks / blocks_x;
        const long long min_steps = (G == 1) ? 4 : 2;     // keep every split busy for a while
        if (splits > P.steps / min_steps) splits = P.steps / min_steps;
        if (splits > 1024) splits = 1024;
        if (splits < 1) splits = 1;
    }
    if (blocks_x > target_blocks * 4) blocks_x = target_blocks * 4;  // grid-stride beyond
    P.steps_per_split = (P.steps + splits - 1) / splits;
    splits = (P.steps + P.steps_per_split - 1) / P.steps_per_split;
    P.splits = (int)splits;
    P.partial = nullptr;
    if (splits > 1) {
        cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&P.partial), (size_t)splits * P.O * sizeof(double), st);
        if (e != cudaSuccess) return set_error("reduce: workspace allocation failed: %s", cudaGetErrorString(e));
    }
    dim3 grid((unsigned)blocks_x, (unsigned)splits, 1);
    if (G == 32) launch_reduce<32>(P, grid, st);
    else if (G == 8) launch_reduce<8>(P, grid, st);
    else launch_reduce<1>(P, grid, st);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("reduce launch failed: %s", cudaGetErrorString(e));
    count_launch();
    if (splits > 1) {
        long long fb = (P.O + 255) / 256;
        if (fb > sms * 8) fb = sms * 8;
        reduce_finalize_kernel<<<(unsigned)fb, 256, 0, st>>>(P);
        e = cudaGetLastError();
        if (e != cudaSuccess) return set_error("reduce finalize launch failed: %s", cudaGetErrorString(e));
        count_launch();
        cudaFreeAsync(P.partial, st);
    }
    if (S.post_sqrt) {
        sqrt_scalar_kernel<<<1, 1, 0, st>>>(S.out);
        e = cudaGetLastError();
        if (e != cudaSuccess) return set_error("reduce sqrt launch failed: %s", cudaGetErrorString(e));
        count_launch();
    }
    return 0;
}

// ----------------------------------------------------------------------------- SoftmaxCEWithLogits
// reference ops.py:243-255 — one warp per row:
//   lse = max + log(sum(exp(z - max)));  out = -sum(l*z - l*lse);  flag if !allclose(sum(l), 1)
__global__ void __launch_bounds__(256) softmax_ce_kernel(const double* __restrict__ labels, long long l_sr, long long l_sc,
                                                         const double* __restrict__ logits, long long z_sr, long long z_sc,
                                                         double* __restrict__ out, long long o_s, long long rows, long long cols,
                                                         int* __restrict__ flag) {
    const int lane = threadIdx.x & 31;
    const long long wstep = (long long)gridDim.x * 8;
    for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += wstep) {
        const double* z = logits + r * z_sr;
        const double* l = labels + r * l_sr;
        double mx = -INFINITY;
        bool nan_seen = false;
        for (long long c = lane; c < cols; c += 32) {
            const double v = __ldg(z + c * z_sc);
            nan_seen |= (v != v);
            mx = fmax(mx, v);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        if (__any_sync(0xffffffffu, nan_seen)) mx = NAN;      // np.max propagates NaN
        double se = 0.0, lsum = 0.0;
        for (long long c = lane; c < cols; c += 32) {
            se += exp_nonpos(__ldg(z + c * z_sc) - mx);
            lsum += __ldg(l + c * l_sc);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            se += __shfl_xor_sync(0xffffffffu, se, off);
            lsum += __shfl_xor_sync(0xffffffffu, lsum, off);
        }
        const double lse = mx + log(se);
        double s = 0.0;
        for (long long c = lane; c < cols; c += 32) {
            const double lv = __ldg(l + c * l_sc);
            s += lv * __ldg(z + c * z_sc) - lv * lse;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) {
            out[r * o_s] = -s;
            // np.allclose(a, 1): |a - 1| <= atol + rtol*|1| with atol=1e-8, rtol=1e-5; NaN fails
            if (!(fabs(lsum - 1.0) <= 1e-8 + 1e-5)) atomicOr(flag, 1);
        }
    }
}

// Wide rows: one CTA per row, the row of logits AND labels held in registers (K quads per thread each), so HBM is
// read exactly once.  K quads per thread, NT threads per CTA: cols <= NT * 4 * K.  Single pass ("online" log-sum-exp): every warp reduces
// its elements to (max m_w, sum exp(z - m_w), sum l, sum l*z) - the warp max first, so each element costs one exp -
// and after ONE shared-memory exchange warp 0 rescales the per-warp sums to the row max (one exp per warp) and adds
// them: a row costs one block barrier instead of four dependent phases.  out = -(sum l*z - lse * sum l), the reference's
// -sum(l*z - l*lse) re-associated (error ~1e-16 of the terms; parity bar 1e-12).
template <int K, int NT>
__global__ void __launch_bounds__(NT) softmax_ce_wide_kernel(const double* __restrict__ labels, long long l_sr, const double* __restrict__ logits,
                                                             long long z_sr, double* __restrict__ out, long long o_s, long long rows,
                                                             unsigned cols, int* __restrict__ flag, int exp_tab) {
    constexpr int NW = NT / 32;
    __shared__ double sm[4][NW];
    const long long r = blockIdx.x;
    const double* z = logits + r * z_sr;
    const double* l = labels + r * l_sr;
    Quad zq[K], lq[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const unsigned c = (k * (unsigned)NT + threadIdx.x) * 4u;
        if (c + 3 < cols) {
            zq[k] = ldq(z + c);
            lq[k] = ldq(l + c);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                zq[k].v[e] = c + e < cols ? __ldg(z + c + e) : -INFINITY;   // neutral: never the max, exp() = 0, label 0
                lq[k].v[e] = c + e < cols ? __ldg(l + c + e) : 0.0;
            }
        }
    }
    double m = -INFINITY;
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int e = 0; e < 4; ++e) m = fmax(m, zq[k].v[e]);
    }
    // the WARP's max first (plain fmax shuffles), so that each element costs exactly one exp
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
    // fmax skips NaN, exp_nonpos does not: a NaN logit makes `se` NaN without a per-element test (np.max propagates NaN).
    // A warp that holds nothing but -inf / padding subtracts 0 instead of its max (-inf - -inf would be NaN): se = 0.
    const double mm = m == -INFINITY ? 0.0 : m;
    double se = 0.0, lsum = 0.0, dot = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const unsigned c = (k * (unsigned)NT + threadIdx.x) * 4u + e;
            dot += c < cols ? lq[k].v[e] * zq[k].v[e] : 0.0;   // padding is -inf -> exp = 0
            lsum += lq[k].v[e];
        }
    }
    if (exp_tab) {                                             // (uniform branch: one of two straight-line copies; the 4K evaluations interleave)
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int e = 0; e < 4; ++e) se += exp_nonpos_tab(zq[k].v[e] - mm);
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int e = 0; e < 4; ++e) se += exp_nonpos(zq[k].v[e] - mm);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        se += __shfl_xor_sync(0xffffffffu, se, off);
        lsum += __shfl_xor_sync(0xffffffffu, lsum, off);
        dot += __shfl_xor_sync(0xffffffffu, dot, off);
    }
    if ((threadIdx.x & 31) == 0) {
        const int w = threadIdx.x >> 5;
        sm[0][w] = m; sm[1][w] = se; sm[2][w] = lsum; sm[3][w] = dot;
    }
    __syncthreads();
    if (threadIdx.x < 32) {                    // warp 0 combines the NW per-warp results (log-sum-exp pairs rescaled)
        const int w = threadIdx.x;
        double M = w < NW ? sm[0][w] : -INFINITY, S = w < NW ? sm[1][w] : 0.0;
        double L = w < NW ? sm[2][w] : 0.0, D = w < NW ? sm[3][w] : 0.0;
        // rescale every warp's sum to the row max with ONE exp per lane (max first, then a plain sum): the chain the
        // other warps wait behind is 8 shuffles + 1 exp instead of log2(NW) dependent (exp, exp) pairs
        double Mx = M;
#pragma unroll
        for (int off = NW / 2; off > 0; off >>= 1) Mx = fmax(Mx, __shfl_xor_sync(0xffffffffu, Mx, off));
        S = (M == -INFINITY) ? S : S * exp_nonpos(M - Mx);    // (S is 0, or NaN when the warp saw a NaN, for M = -inf)
        M = Mx;
#pragma unroll
        for (int off = NW / 2; off > 0; off >>= 1) {
            S += __shfl_xor_sync(0xffffffffu, S, off);
            L += __shfl_xor_sync(0xffffffffu, L, off);
            D += __shfl_xor_sync(0xffffffffu, D, off);
        }
        if (w == 0) {
            if (M == -INFINITY) S = NAN;       // a row of -inf: numpy computes exp(-inf - -inf) = nan
            const double lse = M + log(S);
            out[r * o_s] = -(D - lse * L);
            if (!(fabs(L - 1.0) <= 1e-8 + 1e-5)) atomicOr(flag, 1);
        }
    }
}

// Narrow rows (cols <= 512): one WARP per row, the row of logits and labels in registers (K quads per lane each), every
// reduction a shuffle - no shared memory, no block barrier, and a CTA's 8 warps work on 8 rows whose loads are all in flight
// together.  (The one-CTA-per-row kernel above at 512 columns: half its threads hold padding and every row pays a
// __syncthreads - 0.39 of the HBM roof, profiles/r01_kernel_microbench_v9.log; C5's per-step loss and C1/C3's 10-class
// losses are this shape class.)  Same arithmetic as the wide kernel with NW = 1.
template <int K>
__global__ void __launch_bounds__(256) softmax_ce_warp_kernel(const double* __restrict__ labels, long long l_sr, const double* __restrict__ logits,
                                                              long long z_sr, double* __restrict__ out, long long o_s, long long rows,
                                                              unsigned cols, int* __restrict__ flag, int exp_tab) {
    const unsigned lane = threadIdx.x & 31;
    const long long wstep = (long long)gridDim.x * 8;
    for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += wstep) {
        const double* z = logits + r * z_sr;
        const double* l = labels + r * l_sr;
        Quad zq[K], lq[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const unsigned c = (k * 32u + lane) * 4u;
            if (c + 3 < cols) {
                zq[k] = ldq(z + c);
                lq[k] = ldq(l + c);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    zq[k].v[e] = c + e < cols ? __ldg(z + c + e) : -INFINITY;   // neutral: never the max, exp() = 0, label 0
                    lq[k].v[e] = c + e < cols ? __ldg(l + c + e) : 0.0;
                }
            }
        }
        double m = -INFINITY;
#pragma unroll
        for (int k = 0; k < K; ++k) {
#pragma unroll
            for (int e = 0; e < 4; ++e) m = fmax(m, zq[k].v[e]);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
        const double mm = m == -INFINITY ? 0.0 : m;      // (fmax skips NaN, exp_nonpos does not: a NaN logit makes `se` NaN)
        double se = 0.0, lsum = 0.0, dot = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const unsigned c = (k * 32u + lane) * 4u + e;
                dot += c < cols ? lq[k].v[e] * zq[k].v[e] : 0.0;
                lsum += lq[k].v[e];
            }
        }
        if (exp_tab) {
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int e = 0; e < 4; ++e) se += exp_nonpos_tab(zq[k].v[e] - mm);
        } else {
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int e = 0; e < 4; ++e) se += exp_nonpos(zq[k].v[e] - mm);
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
            out[r * o_s] = -(dot - lse * lsum);
            if (!(fabs(lsum - 1.0) <= 1e-8 + 1e-5)) atomicOr(flag, 1);
        }
    }
}

// (A TMA-staged persistent variant - one 512-thread CTA per SM, rows prefetched into a shared-memory ring with 1-D bulk
// copies - was measured at 0.47 of the HBM roof against 0.66 for the kernel above: with 64 KB per row only one CTA fits
// per SM, and the per-row dependent chain (max -> exp -> sum -> log) then has no second CTA to overlap with.)

int sce_exp_tab() {                               // ADB_SCE_EXP=taylor: the degree-13 exp_nonpos instead of the table variant (A/B measurements)
    static int mode = -1;
    if (mode < 0) { const char* e = getenv("ADB_SCE_EXP"); mode = (e && !strcmp(e, "taylor")) ? 0 : 1; }
    return mode;
}

static bool sce_ring_enabled() {                 // ADB_SCE_RING=0 pins the register-resident kernels (A/B measurements, tests)
    const char* e = getenv("ADB_SCE_RING");
    return !(e && e[0] == '0');
}

int softmax_ce_run(const adb_tensor* labels, const adb_tensor* logits, const adb_tensor* out, int* flag, cudaStream_t st) {
    if (labels->rank != 2 || logits->rank != 2) return set_error("softmax_ce: labels and logits must be 2-D (got %d, %d)", labels->rank, logits->rank);
    const long long rows = logits->shape[0], cols = logits->shape[1];
    if (labels->shape[0] != rows || labels->shape[1] != cols) return set_error("softmax_ce: labels %lldx%lld vs logits %lldx%lld", (long long)labels->shape[0], (long long)labels->shape[1], rows, cols);
    if (out->rank != 1 || out->shape[0] != rows) return set_error("softmax_ce: out must have %lld elements", rows);
    if (rows == 0) return 0;
    const bool quad_ok = labels->stride[1] == 1 && logits->stride[1] == 1 && aligned32(labels->ptr) && aligned32(logits->ptr) &&
                         labels->stride[0] % 4 == 0 && logits->stride[0] % 4 == 0;
    if (quad_ok && sce_ring_enabled()) {
        int r = softmax_ce_ring_try(labels, logits, out, flag, st);           // persistent ring kernels (softmax_ce_ring.cu)
        if (r < 0) r = softmax_ce_ring_narrow_try(labels, logits, out, flag, st);
        if (r >= 0) return r;
    }
    static int warp_rows = -1;                    // ADB_SCE_WARP=0: the round-1 dispatch (A/B measurements)
    if (warp_rows < 0) { const char* e = getenv("ADB_SCE_WARP"); warp_rows = (e && e[0] == '0') ? 0 : 1; }
    // (measured, profiles/r02_softmax_ce_warp.log: 0.16 -> 0.32 of the HBM roof at 64 columns, 0.36 -> 0.61 at 128, 0.46 -> 0.74 at 256,
    // 0.39 -> 0.71 at 512; beyond 512 columns a lane would hold 8 + 8 quads - 183 registers, one CTA per SM - and the
    // one-CTA-per-row kernel below wins: 0.56 at 768, 0.71 at 1024)
    if (quad_ok && warp_rows && cols <= 512) {
        long long blocks = (rows + 7) / 8;
        if (blocks > 148 * 32) blocks = 148 * 32;
#define ADB_SCEW(KK) softmax_ce_warp_kernel<KK><<<(unsigned)blocks, 256, 0, st>>>(labels->ptr, labels->stride[0], logits->ptr, logits->stride[0], out->ptr, \
                                                                                   out->stride[0], rows, (unsigned)cols, flag, sce_exp_tab())
        if (cols <= 128) ADB_SCEW(1);
        else if (cols <= 256) ADB_SCEW(2);
        else ADB_SCEW(4);
#undef ADB_SCEW
    } else if (quad_ok && cols >= 512 && cols <= 8192 && rows < (1LL << 31)) {
        const unsigned g = (unsigned)rows;
#define ADB_SCE(KK, NTT) softmax_ce_wide_kernel<KK, NTT><<<g, NTT, 0, st>>>(labels->ptr, labels->stride[0], logits->ptr, logits->stride[0], out->ptr, \
                                                                            out->stride[0], rows, (unsigned)cols, flag, sce_exp_tab())
        if (cols <= 1024) ADB_SCE(1, 256);
        else if (cols <= 2048) ADB_SCE(2, 256);
        else if (cols <= 4096) ADB_SCE(2, 512);
        else ADB_SCE(4, 512);
#undef ADB_SCE
    } else {
        long long blocks = (rows + 7) / 8;
        if (blocks > 148 * 16) blocks = 148 * 16;
        softmax_ce_kernel<<<(unsigned)blocks, 256, 0, st>>>(labels->ptr, labels->stride[0], labels->stride[1], logits->ptr, logits->stride[0],
                                                            logits->stride[1], out->ptr, out->stride[0], rows, cols, flag);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("softmax_ce launch failed: %s", cudaGetErrorString(e));
    count_launch();
    return 0;
}

}  // namespace adb

extern "C" int adb_softmax_ce(const adb_tensor* labels, const adb_tensor* logits, const adb_tensor* out, int* flag, void* stream) {
    return adb::softmax_ce_run(labels, logits, out, flag, reinterpret_cast<cudaStream_t>(stream));
}
