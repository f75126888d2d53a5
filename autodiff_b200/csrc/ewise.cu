// Fused elementwise programs for sm_100a — replaces the numpy ufunc call sites of the reference's
// elementwise `_eval` bodies (autodiff/core/ops.py:53 Add, :75 Mul, :96 Negate, :116 Recipr, :228 ReLU,
// :274 SigmoidCE, :290 Pow, :309 Log, :339 Exp, :354 Sigmoid, :369 sqrt, :391 NormalDistribution) and the
// data-movement nodes (reshape.py:35 Concat, :94 Slice, :99-100 scatter, :126-129 Pad) which are the
// same kernel running the two-instruction program [LOAD_IN 0; STORE_OUT 0].
//
// A program is a short accumulator-machine bytecode that REPLAYS THE REFERENCE'S fp64 OPERATION ORDER
// (1/(x+eps), x*(x>0), left-to-right n-ary sums/products ...), so fused chains stay bit-compatible with
// the unfused reference up to libm ulp differences.  Each thread evaluates the whole program on 4
// consecutive elements of the innermost dimension (interpreter overhead amortised 4x, 2x 128-bit
// loads/stores per operand), operands broadcast through zero strides.  HBM-bound by construction:
// every operand element is read once and every output element written once.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "adb_internal.h"

namespace adb {

constexpr int EW_RANK = 6;
constexpr int EW_TEMPS = 8;

// q = n / d for n < 2^31 without a divide: d == 1 ? n : umulhi(n, mul) >> shr   (round-up magic number)
struct FastDiv {
    unsigned mul, shr, d, _pad;
};
static FastDiv make_fastdiv(unsigned d) {
    FastDiv f;
    f.d = d; f._pad = 0;
    if (d <= 1) { f.mul = 0; f.shr = 0; return f; }
    unsigned lg = 0;
    while ((1ull << lg) < d) ++lg;
    const unsigned p = 31 + lg;
    f.mul = (unsigned)(((1ull << p) + d - 1) / d);
    f.shr = p - 32;
    return f;
}
__device__ __forceinline__ unsigned fastdiv(unsigned n, const FastDiv& f) { return f.d == 1 ? n : (__umulhi(n, f.mul) >> f.shr); }

struct EwParams {
    int n_in, n_out, n_instr, nd;
    long long dim[EW_RANK];  // dim[0] innermost (collapsed iteration space of the outputs)
    FastDiv div_chunks0;     // divisor ceil(dim[0]/4)
    FastDiv div_dim[EW_RANK];
    unsigned long long total;  // work items = chunks0 * prod(dim[1..])
    long long in_stride[ADB_EW_MAX_IN][EW_RANK];
    long long out_stride[ADB_EW_MAX_OUT][EW_RANK];
    const double* in[ADB_EW_MAX_IN];
    double* out[ADB_EW_MAX_OUT];
    unsigned in_vec, out_vec;  // bit i: 16-byte aligned unit-stride access along dim0 is legal
    adb_ew_instr code[ADB_EW_MAX_INSTR];
};

template <int V>
struct Vec {
    double v[V];
};

// A thread owns V/2 element PAIRS along the innermost dimension, PS elements apart:
//   PS = 64 ("lane-contiguous": the 32 lanes of a warp cover 512 contiguous bytes per 128-bit access; used when
//            the innermost extent is long) or PS = 2 ("thread-contiguous": V consecutive elements per thread).
template <int V, int PS>
__device__ __forceinline__ Vec<V> load_in(const EwParams& P, int i, long long base, long long pos0) {
    Vec<V> r;
    const double* p = P.in[i] + base;
    const long long s0 = P.in_stride[i][0];
    const long long dim0 = P.dim[0];
    if (s0 == 0) {
        const double x = __ldg(p);
#pragma unroll
        for (int e = 0; e < V; ++e) r.v[e] = x;
    } else if ((P.in_vec >> i) & 1u) {
#pragma unroll
        for (int j = 0; j < V / 2; ++j) {
            const long long pos = pos0 + j * PS;
            if (pos + 1 < dim0) {
                const double2 a = __ldg(reinterpret_cast<const double2*>(p + j * PS));
                r.v[2 * j] = a.x;
                r.v[2 * j + 1] = a.y;
            } else {
                r.v[2 * j] = pos < dim0 ? __ldg(p + j * PS) : 0.0;
                r.v[2 * j + 1] = 0.0;
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < V / 2; ++j) {
            const long long pos = pos0 + j * PS;
            r.v[2 * j] = pos < dim0 ? __ldg(p + (j * PS) * s0) : 0.0;
            r.v[2 * j + 1] = pos + 1 < dim0 ? __ldg(p + (j * PS + 1) * s0) : 0.0;
        }
    }
    return r;
}

template <int V, int PS>
__device__ __forceinline__ void store_out(const EwParams& P, int jout, long long base, long long pos0, const Vec<V>& a) {
    double* p = P.out[jout] + base;
    const long long s0 = P.out_stride[jout][0];
    const long long dim0 = P.dim[0];
    const bool vec = (P.out_vec >> jout) & 1u;
#pragma unroll
    for (int j = 0; j < V / 2; ++j) {
        const long long pos = pos0 + j * PS;
        if (vec && pos + 1 < dim0) {
            __stcs(reinterpret_cast<double2*>(p + j * PS), make_double2(a.v[2 * j], a.v[2 * j + 1]));
        } else {
            if (pos < dim0) p[(j * PS) * s0] = a.v[2 * j];
            if (pos + 1 < dim0) p[(j * PS + 1) * s0] = a.v[2 * j + 1];
        }
    }
}

#define EW_FOR for (int e = 0; e < V; ++e)

__device__ __noinline__ double ew_pow(double a, double b) { return pow(a, b); }   // keeps pow's registers out of the main loop

// V elements per thread per step (V = 8: four 128-bit loads in flight per operand), NT register temps.
// LC: lane-contiguous pair layout (work item = one warp-wide tile of 32*V elements) vs thread-contiguous.
template <int V, int NT, bool BIG, bool LC>
__global__ void __launch_bounds__(256, (V == 8 && NT == 0) ? 3 : 2) ewise_kernel(const __grid_constant__ EwParams P) {
    constexpr int PS = LC ? 64 : 2;
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long nthreads = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long step = LC ? nthreads / 32 : nthreads;
    const unsigned lane = threadIdx.x & 31u;
    for (unsigned long long w = LC ? tid / 32 : tid; w < P.total; w += step) {
        long long idx[EW_RANK];
        if (BIG) {      // >= 2^31 work items: plain 64-bit divisions
            const long long chunks0 = (long long)P.div_chunks0.d;
            long long o = (long long)w / chunks0;
            idx[0] = ((long long)w - o * chunks0) * (LC ? 32 * V : V) + (LC ? lane * 2 : 0);
#pragma unroll
            for (int k = 1; k < EW_RANK; ++k) {
                const long long d = P.dim[k];
                const long long qn = o / d;
                idx[k] = o - qn * d;
                o = qn;
            }
        } else {
            const unsigned w32 = (unsigned)w;
            unsigned o = fastdiv(w32, P.div_chunks0);
            idx[0] = (long long)(w32 - o * P.div_chunks0.d) * (LC ? 32 * V : V) + (LC ? lane * 2 : 0);
#pragma unroll
            for (int k = 1; k < EW_RANK; ++k) {
                if (k < P.nd) {
                    const unsigned qn = fastdiv(o, P.div_dim[k]);
                    idx[k] = o - qn * P.div_dim[k].d;
                    o = qn;
                } else {
                    idx[k] = 0;
                }
            }
        }
        const long long pos0 = idx[0];

        Vec<V> acc;
        Vec<V> t[NT > 0 ? NT : 1];
#pragma unroll
        for (int e = 0; e < V; ++e) acc.v[e] = 0.0;

#pragma unroll 1
        for (int pc = 0; pc < P.n_instr; ++pc) {  // interpreter loop (kept rolled)
            const int op = P.code[pc].op;
            const int arg = P.code[pc].arg;
            const double imm = P.code[pc].imm;
            Vec<V> x;  // second operand for binary ops
            const int src = op & 3;  // 0 = none/unary, 1 = input tensor, 2 = immediate, 3 = temp
            if (src == 1) {
                long long base = 0;
#pragma unroll
                for (int k = 0; k < EW_RANK; ++k)
                    if (k < P.nd) base += idx[k] * P.in_stride[arg][k];
                x = load_in<V, PS>(P, arg, base, pos0);
            } else if (src == 2) {
#pragma unroll
                EW_FOR x.v[e] = imm;
            } else if (src == 3) {
                // temps live in registers: selected by a uniform branch, never by dynamic indexing
                switch (arg) {
                    case 0: if (NT > 0) x = t[0]; break;
                    case 1: if (NT > 1) x = t[1 % (NT > 0 ? NT : 1)]; break;
                    case 2: if (NT > 2) x = t[2 % (NT > 0 ? NT : 1)]; break;
                    case 3: if (NT > 3) x = t[3 % (NT > 0 ? NT : 1)]; break;
                    case 4: if (NT > 4) x = t[4 % (NT > 0 ? NT : 1)]; break;
                    case 5: if (NT > 5) x = t[5 % (NT > 0 ? NT : 1)]; break;
                    case 6: if (NT > 6) x = t[6 % (NT > 0 ? NT : 1)]; break;
                    default: if (NT > 7) x = t[7 % (NT > 0 ? NT : 1)]; break;
                }
            }
            switch (op >> 2) {
                case ADB_EW_LOAD >> 2:
#pragma unroll
                    EW_FOR acc.v[e] = x.v[e];
                    break;
                case ADB_EW_ADD >> 2:
#pragma unroll
                    EW_FOR acc.v[e] = acc.v[e] + x.v[e];
                    break;
                case ADB_EW_MUL >> 2:
#pragma unroll
                    EW_FOR acc.v[e] = acc.v[e] * x.v[e];
                    break;
                case ADB_EW_DIV >> 2:
#pragma unroll
                    EW_FOR acc.v[e] = acc.v[e] / x.v[e];
                    break;
                case ADB_EW_POW >> 2:
#pragma unroll
                    EW_FOR acc.v[e] = ew_pow(acc.v[e], x.v[e]);
                    break;
                case ADB_EW_RPOW >> 2:
#pragma unroll
                    EW_FOR acc.v[e] = ew_pow(x.v[e], acc.v[e]);
                    break;
                case ADB_EW_MAX >> 2:
#pragma unroll
                    EW_FOR acc.v[e] = fmax(acc.v[e], x.v[e]);
                    break;
                case ADB_EW_STORE_T >> 2:
                    switch (arg) {
                        case 0: if (NT > 0) t[0] = acc; break;
                        case 1: if (NT > 1) t[1 % (NT > 0 ? NT : 1)] = acc; break;
                        case 2: if (NT > 2) t[2 % (NT > 0 ? NT : 1)] = acc; break;
                        case 3: if (NT > 3) t[3 % (NT > 0 ? NT : 1)] = acc; break;
                        case 4: if (NT > 4) t[4 % (NT > 0 ? NT : 1)] = acc; break;
                        case 5: if (NT > 5) t[5 % (NT > 0 ? NT : 1)] = acc; break;
                        case 6: if (NT > 6) t[6 % (NT > 0 ? NT : 1)] = acc; break;
                        default: if (NT > 7) t[7 % (NT > 0 ? NT : 1)] = acc; break;
                    }
                    break;
                case ADB_EW_STORE_OUT >> 2: {
                    long long base = 0;
#pragma unroll
                    for (int k = 0; k < EW_RANK; ++k)
                        if (k < P.nd) base += idx[k] * P.out_stride[arg][k];
                    store_out<V, PS>(P, arg, base, pos0, acc);
                    break;
                }
                case ADB_EW_NEG >> 2:
#pragma unroll
                    EW_FOR acc.v[e] = -acc.v[e];
                    break;
                case ADB_EW_RECIP_EPS >> 2:  // reference ops.py:116: 1 / (x + eps)
#pragma unroll
                    EW_FOR acc.v[e] = 1.0 / (acc.v[e] + imm);
                    break;
                case ADB_EW_LOG_EPS >> 2:  // reference ops.py:309: log(x + eps)
#pragma unroll
                    EW_FOR acc.v[e] = log(acc.v[e] + imm);
                    break;
                case ADB_EW_EXP >> 2:
#pragma unroll
                    EW_FOR acc.v[e] = exp(acc.v[e]);
                    break;
                case ADB_EW_RELU >> 2:  // reference ops.py:228: val * (val > 0)
#pragma unroll
                    EW_FOR acc.v[e] = acc.v[e] * (acc.v[e] > 0.0 ? 1.0 : 0.0);
                    break;
                case ADB_EW_SIGMOID >> 2:  // reference ops.py:354: 1 / (1 + exp(-x))
#pragma unroll
                    EW_FOR acc.v[e] = 1.0 / (1.0 + exp(-acc.v[e]));
                    break;
                case ADB_EW_SQRT >> 2:
#pragma unroll
                    EW_FOR acc.v[e] = sqrt(acc.v[e]);
                    break;
                case ADB_EW_ABS >> 2:
#pragma unroll
                    EW_FOR acc.v[e] = fabs(acc.v[e]);
                    break;
                case ADB_EW_SQUARE >> 2:
#pragma unroll
                    EW_FOR acc.v[e] = acc.v[e] * acc.v[e];
                    break;
                case ADB_EW_LOG >> 2:
#pragma unroll
                    EW_FOR acc.v[e] = log(acc.v[e]);
                    break;
                case ADB_EW_RECIP >> 2:
#pragma unroll
                    EW_FOR acc.v[e] = 1.0 / acc.v[e];
                    break;
                default:
                    break;
            }
        }
    }
}

static int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// Collapse the iteration space: drop size-1 dims, merge adjacent dims that are jointly linear in every tensor.
int ewise_run(const adb_ew_instr* code, int n_instr, int n_in, const adb_tensor* ins, int n_out, const adb_tensor* outs,
              cudaStream_t st) {
    if (n_in < 0 || n_in > ADB_EW_MAX_IN) return set_error("ewise: %d inputs (max %d)", n_in, ADB_EW_MAX_IN);
    if (n_out < 1 || n_out > ADB_EW_MAX_OUT) return set_error("ewise: %d outputs (max %d)", n_out, ADB_EW_MAX_OUT);
    if (n_instr < 1 || n_instr > ADB_EW_MAX_INSTR) return set_error("ewise: %d instructions (max %d)", n_instr, ADB_EW_MAX_INSTR);
    const int rank = outs[0].rank;
    if (rank < 0 || rank > ADB_MAX_RANK) return set_error("ewise: bad rank %d", rank);
    for (int j = 1; j < n_out; ++j) {
        if (outs[j].rank != rank) return set_error("ewise: outputs must share one shape");
        for (int k = 0; k < rank; ++k)
            if (outs[j].shape[k] != outs[0].shape[k]) return set_error("ewise: outputs must share one shape");
    }
    // validate opcodes / operand indices on the host so the kernel never indexes out of range
    for (int i = 0; i < n_instr; ++i) {
        const int op = code[i].op, src = op & 3, arg = code[i].arg;
        if (op < 0 || (op >> 2) > (ADB_EW_RECIP >> 2)) return set_error("ewise: bad opcode %d at %d", op, i);
        if (src == 1 && (arg < 0 || arg >= n_in)) return set_error("ewise: instr %d reads input %d of %d", i, arg, n_in);
        if ((src == 3 || (op >> 2) == (ADB_EW_STORE_T >> 2)) && (arg < 0 || arg >= EW_TEMPS)) return set_error("ewise: instr %d uses temp %d", i, arg);
        if ((op >> 2) == (ADB_EW_STORE_OUT >> 2) && (arg < 0 || arg >= n_out)) return set_error("ewise: instr %d stores output %d of %d", i, arg, n_out);
    }
    // inputs broadcast against the output shape, numpy style (right-aligned)
    long long shape[ADB_MAX_RANK];
    long long istr[ADB_EW_MAX_IN][ADB_MAX_RANK];
    long long ostr[ADB_EW_MAX_OUT][ADB_MAX_RANK];
    long long total = 1;
    for (int k = 0; k < rank; ++k) { shape[k] = outs[0].shape[k]; total *= shape[k]; }
    if (total == 0) return 0;
    for (int i = 0; i < n_in; ++i) {
        const int r = ins[i].rank;
        if (r > rank) return set_error("ewise: input %d has rank %d > output rank %d", i, r, rank);
        for (int k = 0; k < rank; ++k) {
            const int ik = k - (rank - r);
            if (ik < 0) { istr[i][k] = 0; continue; }
            const long long d = ins[i].shape[ik];
            if (d == shape[k]) istr[i][k] = (d == 1) ? 0 : ins[i].stride[ik];
            else if (d == 1) istr[i][k] = 0;
            else return set_error("ewise: input %d dim %d (%lld) does not broadcast to %lld", i, ik, d, shape[k]);
        }
    }
    for (int j = 0; j < n_out; ++j)
        for (int k = 0; k < rank; ++k) ostr[j][k] = shape[k] == 1 ? 0 : outs[j].stride[k];

    // collapse, walking from the innermost dim outwards
    EwParams P;
    P.n_in = n_in; P.n_out = n_out; P.n_instr = n_instr; P.nd = 1;
    int nd = 0;
    for (int k = rank - 1; k >= 0; --k) {
        if (shape[k] == 1) continue;
        bool merged = false;
        if (nd > 0) {
            const int c = nd - 1;
            bool ok = true;
            for (int i = 0; i < n_in && ok; ++i) ok = istr[i][k] == P.in_stride[i][c] * P.dim[c];
            for (int j = 0; j < n_out && ok; ++j) ok = ostr[j][k] == P.out_stride[j][c] * P.dim[c];
            if (ok) { P.dim[c] *= shape[k]; merged = true; }
        }
        if (!merged) {
            if (nd == EW_RANK) return set_error("ewise: iteration space does not collapse to %d dims", EW_RANK);
            P.dim[nd] = shape[k];
            for (int i = 0; i < n_in; ++i) P.in_stride[i][nd] = istr[i][k];
            for (int j = 0; j < n_out; ++j) P.out_stride[j][nd] = ostr[j][k];
            ++nd;
        }
    }
    for (int k = nd; k < EW_RANK; ++k) {
        P.dim[k] = 1;
        for (int i = 0; i < n_in; ++i) P.in_stride[i][k] = 0;
        for (int j = 0; j < n_out; ++j) P.out_stride[j][k] = 0;
    }
    P.in_vec = 0; P.out_vec = 0;
    for (int i = 0; i < n_in; ++i) {
        P.in[i] = ins[i].ptr;
        bool v = P.in_stride[i][0] == 1 && (reinterpret_cast<uintptr_t>(ins[i].ptr) & 15) == 0;
        for (int k = 1; k < EW_RANK && v; ++k) v = (P.in_stride[i][k] & 1) == 0;
        if (v) P.in_vec |= 1u << i;
    }
    for (int j = 0; j < n_out; ++j) {
        P.out[j] = outs[j].ptr;
        bool v = P.out_stride[j][0] == 1 && (reinterpret_cast<uintptr_t>(outs[j].ptr) & 15) == 0;
        for (int k = 1; k < EW_RANK && v; ++k) v = (P.out_stride[j][k] & 1) == 0;
        if (v) P.out_vec |= 1u << j;
    }
    for (int i = 0; i < n_instr; ++i) P.code[i] = code[i];

    int max_temp = -1;
    for (int i = 0; i < n_instr; ++i)
        if ((code[i].op & 3) == 3 || (code[i].op >> 2) == (ADB_EW_STORE_T >> 2)) max_temp = code[i].arg > max_temp ? code[i].arg : max_temp;
    const int V = max_temp < 2 ? 8 : 4;
    const bool lc = P.dim[0] >= 32 * V;                       // long rows: warp-wide coalesced tiles
    const long long per_item = lc ? 32 * V : V;
    const long long chunks0 = (P.dim[0] + per_item - 1) / per_item;
    long long work = chunks0;
    for (int k = 1; k < EW_RANK; ++k) work *= P.dim[k];
    P.nd = nd < 1 ? 1 : nd;
    P.total = (unsigned long long)work;
    bool big = work >= (1LL << 31);
    for (int k = 0; k < EW_RANK; ++k) big = big || P.dim[k] >= (1LL << 31);
    P.div_chunks0 = make_fastdiv((unsigned)(chunks0 < (1LL << 32) ? chunks0 : 0));
    P.div_chunks0.d = (unsigned)chunks0;
    if (chunks0 >= (1LL << 32)) return set_error("ewise: innermost extent too large");
    if (!big)
        for (int k = 0; k < EW_RANK; ++k) P.div_dim[k] = make_fastdiv((unsigned)P.dim[k]);
    const long long threads_needed = lc ? work * 32 : work;
    long long blocks = (threads_needed + 255) / 256;
    const long long cap = (long long)sm_count() * 16;   // resident CTAs, grid-stride beyond
    if (blocks > cap) blocks = cap;
    const unsigned g = (unsigned)blocks;
#define ADB_EW_LAUNCH(VV, NTT)                                                               \
    do {                                                                                     \
        if (big) { if (lc) ewise_kernel<VV, NTT, true, true><<<g, 256, 0, st>>>(P); else ewise_kernel<VV, NTT, true, false><<<g, 256, 0, st>>>(P); } \
        else { if (lc) ewise_kernel<VV, NTT, false, true><<<g, 256, 0, st>>>(P); else ewise_kernel<VV, NTT, false, false><<<g, 256, 0, st>>>(P); } \
    } while (0)
    if (max_temp < 0) ADB_EW_LAUNCH(8, 0);
    else if (max_temp < 2) ADB_EW_LAUNCH(8, 2);
    else ADB_EW_LAUNCH(4, EW_TEMPS);
#undef ADB_EW_LAUNCH
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("ewise launch failed: %s", cudaGetErrorString(e));
    count_launch();
    return 0;
}

}  // namespace adb

extern "C" int adb_ewise(const adb_ew_instr* code, int n_instr, int n_in, const adb_tensor* ins, int n_out,
                         const adb_tensor* outs, void* stream) {
    return adb::ewise_run(code, n_instr, n_in, ins, n_out, outs, reinterpret_cast<cudaStream_t>(stream));
}
