// Device side of the fused elementwise programs (see ewise.cu for the host side and the reference call sites).
//
// One thread = one QUAD (4 consecutive fp64 elements of the innermost dimension, one 256-bit access per operand),
// one-shot grid.  Measured on B200 (tools/probe_stream.cu, profiles/r01_probe_stream.log): one-shot grids stream at
// 6.6-7.0 TB/s where a persistent grid-stride loop tops out near 5.7 TB/s once stores are involved, and bandwidth
// follows the bytes of loads in flight per SM (24 KB -> 5.2, 32 KB -> 5.9, 64 KB -> 6.7+ TB/s) - so the kernels
// below fetch every input operand of the quad first, back to back, and keep the per-thread register and
// instruction footprint small so that many CTAs are resident.
//
// The same operation table (ew_apply) is executed two ways:
//   * EwStatic<Prog>: the program is a TYPE (ewise_catalog.inc, generated from the programs the BASELINE graphs
//     actually run), so the kernel is straight-line code with operands in named registers;
//   * EwInterp<NIN, NT>: a uniform-branch bytecode interpreter for any other program.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <utility>

#include "adb_internal.h"

namespace adb {

constexpr int EW_RANK = 6;
constexpr int EW_TEMPS = 8;
constexpr int EW_V = 4;   // elements per thread

// q = n / d for n < 2^31 without a divide: d == 1 ? n : umulhi(n, mul) >> shr   (round-up magic number)
struct FastDiv {
    unsigned mul, shr, d, _pad;
};
inline FastDiv make_fastdiv(unsigned d) {
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

enum { EW_BCAST = 0, EW_V256 = 1, EW_V128 = 2, EW_SCALAR = 3 };

struct EwParams {
    int n_in, n_out, n_instr, nd;
    long long dim[EW_RANK];    // host-side collapsed extents, dim[0] innermost (the kernel reads dim0 / div_*)
    unsigned dim0;             // innermost extent (elements)
    unsigned total;            // quads in this launch
    FastDiv div_chunks0;       // ceil(dim0 / 4)
    FastDiv div_dim[EW_RANK];  // extents of dims 1..
    long long in_stride[ADB_EW_MAX_IN][EW_RANK];
    long long out_stride[ADB_EW_MAX_OUT][EW_RANK];
    const double* in[ADB_EW_MAX_IN];
    double* out[ADB_EW_MAX_OUT];
    unsigned in_mode, out_mode;  // 2 bits per tensor: how a full quad is accessed (EW_BCAST / EW_V256 / EW_V128 / EW_SCALAR)
    unsigned off32;              // every element offset of every tensor fits 31 bits and no stride is negative: 32-bit index math
    adb_ew_instr code[ADB_EW_MAX_INSTR];
};

struct Vec {
    double v[EW_V];
};
#define EW_FOR _Pragma("unroll") for (int e = 0; e < EW_V; ++e)

__device__ __forceinline__ Vec ldg256(const double* p) {   // sm_100: 256-bit global load, one full sector per lane
    Vec r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3]) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg256(double* p, const Vec& a) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p), "d"(a.v[0]), "d"(a.v[1]), "d"(a.v[2]), "d"(a.v[3]) : "memory");
}

template <int ND>
__device__ __forceinline__ long long ew_offset(const EwParams& P, const long long* stride, const unsigned (&idx)[ND]) {
    if (ND > 1 && P.off32) {     // (uniform branch) one 32-bit IMAD per dim instead of a 64-bit multiply-add chain
        unsigned off = 0;
#pragma unroll
        for (int k = 0; k < ND; ++k) off += idx[k] * (unsigned)stride[k];
        return (long long)off;
    }
    long long base = 0;
#pragma unroll
    for (int k = 0; k < ND; ++k) base += (long long)idx[k] * stride[k];
    return base;
}

template <int ND>
__device__ __forceinline__ Vec load_in(const EwParams& P, int i, const unsigned (&idx)[ND], unsigned nvalid) {
    Vec r;
    const double* p = P.in[i] + ew_offset<ND>(P, P.in_stride[i], idx);
    const unsigned mode = (P.in_mode >> (2 * i)) & 3u;
    if (mode == EW_BCAST) {
        const double x = __ldg(p);
        EW_FOR r.v[e] = x;
    } else if (nvalid == EW_V && mode == EW_V256) {
        r = ldg256(p);
    } else if (nvalid == EW_V && mode == EW_V128) {
        const double2 a = __ldg(reinterpret_cast<const double2*>(p));
        const double2 b = __ldg(reinterpret_cast<const double2*>(p) + 1);
        r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y;
    } else {
        const long long s0 = P.in_stride[i][0];
        EW_FOR r.v[e] = e < (int)nvalid ? __ldg(p + e * s0) : 0.0;
    }
    return r;
}

template <int ND>
__device__ __forceinline__ void store_out(const EwParams& P, int j, const unsigned (&idx)[ND], unsigned nvalid, const Vec& a) {
    double* p = P.out[j] + ew_offset<ND>(P, P.out_stride[j], idx);
    const unsigned mode = (P.out_mode >> (2 * j)) & 3u;
    if (nvalid == EW_V && mode == EW_V256) {
        stg256(p, a);
    } else if (nvalid == EW_V && mode == EW_V128) {
        reinterpret_cast<double2*>(p)[0] = make_double2(a.v[0], a.v[1]);
        reinterpret_cast<double2*>(p)[1] = make_double2(a.v[2], a.v[3]);
    } else {
        const long long s0 = P.out_stride[j][0];
        EW_FOR if (e < (int)nvalid) p[e * s0] = a.v[e];
    }
}

// Out-of-line transcendental bodies for the interpreter: inlined, their temporaries would be live four times (once per
// element) in the interpreter loop and push every program - plain copies included - to 70+ registers.
static __device__ __noinline__ double ew_pow(double a, double b) { return pow(a, b); }
static __device__ __noinline__ double ew_exp_ool(double a) { return exp(a); }
static __device__ __noinline__ double ew_log_ool(double a) { return log(a); }
template <bool INL> __device__ __forceinline__ double ew_exp(double a) { if (INL) return exp(a); else return ew_exp_ool(a); }
template <bool INL> __device__ __forceinline__ double ew_log(double a) { if (INL) return log(a); else return ew_log_ool(a); }

// THE operation table.  KIND = opcode >> 2 (include/adb200.h); every body replays the reference's fp64 expression.
template <int KIND, bool INL>
__device__ __forceinline__ void ew_apply(Vec& acc, const Vec& x, double imm) {
    if constexpr (KIND == (ADB_EW_LOAD >> 2)) { EW_FOR acc.v[e] = x.v[e]; }
    else if constexpr (KIND == (ADB_EW_ADD >> 2)) { EW_FOR acc.v[e] = acc.v[e] + x.v[e]; }
    else if constexpr (KIND == (ADB_EW_MUL >> 2)) { EW_FOR acc.v[e] = acc.v[e] * x.v[e]; }
    else if constexpr (KIND == (ADB_EW_DIV >> 2)) { EW_FOR acc.v[e] = acc.v[e] / x.v[e]; }
    else if constexpr (KIND == (ADB_EW_POW >> 2)) { EW_FOR acc.v[e] = ew_pow(acc.v[e], x.v[e]); }
    else if constexpr (KIND == (ADB_EW_RPOW >> 2)) { EW_FOR acc.v[e] = ew_pow(x.v[e], acc.v[e]); }
    else if constexpr (KIND == (ADB_EW_MAX >> 2)) { EW_FOR acc.v[e] = fmax(acc.v[e], x.v[e]); }
    else if constexpr (KIND == (ADB_EW_NEG >> 2)) { EW_FOR acc.v[e] = -acc.v[e]; }
    else if constexpr (KIND == (ADB_EW_RECIP_EPS >> 2)) { EW_FOR acc.v[e] = 1.0 / (acc.v[e] + imm); }          // ops.py:116
    else if constexpr (KIND == (ADB_EW_LOG_EPS >> 2)) { EW_FOR acc.v[e] = ew_log<INL>(acc.v[e] + imm); }        // ops.py:309
    else if constexpr (KIND == (ADB_EW_EXP >> 2)) { EW_FOR acc.v[e] = ew_exp<INL>(acc.v[e]); }
    else if constexpr (KIND == (ADB_EW_RELU >> 2)) { EW_FOR acc.v[e] = acc.v[e] * (acc.v[e] > 0.0 ? 1.0 : 0.0); }  // ops.py:228
    else if constexpr (KIND == (ADB_EW_SIGMOID >> 2)) { EW_FOR acc.v[e] = 1.0 / (1.0 + ew_exp<INL>(-acc.v[e])); }   // ops.py:354
    else if constexpr (KIND == (ADB_EW_SQRT >> 2)) { EW_FOR acc.v[e] = sqrt(acc.v[e]); }
    else if constexpr (KIND == (ADB_EW_ABS >> 2)) { EW_FOR acc.v[e] = fabs(acc.v[e]); }
    else if constexpr (KIND == (ADB_EW_SQUARE >> 2)) { EW_FOR acc.v[e] = acc.v[e] * acc.v[e]; }
    else if constexpr (KIND == (ADB_EW_LOG >> 2)) { EW_FOR acc.v[e] = ew_log<INL>(acc.v[e]); }
    else if constexpr (KIND == (ADB_EW_RECIP >> 2)) { EW_FOR acc.v[e] = 1.0 / acc.v[e]; }
}

// ------------------------------------------------------------------------------------------------ interpreter
// uniform (warp-wide) register select: never dynamic indexing, so the arrays stay in registers
template <int N>
__device__ __forceinline__ Vec reg_select(const Vec (&a)[N], int k) {
    Vec x = a[0];
#pragma unroll
    for (int i = 1; i < N; ++i)
        if (k == i) x = a[i];
    return x;
}
template <int N>
__device__ __forceinline__ void reg_assign(Vec (&a)[N], int k, const Vec& x) {
#pragma unroll
    for (int i = 0; i < N; ++i)
        if (k == i) a[i] = x;
}

template <int NIN_, int NT_>
struct EwInterp {
    static constexpr int NIN = NIN_;
    static constexpr bool STATIC = false;
    template <int ND>
    __device__ __forceinline__ static void run(const EwParams& P, const Vec (&in)[NIN], const unsigned (&idx)[ND], unsigned nvalid) {
        constexpr int NT = NT_ > 0 ? NT_ : 1;
        Vec acc;
        Vec t[NT];
        EW_FOR acc.v[e] = 0.0;
#pragma unroll 1
        for (int pc = 0; pc < P.n_instr; ++pc) {
            const int op = P.code[pc].op;
            const int arg = P.code[pc].arg;
            const double imm = P.code[pc].imm;
            Vec x;
            const int src = op & 3;  // 0 = none/unary, 1 = input tensor, 2 = immediate, 3 = temp
            if (src == 1) {
                x = reg_select<NIN>(in, arg);
            } else if (src == 2) {
                EW_FOR x.v[e] = imm;
            } else if (src == 3) {
                if (NT_ > 0) x = reg_select<NT>(t, arg);
            }
            switch (op >> 2) {
                case ADB_EW_STORE_T >> 2:
                    if (NT_ > 0) reg_assign<NT>(t, arg, acc);
                    break;
                case ADB_EW_STORE_OUT >> 2:
                    store_out<ND>(P, arg, idx, nvalid, acc);
                    break;
#define EW_CASE(K) case (K) >> 2: ew_apply<((K) >> 2), false>(acc, x, imm); break;
                EW_CASE(ADB_EW_LOAD) EW_CASE(ADB_EW_ADD) EW_CASE(ADB_EW_MUL) EW_CASE(ADB_EW_DIV) EW_CASE(ADB_EW_POW)
                EW_CASE(ADB_EW_RPOW) EW_CASE(ADB_EW_MAX) EW_CASE(ADB_EW_NEG) EW_CASE(ADB_EW_RECIP_EPS) EW_CASE(ADB_EW_LOG_EPS)
                EW_CASE(ADB_EW_EXP) EW_CASE(ADB_EW_RELU) EW_CASE(ADB_EW_SIGMOID) EW_CASE(ADB_EW_SQRT) EW_CASE(ADB_EW_ABS)
                EW_CASE(ADB_EW_SQUARE) EW_CASE(ADB_EW_LOG) EW_CASE(ADB_EW_RECIP)
#undef EW_CASE
                default:
                    break;
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------ static programs
template <int OP, int ARG>
struct EwI {};
template <class... Is>
struct EwProg {};

constexpr int ew_cmax(int a, int b) { return a > b ? a : b; }
template <class... T>
constexpr int ew_cmax(int a, int b, T... rest) { return ew_cmax(ew_cmax(a, b), rest...); }

template <class Prog>
struct EwStatic;
template <int... OP, int... ARG>
struct EwStatic<EwProg<EwI<OP, ARG>...>> {
    static constexpr int N = sizeof...(OP);
    static constexpr int NIN = ew_cmax(1, 1, (((OP & 3) == 1) ? ARG + 1 : 0)...);
    static constexpr int NT = ew_cmax(1, 1, ((((OP & 3) == 3) || ((OP >> 2) == (ADB_EW_STORE_T >> 2))) ? ARG + 1 : 0)...);
    static constexpr bool STATIC = true;

    template <int ND, int OPK, int ARGK, int PC>
    __device__ __forceinline__ static void step(const EwParams& P, const Vec (&in)[NIN], Vec& acc, Vec (&t)[NT], const unsigned (&idx)[ND],
                                                unsigned nvalid) {
        constexpr int src = OPK & 3, kind = OPK >> 2;
        Vec x;
        if constexpr (src == 1) x = in[ARGK];
        else if constexpr (src == 2) { const double imm = P.code[PC].imm; EW_FOR x.v[e] = imm; }
        else if constexpr (src == 3) x = t[ARGK];
        else { EW_FOR x.v[e] = 0.0; }
        if constexpr (kind == (ADB_EW_STORE_T >> 2)) t[ARGK] = acc;
        else if constexpr (kind == (ADB_EW_STORE_OUT >> 2)) store_out<ND>(P, ARGK, idx, nvalid, acc);
        else ew_apply<kind, true>(acc, x, P.code[PC].imm);
    }
    template <int ND, size_t... PC>
    __device__ __forceinline__ static void run_seq(std::index_sequence<PC...>, const EwParams& P, const Vec (&in)[NIN], const unsigned (&idx)[ND],
                                                   unsigned nvalid) {
        Vec acc;
        Vec t[NT];
        EW_FOR acc.v[e] = 0.0;
        (step<ND, OP, ARG, (int)PC>(P, in, acc, t, idx, nvalid), ...);
    }
    template <int ND>
    __device__ __forceinline__ static void run(const EwParams& P, const Vec (&in)[NIN], const unsigned (&idx)[ND], unsigned nvalid) {
        run_seq<ND>(std::make_index_sequence<N>{}, P, in, idx, nvalid);
    }
};

// ------------------------------------------------------------------------------------------------ the kernel
template <int ND, class Exec>
__global__ void __launch_bounds__(256) ewise_kernel(const __grid_constant__ EwParams P) {
    const unsigned w = blockIdx.x * 256u + threadIdx.x;
    if (w >= P.total) return;
    unsigned idx[ND];
    if constexpr (ND == 1) {
        idx[0] = w * (unsigned)EW_V;
    } else {
        unsigned o = fastdiv(w, P.div_chunks0);
        idx[0] = (w - o * P.div_chunks0.d) * (unsigned)EW_V;
#pragma unroll
        for (int k = 1; k < ND - 1; ++k) {
            const unsigned qn = fastdiv(o, P.div_dim[k]);
            idx[k] = o - qn * P.div_dim[k].d;
            o = qn;
        }
        idx[ND - 1] = o;
    }
    const unsigned left = P.dim0 - idx[0];
    const unsigned nvalid = left < (unsigned)EW_V ? left : (unsigned)EW_V;

    Vec in[Exec::NIN];
#pragma unroll
    for (int i = 0; i < Exec::NIN; ++i) {
        if (Exec::STATIC || i < P.n_in) {
            in[i] = load_in<ND>(P, i, idx, nvalid);
        } else {
            EW_FOR in[i].v[e] = 0.0;
        }
    }
    Exec::template run<ND>(P, in, idx, nvalid);
}

typedef void (*EwLaunchFn)(const EwParams& P, unsigned grid, cudaStream_t st);
template <int ND, class Exec>
void ew_launch(const EwParams& P, unsigned grid, cudaStream_t st) {
    ewise_kernel<ND, Exec><<<grid, 256, 0, st>>>(P);
}

// A statically compiled program: its (op, arg) signature and one launcher per collapsed rank class (1, 2, any).
struct EwCatalogEntry {
    int n_instr;
    const int* sig;   // 2 * n_instr ints: op, arg
    EwLaunchFn fn[3];
};
const EwCatalogEntry* ew_catalog(int* count);   // ewise_catalog.cu

}  // namespace adb
