// fp64 math helpers shared by the reduction kernels.
#pragma once
#include <cuda_runtime.h>

namespace adb {

// exp(x) for x <= 0 (the only argument a max-shifted log-sum-exp produces; -inf -> 0, NaN -> NaN), <= 0.81 ulp against
// a 60-digit reference over [-700, 0] (checked on the host with exact FMA emulation).  CUDA's exp() carries a branch for
// results near the denormal range, which keeps the compiler from interleaving the independent evaluations of a row; with
// 3-8 warps per scheduler the FP64 pipe then idles on its own latency (ncu: 2.7 of 6.1 stall cycles per issue were
// fixed-latency dependency waits in the softmax-CE kernels).  This version is straight-line code:
//   n = rint(x * log2 e) by the 1.5 * 2^52 trick, r = x - n * ln2 (Cody-Waite, hi + lo), Taylor degree 13 in r
//   (|r| <= 0.3466: truncation 4e-18), scaled by 2^n built in the exponent field.  Results below exp(-700) ~ 1e-304
//   flush to 0: next to the row maximum's exp(0) = 1 they are 288 orders of magnitude under the 1e-12 parity bar.
__device__ __forceinline__ double exp_nonpos(double x) {
    const double z = fma(x, 1.4426950408889634, 6755399441055744.0);
    const int i = __double2loint(z);
    const double n = z - 6755399441055744.0;
    double r = fma(n, -6.93147180369123816490e-01, x);
    r = fma(n, -1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;
    p = fma(p, r, 2.08767569878681e-09);
    p = fma(p, r, 2.505210838544172e-08);
    p = fma(p, r, 2.755731922398589e-07);
    p = fma(p, r, 2.7557319223985893e-06);
    p = fma(p, r, 2.48015873015873e-05);
    p = fma(p, r, 0.0001984126984126984);
    p = fma(p, r, 0.001388888888888889);
    p = fma(p, r, 0.008333333333333333);
    p = fma(p, r, 0.041666666666666664);
    p = fma(p, r, 0.16666666666666666);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double v = p * __hiloint2double((i + 1023) << 20, 0);
    return x < -700.0 ? 0.0 : v;
}

// The same function with a 32-entry table: exp(x) = 2^q * 2^(j/32) * exp(r), N = rint(x * 32 / ln 2) = 32 q + j, |r| <= ln2 / 64, so a
// degree-6 Taylor polynomial is enough (truncation r^7 / 7! < 5e-18) - 12 FP64 instructions instead of 19.  The log-sum-exp
// kernels for narrow rows are bound by instruction issue, not by HBM (ncu, profiles/r02_ncu_summary.md section 4: 65 % of the issue
// slots, FP64 pipe 52 %, DRAM 61 %).  Table entries are the correctly rounded 2^(j/32); error < 1.8 ulp against a 60-digit
// reference (tests/test_math_cpu.py, exact-FMA emulation) - three orders of magnitude inside the 1e-12 parity bar.
__device__ const double exp_tab32[32] = {
    1.0, 1.0218971486541166, 1.0442737824274138, 1.0671404006768237, 1.0905077326652577, 1.1143867425958924, 1.1387886347566916,
    1.1637248587775775, 1.189207115002721, 1.215247359980469, 1.241857812073484, 1.2690509571917332, 1.2968395546510096,
    1.3252366431597413, 1.3542555469368927, 1.383909881963832, 1.4142135623730951, 1.4451808069770467, 1.4768261459394993,
    1.5091644275934228, 1.5422108254079407, 1.5759808451078865, 1.6104903319492543, 1.645755478153965, 1.681792830507429,
    1.718619298122478, 1.7562521603732995, 1.7947090750031072, 1.8340080864093424, 1.8741676341103, 1.9152065613971474,
    1.9571441241754002};
__device__ __forceinline__ double exp_nonpos_tab(double x) {
    const double z = fma(x, 46.16624130844683, 6755399441055744.0);
    const int i = __double2loint(z);
    const double n = z - 6755399441055744.0;
    double r = fma(n, -0.021660849392446835, x);
    r = fma(n, -5.145609244655338e-14, r);
    double p = 0.001388888888888889;
    p = fma(p, r, 0.008333333333333333);
    p = fma(p, r, 0.041666666666666664);
    p = fma(p, r, 0.16666666666666666);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double t = __ldg(exp_tab32 + (i & 31));
    const double v = (t * p) * __hiloint2double(((i >> 5) + 1023) << 20, 0);
    return x < -700.0 ? 0.0 : v;
}

}  // namespace adb
