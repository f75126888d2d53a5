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

}  // namespace adb
