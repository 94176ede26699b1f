// microclimf_b200 — branch-free FP64 elementary functions for the hot loops.
//
// The grid solver is bound by the FP64 pipe and by instruction issue, and the CUDA math
// library's double-precision exp / log / pow / division carry special-case branches, slow-path calls
// and 64-bit immediates that more than double the instruction count and spill the loop bodies out of
// the 32 KB L1.5 instruction cache.  These replacements are straight-line code:
//
//   mrcp, mdiv : MUFU.RCP64H seed (rel. err 2^-23) + one cubic Newton step         (<= 2 ulp)
//   msqrt      : MUFU.RSQ64H seed + two coupled Newton steps                       (<= 1 ulp)
//   mexp       : Cody-Waite reduction by ln2 (hi/lo), degree-11 polynomial         (<= 1 ulp)
//   mlog       : exponent/mantissa split, log(1+f) = f - s(f - zP(z)), s = f/(2+f)  (<= 1 ulp)
//   mpow       : exp(y log x)                                                        (~1e-14 relative)
//
// Coefficients come from tools/gen_math_coeffs.py (mpmath Chebyshev fits, verified there against
// mpmath).  The parity bar is 1e-6 (tests/); these are accurate to ~1e-15.
//
// Domain contract (every call site in mcf_physics.cuh is annotated):
//   * mrcp / mdiv: divisor finite, normal, non-zero.  A zero or infinite divisor yields NaN (not
//     +-inf / 0), so it is only used where the reference's own result for that case is discarded or NaN.
//   * mexp: any x; x < -708 returns exp(-708) ~ 3e-308 (instead of a denormal or 0), x > 709 returns
//     exp(709); NaN propagates.
//   * mlog: x > 0 finite normal; mlog(0) returns ~ -709.8 (callers that need -inf use log()).
//   * msqrt: x >= 0 finite normal or exactly 0.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mcf {

__device__ __forceinline__ double rcp_seed(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}
__device__ __forceinline__ double rsqrt_seed(double x) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}

// 1 / x
__device__ __forceinline__ double mrcp(double x) {
    double r = rcp_seed(x);
    double e = fma(-x, r, 1.0);
    double t = fma(e, e, e);
    return fma(r, t, r);
}
// a / b
__device__ __forceinline__ double mdiv(double a, double b) { return a * mrcp(b); }

// sqrt(x), x >= 0
__device__ __forceinline__ double msqrt(double x) {
    double y = rsqrt_seed(x);       // ~2^-23
    double g = x * y;               // ~sqrt(x)
    double h = 0.5 * y;
    double r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    double d = fma(-g, g, x);
    g = fma(d, h, g);
    return (x == 0.0) ? 0.0 : g;
}

constexpr double kLn2Hi = 0.6931471675634384;      // 27 trailing zero bits: k * kLn2Hi is exact for |k| < 2^26
constexpr double kLn2Lo = 1.2996506893889889e-08;
constexpr double kLog2e = 1.4426950408889634;

// exp(r) for |r| <= ln2/2 (Estrin evaluation: 4 dependent levels instead of 11)
__device__ __forceinline__ double exp_poly(double r) {
    const double r2 = r * r;
    const double r4 = r2 * r2;
    const double a01 = fma(1.0, r, 1.0);
    const double a23 = fma(0.1666666666666668, r, 0.5000000000000019);
    const double a45 = fma(0.008333333333319601, r, 0.0416666666664881);
    const double a67 = fma(0.00019841269890047113, r, 0.0013888888952314775);
    const double a89 = fma(2.755724091857897e-06, r, 2.4801485482328494e-05);
    const double aab = fma(2.5110037605963777e-08, r, 2.763263963904103e-07);
    const double b0 = fma(a23, r2, a01);
    const double b1 = fma(a67, r2, a45);
    const double b2 = fma(aab, r2, a89);
    const double r8 = r4 * r4;
    const double c0 = fma(b1, r4, b0);
    return fma(b2, r8, c0);
}

// multiply p (0.5 < p < 2) by 2^k, |k| <= 1021, through the exponent field
__device__ __forceinline__ double scale2(double p, int k) {
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

__device__ __forceinline__ double mexp(double x) {
    x = (x < -708.0) ? -708.0 : x;
    x = (x > 709.0) ? 709.0 : x;
    const double t = fma(x, kLog2e, 6755399441055744.0); // 1.5 * 2^52: round-to-nearest integer in the low word
    const int k = __double2loint(t);
    const double kf = t - 6755399441055744.0;
    double r = fma(kf, -kLn2Hi, x);
    r = fma(kf, -kLn2Lo, r);
    return scale2(exp_poly(r), k);
}

// 2^x
__device__ __forceinline__ double mexp2(double x) {
    x = (x < -1021.0) ? -1021.0 : x;
    x = (x > 1023.0) ? 1023.0 : x;
    const double t = x + 6755399441055744.0;
    const int k = __double2loint(t);
    const double kf = t - 6755399441055744.0;
    const double r = (x - kf) * 0.6931471805599453;
    return scale2(exp_poly(r), k);
}

__device__ __forceinline__ double mlog(double x) {
    int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    int e = (hi >> 20) - 1023;
    hi = (hi & 0x000FFFFF) | 0x3FF00000;               // mantissa in [1, 2)
    const bool big = hi >= 0x3FF6A09F;                   // >= ~sqrt(2): use m/2 in [sqrt(2)/2, 1)
    hi = big ? hi - 0x00100000 : hi;
    e = big ? e + 1 : e;
    const double m = __hiloint2double(hi, lo);
    const double f = m - 1.0;
    const double s = f * mrcp(2.0 + f);
    const double z = s * s;
    const double z2 = z * z;
    // P(z) degree 6, Estrin
    const double p01 = fma(0.39999999999886615, z, 0.666666666666667);
    const double p23 = fma(0.2222221019926421, z, 0.28571428631764334);
    const double p45 = fma(0.15329500754204178, z, 0.18182956608063458);
    const double q0 = fma(p23, z2, p01);
    const double q1 = fma(0.14643628601909797, z2, p45);
    const double P = fma(q1, z2 * z2, q0);
    const double R = z * P;
    const double ef = (double)e;
    const double lm = f - s * (f - R);
    return fma(ef, kLn2Hi, fma(ef, kLn2Lo, lm));
}

// x^y for x > 0
__device__ __forceinline__ double mpow(double x, double y) { return mexp(y * mlog(x)); }

} // namespace mcf
