// microclimf_b200 — branch-free FP64 elementary functions for the hot loops.
//
// The grid solver is bound by the FP64 pipe and by instruction issue, and the CUDA math
// library's double-precision exp / log / pow / division carry special-case branches, slow-path calls
// and 64-bit immediates that more than double the instruction count and spill the loop bodies out of
// the 32 KB L1.5 instruction cache.  These replacements are straight-line code:
//
//   mrcp, mdiv : MUFU.RCP64H seed (~20 bits) + one Newton step                      (<= 1e-12)
//   msqrt      : MUFU.RSQ64H seed + one coupled Newton step + Heron correction     (<= 1 ulp)
//   mexp       : Cody-Waite reduction by ln2 (hi/lo), degree-9 polynomial          (<= 8e-14)
//   mlog       : exponent/mantissa split, log(1+f) = f - s(f - zP(z)), s = f/(2+f)  (<= 1 ulp)
//   mpow       : exp(y log x)                                                        (~1e-14 relative)
//
// Coefficients come from tools/gen_math_coeffs.py (mpmath Chebyshev fits, verified there against
// mpmath).  The parity bar is 1e-6 (tests/); these are accurate to ~1e-12 or better, which leaves a
// margin of >= 1e3 after the largest error amplification in the physics (Penman-Monteith: ~1e3).
//
// Domain contract (every call site in mcf_physics.cuh is annotated):
//   * mrcp / mdiv: divisor finite, normal, non-zero.  A zero or infinite divisor yields NaN (not
//     +-inf / 0), so it is only used where the reference's own result for that case is discarded or NaN.
//   * mexp: any x; x < -708 returns exp(-708) ~ 3e-308 (instead of a denormal or 0), x > 709 returns
//     exp(709); NaN propagates.  mexp_lo drops the upper clamp (x <= 700), mexp_nc both (|x| <= 700).
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
    double r = rcp_seed(x);       // relative error ~2^-20 (the seed looks at the high word only)
    double e = fma(-x, r, 1.0);
    return fma(r, e, r);          // one Newton step: <= 2^-40 ~ 1e-12
}
// a / b
__device__ __forceinline__ double mdiv(double a, double b) { return a * mrcp(b); }

// sqrt(x), x >= 0
__device__ __forceinline__ double msqrt(double x) {
    double y = rsqrt_seed(x);       // ~2^-23
    double g = x * y;               // ~sqrt(x)
    double h = 0.5 * y;
    double r = fma(-h, g, 0.5);
    g = fma(g, r, g);               // ~2^-45
    h = fma(h, r, h);
    double d = fma(-g, g, x);       // Heron correction: <= 1 ulp
    g = fma(d, h, g);
    return (x == 0.0) ? 0.0 : g;
}

// Polynomial coefficients live in constant memory: ptxas then feeds them to DFMA as uniform-register
// operands (LDCU.128 loads two at a time and keeps them across call sites) instead of materialising
// every 64-bit immediate with two moves per use, which was ~20 % of all issued instructions.  For that
// to work each DFMA may carry only ONE constant, hence the even/odd Horner split below (which also gives
// two independent dependency chains).
__constant__ double kMathC[32] = {
    // exp(r) = 1 + r + r^2 q(r), q of degree 7 (Chebyshev fit on |r| <= ln2/2, max rel. error 7.4e-14)
    // [0..3]  even part q6 q4 q2 q0            [4..7] odd part q7 q5 q3 q1
    2.4867870179687727e-05, 0.0013888839110572009, 0.04166666678626573, 0.4999999999995511,
    2.7617564785876086e-06, 0.00019841224599656011, 0.00833333334420298, 0.16666666666662586,
    0, 0,
    // [10] log2(e)   [11] ln2 hi (27 trailing zero bits)   [12] ln2 lo   [13] ln2
    1.4426950408889634, 0.6931471675634384, 1.2996506893889889e-08, 0.6931471805599453,
    // [14..17] log P(z) even part L6 L4 L2 L0      [18..20] odd part L5 L3 L1
    0.14643628601909797, 0.18182956608063458, 0.28571428631764334, 0.666666666666667,
    0.15329500754204178, 0.2222221019926421, 0.39999999999886615,
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

// sin / cos on |r| <= pi/4:  sin r = r + r^3 S(r^2),  cos r = 1 - r^2/2 + r^4 C(r^2), S and C of degree 5
// (Chebyshev fits, max error 1.2e-16); highest coefficient first.
__constant__ double kTrigC[16] = {
    1.59153232122714e-10, -2.5051092507061385e-08, 2.755731591191116e-06, -0.00019841269836387345,
    0.0083333333333307, -0.16666666666666666,
    -1.1380876948169717e-11, 2.087612165887116e-09, -2.755731715246704e-07, 2.4801587298533456e-05,
    -0.0013888888888887241, 0.041666666666666664,
    // [12] 2/pi   [13] pi/2 hi (27 trailing zero bits)   [14] pi/2 mid   [15] pi/2 lo
    0.6366197723675814, 1.5707963109016418, 1.5893254712295857e-08, 6.123233995736766e-17};

constexpr double kMagic = 6755399441055744.0; // 1.5 * 2^52 (zero low word: encodable as a DFMA immediate)

// exp(r) for |r| <= ln2/2:  1 + r + r^2 (E(r^2) + r O(r^2)), degree 9
__device__ __forceinline__ double exp_poly(double r) {
    const double r2 = r * r;
    double e = kMathC[0], o = kMathC[4];
    e = fma(e, r2, kMathC[1]);
    o = fma(o, r2, kMathC[5]);
    e = fma(e, r2, kMathC[2]);
    o = fma(o, r2, kMathC[6]);
    e = fma(e, r2, kMathC[3]);
    o = fma(o, r2, kMathC[7]);
    const double q = fma(o, r, e);
    return fma(q, r2, r + 1.0);
}

// multiply p (0.5 < p < 2) by 2^k, |k| <= 1021, through the exponent field
__device__ __forceinline__ double scale2(double p, int k) {
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// exp(x) for |x| <= 700 (no range clamps: out-of-range or non-finite x gives garbage or NaN, never a trap)
__device__ __forceinline__ double mexp_nc(double x) {
    const double t = fma(x, kMathC[10], kMagic); // round-to-nearest integer lands in the low word
    const int k = __double2loint(t);
    const double kf = t - kMagic;
    // one-constant reduction: the product is exact inside the FMA, so the only error is ln2's own rounding,
    // |k| * 8e-17 in r, i.e. <= 8e-14 relative in the result for |x| <= 700 (inside the 1e-12 budget)
    const double r = fma(kf, -kMathC[13], x);
    return scale2(exp_poly(r), k);
}
// exp(x) for x <= 700: arguments below -708 (including -inf) return exp(-708) ~ 3e-308
__device__ __forceinline__ double mexp_lo(double x) {
    x = (x < -708.0) ? -708.0 : x;
    return mexp_nc(x);
}
// exp(x), any x
__device__ __forceinline__ double mexp(double x) {
    x = (x < -708.0) ? -708.0 : x;
    x = (x > 709.0) ? 709.0 : x;
    return mexp_nc(x);
}

// 2^x for |x| <= 1000 (no clamps)
__device__ __forceinline__ double mexp2_nc(double x) {
    const double t = x + kMagic;
    const int k = __double2loint(t);
    const double kf = t - kMagic;
    const double r = (x - kf) * kMathC[13];
    return scale2(exp_poly(r), k);
}
__device__ __forceinline__ double mexp2(double x) {
    x = (x < -1021.0) ? -1021.0 : x;
    x = (x > 1023.0) ? 1023.0 : x;
    return mexp2_nc(x);
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
    // P(z) = E(z^2) + z O(z^2), degree 6
    double pe = kMathC[14], po = kMathC[18];
    pe = fma(pe, z2, kMathC[15]);
    po = fma(po, z2, kMathC[19]);
    pe = fma(pe, z2, kMathC[16]);
    po = fma(po, z2, kMathC[20]);
    pe = fma(pe, z2, kMathC[17]);
    const double P = fma(po, z, pe);
    const double R = z * P;
    const double ef = (double)e;
    const double lm = f - s * (f - R);
    return fma(ef, kMathC[13], lm); // e * ln2 + log(m): ln2's rounding contributes <= 1.1e-16 relative
}

// sin(x) and cos(x) for |x| < ~1e5 (two-term Cody-Waite reduction by pi/2; no Payne-Hanek path)
__device__ __forceinline__ void msincos(double x, double* sn, double* cs) {
    const double t = fma(x, kTrigC[12], kMagic);
    const int k = __double2loint(t);
    const double kf = t - kMagic;
    double r = fma(kf, -kTrigC[13], x); // two-term Cody-Waite: exact for |k| < 2^26, residual error k * 6e-17
    r = fma(kf, -kTrigC[14], r);
    const double z = r * r;
    double ps = kTrigC[0], pc = kTrigC[6];
    ps = fma(ps, z, kTrigC[1]);
    pc = fma(pc, z, kTrigC[7]);
    ps = fma(ps, z, kTrigC[2]);
    pc = fma(pc, z, kTrigC[8]);
    ps = fma(ps, z, kTrigC[3]);
    pc = fma(pc, z, kTrigC[9]);
    ps = fma(ps, z, kTrigC[4]);
    pc = fma(pc, z, kTrigC[10]);
    ps = fma(ps, z, kTrigC[5]);
    pc = fma(pc, z, kTrigC[11]);
    const double s0 = fma(r * z, ps, r);
    const double c0 = fma(z * z, pc, fma(z, -0.5, 1.0));
    // quadrant: k mod 4 = 0: (s, c), 1: (c, -s), 2: (-s, -c), 3: (-c, s)
    const bool swap = (k & 1) != 0;
    double ss = swap ? c0 : s0;
    double cc = swap ? s0 : c0;
    if (k & 2) ss = -ss;
    if ((k + 1) & 2) cc = -cc;
    *sn = ss;
    *cs = cc;
}

// x^y for x > 0, |y log x| <= 700
__device__ __forceinline__ double mpow(double x, double y) { return mexp_nc(y * mlog(x)); }

} // namespace mcf
