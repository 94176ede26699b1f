// microclimf_b200 — branch-free FP64 elementary functions for the hot loops.
//
// The grid solver is bound by the FP64 pipe and by instruction issue, and the CUDA math
// library's double-precision exp / log / pow / division carry special-case branches, slow-path calls
// and 64-bit immediates that more than double the instruction count and spill the loop bodies out of
// the 32 KB L1.5 instruction cache.  These replacements are straight-line code:
//
//   mrcp, mdiv : MUFU.RCP64H seed (~20 bits) + one Newton step                      (<= 1e-12)
//   msqrt      : MUFU.RSQ64H seed + one coupled Newton step + Heron correction     (<= 1 ulp)
//   mexp       : reduction by ln2/32, 32-entry table of 2^(j/32), degree-5 polynomial          (<= 8e-14)
//   mlog       : exponent/mantissa split, 64-entry table {1/c, log c}, log1p(m/c - 1), degree 6 (<= 3e-16 absolute
//                for |log x| < 1, relative beyond: every call site feeds an exponential or a sum)
//   (-DMCF_MATH_POLY: the table-free degree-9 exp and f - s(f - zP(z)) log they replaced)
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
// 1 / x to ~1 ulp (two Newton steps): for the few quotients whose error a later cancellation amplifies
__device__ __forceinline__ double mrcp2(double x) {
    double r = rcp_seed(x);
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
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
#ifndef MCF_SQRT_FAST
    double d = fma(-g, g, x);       // Heron correction: <= 1 ulp
    g = fma(d, h, g);
#endif
    return (x == 0.0) ? 0.0 : g;
}

// 1 / sqrt(x), x > 0 finite normal: MUFU.RSQ64H seed (~2^-23) + one Newton step (~2^-45)
__device__ __forceinline__ double mrsqrt(double x) {
    const double y = rsqrt_seed(x);
    const double e = fma(-x * y, y, 1.0);
    return fma(y * 0.5, e, y);
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
    // table-driven variants: [21] 1/6  [22] 1/24  [23] 1/120  [24] 32/ln2  [25] ln2/32  [26] 1/3  [27] 1/5  [28] -1/6
    0.16666666666666666, 0.041666666666666664, 0.008333333333333333, 46.16624130844683, 0.02166084939249829,
    0.3333333333333333, 0.2, -0.16666666666666666,
    0, 0, 0};

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

// Table-driven exp / log (default; -DMCF_MATH_POLY selects the table-free polynomials below).  The physics is a
// long dependent chain, so the hot loops are bound by FP64 latency as much as by FP64 issue: a
// 32-entry table of 2^(j/32) shrinks exp's reduced argument to |r| <= ln2/64, where a degree-5 polynomial in
// Estrin form (dependency depth 3) replaces the degree-9 one, and a 64-entry table of {1/c_j, log c_j} turns
// log(m) into log1p(m/c_j - 1) with |r| < 2^-7 and no division.  9 and 10 FP64 instructions instead of 13 and 18,
// about half the dependency depth.  The tables (256 B + 1 KB, generated and verified against mpmath by
// tools/gen_math_tables.py: exp <= 2.6e-14 relative, log <= 2.4e-16 of max(1, |log x|)) stay L1-resident.
__device__ const double kExpTab[32] = {
    1.0, 1.0218971486541166, 1.0442737824274138, 1.0671404006768237,
    1.0905077326652577, 1.1143867425958924, 1.1387886347566916, 1.1637248587775775,
    1.189207115002721, 1.215247359980469, 1.241857812073484, 1.2690509571917332,
    1.2968395546510096, 1.3252366431597413, 1.3542555469368927, 1.383909881963832,
    1.4142135623730951, 1.4451808069770467, 1.4768261459394993, 1.5091644275934228,
    1.5422108254079407, 1.5759808451078865, 1.6104903319492543, 1.645755478153965,
    1.681792830507429, 1.718619298122478, 1.7562521603732995, 1.7947090750031072,
    1.8340080864093424, 1.8741676341103, 1.9152065613971474, 1.9571441241754002,
};
__device__ const double2 kLogTab[64] = { // {1 / c_j, log c_j}, c_j = 1 + (j + 0.5) / 64
    {0.9922480620155039, 0.007782140442054963}, {0.9770992366412213, 0.023167059281534418},
    {0.9624060150375939, 0.03831886430213666}, {0.9481481481481482, 0.05324451451881224},
    {0.9343065693430657, 0.06795066190850778}, {0.920863309352518, 0.08244366921107454},
    {0.9078014184397163, 0.09672962645855114}, {0.8951048951048951, 0.11081436634029011},
    {0.8827586206896552, 0.12470347850095725}, {0.8707482993197279, 0.1384023228591192},
    {0.8590604026845637, 0.151916042025842}, {0.847682119205298, 0.16524957289530717},
    {0.8366013071895425, 0.17840765747281825}, {0.8258064516129032, 0.19139485299962947},
    {0.8152866242038217, 0.20421554142869083}, {0.8050314465408805, 0.2168739383006143},
    {0.7950310559006211, 0.2293741010648459}, {0.7852760736196319, 0.24171993688714513},
    {0.7757575757575758, 0.25391520998096345}, {0.7664670658682635, 0.2659635484971379},
    {0.757396449704142, 0.2778684510034563}, {0.7485380116959064, 0.2896332925830427},
    {0.7398843930635838, 0.30126133057816185}, {0.7314285714285714, 0.3127557100038969},
    {0.7231638418079096, 0.324119468654212}, {0.7150837988826816, 0.3353555419211378},
    {0.7071823204419889, 0.3464667673462086}, {0.6994535519125683, 0.3574558889218038},
    {0.6918918918918919, 0.36832556115870757}, {0.6844919786096256, 0.3790783529349695},
    {0.6772486772486772, 0.38971675114002524}, {0.6701570680628273, 0.40024316412701266},
    {0.6632124352331606, 0.4106599249852683}, {0.6564102564102564, 0.42096929464412963},
    {0.649746192893401, 0.43117346481837143}, {0.6432160804020101, 0.4412745608048752},
    {0.6368159203980099, 0.4512746441394586}, {0.6305418719211823, 0.46117571512217015},
    {0.624390243902439, 0.470979715218791}, {0.6183574879227053, 0.48068852934575196},
    {0.6124401913875598, 0.4903039880451939}, {0.6066350710900474, 0.49982786955644926},
    {0.6009389671361502, 0.5092619017898079}, {0.5953488372093023, 0.5186077642080457},
    {0.5898617511520737, 0.5278670896208424}, {0.5844748858447488, 0.5370414658968837},
    {0.579185520361991, 0.5461324375981356}, {0.5739910313901345, 0.5551415075405016},
    {0.5688888888888889, 0.564070138284803}, {0.5638766519823789, 0.5729197535617854},
    {0.5589519650655022, 0.5816917396346225}, {0.5541125541125541, 0.5903874466021763},
    {0.5493562231759657, 0.5990081896460834}, {0.5446808510638298, 0.6075552502245418},
    {0.540084388185654, 0.616029877215514}, {0.5355648535564853, 0.6244332880118936},
    {0.5311203319502075, 0.6327666695710378}, {0.5267489711934157, 0.6410311794209312},
    {0.5224489795918368, 0.6492279466251097}, {0.5182186234817814, 0.65735807270836},
    {0.5140562248995983, 0.6654226325450905}, {0.5099601593625498, 0.6734226752121667},
    {0.5059288537549407, 0.6813592248079031}, {0.5019607843137255, 0.689233281238809},
};

// Where the two tables are read from (template parameter TAB of the functions below):
//   0  global memory through the read-only path (__ldg): L1-resident in kernels that leave the L1 its capacity;
//   1  the first 12 KB of the kernel's DYNAMIC shared memory, which the kernel has filled with math_tables_to_smem():
//      conflict-free replicas — the 2^(j/32) table [32][16 copies] of doubles (a 64-bit LDS is served per half-warp: lane
//      c reads copy c & 15, 16 distinct bank pairs) and the {1/c, log c} table [64][8 copies] of double2 (a 128-bit LDS is
//      served per quarter-warp: lane c reads copy c & 7).  For kernels whose shared-memory carve-out leaves a 28 KB L1,
//      where the data-dependent __ldg gathers miss (k_grid_pair: 7 % of its stall samples, profiles/r02_kpair_v2).
extern __shared__ __align__(128) unsigned char mcf_dyn_smem[]; // every extern __shared__ array of a kernel starts here
constexpr int kMathSmemExpBytes = 32 * 16 * 8, kMathSmemLogBytes = 64 * 8 * 16;
constexpr int kMathSmemBytes = kMathSmemExpBytes + kMathSmemLogBytes;
template <int TAB>
__device__ __forceinline__ double exp_tab_entry(int j) {
    if (TAB == 0) return __ldg(&kExpTab[j]);
    return reinterpret_cast<const double*>(mcf_dyn_smem)[(j << 4) | (threadIdx.x & 15)];
}
template <int TAB>
__device__ __forceinline__ double2 log_tab_entry(int j) {
    if (TAB == 0) return __ldg(&kLogTab[j]);
    return reinterpret_cast<const double2*>(mcf_dyn_smem + kMathSmemExpBytes)[(j << 3) | (threadIdx.x & 7)];
}
// all threads of the CTA; the caller synchronises afterwards
__device__ __forceinline__ void math_tables_to_smem() {
    double* e = reinterpret_cast<double*>(mcf_dyn_smem);
    for (int i = threadIdx.x; i < 32 * 16; i += blockDim.x) e[i] = kExpTab[i >> 4];
    double2* l = reinterpret_cast<double2*>(mcf_dyn_smem + kMathSmemExpBytes);
    for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) l[i] = kLogTab[i >> 3];
}

// exp(r) 2^(k/32) for |r| <= ln2/64
template <int TAB = 0>
__device__ __forceinline__ double exp_tab(double r, int k) {
    const double r2 = r * r;
    const double a = fma(r, kMathC[21], 0.5);
    const double b = fma(r, kMathC[23], kMathC[22]);
    const double c = fma(r2, b, a);
    const double p = fma(r2, c, r + 1.0);
    const double t = exp_tab_entry<TAB>(k & 31);
    return __hiloint2double(__double2hiint(t * p) + ((k >> 5) << 20), __double2loint(t * p));
}

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
template <int TAB = 0>
__device__ __forceinline__ double mexp_nc(double x) {
#ifndef MCF_MATH_POLY
    const double t = fma(x, kMathC[24], kMagic); // round(32 x / ln2) lands in the low word
    const int k = __double2loint(t);
    const double kf = t - kMagic;
    // one-constant reduction (exact product inside the FMA): error |k| * 2.4e-18 in r, <= 8e-14 relative for |x| <= 700
    const double r = fma(kf, -kMathC[25], x);
    return exp_tab<TAB>(r, k);
#else
    const double t = fma(x, kMathC[10], kMagic); // round-to-nearest integer lands in the low word
    const int k = __double2loint(t);
    const double kf = t - kMagic;
    // one-constant reduction: the product is exact inside the FMA, so the only error is ln2's own rounding,
    // |k| * 8e-17 in r, i.e. <= 8e-14 relative in the result for |x| <= 700 (inside the 1e-12 budget)
    const double r = fma(kf, -kMathC[13], x);
    return scale2(exp_poly(r), k);
#endif
}
// exp(x) for x <= 700: arguments below -708 (including -inf) return exp(-708) ~ 3e-308
template <int TAB = 0>
__device__ __forceinline__ double mexp_lo(double x) {
    x = (x < -708.0) ? -708.0 : x;
    return mexp_nc<TAB>(x);
}
// exp(x), any x
template <int TAB = 0>
__device__ __forceinline__ double mexp(double x) {
    x = (x < -708.0) ? -708.0 : x;
    x = (x > 709.0) ? 709.0 : x;
    return mexp_nc<TAB>(x);
}

// 2^x for |x| <= 1000 (no clamps)
template <int TAB = 0>
__device__ __forceinline__ double mexp2_nc(double x) {
#ifndef MCF_MATH_POLY
    const double t = fma(x, 32.0, kMagic);
    const int k = __double2loint(t);
    const double kf = t - kMagic;
    const double r = fma(kf, -0.03125, x) * kMathC[13]; // (x - k/32) ln2, the subtraction is exact
    return exp_tab<TAB>(r, k);
#else
    const double t = x + kMagic;
    const int k = __double2loint(t);
    const double kf = t - kMagic;
    const double r = (x - kf) * kMathC[13];
    return scale2(exp_poly(r), k);
#endif
}
template <int TAB = 0>
__device__ __forceinline__ double mexp2(double x) {
    x = (x < -1021.0) ? -1021.0 : x;
    x = (x > 1023.0) ? 1023.0 : x;
    return mexp2_nc<TAB>(x);
}

template <int TAB = 0>
__device__ __forceinline__ double mlog(double x) {
#ifndef MCF_MATH_POLY
    const int hi = __double2hiint(x);
    const int e = (hi >> 20) - 1023;
    const double m = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, __double2loint(x)); // mantissa in [1, 2)
    const double2 tc = log_tab_entry<TAB>((hi >> 14) & 63);                                  // {1 / c_j, log c_j}
    const double r = fma(m, tc.x, -1.0); // m / c_j - 1, |r| < 2^-7 (single rounding)
    const double r2 = r * r;
    // log1p(r) = r + r^2 (-1/2 + r/3 - r^2/4 + r^3/5 - r^4/6), Estrin form
    const double a = fma(r, kMathC[26], -0.5);
    const double b = fma(r, kMathC[27], -0.25);
    const double c = fma(r2, kMathC[28], b);
    const double q = fma(r2, c, a);
    const double l1p = fma(r2, q, r);
    return fma((double)e, kMathC[13], tc.y + l1p);
#else
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
#endif
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
template <int TAB = 0>
__device__ __forceinline__ double mpow(double x, double y) { return mexp_nc<TAB>(y * mlog<TAB>(x)); }

} // namespace mcf
