// microclimf_b200 — device physics of the grid solver (FP64).
//
// This is a re-derivation of the per-cell, per-hour physics behind the reference drivers
// runmicro{1..4}Cpp (src/microclimfCpp.cpp:2052-3223), organised for one-thread-per-cell execution:
//
//   * everything that depends only on the cell (vegetation layer, soil, terrain) is folded ONCE into a
//     CellInv record: the two-stream diffuse solution (ref twostreamdif :1034-1084), wind profile
//     logarithms (windtiCpp/windCpp :1179-1218), soil constants (soilpfun :628), stomatal class
//     (stomparamsCpp :391-440), the canopy-profile integrals of rhcanopy (:1365-1380) ...
//   * everything that depends only on the hour (modes 1/3) is folded ONCE into an HourRec by a prep
//     kernel: solar geometry (solpositionCpp :48-83), the Penman-Monteith air terms (:1223-1232) ...
//   * the remaining per-cell-hour work is written with hoisted logarithms / reciprocals: general
//     pow(a, b) with a cell- or hour-invariant base becomes exp(b * log a), pow(., 2|3|4) become
//     multiplies.  These change results only at rounding level (<= 1e-12 relative); the parity bar is
//     1e-6 (tests/).  Branch structure, comparison directions and NaN behaviour follow the reference
//     (SURVEY.md Appendix B): clamps are written as `x = (x > hi) ? hi : x`, never fmin/fmax.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "mcf_math.cuh"

namespace mcf {

constexpr double kPi = 3.14159265358979323846;
constexpr double kToRad = 3.14159265358979323846 / 180.0;
constexpr double kSb = 5.67e-8;
constexpr double kThetaM = 0.365;
constexpr double kKa = 0.4;
constexpr double kOmdy = (2.0 * 3.14159265358979323846) / (24.0 * 3600.0);
constexpr double kEm = 0.97;

// ---------------------------------------------------------------------------------------------
// Per-hour record (256 B, so a day is one 6 KB bulk copy into shared memory).
// ---------------------------------------------------------------------------------------------
struct __align__(16) HourRec {
    // forcing + point model (ref: climdata / pointm columns, src/microclimfCpp.cpp:2062-2082)
    double tc, es, ea, tdew, pk, Rsw, Rdif, Rlw, u2, umu, soilmp, Gp, dtrp, muGp, kp; // 15
    // solar geometry
    double cosz, sinz;     // cos / sin of the zenith angle (unclamped)
    double cosazi, sinazi; // cos / sin of the solar azimuth
    double tan_sa;         // tan(pi/2 - zenith): horizon test (:2222-2223)
    double tanzc, coszc;   // tan / cos of min(zenith, pi/2): cankCpp (:106-119)
    double kq_tan, kq_cos; // tan / cos of min(zend [DEGREES taken as radians], pi/2): the cankCpp call
                           // inside TVaboveground (:1425) — reproduced as written
    double zend;           // zenith in degrees (solarindexCpp's zend > 90 test when shadowmask=false)
    // Penman-Monteith air terms (:1223-1232)
    double De;  // satvap(tc+0.5) - satvap(tc-0.5)
    double gr4; // 4*0.97*sb*(tc+273.15)^3 / 29.3
    double Rem; // 0.97*sb*(tc+273.15)^4
    double la;  // latent heat of vapourisation
    int32_t sindex, windex; // horizon / wind-shelter sector (:2166-2167)
    // hour-invariant quotients, formed once with IEEE division so that the hot loops only multiply
    double inv_pk;    // 1 / pk
    double invRT;     // 1 / (8.31 (tc + 273.15))                   (soiltempG0 :1268)
    double inv_dtrp;  // 1 / dtrp                                    (soiltemp_hrCpp :1284)
    double muGp_kp;   // muGp / kp                                   (soiltemp_hrCpp :1285)
    double Rbeam0;    // (Rsw - Rdif) / cos(zenith), uncapped        (twostreamCpp :1122, :1152)
    double pmmu;      // la * 43 / pk                                (TVaboveground :1456)
    double inv_pmmu;
    double k1;        // 1 / (2 coszc): cankCpp's x == 1 branch      (:111)
    double kq1;       // 1 / (2 kq_cos): the same for the degrees-as-radians call (:1425)
};
static_assert(sizeof(HourRec) == 320, "HourRec must be 320 bytes");

// ---------------------------------------------------------------------------------------------
// FP64 literals of the hour loops.  An FP64 immediate whose low 32 bits are not zero is materialised by two UMOVs (or two
// IMAD.MOVs) at every use: 65 + 31 of the 1,414 instructions per cell-hour (profiles/r02_kpair_v3.txt).  From constant
// memory the same values are plain c[bank][offset] operands of the DFMA / DMUL / DSETP that use them.  Values as the
// reference writes them (compile-time folded exactly as the expressions they replace).
// ---------------------------------------------------------------------------------------------
#define MCF_LITS(X)                                                                                                    \
    X(kelvin, 273.15) X(em, 0.97) X(emsb, 0.97 * 5.67e-8) X(emhalf, 0.97 * 0.5) X(th_hi, 0.9999) X(th_lo, 0.0001)      \
    X(ws_lo, 0.05) X(uf_lo, 0.001) X(gha_lo, 0.0001) X(cpa, 29.3) X(tr_hi, 0.999) X(alb_lo, 0.01) X(theta_m, 0.365)    \
    X(pct, 0.01) X(et_lo, 0.001) X(plf_a, 0.8753) X(plf_b, 1.7126) X(gs_na, 9999.99) X(gh_a, 0.135) X(gh_b, 1.4)        \
    X(gs_cap, 999.99) X(hlf_a, 1.09767) X(hlf_b, 0.2672778) X(gmin_a, 0.0463) X(gmin_b, 0.2) X(gmin_lo, 0.05)          \
    X(r_lo, 0.001) X(cp, 29.3 * 43.0) X(inv_cp, 1.0 / (29.3 * 43.0)) X(sv_a, 17.27) X(sv_b, 237.3) X(sv_c, 0.61078)     \
    X(wet_a, 0.018) X(c2_a, 1.06) X(two_omdy, 2.0 / ((2.0 * 3.14159265358979323846) / (24.0 * 3600.0))) X(g_cap, 0.6)
struct PhysLit {
#define X(n, v) double n;
    MCF_LITS(X)
#undef X
};
__constant__ PhysLit kL = {
#define X(n, v) v,
    MCF_LITS(X)
#undef X
};

// ---------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double sq(double x) { return x * x; }
__device__ __forceinline__ double pow4(double x) { double y = x * x; return y * y; }
__device__ __forceinline__ double radem4(double tc) { return pow4(tc + kL.kelvin); } // ref radem :24

// ref satvapCpp :480-490
__device__ __forceinline__ double satvap(double tc) {
    return (tc > 0) ? 0.61078 * exp(17.27 * tc / (tc + 237.3)) : 0.61078 * exp(21.875 * tc / (tc + 265.5));
}
// same, through the branch-free mexp / mrcp (hot loops; tc + 237.3 and tc + 265.5 never vanish)
template <int TAB = 0>
__device__ __forceinline__ double satvap_m(double tc) {
    const bool w = tc > 0;
    const double a = w ? kL.sv_a : 21.875;
    const double b = w ? kL.sv_b : 265.5;
    return kL.sv_c * mexp_nc<TAB>(a * tc * mrcp(tc + b)); // |argument| < 20 for any air / surface temperature
}
__device__ __forceinline__ double latent(double tc) { // ref :1227-1232
    return (tc >= 0) ? 45068.7 - 42.8428 * tc : 51078.69 - 4.338 * tc - 0.06367 * tc * tc;
}

// ref juldayCpp :28-37 (int arithmetic; yadj/100 is integer division)
__device__ __forceinline__ int julday(int year, int month, int day) {
    double dd = day + 0.5;
    int madj = month + (month < 3) * 12;
    int yadj = year + (month < 3) * -1;
    double j = trunc(365.25 * (yadj + 4716)) + trunc(30.6001 * (madj + 1)) + dd - 1524.5;
    int b = (int)(2 - trunc((double)(yadj / 100)) + trunc(trunc((double)(yadj / 100)) / 4));
    return (int)(j + (j > 2299160) * b);
}

struct SolPos {
    double zend, zenr, azid;
};
// ref soltimeCpp :39-46 and solpositionCpp :48-83
__device__ __forceinline__ SolPos solposition(double lat, double lon, int year, int month, int day, double lt) {
    int jd = julday(year, month, day);
    double m = 6.24004077 + 0.01720197 * (jd - 2451545.0);
    double eot = -7.659 * sin(m) + 9.863 * sin(2 * m + 3.5932);
    double st = lt + (4.0 * lon + eot) / 60.0;
    double latr = lat * kPi / 180.0;
    double tt = 0.261799 * (st - 12);
    double dec = (kPi * 23.5 / 180) * cos(2 * kPi * ((jd - 159.5) / 365.25));
    double sd, cd, sl, cl, stt, ctt;
    sincos(dec, &sd, &cd);
    sincos(latr, &sl, &cl);
    sincos(tt, &stt, &ctt);
    double coh = sd * sl + cd * cl * ctt;
    double z = acos(coh) * (180 / kPi);
    double sh = coh;
    double hh = atan(sh / sqrt(1 - sh * sh));
    double sazi = cd * stt / cos(hh);
    double num = sl * cd * ctt - cl * sd;
    double cazi = num / sqrt(sq(cd * stt) + sq(num));
    double sqt = 1 - sazi * sazi;
    if (sqt < 0) sqt = 0;
    double azi = 180 + (180 * atan(sazi / sqrt(sqt))) / kPi;
    if (cazi < 0) {
        if (sazi < 0) azi = 180 - azi;
        else azi = 540 - azi;
    }
    SolPos s;
    s.zend = z;
    s.zenr = z * kToRad;
    s.azid = azi;
    return s;
}

// Fill the geometry / air-term part of an HourRec from (solar position, tc).  Shared by the prep kernel
// (modes 1/3: once per hour) and the array-climate kernels (modes 2/4: per cell-hour).
__device__ __forceinline__ void hour_geometry(HourRec& h, const SolPos& s) {
    sincos(s.zenr, &h.sinz, &h.cosz);
    sincos(s.azid * kToRad, &h.sinazi, &h.cosazi);
    h.tan_sa = tan((kPi / 2.0) - s.zenr);
    double zc = (s.zenr > (kPi / 2.0)) ? (kPi / 2.0) : s.zenr;
    h.tanzc = tan(zc);
    h.coszc = cos(zc);
    double zq = (s.zend > (kPi / 2.0)) ? (kPi / 2.0) : s.zend;
    h.kq_tan = tan(zq);
    h.kq_cos = cos(zq);
    h.k1 = 1.0 / (2.0 * h.coszc);
    h.kq1 = 1.0 / (2.0 * h.kq_cos);
    h.zend = s.zend;
    h.sindex = ((int)round(s.azid / 15.0)) % 24;
    h.Rbeam0 = (h.Rsw - h.Rdif) / h.cosz; // forcing fields are filled before the geometry
}
__device__ __forceinline__ void hour_airterms(HourRec& h) {
    h.De = satvap(h.tc + 0.5) - satvap(h.tc - 0.5);
    double tk = h.tc + 273.15;
    h.gr4 = (4 * kEm * kSb * (tk * tk * tk)) / 29.3;
    h.Rem = kEm * kSb * pow4(tk);
    h.la = latent(h.tc);
    h.inv_pk = 1.0 / h.pk;
    h.invRT = 1.0 / (8.31 * tk);
    h.inv_dtrp = 1.0 / h.dtrp;
    h.muGp_kp = h.muGp / h.kp;
    h.pmmu = h.la * (43.0 / h.pk);
    h.inv_pmmu = 1.0 / h.pmmu;
}

// ---------------------------------------------------------------------------------------------
// Per-cell (per vegetation layer) invariants
// ---------------------------------------------------------------------------------------------
struct CellIn { // raw static inputs of one cell / layer
    double hgt, pai, x, gsmax, lref, ltra, clump, leafd, paia, leafden;
    double Smin, Smax, gref, soilb, psie, Vq, Vm, Mc, rho, slope, aspect, tadd, svfa;
};

struct CellInv {
    // terrain / solar index (ref solarindexCpp :85-102): si = cosz*cs + sinz*(cosazi*ssca + sinazi*sssa)
    double cs, ssca, sssa, svfa;
    // soil moisture redistribution (ref soildCpp :1021-1032) in closed form: E = exp(-tadd)
    double Smin, rge, Etadd;
    // canopy extinction (ref cankCpp :104-132)
    double x, kden; // kden = x + 1.774*(x+1.182)^-0.733
    int xflag;      // 1: x == 1, 2: x == 0, 3: x is inf, 0: general
    // two-stream, diffuse part (ref twostreamdifCpp :134-162, twostreamdif :1034-1084)
    double pai, pait, paiaa, om, omp, a, gma, Jdel, h, gref, u1, u2, S1, invS1, invD1, invD2;
    double logclump, loggi, trdn, trdu, amx, albd, Rddn_g, Rdup_z, Rddn_z, Ehm, Ehp; // Ehm/p = exp(-/+h*paiaa)
    double trdif;   // LW gap transmission (ref :1166)
    // wind (ref windtiCpp :1179, windCpp :1189-1218, gturbCpp :373)
    double ufs_coef; // ka / log((zref-d)/zm)
    double gHa_coef; // ka*43 / log((zref-d)/(z0-d))
    double uz_coef;  // uz = uf * uz_coef (before the uref cap)
    // soil thermal (ref soilpfun :628, soilcondCpp :1249) and hydraulic
    double c1, c3, c4, rho, Smax, psie_abs, soilb;
    double cs0, c14; // 2400 rho / 2.64 (volumetric heat capacity of the solids, ref soilcondCpp :1251) and c1 - c4
    // stomata (ref stomparamsCpp :391-440, stomcondCpp :442-458)
    double gsmax, Rsmx, inv02Rsmx, psiw0, kk, rat, inv_stomden; // 1/(exp(-kk*psiw0)-1)
    // canopy / above-ground temperature model (ref TVabove :1298, TVbelow :1381, leaftemp :1333)
    double hgt, d, zm;
    double one_m_lnr;     // 1 - log((zq-d)/zh)/log((zref-d)/zh) at zq = reqhgt (above canopy) or hgt (below)
    int prof_above;       // zq > d + zh
    int above;            // reqhgt >= hgt
    double e_mpai;        // 1 - exp(-pai)
    double shade_fac;     // ((1-exp(-pai))/pai) * (1-omp)      (ref canopycondCpp :469)
    double e_paia, e_paig; // exp(-paia), exp(-(pai-paia))
    double leafd, leafden, nearcoef; // nearcoef = 3.047519 + 0.128642*log(pai)
    double a2h, inth_h, inth_z, zq, hmz; // rhcanopy pieces; zq = reqhgt, hmz = hgt - reqhgt
    double Hf0;           // mincondCpp's Hf for gs = 999.99 (first call in leaftemp :1348)
    double Hf500;         // mincondCpp's Hf for rs = 500 (gs = 0: every night hour)
    // reciprocals of cell invariants (IEEE division, once per cell)
    double inv_rge, inv_Smax, inv_kden, inv_leafd, inv_hgt, inv_a2h;
};

// The fields of CellInv the hour loops read (the rest only feeds cell_setup itself), as an X-macro: the pair kernel
// (k_grid_pair, mcf_kernels_pair.inl) keeps them in SHARED memory instead of ~150 registers per thread.  Layout
// [field pair][cell][2]: fields 2i and 2i + 1 of a cell are one aligned 16-byte element, and the list is ordered so that
// neighbours are used together — the compiler then fetches a pair with one LDS.128 (a quarter-warp's 8 cells are 128
// contiguous bytes: conflict-free).
#define MCF_INV_D(X)                                                                                                   \
    X(cs) X(ssca) X(sssa) X(svfa) X(Smin) X(inv_rge) X(Etadd) X(rge) X(pai) X(x) X(inv_kden) X(a) X(gma) X(om)         \
    X(Jdel) X(pait) X(u1) X(h) X(invD1) X(invS1) X(S1) X(u2) X(gref) X(invD2) X(logclump) X(loggi) X(paiaa) X(trdn)    \
    X(amx) X(Ehm) X(Ehp) X(trdu) X(Rddn_g) X(albd) X(Rddn_z) X(Rdup_z) X(trdif) X(ufs_coef) X(uz_coef) X(gHa_coef)     \
    X(psie_abs) X(soilb) X(inv_Smax) X(rho) X(cs0) X(c1) X(c14) X(c3) X(omp) X(shade_fac) X(rat) X(psiw0) X(kk)        \
    X(inv_stomden) X(gsmax) X(Rsmx) X(inv02Rsmx) X(one_m_lnr) X(e_paig) X(e_paia) X(inv_leafd) X(Hf0) X(Hf500)         \
    X(inv_a2h) X(inth_h) X(inth_z) X(inv_hgt) X(zq) X(hmz) X(e_mpai) X(nearcoef) X(leafden) X(c4)
enum {
#define X(n) INVD_##n,
    MCF_INV_D(X)
#undef X
    kInvD
};
// A CellInv whose fields live in shared memory: member `n` reads element [INVD_n][this thread's cell] where it is used.
// The physics templates take either this or the register-resident CellInv.
#ifndef MCF_INV_PAIRED
#define MCF_INV_PAIRED 0
#endif
// element index of field F in a thread's view of the table (the thread's base pointer carries its cell)
template <int F, int STRIDE>
__device__ __forceinline__ constexpr int inv_index() {
    return MCF_INV_PAIRED ? (F >> 1) * (2 * STRIDE) + (F & 1) : F * STRIDE;
}
constexpr int kInvCellStep = MCF_INV_PAIRED ? 2 : 1; // doubles between neighbouring cells
template <int F, int STRIDE>
struct InvD {
    const double* b;
    __device__ __forceinline__ operator double() const { return b[inv_index<F, STRIDE>()]; }
};
template <int SHIFT, int MASK> // the three small integers share one word per cell
struct InvI {
    const int* b;
    __device__ __forceinline__ operator int() const { return (b[0] >> SHIFT) & MASK; }
};
template <int STRIDE>
struct CellInvS {
#define X(n) InvD<INVD_##n, STRIDE> n;
    MCF_INV_D(X)
#undef X
    InvI<0, 3> xflag;
    InvI<2, 1> prof_above;
    InvI<3, 1> above;
    __device__ __forceinline__ CellInvS(const double* d, const int* i)
        :
#define X(n) n{d},
          MCF_INV_D(X)
#undef X
          xflag{i}, prof_above{i}, above{i} {}
    // cell_setup's result -> this thread's column (d, i as above, writable)
    static __device__ __forceinline__ void store(const CellInv& v, double* d, int* i) {
#define X(n) d[inv_index<INVD_##n, STRIDE>()] = v.n;
        MCF_INV_D(X)
#undef X
        i[0] = v.xflag | (v.prof_above << 2) | (v.above << 3);
    }
};
// which copy of the math tables the physics templates read (mcf_math.cuh): shared-memory replicas next to shared-memory
// invariants (the kernel has then given most of the L1 away), the global tables otherwise
template <class V>
struct MathTab {
    static constexpr int value = 0;
};
#ifndef MCF_GRID_SMEM_TABLES
#define MCF_GRID_SMEM_TABLES 0
#endif
template <>
struct MathTab<CellInv> { // k_grid: the register build of the grid kernel
    static constexpr int value = MCF_GRID_SMEM_TABLES;
};
#ifndef MCF_PAIR_SMEM_TABLES
#define MCF_PAIR_SMEM_TABLES 1
#endif
template <int STRIDE>
struct MathTab<CellInvS<STRIDE>> {
    static constexpr int value = MCF_PAIR_SMEM_TABLES;
};

// ref zeroplanedisCpp :294
__device__ __forceinline__ double zeroplanedis(double h, double pai) {
    if (pai < 0.001) pai = 0.001;
    return (1.0 - (1.0 - exp(-sqrt(7.5 * pai))) / sqrt(7.5 * pai)) * h;
}
// ref roughlengthCpp :302 (psi_h = 0)
__device__ __forceinline__ double roughlength(double h, double pai, double d) {
    double Be = sqrt(0.003 + (0.2 * pai) / 2);
    double zm = (h - d) * exp(-kKa / Be) * exp(kKa * 0.0);
    if (zm > (0.9 * (h - d))) zm = 0.9 * (h - d);
    if (zm < 0.0005) zm = 0.0005;
    return zm;
}
// ref rhcanopy :1365-1380: the z-dependent integral (everything except the uf factor)
__device__ __forceinline__ double rh_integral(double h, double z) {
    if (z == h) return 4.293251 * h;
    double s, c;
    sincos((kPi * z) / h, &s, &c);
    double cp1 = c + 1.0;
    return (2.0 * h * ((48 * atan((sqrt(5.0) * s) / cp1)) / pow(5.0, 1.5) +
                       (32.0 * s) / (cp1 * ((25.0 * s * s) / (cp1 * cp1) + 5.0)))) / kPi;
}

static __device__ __noinline__ void cell_setup(const CellIn& c, double reqhgt2, double zref, double lat, CellInv& v) {
    // --- terrain
    double ss, sa, ca;
    sincos(c.slope * kToRad, &ss, &v.cs);
    sincos(c.aspect * kToRad, &sa, &ca);
    if (c.slope == 0.0) { v.cs = 1.0; ss = 0.0; } // ref :92-93 (exactly cos(zen))
    v.ssca = ss * ca;
    v.sssa = ss * sa;
    v.svfa = c.svfa;
    // --- soil moisture
    v.Smin = c.Smin;
    v.rge = c.Smax - c.Smin;
    v.Etadd = exp(-c.tadd);
    // --- extinction
    v.x = c.x;
    v.kden = c.x + 1.774 * pow(c.x + 1.182, -0.733);
    v.xflag = (c.x == 1.0) ? 1 : (isinf(c.x) ? 3 : ((c.x == 0.0) ? 2 : 0));
    // --- two-stream diffuse (ref :134-162)
    v.pai = c.pai;
    v.gref = c.gref;
    v.pait = c.pai / (1.0 - c.clump);
    v.om = c.lref + c.ltra;
    v.omp = 0.5 * v.om;
    v.a = 1.0 - v.om;
    double del = c.lref - c.ltra;
    double J = 1.0 / 3.0;
    if (c.x != 1.0) {
        double mla = 9.65 * pow(3.0 + c.x, -1.65);
        if (mla > kPi / 2.0) mla = kPi / 2.0;
        J = cos(mla) * cos(mla);
    }
    v.Jdel = J * del;
    v.gma = 0.5 * (v.om + J * del);
    v.h = sqrt(v.a * v.a + 2.0 * v.a * v.gma);
    v.S1 = exp(-v.h * v.pait);
    v.invS1 = 1.0 / v.S1;
    v.u1 = v.a + v.gma * (1.0 - 1.0 / c.gref);
    v.u2 = v.a + v.gma * (1.0 - c.gref);
    double D1 = (v.a + v.gma + v.h) * (v.u1 - v.h) * 1.0 / v.S1 - (v.a + v.gma - v.h) * (v.u1 + v.h) * v.S1;
    double D2 = (v.u2 + v.h) * 1.0 / v.S1 - (v.u2 - v.h) * v.S1;
    v.invD1 = 1.0 / D1;
    v.invD2 = 1.0 / D2;
    double p1 = (v.gma / (D1 * v.S1)) * (v.u1 - v.h);
    double p2 = (-v.gma * v.S1 / D1) * (v.u1 + v.h);
    double p3 = (1.0 / (D2 * v.S1)) * (v.u2 + v.h);
    double p4 = (-v.S1 / D2) * (v.u2 - v.h);
    // --- gap fractions and normalised diffuse fluxes (ref :1052-1082)
    double gi = 0.0;
    if (c.clump > 0.0) gi = pow(c.clump, c.paia / c.pai);
    if (gi > 0.99) gi = 0.99;
    double giu = 0.0;
    if (c.clump > 0.0) giu = pow(c.clump, (c.pai - c.paia) / c.pai);
    if (giu > 0.99) giu = 0.99;
    double trd = gi * gi;
    v.trdn = c.clump * c.clump;
    v.trdu = giu * giu;
    v.paiaa = c.paia / (1.0 - gi);
    v.logclump = log(c.clump); // -inf for clump == 0: exp(Kc * -inf) = 0 = pow(0, Kc)
    v.loggi = log(gi);
    v.amx = c.gref;
    if (v.amx < c.lref) v.amx = c.lref;
    double albd = (1.0 - v.trdn * v.trdn) * (p1 + p2) + v.trdn * v.trdn * c.gref;
    if (albd > v.amx) albd = v.amx;
    if (albd < 0.01) albd = 0.01;
    v.albd = albd;
    double r = (1.0 - v.trdn) * (p3 * exp(-v.h * v.pait) + p4 * exp(v.h * v.pait)) + v.trdn;
    if (r > 1.0) r = 1.0;
    if (r < 0.0) r = 0.0;
    v.Rddn_g = r;
    v.Ehm = exp(-v.h * v.paiaa);
    v.Ehp = exp(v.h * v.paiaa);
    r = (1.0 - v.trdu * v.trdn) * (p1 * v.Ehm + p2 * v.Ehp) + v.trdu * v.trdn * c.gref;
    if (r > 1.0) r = 1.0;
    if (r < 0.0) r = 0.0;
    v.Rdup_z = r;
    r = (1.0 - trd) * (p3 * v.Ehm + p4 * v.Ehp) + trd;
    if (r > 1.0) r = 1.0;
    if (r < 0.0) r = 0.0;
    v.Rddn_z = r;
    v.trdif = (1.0 - v.trdn) * exp(-v.pait) + v.trdn;
    // --- wind (ref :1179-1218)
    double d = zeroplanedis(c.hgt, c.pai);
    double zm = roughlength(c.hgt, c.pai, d);
    if (zm < 1e-6) zm = 1e-6;
    double aw = c.pai / c.hgt;
    v.hgt = c.hgt;
    v.d = d;
    v.zm = zm;
    v.ufs_coef = kKa / log((zref - d) / zm);
    double z0 = 0.2 * zm + d;
    v.gHa_coef = (kKa * 43) / (log((zref - d) / (z0 - d)) + 0);
    v.above = (reqhgt2 >= c.hgt) ? 1 : 0;
    if (v.above) {
        v.uz_coef = log((reqhgt2 - d) / zm) / kKa;
    } else {
        // uh = (uf/ka) log((h-d)/zm), floored at uf; Be = uf/uh floored at 0.001 — all ratios of uf
        double uhc = log((c.hgt - d) / zm) / kKa;
        if (uhc < 1.0) uhc = 1.0;
        double Be = 1.0 / uhc;
        if (Be < 0.001) Be = 0.001;
        double Lc = pow(0.25 * aw, -1.0);
        double Lm = 2 * (Be * Be * Be) * Lc;
        v.uz_coef = uhc * exp(Be * (reqhgt2 - c.hgt) / Lm);
    }
    // --- soil (ref :628-636)
    double frs = c.Vm + c.Vq;
    v.c1 = (0.57 + 1.73 * c.Vq + 0.93 * c.Vm) / (1.0 - 0.74 * c.Vq - 0.49 * c.Vm) - 2.8 * frs * (1.0 - frs);
    v.c3 = 1.0 + 2.6 * pow(c.Mc, -0.5);
    v.c4 = 0.03 + 0.7 * frs * frs;
    v.rho = c.rho;
    v.cs0 = 2400 * c.rho / 2.64;
    v.c14 = v.c1 - v.c4;
    v.Smax = c.Smax;
    v.psie_abs = fabs(c.psie);
    v.soilb = c.soilb;
    // --- stomatal class (ref :391-440)
    double Rsmx = 420.0, psiw0 = -3.1, kk = 0.34, rat = 0.9;
    double alat = fabs(lat);
    if (c.hgt < 1.0 && alat < 22.5) { Rsmx = 450.0; psiw0 = -2.7; kk = 0.39; rat = 0.9; }
    if (c.hgt >= 1.0 && c.hgt < 7.0) { Rsmx = 430.0; psiw0 = -4.0; kk = 0.28; rat = 0.75; }
    if (c.hgt >= 7.0) {
        if (alat < 22.5) { Rsmx = 500.0; psiw0 = -1.75; kk = 0.67; rat = 0.4; }
        else if (c.x < 0.8 || alat > 58.0) { Rsmx = 420.0; psiw0 = -4.09; kk = 0.29; rat = 0.6; }
        else { Rsmx = 500.0; psiw0 = -2.51; kk = 0.46; rat = 0.45; }
    }
    v.gsmax = c.gsmax;
    v.Rsmx = Rsmx;
    v.inv02Rsmx = 1.0 / (0.2 * Rsmx);
    v.psiw0 = psiw0;
    v.kk = kk;
    v.rat = rat;
    v.inv_stomden = 1.0 / (exp(-kk * psiw0) - 1.0);
    // --- above-ground temperature model
    double zh = 0.2 * zm;
    double zq = v.above ? reqhgt2 : c.hgt; // TVabove is evaluated at reqhgt (above canopy) or at hgt (:1435, :1449)
    v.prof_above = (zq > (d + zh)) ? 1 : 0;
    v.one_m_lnr = 1 - log((zq - d) / zh) / log((zref - d) / zh);
    v.e_mpai = 1.0 - exp(-c.pai);
    v.shade_fac = ((1.0 - exp(-c.pai)) / c.pai) * (1.0 - v.omp);
    v.e_paia = exp(-c.paia);
    v.e_paig = exp(-(c.pai - c.paia));
    v.leafd = c.leafd;
    v.leafden = c.leafden;
    v.nearcoef = 3.047519 + 0.128642 * log(c.pai);
    double a2 = 0.4 * (1.0 - (d / c.hgt)) / (1.25 * 1.25);
    v.a2h = a2 * c.hgt;
    v.inth_h = 4.293251 * c.hgt;
    v.inth_z = v.above ? 0.0 : rh_integral(c.hgt, reqhgt2);
    v.zq = reqhgt2;
    v.hmz = c.hgt - reqhgt2;
    double Hlf0 = 1.09767 * pow(1 / 999.99, 0.2672778);
    v.Hf0 = -1.0 / (1.0 + exp(2.0 - Hlf0));
    v.Hf500 = -1.0 / (1.0 + exp(2.0 - 1.09767 * pow(500.0, 0.2672778)));
    v.inv_rge = 1.0 / v.rge;
    v.inv_Smax = 1.0 / c.Smax;
    v.inv_kden = 1.0 / v.kden;
    v.inv_leafd = 1.0 / c.leafd;
    v.inv_hgt = 1.0 / c.hgt;
    v.inv_a2h = 1.0 / v.a2h;
}

// ---------------------------------------------------------------------------------------------
// Per-cell-hour physics
// ---------------------------------------------------------------------------------------------

// Pass-1 products that pass 2 needs again (the day stash): 6 doubles per hour.  Everything else pass 2
// uses (longwave, conductances) is cheap to recompute from cell/hour invariants.
//   radabs  : ground absorbed SW + LW                        (ref soilmodelG0.radabs)
//   surfwet : soil surface wetness                           (ref soilmodelG0.surfwet)
//   radCsw  : canopy absorbed SW                             (ref radmodel2.radCsw)
//   Lhalf   : 0.5*(Rddown + Rdup + k cosz Rbdown), so that radLsw = (1-om)*Lhalf and
//             radLpar = (1-omp)*Lhalf                        (ref :1142-1143)
//   soild   : redistributed soil moisture                    (ref soildCpp :1021-1032)
//   uf      : friction velocity; uz and gHa follow from it   (ref windCpp :1196-1217) — 2.4 % faster than
//             recomputing both in pass 2, and pass 2 no longer reads the wind-shelter sector layer
#ifndef MCF_STASH4
#define MCF_STASH6 1 // default; -DMCF_STASH4 restores the four-variable stash (pass 2 recomputes soil moisture and wind)
#endif
#ifdef MCF_STASH6
constexpr int kStashVars = 6; // + soild, uf: pass 2 rebuilds soil moisture and wind from two loads instead of recomputing them
#else
constexpr int kStashVars = 4;
#endif

// ref soildCpp :1021-1032, closed form of logistic(logit(theta) + tadd)
template <class V>
__device__ __forceinline__ double soil_distribute(const V& v, double soilmp) {
    double theta = (soilmp - v.Smin) * v.inv_rge;
    if (theta > kL.th_hi) theta = kL.th_hi;
    if (theta < kL.th_lo) theta = kL.th_lo;
    double sm = mdiv(theta, theta + (1.0 - theta) * v.Etadd); // divisor in (0, max(1, Etadd)]
    return sm * v.rge + v.Smin;
}

struct Wind {
    double uf, uz, gHa;
};
// ref windCpp :1189-1218 with the cell-invariant logarithms folded
template <class V>
__device__ __forceinline__ Wind wind_hour(const V& v, double u2, double umu, double ws) {
    Wind w;
    if (isnan(ws)) ws = 1.0;
    if (ws < kL.ws_lo) ws = kL.ws_lo;
    double ufs = u2 * v.ufs_coef;
    w.uf = ufs * umu * ws;
    if (w.uf < kL.uf_lo) w.uf = kL.uf_lo;
    w.uz = w.uf * v.uz_coef;
    if (w.uz > u2) w.uz = u2;
    w.gHa = w.uf * v.gHa_coef;
    if (w.gHa < kL.gha_lo) w.gHa = kL.gha_lo;
    return w;
}

// Penman-Monteith surface temperature (ref PenmanMonteith2Cpp :1220-1247), air terms from HourRec.
struct PM {
    double Ts, H, L, mu;
};
__device__ __forceinline__ double pm_ts(const HourRec& h, double dTmx, double Rabs, double gHa, double gV, double G,
                                        double surfwet, double& m_out) {
    double gHr = gHa + h.gr4;
    double m = h.la * (gV * h.inv_pk);
    double L = m * (h.es - h.ea) * surfwet;
    double dT = mdiv(Rabs - h.Rem - L - G, kL.cpa * gHr + m * h.De); // divisor > 0
    if (dT > dTmx) dT = dTmx;
    if (dT > 80.0) dT = 80.0;
    double Ts = dT + h.tc;
    if (Ts < h.tdew) Ts = h.tdew;
    m_out = m;
    return Ts;
}

struct Rad {
    double radGsw, radCsw, Rbdown, Rddown, Rdup, Lhalf;
};

// ref cankCpp :104-132 + twostreamdirCpp :164-185 + twostreamCpp :1086-1163 (shortwave part, Rsw > 0)
template <class V>
__device__ __forceinline__ Rad shortwave(const V& v, const HourRec& h, double si) {
    constexpr int MT = MathTab<V>::value;
    Rad o;
    const double Rsw = h.Rsw, Rdif = h.Rdif;
    const double cosz = h.cosz;
    if (v.pai > 0.0) {
        // canopy extinction coefficient
        double k = msqrt(v.x * v.x + h.tanzc * h.tanzc) * v.inv_kden; // x == inf: NaN, replaced below
        k = (v.xflag == 1) ? h.k1 : k;
        k = (v.xflag == 3) ? 1.0 : k;
        k = (v.xflag == 2) ? h.tanzc : k;
        if (k > 6000.0) k = 6000.0;
        // full-precision reciprocal: kd = k cos(z) / si enters sig = kd^2 + gma^2 - (a + gma)^2 below, which vanishes at
        // kd = h (the two-stream solution's removable singularity) — near it the 1e-12 of the one-step reciprocal was
        // amplified beyond the parity bar (1 value in 8,000 fuzzed problems, profiles/r02_fuzz.txt)
        const double isi = mrcp2(si); // NaN for si == 0: both uses are replaced below, as the reference's inf is
        double kd = k * h.coszc * isi;
        if (si == 0) kd = 1.0;
        double Kc = isi;
        if (si == 0.0) Kc = 600.0;
        // direct-beam two-stream parameters
        const double apg = v.a + v.gma;
        double sig = kd * kd + v.gma * v.gma - apg * apg;
        double ss = 0.5 * (v.om + v.Jdel * mrcp2(kd)) * kd;
        double sstr = v.om * kd - ss;
        double S2 = mexp_lo<MT>(-kd * v.pait);
        double p5 = -ss * (apg - kd) - v.gma * sstr;
        double isig = mrcp2(sig); // p5 .. p10 are differences of large terms for thin canopies: keep their inputs at 1 ulp
        double p5s = p5 * isig;
        double v1 = ss - p5s * (apg + kd);
        double v2 = ss - v.gma - p5s * (v.u1 + kd);
        double p6 = v.invD1 * ((v1 * v.invS1) * (v.u1 - v.h) - (apg - v.h) * S2 * v2);
        double p7 = -v.invD1 * ((v1 * v.S1) * (v.u1 + v.h) - (apg + v.h) * S2 * v2);
        double p8 = sstr * (apg + kd) - v.gma * ss;
        double p8s = -p8 * isig; // p8 / (-sig)
        double v3 = (sstr + v.gma * v.gref - p8s * (v.u2 - kd)) * S2;
        double p9 = -v.invD2 * ((p8s * v.invS1) * (v.u2 + v.h) + v3);
        double p10 = v.invD2 * ((p8s * v.S1) * (v.u2 - v.h) + v3);
        // gap transmissions
        double trbn = mexp_lo<MT>(Kc * v.logclump); // clump == 0: logclump = -inf, mexp -> 3e-308 (the reference: 0)
        if (trbn > kL.tr_hi) trbn = kL.tr_hi;
        if (trbn < 0.0) trbn = 0.0;
        double trb = mexp_lo<MT>(Kc * v.loggi);
        if (trb > kL.tr_hi) trb = kL.tr_hi;
        if (trb < 0.0) trb = 0.0;
        double S2a = mexp_lo<MT>(-kd * v.paiaa);
        // black-sky albedo
        double albb = (1.0 - v.trdn * trbn) * (p5s + p6 + p7) + v.trdn * trbn * v.gref;
        if (albb > v.amx) albb = v.amx;
        if (albb < kL.alb_lo) albb = kL.alb_lo;
        double Rdbdn_g = (1.0 - trbn) * (p8s * S2 + p9 * v.S1 + p10 * v.invS1);
        if (Rdbdn_g > v.amx) Rdbdn_g = v.amx;
        if (Rdbdn_g < 0.0) Rdbdn_g = 0.0;
        double Rdbup_z = (1.0 - v.trdu * trbn) * (p5s * S2a + p6 * v.Ehm + p7 * v.Ehp) + v.trdu * trbn * v.gref;
        if (Rdbup_z > v.amx) Rdbup_z = v.amx;
        if (Rdbup_z < 0.0) Rdbup_z = 0.0;
        double Rdbdn_z = (1.0 - trb) * (p8s * S2a + p9 * v.Ehm + p10 * v.Ehp);
        if (Rdbdn_z > v.amx) Rdbdn_z = v.amx;
        if (Rdbdn_z < 0.0) Rdbdn_z = 0.0;
        // incident flux
        double Rbeam = h.Rbeam0;
        if (Rbeam > 1352.0) Rbeam = 1352.0;
        double Rb = Rbeam * cosz;
        double trg = trb + (1 - trb) * S2;
        double Rbc = (trg * si + (1 - trg) * cosz) * Rbeam;
        double Rbdn_g = trbn + (1.0 - trbn) * S2;
        if (Rbdn_g > 1.0) Rbdn_g = 1.0;
        if (Rbdn_g < 0.0) Rbdn_g = 0.0;
        double Rds = Rdif * v.svfa;
        o.radGsw = (1.0 - v.gref) * (v.Rddn_g * Rds + Rdbdn_g * Rb + Rbdn_g * Rbeam * si);
        double maxg = (1.0 - v.gref) * (Rds + Rbeam * si);
        if (o.radGsw > maxg) o.radGsw = maxg;
        o.radCsw = (1.0 - v.albd) * Rds + (1.0 - albb) * Rbc;
        o.Rbdown = (trb + (1.0 - trb) * S2a) * Rbeam;
        o.Rddown = v.Rddn_z * Rds + Rdbdn_z * Rb;
        o.Rdup = v.Rdup_z * Rds + Rdbup_z * Rb;
        o.Lhalf = 0.5 * (o.Rddown + o.Rdup + k * cosz * o.Rbdown); // ref :1142-1143
    } else {
        o.Rbdown = h.Rbeam0;
        o.Rddown = Rdif * v.svfa;
        o.Rdup = v.gref * (Rdif * v.svfa + (Rsw - Rdif));
        o.radGsw = (1.0 - v.gref) * (v.svfa * Rdif + si * o.Rbdown);
        o.radCsw = o.radGsw;
        o.Lhalf = 0.0;
    }
    return o;
}

// Soil-moisture limitation of stomatal conductance: the theta-only factor of stomcondCpp (:450-455),
// gs2 = mu * gsmax.  Shared by the (up to) three stomcondCpp calls of one cell-hour.
template <class V>
__device__ __forceinline__ double stom_gs2(const V& v, double theta) {
    constexpr int MT = MathTab<V>::value;
    double thetan = v.rat * theta + (1 - v.rat) * kL.theta_m;
    double Se = thetan * v.inv_Smax;
    if (Se > 1.0) Se = 1.0;
    double psiw = -v.psie_abs * mexp_nc<MT>(-v.soilb * mlog<MT>(Se)) * kL.pct; // pow(Se, -b), Se in (0, 1]
    if (psiw < v.psiw0) psiw = v.psiw0;
    double mu = 1.0 - (mexp_nc<MT>(-v.kk * psiw) - 1.0) * v.inv_stomden;
    return mu * v.gsmax;
}
// ref stomcondCpp :442-458 given gs2
template <class V>
__device__ __forceinline__ double stomcond(const V& v, double Rswabs, double gs2) {
    constexpr int MT = MathTab<V>::value;
    if (Rswabs <= 0.0) return 0.0;
    // Rswabs >= Rsmx (always the case for sunlit leaves: kq is ~6000) clamps the exponent to 0 and
    // 2^0 = 1 exactly, so the exponential is only evaluated below saturation
    double gs = v.gsmax;
    if (Rswabs < v.Rsmx) gs = v.gsmax * mexp2_nc<MT>(-(v.Rsmx - Rswabs) * v.inv02Rsmx); // argument in (-5, 0)
    if (gs > gs2) gs = gs2;
    return gs;
}

struct Above {
    double Tz, tleaf, rh, lwdn, lwup;
};

// ref TVaboveground :1411-1472 (with TVabove :1298, leaftemp :1333, TVbelow :1381, rhcanopy :1365)
template <class V>
__device__ __forceinline__ Above above_ground(const V& v, const HourRec& h, double dTmx, double soilm, double Tg,
                                              double G, const Wind& w, double radCsw, double radClw, double Lhalf) {
    constexpr int MT = MathTab<V>::value;
    Above out;
    const double tc = h.tc, ea = h.ea, Rlw = h.Rlw;
    // ground and surface wetness
    double esTg = satvap_m<MT>(Tg);
    double eT = esTg - ea;
    if (eT < kL.et_lo) eT = kL.et_lo;
    double plf = kL.plf_a - kL.plf_b * mlog<MT>(eT);
    double gwet = mrcp(1.0 + mexp_nc<MT>(-plf)); // plf in [-6, 13]
    double surfwet = (soilm - v.Smin) * v.inv_rge;
    if (surfwet > gwet) gwet = surfwet;
    // canopy conductance (ref canopycondCpp :460-477) with k from the degrees-as-radians cankCpp call
    double gV = 0.0;
    double gs2 = 0.0;
    bool have_gs2 = false;
    // leaf reflectance / transmittance NA (bare cells of real rasters keep NA there, R/internal.R:1052-1058):
    // canopycondCpp skips its body and returns Gs = 9999.99 (ref :463-464), whatever pai is
    const bool om_na = isnan(v.omp);
    double gS = om_na ? kL.gs_na : 0.0;
    if (v.pai != 0.0 && !om_na) { // pai == 0: Rshade_abs is 0/0 = NaN in the reference, so gS is NaN and gV stays 0
        double kq = msqrt(v.x * v.x + h.kq_tan * h.kq_tan) * v.inv_kden;
        kq = (v.xflag == 1) ? h.kq1 : kq;
        kq = (v.xflag == 3) ? 1.0 : kq;
        kq = (v.xflag == 2) ? h.kq_tan : kq;
        if (kq > 6000.0) kq = 6000.0;
        double Rshade_abs = h.Rdif * v.shade_fac; // NaN for pai == 0 (0/0), as in the reference
        double Rsun_abs = (h.Rsw - h.Rdif) * kq * (1 - v.omp) + Rshade_abs;
        if (Rshade_abs <= 0.0 && Rsun_abs <= 0.0) {
            gS = 0.0; // both stomcondCpp calls return 0
        } else {
            // exp(-kq pai) underflows to exactly 0 in the reference once kq pai > 745 (kq is ~6000 whenever
            // the zenith-in-degrees argument is clamped, i.e. almost always)
            const double kp_ = kq * v.pai;
            double esun = 0.0;
            if (kp_ < 700.0) esun = mexp_nc<MT>(-kp_);
            double P_sun = (1.0 - esun) * mrcp(kq);
            double P_shade = v.pai - P_sun;
            gs2 = stom_gs2(v, soilm);
            have_gs2 = true;
            double gs_sun, gs_shade;
            if (isnan(Rsun_abs)) gs_sun = Rsun_abs; // NaN propagates through stomcondCpp (no comparison is true)
            else gs_sun = stomcond(v, Rsun_abs, gs2);
            if (isnan(Rshade_abs)) gs_shade = Rshade_abs;
            else gs_shade = stomcond(v, Rshade_abs, gs2);
            gS = gs_sun * P_sun + gs_shade * P_shade;
        }
    }
    if (gS > 0.0) gV = mdiv(w.gHa * gS, w.gHa + gS); // 1 / (1/gHa + 1/gS)
    // canopy temperature
    double Rabs = radCsw + radClw;
    double m;
    double Tcan = pm_ts(h, dTmx, Rabs, w.gHa, gV, G, surfwet, m);
    double esTcan = satvap_m<MT>(Tcan);
    double ez;
    if (v.above) {
        // ref TVabove :1298-1313 at reqhgt
        if (v.prof_above) {
            out.Tz = tc + (Tcan - tc) * v.one_m_lnr;
            ez = ea + (esTcan - ea) * surfwet * v.one_m_lnr;
        } else {
            out.Tz = Tcan;
            ez = ea + (esTcan - ea) * surfwet;
        }
        out.tleaf = Tcan;
        out.lwup = kL.emsb * radem4(Tcan);
        out.lwdn = Rlw;
    } else {
        double pmH = kL.cpa * w.gHa * (Tcan - tc);
        double pmL = m * (esTcan - ea) * surfwet;
        double pmmu = h.pmmu;
        // ---- leaf temperature (ref leaftemp :1333-1364)
        double lwcan = kL.emsb * radem4(Tcan);
        double lwgro = kL.emsb * radem4(Tg);
        out.lwup = v.e_paig * lwgro + (1 - v.e_paig) * lwcan;
        out.lwdn = v.e_paia * Rlw + (1 - v.e_paia) * lwcan;
        double lwabs = kL.emhalf * (out.lwup + out.lwdn);
        // radLsw / radLpar are set to 0, not computed, at night and for pai == 0 (ref twostreamCpp :1147-1163): with NA
        // leaf reflectance the product (1 - om) * 0 would be NaN where the reference has 0
        const bool lit = (h.Rsw > 0.0) && (v.pai > 0.0);
        const double radLsw = lit ? (1.0 - v.om) * Lhalf : 0.0;
        double leafabs = radLsw + lwabs;
        double gh = kL.gh_a * msqrt(w.uz * v.inv_leafd) * kL.gh_b;
        double Rnetl = leafabs - lwcan;
        // leaftemp calls mincondCpp twice (gs = 999.99, then the stomatal gs; ref :1348, :1354) and keeps the larger
        // gmin.  gmin = 0.0463 (|Hf| |Rnet| / leafd)^0.2 is monotone in |Hf|, so one power of the larger |Hf| gives
        // max(gmin1, gmin2) exactly; at night (gs = 0, rs = 500) the second |Hf| is a constant.
        double Hfmag = fabs(v.Hf0);
        const bool stom = v.gsmax < kL.gs_cap;
        double gs = 0.0;
        if (stom) {
            double radLpar = lit ? (1.0 - v.omp) * Lhalf : 0.0;
            if (radLpar > 0.0) {
                if (!have_gs2) gs2 = stom_gs2(v, soilm);
                gs = stomcond(v, radLpar, gs2);
            }
            double Hf2 = v.Hf500;
            if (gs > 0.0) {
                double rs = mrcp(gs);
                if (rs > 500.0) rs = 500.0;
                const double Hlf = kL.hlf_a * mpow<MT>(rs, kL.hlf_b);
                Hf2 = -mrcp(1.0 + mexp_nc<MT>(2.0 - Hlf));
            }
            const double m2 = fabs(Hf2);
            if (Hfmag < m2) Hfmag = m2;
        }
        // H == 0: mlog<MT>(0) ~ -709, so the power is ~1e-62 instead of 0; either way gmin takes its floor
        double gmin = kL.gmin_a * mpow<MT>((Hfmag * fabs(Rnetl)) * v.inv_leafd, kL.gmin_b);
        if (gmin < kL.gmin_lo) gmin = kL.gmin_lo;
        if (gh < gmin) gh = gmin;
        double gVl = gh;
        if (stom) {
            gVl = 0.0;
            if (gs > 0.0) gVl = mdiv(gh * gs, gh + gs); // 1 / (1/gh + 1/gs)
        }
        double ml;
        double tleaf = pm_ts(h, dTmx, leafabs, gh, gVl, 0.0, surfwet, ml);
        double esTl = satvap_m<MT>(tleaf);
        double lfH = kL.cpa * gh * (tleaf - tc);
        double lfL = ml * (esTl - ea) * surfwet;
        out.tleaf = tleaf;
        // ---- canopy-top state (ref TVabove at hgt, :1449)
        double Th, eh;
        if (v.prof_above) {
            Th = tc + (Tcan - tc) * v.one_m_lnr;
            eh = ea + (esTcan - ea) * surfwet * v.one_m_lnr;
        } else {
            Th = Tcan;
            eh = ea + (esTcan - ea) * surfwet;
        }
        // ---- diffusivities (ref TVbelow :1385-1390, rhcanopy :1365-1380)
        double mu_r = v.inv_a2h * mrcp(w.uf); // (uf / a2h) / uf^2
        double Rc = v.inth_h * mu_r;
        if (Rc < kL.r_lo) Rc = kL.r_lo;
        double Rz = v.inth_z * mu_r;
        if (Rz < kL.r_lo) Rz = kL.r_lo;
        // The reference weights ground, canopy-top and canopy sources with the conductances Kg, Kh, Kc and divides by their
        // sum (four divisions in a row).  With the resistances A = 1/Kc, B = 1/Kg, C = 1/Kh the normalised weights are
        // AC, AB, BC over AB + AC + BC: one reciprocal, all terms positive (no cancellation).  C == 0 (both resistance
        // integrals at their floor): Kh = 1/0 = inf and the reference's inf / inf is NaN — kept.
        const double iKc = Rc * v.inv_hgt;          // A
        const double rB = Rz * v.zq;                // B
        const double rC = (Rc - Rz) * v.hmz;        // C
        const double wAB = iKc * rB, wAC = iKc * rC, wBC = rB * rC;
        double iden = mrcp(wAB + wAC + wBC);
        if (rC == 0.0) iden = (double)NAN;
        const double wG = wAC * iden, wH = wAB * iden, wC = wBC * iden;
        // ---- temperature below canopy (ref :1447-1453)
        {
            const double cp = kL.cp;
            double Flux = pmH * v.e_mpai;
            double SH = Th * cp;
            double SG = Tg * cp;
            double mxnear = fabs(tleaf - Th) * cp;
            double SC = SH + Flux * iKc;
            double farg = wG * SG + wH * SH + wC * SC;
            double nearf = v.nearcoef * (lfH * v.leafden);
            if (fabs(nearf) > mxnear) nearf = (nearf > 0.0) ? mxnear : -mxnear;
            if (isnan(nearf)) nearf = 0;
            out.Tz = (nearf + farg) * kL.inv_cp;
        }
        // ---- vapour pressure below canopy (ref :1455-1460)
        {
            double Flux = pmL * v.e_mpai;
            double SH = eh * pmmu;
            double SG = esTg * gwet * pmmu;
            double mxnear = fabs(esTl - eh) * pmmu;
            double SC = SH + Flux * iKc;
            double farg = wG * SG + wH * SH + wC * SC;
            double nearf = v.nearcoef * (lfL * v.leafden);
            if (fabs(nearf) > mxnear) nearf = (nearf > 0.0) ? mxnear : -mxnear;
            if (isnan(nearf)) nearf = 0;
            ez = (nearf + farg) * h.inv_pmmu;
        }
    }
    out.rh = mdiv(ez, satvap_m<MT>(out.Tz)) * 100.0;
    if (out.rh > 100.0) out.rh = 100.0;
    // limits (ref :1467-1470; std::max/min over {tleaf, tc, Tg, Tcan} with their NaN-ignoring fold order)
    double tmx = out.tleaf;
    if (tmx < tc) tmx = tc;
    if (tmx < Tg) tmx = Tg;
    if (tmx < Tcan) tmx = Tcan;
    double tmn = out.tleaf;
    if (tc < tmn) tmn = tc;
    if (Tg < tmn) tmn = Tg;
    if (Tcan < tmn) tmn = Tcan;
    tmx += 2.0;
    tmn -= 2.0;
    if (out.Tz > tmx) out.Tz = tmx;
    if (out.Tz < tmn) out.Tz = tmn;
    return out;
}

} // namespace mcf
