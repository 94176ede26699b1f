/*
 * TEST INFRASTRUCTURE ONLY — plain-C restatement of the reference's grid-model algorithm.
 *
 * Not part of the product path: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs
 * may load this library.  It restates, function by function and in the reference's own evaluation order
 * (standard libm pow/exp/log, no hoisting, no fused multiply-add: built with -ffp-contract=off), the
 * algorithm of ilyamaclean/microclimf v2.0.0 behind runmicro{1..4}Cpp and runbioclim{1..4}Cpp.
 * Citations `cpp:N` are /root/reference/src/microclimfCpp.cpp:N.
 *
 * Pinning: this restatement is checked (tests/test_oracle_cpu.py) against
 *   (1) the UNMODIFIED reference compiled here (oracle/_ref, see oracle/Makefile) on seeded problems, and
 *   (2) the golden vectors in tests/golden/ that were produced by that compiled reference.
 * The reference's own test-suite holds no golden vectors for this path.
 *
 * One deliberate extension: mcf_problem.has_twi_mean lets a caller supply the whole-raster mean that
 * soildCppm subtracts (cpp:993-1004), so that a column band reproduces the whole-raster result (the
 * multi-GPU sharding contract).  With has_twi_mean = 0 the behaviour is the reference's.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "microclimf_b200.h"

#define PI 3.14159265358979323846
#define TORAD (PI / 180.0)
static const double SB = 5.67e-8;    /* Stefan-Boltzmann, reference global `sb`  */
static const double KA = 0.4;        /* von Karman, reference global `ka`        */
static const double THETAM = 0.365;  /* reference global `thetam`                */
#define OMDY ((2.0 * PI) / (24.0 * 3600.0))

static double na_real(void) {
    union { uint64_t u; double d; } v;
    v.u = MCF_NA_REAL_BITS;
    return v.d;
}

/* ---------------------------------------------------------------------------------------------- */
/* solar geometry                                                                                   */
/* ---------------------------------------------------------------------------------------------- */
static double radem(double tc) { return pow(tc + 273.15, 4.0); } /* cpp:24 */

static int julday(int year, int month, int day) { /* cpp:28-37 */
    double dd = day + 0.5;
    int madj = month + (month < 3) * 12;
    int yadj = year + (month < 3) * -1;
    double j = trunc(365.25 * (yadj + 4716)) + trunc(30.6001 * (madj + 1)) + dd - 1524.5;
    int b = (int)(2 - trunc((double)(yadj / 100)) + trunc(trunc((double)(yadj / 100)) / 4));
    return (int)(j + (j > 2299160) * b);
}

typedef struct { double zend, zenr, azid; } sol_t;

static sol_t solposition(double lat, double lon, int year, int month, int day, double lt) { /* cpp:39-83 */
    int jd = julday(year, month, day);
    double m = 6.24004077 + 0.01720197 * (jd - 2451545.0);
    double eot = -7.659 * sin(m) + 9.863 * sin(2 * m + 3.5932);
    double st = lt + (4.0 * lon + eot) / 60.0;
    double latr = lat * PI / 180.0;
    double tt = 0.261799 * (st - 12);
    double dec = (PI * 23.5 / 180) * cos(2 * PI * ((jd - 159.5) / 365.25));
    double coh = sin(dec) * sin(latr) + cos(dec) * cos(latr) * cos(tt);
    double z = acos(coh) * (180 / PI);
    double sh = sin(dec) * sin(latr) + cos(dec) * cos(latr) * cos(tt);
    double hh = atan(sh / sqrt(1 - sh * sh));
    double sazi = cos(dec) * sin(tt) / cos(hh);
    double cazi = (sin(latr) * cos(dec) * cos(tt) - cos(latr) * sin(dec)) /
                  sqrt(pow(cos(dec) * sin(tt), 2) + pow(sin(latr) * cos(dec) * cos(tt) - cos(latr) * sin(dec), 2));
    double sqt = 1 - sazi * sazi;
    if (sqt < 0) sqt = 0;
    double azi = 180 + (180 * atan(sazi / sqrt(sqt))) / PI;
    if (cazi < 0) {
        if (sazi < 0) azi = 180 - azi;
        else azi = 540 - azi;
    }
    sol_t s;
    s.zend = z;
    s.zenr = z * TORAD;
    s.azid = azi;
    return s;
}

static double solarindex(double slope, double aspect, double zend, double azid, int shadowmask) { /* cpp:85-102 */
    double si;
    if (zend > 90.0 && !shadowmask) {
        si = 0;
    } else if (slope == 0.0) {
        si = cos(zend * TORAD);
    } else {
        si = cos(zend * TORAD) * cos(slope * TORAD) + sin(zend * TORAD) * sin(slope * TORAD) * cos((azid - aspect) * TORAD);
    }
    if (si < 0.0) si = 0.0;
    return si;
}

/* ---------------------------------------------------------------------------------------------- */
/* two-stream radiation                                                                             */
/* ---------------------------------------------------------------------------------------------- */
typedef struct { double k, kd, Kc; } kext_t;

static kext_t cank(double zenr, double x, double si) { /* cpp:104-132 */
    kext_t o;
    double k;
    if (zenr > (PI / 2.0)) zenr = PI / 2.0;
    if (si < 0.0) si = 0.0;
    if (x == 1.0) k = 1.0 / (2.0 * cos(zenr));
    else if (isinf(x)) k = 1.0;
    else if (x == 0.0) k = tan(zenr);
    else k = sqrt(x * x + (tan(zenr) * tan(zenr))) / (x + 1.774 * pow((x + 1.182), -0.733));
    if (k > 6000.0) k = 6000.0;
    o.k = k;
    o.kd = k * cos(zenr) / si;
    if (si == 0) o.kd = 1.0;
    o.Kc = 1.0 / si;
    if (si == 0.0) o.Kc = 600.0;
    return o;
}

typedef struct { /* the reference's tirstruct (+ the tsdifstruct members it copies), cpp:1034-1084 */
    double pait, om, omp, a, gma, J, del, h, u1, S1, D1, D2, p1, p2, p3, p4;
    double gi, trdn, trdu, paiaa, amx, albd, Rddn_g, Rdup_z, Rddn_z;
} tir_t;

static tir_t twostream_invariants(double pai, double paia, double x, double lref, double ltra, double clump, double gref) {
    tir_t o;
    o.pait = pai / (1.0 - clump);
    /* twostreamdifCpp, cpp:134-162 */
    o.om = lref + ltra;
    o.a = 1.0 - o.om;
    o.del = lref - ltra;
    o.J = 1.0 / 3.0;
    if (x != 1.0) {
        double mla = 9.65 * pow((3.0 + x), -1.65);
        if (mla > PI / 2.0) mla = PI / 2.0;
        o.J = cos(mla) * cos(mla);
    }
    o.gma = 0.5 * (o.om + o.J * o.del);
    o.h = sqrt(o.a * o.a + 2.0 * o.a * o.gma);
    o.S1 = exp(-o.h * o.pait);
    o.u1 = o.a + o.gma * (1.0 - 1.0 / gref);
    double u2 = o.a + o.gma * (1.0 - gref);
    o.D1 = (o.a + o.gma + o.h) * (o.u1 - o.h) * 1.0 / o.S1 - (o.a + o.gma - o.h) * (o.u1 + o.h) * o.S1;
    o.D2 = (u2 + o.h) * 1.0 / o.S1 - (u2 - o.h) * o.S1;
    o.p1 = (o.gma / (o.D1 * o.S1)) * (o.u1 - o.h);
    o.p2 = (-o.gma * o.S1 / o.D1) * (o.u1 + o.h);
    o.p3 = (1.0 / (o.D2 * o.S1)) * (u2 + o.h);
    o.p4 = (-o.S1 / o.D2) * (u2 - o.h);
    /* twostreamdif, cpp:1050-1083 */
    o.omp = 0.5 * o.om;
    o.gi = 0.0;
    if (clump > 0.0) o.gi = pow(clump, paia / pai);
    if (o.gi > 0.99) o.gi = 0.99;
    double giu = 0.0;
    if (clump > 0.0) giu = pow(clump, (pai - paia) / pai);
    if (giu > 0.99) giu = 0.99;
    double trd = o.gi * o.gi;
    o.trdn = pow(clump, 2.0);
    o.trdu = giu * giu;
    o.paiaa = paia / (1.0 - o.gi);
    o.amx = gref;
    if (o.amx < lref) o.amx = lref;
    o.albd = (1.0 - o.trdn * o.trdn) * (o.p1 + o.p2) + o.trdn * o.trdn * gref;
    if (o.albd > o.amx) o.albd = o.amx;
    if (o.albd < 0.01) o.albd = 0.01;
    o.Rddn_g = (1.0 - o.trdn) * (o.p3 * exp(-o.h * o.pait) + o.p4 * exp(o.h * o.pait)) + o.trdn;
    if (o.Rddn_g > 1.0) o.Rddn_g = 1.0;
    if (o.Rddn_g < 0.0) o.Rddn_g = 0.0;
    o.Rdup_z = (1.0 - o.trdu * o.trdn) * (o.p1 * exp(-o.h * o.paiaa) + o.p2 * exp(o.h * o.paiaa)) + o.trdu * o.trdn * gref;
    if (o.Rdup_z > 1.0) o.Rdup_z = 1.0;
    if (o.Rdup_z < 0.0) o.Rdup_z = 0.0;
    o.Rddn_z = (1.0 - trd) * (o.p3 * exp(-o.h * o.paiaa) + o.p4 * exp(o.h * o.paiaa)) + trd;
    if (o.Rddn_z > 1.0) o.Rddn_z = 1.0;
    if (o.Rddn_z < 0.0) o.Rddn_z = 0.0;
    return o;
}

typedef struct { double p5, p6, p7, p8, p9, p10, sig; } tsdir_t;

static tsdir_t twostream_direct(const tir_t* t, double gref, double kd) { /* cpp:164-185 */
    tsdir_t o;
    double a = t->a, gma = t->gma;
    double sig = kd * kd + gma * gma - pow((a + gma), 2.0);
    double ss = 0.5 * (t->om + t->J * t->del / kd) * kd;
    double sstr = t->om * kd - ss;
    double S2 = exp(-kd * t->pait);
    double u2 = a + gma * (1.0 - gref);
    o.p5 = -ss * (a + gma - kd) - gma * sstr;
    double v1 = ss - (o.p5 * (a + gma + kd)) / sig;
    double v2 = ss - gma - (o.p5 / sig) * (t->u1 + kd);
    o.p6 = (1.0 / t->D1) * ((v1 / t->S1) * (t->u1 - t->h) - (a + gma - t->h) * S2 * v2);
    o.p7 = (-1.0 / t->D1) * ((v1 * t->S1) * (t->u1 + t->h) - (a + gma + t->h) * S2 * v2);
    o.sig = -sig;
    o.p8 = sstr * (a + gma + kd) - gma * ss;
    double v3 = (sstr + gma * gref - (o.p8 / o.sig) * (u2 - kd)) * S2;
    o.p9 = (-1 / t->D2) * ((o.p8 / (o.sig * t->S1)) * (u2 + t->h) + v3);
    o.p10 = (1 / t->D2) * (((o.p8 * t->S1) / o.sig) * (u2 - t->h) + v3);
    return o;
}

typedef struct { double radGsw, radGlw, radCsw, radClw, Rbdown, Rddown, Rdup, radLsw, radLpar, lwout; } rad_t;

static rad_t twostream(double pai, double clump, double gref, double svfa, double si, double tc, double Rsw, double Rdif,
                       double Rlw, double zenr, kext_t kp, const tsdir_t* d, const tir_t* t) { /* cpp:1086-1178 */
    rad_t o;
    if (Rsw > 0.0) {
        double cosz = cos(zenr);
        if (pai > 0.0) {
            double trbn = pow(clump, kp.Kc);
            if (trbn > 0.999) trbn = 0.999;
            if (trbn < 0.0) trbn = 0.0;
            double trb = pow(t->gi, kp.Kc);
            if (trb > 0.999) trb = 0.999;
            if (trb < 0.0) trb = 0.0;
            double albb = (1.0 - t->trdn * trbn) * ((d->p5 / -d->sig) + d->p6 + d->p7) + t->trdn * trbn * gref;
            if (albb > t->amx) albb = t->amx;
            if (albb < 0.01) albb = 0.01;
            double Rdbdn_g = (1.0 - trbn) * ((d->p8 / d->sig) * exp(-kp.kd * t->pait) + d->p9 * exp(-t->h * t->pait) +
                                             d->p10 * exp(t->h * t->pait));
            if (Rdbdn_g > t->amx) Rdbdn_g = t->amx;
            if (Rdbdn_g < 0.0) Rdbdn_g = 0.0;
            double Rdbup_z = (1.0 - t->trdu * trbn) * ((d->p5 / -d->sig) * exp(-kp.kd * t->paiaa) +
                                                       d->p6 * exp(-t->h * t->paiaa) + d->p7 * exp(t->h * t->paiaa)) +
                             t->trdu * trbn * gref;
            if (Rdbup_z > t->amx) Rdbup_z = t->amx;
            if (Rdbup_z < 0.0) Rdbup_z = 0.0;
            double Rdbdn_z = (1.0 - trb) * ((d->p8 / d->sig) * exp(-kp.kd * t->paiaa) + d->p9 * exp(-t->h * t->paiaa) +
                                            d->p10 * exp(t->h * t->paiaa));
            if (Rdbdn_z > t->amx) Rdbdn_z = t->amx;
            if (Rdbdn_z < 0.0) Rdbdn_z = 0.0;
            double Rbeam = (Rsw - Rdif) / cosz;
            if (Rbeam > 1352.0) Rbeam = 1352.0;
            double Rb = Rbeam * cosz;
            double trg = trb + (1 - trb) * exp(-kp.kd * t->pait);
            double Rbc = (trg * si + (1 - trg) * cosz) * Rbeam;
            double Rbdn_g = trbn + (1.0 - trbn) * exp(-kp.kd * t->pait);
            if (Rbdn_g > 1.0) Rbdn_g = 1.0;
            if (Rbdn_g < 0.0) Rbdn_g = 0.0;
            o.radGsw = (1.0 - gref) * (t->Rddn_g * Rdif * svfa + Rdbdn_g * Rb + Rbdn_g * Rbeam * si);
            double maxg = (1.0 - gref) * (Rdif * svfa + Rbeam * si);
            if (o.radGsw > maxg) o.radGsw = maxg;
            o.radCsw = (1.0 - t->albd) * Rdif * svfa + (1.0 - albb) * Rbc;
            o.Rbdown = (trb + (1.0 - trb) * exp(-kp.kd * t->paiaa)) * Rbeam;
            o.Rddown = t->Rddn_z * Rdif * svfa + Rdbdn_z * Rb;
            o.Rdup = t->Rdup_z * Rdif * svfa + Rdbup_z * Rb;
            o.radLsw = 0.5 * (1.0 - t->om) * (o.Rddown + o.Rdup + kp.k * cosz * o.Rbdown);
            o.radLpar = 0.5 * (1.0 - t->omp) * (o.Rddown + o.Rdup + kp.k * cosz * o.Rbdown);
        } else {
            o.Rbdown = (Rsw - Rdif) / cosz;
            o.Rddown = Rdif * svfa;
            o.Rdup = gref * (Rdif * svfa + (Rsw - Rdif));
            o.radGsw = (1.0 - gref) * (svfa * Rdif + si * o.Rbdown);
            o.radCsw = o.radGsw;
            o.radLsw = 0.0;
            o.radLpar = 0.0;
        }
    } else {
        o.Rbdown = 0.0; o.Rddown = 0.0; o.Rdup = 0.0; o.radGsw = 0.0; o.radCsw = 0.0; o.radLsw = 0.0; o.radLpar = 0.0;
    }
    if (pai > 0.0) {
        double trdif = (1.0 - t->trdn) * exp(-t->pait) + t->trdn;
        o.lwout = 0.97 * SB * radem(tc);
        o.radGlw = 0.97 * (trdif * svfa * Rlw + (1.0 - trdif) * o.lwout);
        o.radClw = 0.97 * svfa * Rlw;
    } else {
        o.lwout = 0.97 * SB * radem(tc);
        o.radGlw = 0.97 * svfa * Rlw;
        o.radClw = o.radGlw;
    }
    return o;
}

/* ---------------------------------------------------------------------------------------------- */
/* wind                                                                                             */
/* ---------------------------------------------------------------------------------------------- */
typedef struct { double d, zm, a; } tiw_t;

static tiw_t wind_invariants(double h, double pai) { /* windtiCpp cpp:1179-1187, zeroplanedisCpp :294, roughlengthCpp :302 */
    tiw_t o;
    double p = pai;
    if (p < 0.001) p = 0.001;
    o.d = (1.0 - (1.0 - exp(-sqrt(7.5 * p))) / sqrt(7.5 * p)) * h;
    double Be = sqrt(0.003 + (0.2 * pai) / 2);
    double zm = (h - o.d) * exp(-KA / Be) * exp(KA * 0.0);
    if (zm > (0.9 * (h - o.d))) zm = 0.9 * (h - o.d);
    if (zm < 0.0005) zm = 0.0005;
    if (zm < 1e-6) zm = 1e-6;
    o.zm = zm;
    o.a = pai / h;
    return o;
}

typedef struct { double uf, uz, gHa; } wind_t;

static wind_t wind(double reqhgt, double zref, double h, double uref, double umu, double ws, tiw_t t) { /* cpp:1189-1218 */
    wind_t o;
    if (isnan(ws)) ws = 1.0;
    if (ws < 0.05) ws = 0.05;
    double ufs = (KA * uref) / log((zref - t.d) / t.zm);
    o.uf = ufs * umu * ws;
    if (o.uf < 0.001) o.uf = 0.001;
    o.uz = o.uf;
    if (reqhgt > 0) {
        if (reqhgt >= h) {
            o.uz = (o.uf / KA) * log((reqhgt - t.d) / t.zm);
        } else {
            double uh = (o.uf / KA) * log((h - t.d) / t.zm);
            if (uh < o.uf) uh = o.uf;
            double Be = o.uf / uh;
            if (Be < 0.001) Be = 0.001;
            double Lc = pow(0.25 * t.a, -1.0);
            double Lm = 2 * pow(Be, 3.0) * Lc;
            o.uz = uh * exp(Be * (reqhgt - h) / Lm);
        }
        if (o.uz > uref) o.uz = uref;
    }
    /* gturbCpp(uf, d, zm, zref, 43, 0, 0.0001), cpp:373-380 */
    double z0 = 0.2 * t.zm + t.d;
    double ln = log((zref - t.d) / (z0 - t.d));
    double g = (KA * 43 * o.uf) / (ln + 0);
    if (g < 0.0001) g = 0.0001;
    o.gHa = g;
    return o;
}

/* ---------------------------------------------------------------------------------------------- */
/* energy balance                                                                                   */
/* ---------------------------------------------------------------------------------------------- */
static double satvap(double tc) { /* cpp:480-490 */
    if (tc > 0) return 0.61078 * exp(17.27 * tc / (tc + 237.3));
    return 0.61078 * exp(21.875 * tc / (tc + 265.5));
}

typedef struct { double Ts, H, L, Rem, mu; } pm_t;

static pm_t penman_monteith(double Rabs, double gHa, double gV, double tc, double mxtc, double pk, double ea, double es,
                            double G, double surfwet, double tdew) { /* PenmanMonteith2Cpp cpp:1220-1247 */
    double De = satvap(tc + 0.5) - satvap(tc - 0.5);
    double gHr = gHa + (4 * 0.97 * SB * pow(tc + 273.15, 3.0)) / 29.3;
    double Rem = 0.97 * SB * radem(tc);
    double la;
    if (tc >= 0) la = 45068.7 - 42.8428 * tc;
    else la = 51078.69 - 4.338 * tc - 0.06367 * tc * tc;
    double m = la * (gV / pk);
    double L = m * (es - ea) * surfwet;
    double dT = (Rabs - Rem - L - G) / (29.3 * gHr + m * De);
    double dTmx = -0.6273 * mxtc + 49.79;
    if (dT > dTmx) dT = dTmx;
    if (dT > 80.0) dT = 80.0;
    pm_t o;
    o.Ts = dT + tc;
    if (o.Ts < tdew) o.Ts = tdew;
    o.H = 29.3 * gHa * (o.Ts - tc);
    o.L = m * (satvap(o.Ts) - ea) * surfwet;
    o.Rem = 0.97 * SB * radem(o.Ts);
    o.mu = la * (43.0 / pk);
    return o;
}

typedef struct { double c1, c3, c4; } soilk_t;
typedef struct { double Smax, Smin, soilb, psi_e, rho; } soilp_t;

static soilk_t soil_constants(double Vm, double Vq, double Mc) { /* soilpfun cpp:628-636 */
    soilk_t o;
    double frs = Vm + Vq;
    o.c1 = (0.57 + 1.73 * Vq + 0.93 * Vm) / (1.0 - 0.74 * Vq - 0.49 * Vm) - 2.8 * frs * (1.0 - frs);
    o.c3 = 1.0 + 2.6 * pow(Mc, -0.5);
    o.c4 = 0.03 + 0.7 * frs * frs;
    return o;
}

static double soil_distribute(double soilm, double Smin, double Smax, double tadd) { /* soildCpp cpp:1021-1032 */
    double rge = Smax - Smin;
    double theta = (soilm - Smin) / rge;
    if (theta > 0.9999) theta = 0.9999;
    if (theta < 0.0001) theta = 0.0001;
    double lt = log(theta / (1 - theta));
    double sm = lt + tadd;
    sm = 1 / (1 + exp(-sm));
    return sm * rge + Smin;
}

/* ---------------------------------------------------------------------------------------------- */
/* vegetation conductances, leaf and air temperature                                               */
/* ---------------------------------------------------------------------------------------------- */
typedef struct { double Rsmx, psiw0, kk, rat; } stomp_t;

static stomp_t stomatal_class(double hgt, double lat, double x) { /* stomparamsCpp cpp:391-440 */
    stomp_t o = {420.0, -3.1, 0.34, 0.9};
    if (hgt < 1.0 && fabs(lat) < 22.5) { o.Rsmx = 450.0; o.psiw0 = -2.7; o.kk = 0.39; o.rat = 0.9; }
    if (hgt >= 1.0 && hgt < 7.0) { o.Rsmx = 430.0; o.psiw0 = -4.0; o.kk = 0.28; o.rat = 0.75; }
    if (hgt >= 7.0) {
        if (fabs(lat) < 22.5) { o.Rsmx = 500.0; o.psiw0 = -1.75; o.kk = 0.67; o.rat = 0.4; }
        else if (x < 0.8 || fabs(lat) > 58.0) { o.Rsmx = 420.0; o.psiw0 = -4.09; o.kk = 0.29; o.rat = 0.6; }
        else { o.Rsmx = 500.0; o.psiw0 = -2.51; o.kk = 0.46; o.rat = 0.45; }
    }
    return o;
}

static double stomcond(double Rswabs, double theta, double gsmax, double Smax, double psi_e, double b, stomp_t s) {
    /* stomcondCpp cpp:442-458 with psiwfromthetaCpp cpp:382-389 */
    if (Rswabs <= 0.0) return 0.0;
    if (Rswabs > s.Rsmx) Rswabs = s.Rsmx;
    double gs = gsmax * pow(2.0, -(s.Rsmx - Rswabs) / (0.2 * s.Rsmx));
    double thetan = s.rat * theta + (1 - s.rat) * THETAM;
    psi_e = fabs(psi_e);
    double Se = thetan / Smax;
    if (Se > 1.0) Se = 1.0;
    double psiw = -psi_e * pow(Se, -b) * 0.01;
    if (psiw < s.psiw0) psiw = s.psiw0;
    double mu = 1.0 - (exp(-s.kk * psiw) - 1.0) / (exp(-s.kk * s.psiw0) - 1.0);
    double gs2 = mu * gsmax;
    if (gs > gs2) gs = gs2;
    return gs;
}

static double canopycond(double Rsw, double Rdif, double k, double om, double theta, double gsmax, double PAI,
                         double Smax, double psi_e, double b, stomp_t s) { /* cpp:460-477 */
    double Gs = 9999.99;
    if (!isnan(om)) {
        double P_sun = (1.0 - exp(-k * PAI)) / k;
        double P_shade = PAI - P_sun;
        double Rshade_abs = Rdif * ((1.0 - exp(-PAI)) / PAI) * (1.0 - om);
        double Rsun_abs = (Rsw - Rdif) * k * (1 - om) + Rshade_abs;
        double gs_sun = stomcond(Rsun_abs, theta, gsmax, Smax, psi_e, b, s);
        double gs_shade = stomcond(Rshade_abs, theta, gsmax, Smax, psi_e, b, s);
        Gs = gs_sun * P_sun + gs_shade * P_shade;
    }
    return Gs;
}

static void tv_above(double reqhgt, double zref, double d, double zm, double T0, double tc, double ea, double surfwet,
                     double* Tz, double* ez) { /* TVabove cpp:1298-1313 */
    double zh = 0.2 * zm;
    double estl = satvap(T0);
    if (reqhgt > (d + zh)) {
        double lnr = log((reqhgt - d) / zh) / log((zref - d) / zh);
        *Tz = tc + (T0 - tc) * (1 - lnr);
        *ez = ea + (estl - ea) * surfwet * (1 - lnr);
    } else {
        *Tz = T0;
        *ez = ea + (estl - ea) * surfwet;
    }
}

static double mincond(double leafabs, double gs, double tc, double leafd) { /* mincondCpp cpp:1316-1331 */
    double Rnet = leafabs - 0.97 * SB * radem(tc);
    double rs = 500.0;
    if (gs > 0.0) rs = 1 / gs;
    if (rs > 500.0) rs = 500.0;
    double Hlf = 1.09767 * pow(rs, 0.2672778);
    double Hf = -1.0 / (1.0 + exp(2.0 - Hlf));
    double H = Hf * Rnet;
    double gmin = 0.0463 * pow(fabs(H) / leafd, 0.2);
    if (gmin < 0.05) gmin = 0.05;
    return gmin;
}

typedef struct { double tleaf, H, L, lwup, lwdn; } leaf_t;

static leaf_t leaftemp(double Tcan, double Tg, double tc, double mxtc, double pk, double ea, double es, double uz,
                       double tdew, double surfwet, double radLsw, double Rlw, double pai, double paia, double leafd,
                       double gsmax, double PARabs, double theta, double Smax, double psi_e, double soilb, stomp_t s) {
    /* cpp:1333-1364 */
    leaf_t o;
    double lwcan = 0.97 * SB * radem(Tcan);
    double lwgro = 0.97 * SB * radem(Tg);
    double paig = pai - paia;
    o.lwup = exp(-paig) * lwgro + (1 - exp(-paig)) * lwcan;
    o.lwdn = exp(-paia) * Rlw + (1 - exp(-paia)) * lwcan;
    double lwabs = 0.97 * 0.5 * (o.lwup + o.lwdn);
    double leafabs = radLsw + lwabs;
    double gh = 0.135 * sqrt(uz / leafd) * 1.4;
    double gmin = mincond(leafabs, 999.99, Tcan, leafd);
    if (gh < gmin) gh = gmin;
    double gV = gh;
    if (gsmax < 999.99) {
        gV = 0.0;
        double gs = stomcond(PARabs, theta, gsmax, Smax, psi_e, soilb, s);
        gmin = mincond(leafabs, gs, Tcan, leafd);
        if (gh < gmin) gh = gmin;
        if (gs > 0.0) gV = 1 / (1 / gh + 1 / gs);
    }
    pm_t pm = penman_monteith(leafabs, gh, gV, tc, mxtc, pk, ea, es, 0.0, surfwet, tdew);
    o.tleaf = pm.Ts;
    o.H = pm.H;
    o.L = pm.L;
    return o;
}

static double rhcanopy(double uf, double h, double d, double z) { /* cpp:1365-1380 */
    double a2 = 0.4 * (1.0 - (d / h)) / pow(1.25, 2);
    double inth = 4.293251 * h;
    if (z != h) {
        inth = (2.0 * h * ((48 * atan((sqrt(5.0) * sin((PI * z) / h)) / (cos((PI * z) / h) + 1))) / pow(5.0, 1.5) +
                           (32.0 * sin((PI * z) / h)) / ((cos((PI * z) / h) + 1) *
                                                         ((25.0 * pow(sin((PI * z) / h), 2.0)) / pow((cos((PI * z) / h) + 1.0), 2.0) + 5.0)))) / PI;
    }
    double mu = uf / (a2 * h) * 1.0 / (uf * uf);
    double rHa = inth * mu;
    if (rHa < 0.001) rHa = 0.001;
    return rHa;
}

static double tv_below(double z, double d, double h, double pai, double uf, double leafden, double Flux, double Fluxz,
                       double SH, double SG, double mxnear) { /* TVbelow cpp:1381-1409 */
    double Rc = rhcanopy(uf, h, d, h);
    double Kc = h / Rc;
    double Kg = 1.0 / rhcanopy(uf, h, d, z);
    double Kh = 1.0 / (Rc - rhcanopy(uf, h, d, z));
    Kg = Kg / z;
    Kh = Kh / (h - z);
    double SC = SH + Flux / Kc;
    double farg = (Kg * SG + Kh * SH + Kc * SC) / (Kg + Kh + Kc);
    double SN = Fluxz * leafden;
    double near = (3.047519 + 0.128642 * log(pai)) * SN;
    if (fabs(near) > mxnear) {
        if (near > 0.0) near = mxnear;
        else near = -mxnear;
    }
    if (isnan(near)) near = 0;
    return near + farg;
}

static double max4(double a, double b, double c, double d) { /* std::max({a,b,c,d}): left fold with operator< */
    double m = a;
    if (m < b) m = b;
    if (m < c) m = c;
    if (m < d) m = d;
    return m;
}
static double min4(double a, double b, double c, double d) {
    double m = a;
    if (b < m) m = b;
    if (c < m) m = c;
    if (d < m) m = d;
    return m;
}

typedef struct { double Tz, tleaf, rh, lwdn, lwup; } above_t;

static above_t tv_aboveground(double reqhgt, double zref, double tc, double pk, double ea, double es, double tdew,
                              double Rsw, double Rdif, double Rlw, double soilm, double hgt, double pai, double paia,
                              double vegx, double leafd, double leafden, double Smin, double Smax, double psi_e,
                              double soilb, double gsmax, double mxtc, stomp_t stomp, const tir_t* tir, double zend,
                              double radCsw, double radClw, double radLsw, double radLpar, tiw_t tiw, wind_t w,
                              double Tg, double G) { /* TVaboveground cpp:1411-1472 */
    above_t o;
    double eT = satvap(Tg) - ea;
    if (eT < 0.001) eT = 0.001;
    double plf = 0.8753 - 1.7126 * log(eT);
    double gwet = 1.0 / (1.0 + exp(-plf));
    double surfwet = (soilm - Smin) / (Smax - Smin);
    if (surfwet > gwet) gwet = surfwet;
    /* the reference passes the zenith in DEGREES where cankCpp expects radians (cpp:1425); reproduced */
    kext_t kp = cank(zend, vegx, cos(zend * TORAD));
    double gS = canopycond(Rsw, Rdif, kp.k, tir->omp, soilm, gsmax, pai, Smax, psi_e, soilb, stomp);
    double gV = 0.0;
    if (gS > 0.0) gV = 1.0 / (1.0 / w.gHa + 1 / gS);
    double Rabs = radCsw + radClw;
    pm_t pm = penman_monteith(Rabs, w.gHa, gV, tc, mxtc, pk, ea, es, G, surfwet, tdew);
    double Tcan = pm.Ts;
    double ez = 0;
    if (reqhgt >= hgt) {
        double tz, e;
        tv_above(reqhgt, zref, tiw.d, tiw.zm, Tcan, tc, ea, surfwet, &tz, &e);
        o.Tz = tz;
        o.tleaf = Tcan;
        o.lwup = 0.97 * SB * radem(Tcan);
        o.lwdn = Rlw;
        ez = e;
    } else {
        leaf_t lf = leaftemp(Tcan, Tg, tc, mxtc, pk, ea, es, w.uz, tdew, surfwet, radLsw, Rlw, pai, paia, leafd, gsmax,
                             radLpar, soilm, Smax, psi_e, soilb, stomp);
        o.tleaf = lf.tleaf;
        double Flux = pm.H * (1.0 - exp(-pai));
        double Fluxz = lf.H;
        double Th, eh;
        tv_above(hgt, zref, tiw.d, tiw.zm, Tcan, tc, ea, surfwet, &Th, &eh);
        double SH = Th * 29.3 * 43.0;
        double SG = Tg * 29.3 * 43.0;
        double mxnear = fabs(o.tleaf - Th) * 29.3 * 43.0;
        o.Tz = tv_below(reqhgt, tiw.d, hgt, pai, w.uf, leafden, Flux, Fluxz, SH, SG, mxnear) / (29.3 * 43.0);
        Flux = pm.L * (1.0 - exp(-pai));
        Fluxz = lf.L;
        SH = eh * pm.mu;
        SG = satvap(Tg) * gwet * pm.mu;
        mxnear = fabs(satvap(o.tleaf) - eh) * pm.mu;
        ez = tv_below(reqhgt, tiw.d, hgt, pai, w.uf, leafden, Flux, Fluxz, SH, SG, mxnear) / pm.mu;
        o.lwdn = lf.lwdn;
        o.lwup = lf.lwup;
    }
    o.rh = (ez / satvap(o.Tz)) * 100.0;
    if (o.rh > 100.0) o.rh = 100.0;
    double tmx = max4(o.tleaf, tc, Tg, Tcan) + 2.0;
    double tmn = min4(o.tleaf, tc, Tg, Tcan) - 2.0;
    if (o.Tz > tmx) o.Tz = tmx;
    if (o.Tz < tmn) o.Tz = tmn;
    return o;
}

/* ---------------------------------------------------------------------------------------------- */
/* below-ground temperature                                                                         */
/* ---------------------------------------------------------------------------------------------- */
static void rolling_mean(const double* x, int m, int n, double* y) { /* maCpp cpp:561-572 */
    for (int i = 0; i < m; ++i) {
        double sum = 0.0;
        for (int j = 0; j < n; ++j) sum += x[(i - j + m) % m];
        y[i] = sum / n;
    }
}

static void rolling_mean_n(const double* x, int m, int n, double* z) { /* manCpp cpp:597-627 */
    if (n <= 48) {
        rolling_mean(x, m, n, z);
        return;
    }
    int numDays = m / 24;
    double* d = (double*)calloc((size_t)(numDays > 0 ? numDays : 1), sizeof(double));
    double* y = (double*)calloc((size_t)(numDays > 0 ? numDays : 1), sizeof(double));
    double* e = (double*)calloc((size_t)m, sizeof(double)); /* zero-initialised like std::vector<double> z(x.size()) */
    for (int i = 0; i < numDays; ++i) {
        double sum = 0.0;
        for (int j = 0; j < 24; ++j) sum += x[i * 24 + j];
        d[i] = sum / 24.0;
    }
    int n2 = n / 24;
    rolling_mean(d, numDays, n2, y);
    for (int i = 0; i < numDays; ++i)
        for (int j = 0; j < 24; ++j) e[i * 24 + j] = y[i];
    rolling_mean(e, m, 24, z);
    free(d);
    free(y);
    free(e);
}

/* hourtodayCpp(..., rephour = true) cpp:517-559: stat 0 = max, 1 = min, 2 = mean; trailing hours stay 0 */
static void hour_to_day(const double* hourly, int m, int stat, double* daily) {
    int numDays = m / 24;
    memset(daily, 0, (size_t)m * sizeof(double));
    for (int i = 0; i < numDays; ++i) {
        double v = hourly[i * 24];
        if (stat == 0) {
            for (int j = 1; j < 24; ++j) v = (v < hourly[i * 24 + j]) ? hourly[i * 24 + j] : v;
        } else if (stat == 1) {
            for (int j = 1; j < 24; ++j) v = (hourly[i * 24 + j] < v) ? hourly[i * 24 + j] : v;
        } else {
            v = 0.0;
            for (int j = 0; j < 24; ++j) v += hourly[i * 24 + j];
            v /= 24;
        }
        for (int j = 0; j < 24; ++j) daily[i * 24 + j] = v;
    }
}

static void below_ground(double reqhgt, const double* Tg, const double* Tgp, const double* Tbp, int T, double meanD,
                         double mat, int hiy, int complete, double* Tz) { /* Tbelowgroundv cpp:1474-1539 */
    memcpy(Tz, Tg, (size_t)T * sizeof(double));
    if (!(reqhgt < 0)) return;
    double nb = -118.35 * reqhgt / meanD;
    int n = (int)round(nb);
    if (complete) {
        if (n < T) {
            rolling_mean_n(Tg, T, n, Tz);
        } else {
            double sumT = 0;
            for (int i = 0; i < T; ++i) sumT = sumT + Tg[i];
            double meanT = sumT / T;
            for (int i = 0; i < T; ++i) Tz[i] = meanT;
        }
        return;
    }
    double* w = (double*)calloc((size_t)8 * T, sizeof(double));
    double *Tzd = w, *Tbpd = w + T, *gmx = w + 2 * T, *gmn = w + 3 * T, *gme = w + 4 * T, *pmx = w + 5 * T, *pmn = w + 6 * T,
           *pme = w + 7 * T;
    hour_to_day(Tbp, T, 2, Tbpd);
    hour_to_day(Tg, T, 0, gmx);
    hour_to_day(Tg, T, 1, gmn);
    hour_to_day(Tg, T, 2, gme);
    hour_to_day(Tgp, T, 0, pmx);
    hour_to_day(Tgp, T, 1, pmn);
    hour_to_day(Tgp, T, 2, pme);
    for (int i = 0; i < T; ++i) {
        double Tbpa = Tbp[i] - Tbpd[i];
        double rat = (gmx[i] - gmn[i]) / (pmx[i] - pmn[i]);
        double dif = gme[i] - pme[i];
        Tzd[i] = rat * Tbpa + Tbpd[i] + dif;
    }
    if (nb > 1.0 && nb <= 24.0) {
        double w1 = 1.0 / nb, w2 = nb / 24.0;
        double wgt = w1 / (w1 + w2);
        for (int i = 0; i < T; ++i) Tz[i] = wgt * Tg[i] + (1 - wgt) * Tzd[i];
    }
    if (nb > 24.0) {
        if (nb < hiy) {
            double w1 = 24.0 / nb, w2 = nb / hiy;
            double wgt = w1 / (w1 + w2);
            for (int i = 0; i < T; ++i) Tz[i] = wgt * Tzd[i] + (1 - wgt) * mat;
        } else {
            for (int i = 0; i < T; ++i) Tz[i] = mat;
        }
    }
    free(w);
}

/* ---------------------------------------------------------------------------------------------- */
/* the grid drivers: runmicro1Cpp cpp:2052, runmicro2Cpp :2340, runmicro3Cpp :2624, runmicro4Cpp :2926  */
/* ---------------------------------------------------------------------------------------------- */
static int fail(char* err, size_t errlen, const char* msg, int code) {
    if (err && errlen) snprintf(err, errlen, "%s", msg);
    return code;
}

int oracle_runmicro(const mcf_problem* p, double* const out[MCF_NOUT], char* err, size_t errlen) {
    if (!p || !out) return fail(err, errlen, "oracle_runmicro: NULL argument", MCF_ERR_ARG);
    if (p->mode < 1 || p->mode > 4) return fail(err, errlen, "oracle_runmicro: mode must be 1..4", MCF_ERR_ARG);
    const int arr = (p->mode == 2 || p->mode == 4), layered = (p->mode >= 3);
    const int R = p->rows, C = p->cols, T = p->tsteps;
    const size_t nc = (size_t)R * C;
    const int nlyr = layered ? p->nlyr : 1;
    int* lst = (int*)calloc((size_t)nlyr, sizeof(int));
    int* lnd = (int*)calloc((size_t)nlyr, sizeof(int));
    if (layered) {
        for (int l = 0; l < nlyr; ++l) { /* cpp:2633-2639 */
            int span = p->lyr_ed[l] - p->lyr_st[l] + 1;
            if (span < 24) {
                free(lst);
                free(lnd);
                return fail(err, errlen, "Too many layers in vegp. Max layers must be <= max days", MCF_ERR_ARG);
            }
            lst[l] = p->lyr_st[l];
            lnd[l] = span / 24;
        }
    } else {
        lst[0] = 0;
        lnd[0] = T / 24; /* cpp:2116 */
    }
    const double NA = na_real();
    for (int v = 0; v < MCF_NOUT; ++v)
        if (out[v])
            for (size_t i = 0; i < nc * (size_t)T; ++i) out[v][i] = NA; /* NumericVector(n, NA_REAL), cpp:2131-2140 */

    /* per-hour tables: solar position (modes 1/3), sector indices, series maximum of tc (cpp:2153-2169) */
    double* zend = (double*)calloc((size_t)T, sizeof(double));
    double* zenr = (double*)calloc((size_t)T, sizeof(double));
    double* azid = (double*)calloc((size_t)T, sizeof(double));
    int* sindex = (int*)calloc((size_t)T, sizeof(int));
    int* windex = (int*)calloc((size_t)T, sizeof(int));
    double mxtc_series = -273.15;
    for (int k = 0; k < T; ++k) {
        if (!arr) {
            sol_t s = solposition(p->lat, p->lon, p->year[k], p->month[k], p->day[k], p->hour[k]);
            zend[k] = s.zend;
            zenr[k] = s.zenr;
            azid[k] = s.azid;
            sindex[k] = ((int)round(s.azid / 15.0)) % 24;
            if (p->temp[k] > mxtc_series) mxtc_series = p->temp[k];
        }
        windex[k] = ((int)round(p->winddir[k] / 45)) % 8;
    }
    int hiy = 365 * 24;
    if (p->year[0] % 4 == 0) hiy = 366 * 24; /* cpp:2171-2172 */
    /* soildCppm cpp:975-1019: tadd = log(twi)/tfact - mean over the non-NA cells */
    double me;
    if (p->has_twi_mean) {
        me = p->twi_mean;
    } else {
        double sum = 0.0;
        long count = 0;
        for (size_t c = 0; c < nc; ++c) {
            if (!isnan(p->twi[c])) {
                double l = log(p->twi[c]) / p->tfact;
                if (!isnan(l)) {
                    sum += l;
                    count++;
                }
            }
        }
        me = sum / count;
    }
    double* Tg = (double*)calloc((size_t)T, sizeof(double));
    double* DD = (double*)calloc((size_t)T, sizeof(double));
    double* Tzv = (double*)calloc((size_t)T, sizeof(double));
    double* Tgpv = (double*)calloc((size_t)T, sizeof(double));
    double* Tbpv = (double*)calloc((size_t)T, sizeof(double));
    double reqhgt2 = p->reqhgt;
    if (reqhgt2 < 0.00001) reqhgt2 = 0.00001; /* cpp:2246-2247 */

    /* the reference iterates rows outermost (cpp:2180-2181); cells are independent, so order is immaterial */
    for (int j = 0; j < C; ++j) {
        for (int i = 0; i < R; ++i) {
            const size_t c = (size_t)i + (size_t)R * j;
            if (isnan(p->hgt[c])) continue; /* first layer's height decides (cpp:2182, :2765) */
            const double tadd = isnan(p->twi[c]) ? NA : log(p->twi[c]) / p->tfact - me;
            soilp_t spa = {p->Smax[c], p->Smin[c], p->soilb[c], p->Psie[c], p->rho[c]};
            soilk_t sk = soil_constants(p->Vm[c], p->Vq[c], p->Mc[c]);
            const double lat = arr ? p->lats[c] : p->lat;
            const double lon = arr ? p->lons[c] : p->lon;
            double mxtc = mxtc_series;
            if (arr) { /* cpp:2467-2471 */
                mxtc = -273.15;
                for (int k = 0; k < T; ++k)
                    if (p->temp[c + nc * k] > mxtc) mxtc = p->temp[c + nc * k];
            }
            memset(Tg, 0, (size_t)T * sizeof(double));
            memset(DD, 0, (size_t)T * sizeof(double));
            for (int l = 0; l < nlyr; ++l) {
                const size_t cl = c + nc * l;
                const double hgt = p->hgt[cl], pai = p->pai[cl], x = p->x[cl], gref = p->gref[c];
                tir_t tir = twostream_invariants(pai, p->paia[cl], x, p->leafr[cl], p->leaft[cl], p->clump[cl], gref);
                stomp_t stomp = stomatal_class(hgt, lat, x);
                tiw_t tiw = wind_invariants(hgt, pai);
                for (int dy = 0; dy < lnd[l]; ++dy) {
                    double Rmx = -999.9, tmx = -999.0, tmn = 999.0; /* cpp:2196-2198 */
                    double surfwet[24], radabs[24], soilmday[24], radCsw[24], radClw[24], radLsw[24], radLpar[24];
                    double uf[24], uzday[24], gHa[24], zendday[24];
                    for (int hr = 0; hr < 24; ++hr) { /* pass 1, cpp:2214-2262 */
                        const int k = dy * 24 + hr + lst[l];
                        const size_t kc = arr ? c + nc * k : (size_t)k;
                        const size_t idx = c + nc * k;
                        double zd, zr, az;
                        int sidx;
                        double si;
                        if (arr) { /* cpp:2497-2503: shadowmask defaults to false */
                            sol_t s = solposition(lat, lon, p->year[k], p->month[k], p->day[k], p->hour[k]);
                            zd = s.zend; zr = s.zenr; az = s.azid;
                            si = solarindex(p->slope[c], p->aspect[c], zd, az, 0);
                            sidx = ((int)round(az / 15)) % 24;
                        } else { /* cpp:2218 */
                            zd = zend[k]; zr = zenr[k]; az = azid[k];
                            si = solarindex(p->slope[c], p->aspect[c], zd, az, 1);
                            sidx = sindex[k];
                        }
                        zendday[hr] = zd;
                        if (si < 0.0) si = 0.0;
                        double ws = p->wsa[(size_t)windex[k] * nc + c];
                        double ha = p->hor[(size_t)sidx * nc + c];
                        double sa = (PI / 2.0) - zr;
                        if (ha > tan(sa)) si = 0.0;
                        double soild = soil_distribute(p->p_soilm[kc], p->Smin[c], p->Smax[c], tadd);
                        soilmday[hr] = soild;
                        if (out[MCF_OUT_SOILM]) out[MCF_OUT_SOILM][idx] = soild;
                        kext_t kpp = cank(zr, x, si);
                        tsdir_t dir = twostream_direct(&tir, gref, kpp.kd);
                        rad_t rad = twostream(pai, p->clump[cl], gref, p->svfa[c], si, p->temp[kc], p->swdown[kc],
                                              p->difrad[kc], p->lwdown[kc], zr, kpp, &dir, &tir);
                        radCsw[hr] = rad.radCsw; radClw[hr] = rad.radClw; radLsw[hr] = rad.radLsw; radLpar[hr] = rad.radLpar;
                        if (out[MCF_OUT_RDIRDOWN]) out[MCF_OUT_RDIRDOWN][idx] = rad.Rbdown;
                        if (out[MCF_OUT_RDIFDOWN]) out[MCF_OUT_RDIFDOWN][idx] = rad.Rddown;
                        if (out[MCF_OUT_RSWUP]) out[MCF_OUT_RSWUP][idx] = rad.Rdup;
                        wind_t wm = wind(reqhgt2, p->zref, hgt, p->windspeed[kc], p->p_umu[kc], ws, tiw);
                        uf[hr] = wm.uf; uzday[hr] = wm.uz; gHa[hr] = wm.gHa;
                        if (out[MCF_OUT_WINDSPEED]) out[MCF_OUT_WINDSPEED][idx] = wm.uz;
                        /* soiltempG0 cpp:1262-1275 */
                        double rabs = rad.radGsw + rad.radGlw;
                        double matric = -fabs(spa.psi_e) * pow(soild / spa.Smax, -spa.soilb);
                        double sw = exp((0.018 * matric) / (8.31 * (p->temp[kc] + 273.15)));
                        if (sw > 1.0) sw = 1.0;
                        pm_t pm0 = penman_monteith(rabs, wm.gHa, wm.gHa, p->temp[kc], mxtc, p->pres[kc], p->ea[kc],
                                                   p->es[kc], 0.0, sw, p->tdew[kc]);
                        double Rval = fabs(rabs - pm0.Rem);
                        if (Rmx < Rval) Rmx = Rval;
                        if (tmx < pm0.Ts) tmx = pm0.Ts;
                        if (tmn > pm0.Ts) tmn = pm0.Ts;
                        surfwet[hr] = sw;
                        radabs[hr] = rabs;
                    }
                    double dtr = tmx - tmn;
                    for (int hr = 0; hr < 24; ++hr) { /* pass 2, cpp:2264-2305 */
                        const int k = dy * 24 + hr + lst[l];
                        const size_t kc = arr ? c + nc * k : (size_t)k;
                        const size_t idx = c + nc * k;
                        /* soiltemp_hrCpp cpp:1277-1296 with soilcondCpp cpp:1249-1260 */
                        double sm = soilmday[hr];
                        double cs = (2400 * spa.rho / 2.64 + 4180.0 * sm);
                        double ph = (spa.rho * (1.0 - sm) + sm) * 1000.0;
                        double c2 = 1.06 * spa.rho * sm;
                        double kk = sk.c1 + c2 * sm - (sk.c1 - sk.c4) * exp(-pow(sk.c3 * sm, 4.0));
                        double kap = kk / (cs * ph);
                        double dd = pow(2.0 * kap / OMDY, 0.5);
                        double dtR = dtr / p->p_dtrp[kc];
                        double Gmu = dtR * (kk * p->p_muGp[kc]) / (p->p_kp[kc] * dd);
                        double G = p->p_G[kc] * Gmu;
                        if (G > 0.6 * Rmx) G = 0.6 * Rmx;
                        if (G < -0.6 * Rmx) G = -0.6 * Rmx;
                        pm_t pmg = penman_monteith(radabs[hr], gHa[hr], gHa[hr], p->temp[kc], mxtc, p->pres[kc], p->ea[kc],
                                                   p->es[kc], G, surfwet[hr], p->tdew[kc]);
                        Tg[k] = pmg.Ts;
                        DD[k] = dd;
                        if (p->reqhgt >= 0.0) {
                            wind_t wv = {uf[hr], uzday[hr], gHa[hr]};
                            above_t tv = tv_aboveground(reqhgt2, p->zref, p->temp[kc], p->pres[kc], p->ea[kc], p->es[kc],
                                                        p->tdew[kc], p->swdown[kc], p->difrad[kc], p->lwdown[kc], sm, hgt, pai,
                                                        p->paia[cl], x, p->leafd[cl], p->leafden[cl], p->Smin[c], p->Smax[c],
                                                        p->Psie[c], p->soilb[c], p->gsmax[cl], mxtc, stomp, &tir, zendday[hr],
                                                        radCsw[hr], radClw[hr], radLsw[hr], radLpar[hr], tiw, wv, pmg.Ts, G);
                            if (out[MCF_OUT_TZ]) out[MCF_OUT_TZ][idx] = (p->reqhgt > 0.0) ? tv.Tz : Tg[k];
                            if (out[MCF_OUT_RLWDOWN]) out[MCF_OUT_RLWDOWN][idx] = tv.lwdn;
                            if (out[MCF_OUT_RLWUP]) out[MCF_OUT_RLWUP][idx] = tv.lwup;
                            if (p->reqhgt > 0.0) {
                                if (out[MCF_OUT_TLEAF]) out[MCF_OUT_TLEAF][idx] = tv.tleaf;
                                if (out[MCF_OUT_RELHUM]) out[MCF_OUT_RELHUM][idx] = tv.rh;
                            }
                        }
                    }
                }
            }
            if (p->reqhgt < 0.0 && out[MCF_OUT_TZ]) { /* cpp:2307-2320 */
                double sumD = 0.0;
                for (int k = 0; k < T; ++k) sumD += DD[k];
                double meanD = sumD / (double)T;
                const double *tgp = p->p_Tg, *tbp = p->p_Tbp;
                if (arr) {
                    for (int k = 0; k < T; ++k) {
                        Tgpv[k] = p->p_Tg ? p->p_Tg[c + nc * k] : 0.0;
                        Tbpv[k] = p->p_Tbp ? p->p_Tbp[c + nc * k] : 0.0;
                    }
                    tgp = Tgpv;
                    tbp = Tbpv;
                } else if (!tgp || !tbp) {
                    tgp = Tgpv; /* zero-filled: only read by the incomplete-series branch */
                    tbp = Tbpv;
                }
                below_ground(p->reqhgt, Tg, tgp, tbp, T, meanD, p->mat, hiy, p->complete, Tzv);
                for (int k = 0; k < T; ++k) out[MCF_OUT_TZ][c + nc * k] = Tzv[k];
            }
        }
    }
    free(lst); free(lnd); free(zend); free(zenr); free(azid); free(sindex); free(windex);
    free(Tg); free(DD); free(Tzv); free(Tgpv); free(Tbpv);
    return MCF_OK;
}

/* ---------------------------------------------------------------------------------------------- */
/* bioclim: runbioclimCpp cpp:3457-3560 over bioclim1..19 cpp:3245-3448; wrappers cpp:3563-3700     */
/* ---------------------------------------------------------------------------------------------- */
static double std_dev(const double* v, int n) { /* calc_std_dev cpp:3227-3244 */
    if (n <= 1) return na_real();
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += v[i];
    double mean = s / n;
    double ss = 0.0;
    for (int i = 0; i < n; ++i) ss += pow(v[i] - mean, 2.0);
    return sqrt(ss / (n - 1));
}

static double quarter_mean(const double* v, const int32_t* q, int nq) { /* always / 72, cpp:3317-3360 */
    double o = 0.0;
    for (int i = 0; i < nq; ++i) o = o + v[q[i]];
    return o / 72.0;
}

int oracle_runbioclim(const mcf_problem* p, const int32_t* wetq, int32_t nwetq, const int32_t* dryq, int32_t ndryq,
                      const int32_t* hotq, int32_t nhotq, const int32_t* colq, int32_t ncolq, int32_t air,
                      double* const bio[MCF_NBIO], char* err, size_t errlen) {
    if (!p || !bio) return fail(err, errlen, "oracle_runbioclim: NULL argument", MCF_ERR_ARG);
    if (p->tsteps < 336) return fail(err, errlen, "oracle_runbioclim: needs 336 hours", MCF_ERR_ARG);
    mcf_problem pp = *p;
    pp.complete = 1; /* cpp:3576-3577 */
    int32_t st14[14], ed14[14];
    if (p->mode >= 3) { /* cpp:3635-3646 */
        for (int i = 0; i < 14; ++i) { st14[i] = i * 24; ed14[i] = i * 24 + 23; }
        pp.nlyr = 14;
        pp.lyr_st = st14;
        pp.lyr_ed = ed14;
    }
    const int R = p->rows, C = p->cols, T = p->tsteps;
    const size_t nc = (size_t)R * C;
    double* Tz = (double*)malloc(nc * T * sizeof(double));
    double* sm = (double*)malloc(nc * T * sizeof(double));
    double* outm[MCF_NOUT] = {0};
    outm[air ? MCF_OUT_TZ : MCF_OUT_TLEAF] = Tz; /* cpp:3568-3573 */
    outm[MCF_OUT_SOILM] = sm;
    int rc = oracle_runmicro(&pp, outm, err, errlen);
    if (rc != MCF_OK) {
        free(Tz);
        free(sm);
        return rc;
    }
    const double NA = na_real();
    for (int b = 0; b < MCF_NBIO; ++b)
        if (bio[b])
            for (size_t c = 0; c < nc; ++c) bio[b][c] = NA; /* bioclimfill cpp:3450-3455 */
    double* tv = (double*)malloc((size_t)T * sizeof(double));
    double* sv = (double*)malloc((size_t)T * sizeof(double));
    for (size_t c = 0; c < nc; ++c) {
        if (isnan(Tz[c])) continue; /* cpp:3507-3508 */
        for (int k = 0; k < T; ++k) {
            tv[k] = Tz[c + nc * k];
            sv[k] = sm[c + nc * k];
        }
        double b1 = 0.0;
        for (int i = 0; i < 288; ++i) b1 = b1 + tv[i];
        b1 = b1 / 288.0;
        double dsum = 0.0, monmean[12];
        for (int d = 0; d < 12; ++d) {
            double tmx = -273.15, tmn = 273.15, ms = 0.0;
            for (int h = 0; h < 24; ++h) {
                double t = tv[d * 24 + h];
                if (t > tmx) tmx = t;
                if (t < tmn) tmn = t;
                ms = ms + t;
            }
            dsum = dsum + (tmx - tmn);
            monmean[d] = ms / 24;
        }
        double b2 = dsum / 12;
        double b4 = std_dev(monmean, 12) * 100.0;
        double b5 = -273.15;
        for (int i = 288; i < 312; ++i)
            if (tv[i] > b5) b5 = tv[i];
        double b6 = 273.15;
        for (int i = 312; i < 336; ++i)
            if (tv[i] < b6) b6 = tv[i];
        double b12 = 0.0;
        for (int i = 0; i < 288; ++i) b12 = b12 + sv[i];
        b12 = b12 / 288.0;
        double b13 = 0.0, b14 = 1.0;
        for (int i = 0; i < T; ++i) {
            if (sv[i] > b13) b13 = sv[i];
            if (sv[i] < b14) b14 = sv[i];
        }
        double b15 = b12 / std_dev(sv, T); /* mean / sd, as written (cpp:3392-3403) */
        double b7 = b5 - b6;
        if (bio[0]) bio[0][c] = b1;
        if (bio[1]) bio[1][c] = b2;
        if (bio[3]) bio[3][c] = b4;
        if (bio[4]) bio[4][c] = b5;
        if (bio[5]) bio[5][c] = b6;
        if (bio[7]) bio[7][c] = quarter_mean(tv, wetq, nwetq);
        if (bio[8]) bio[8][c] = quarter_mean(tv, dryq, ndryq);
        if (bio[9]) bio[9][c] = quarter_mean(tv, hotq, nhotq);
        if (bio[10]) bio[10][c] = quarter_mean(tv, colq, ncolq);
        if (bio[11]) bio[11][c] = b12;
        if (bio[12]) bio[12][c] = b13;
        if (bio[13]) bio[13][c] = b14;
        if (bio[14]) bio[14][c] = b15;
        if (bio[15]) bio[15][c] = quarter_mean(sv, wetq, nwetq);
        if (bio[16]) bio[16][c] = quarter_mean(sv, dryq, ndryq);
        if (bio[17]) bio[17][c] = quarter_mean(sv, hotq, nhotq);
        if (bio[18]) bio[18][c] = quarter_mean(sv, colq, ncolq);
        /* bio7 / bio3 use bio5, bio6, bio2 whether or not those outputs were requested: the reference
         * reads unallocated matrices in that case (undefined behaviour, cpp:3533-3534); here they are
         * the values just computed */
        if (bio[6]) bio[6][c] = b7;
        if (bio[2]) bio[2][c] = b2 / b7;
    }
    free(tv); free(sv); free(Tz); free(sm);
    return MCF_OK;
}
