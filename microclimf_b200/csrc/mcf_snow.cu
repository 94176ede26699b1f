// microclimf_b200 — snow kernels (SURVEY.md §8f NEXT-3), data.frame climate:
//
//   k_snowmodel   gridmodelsnow1  src/microclimfCpp.cpp:4172-4424   per-cell HOURLY RECURRENCE of the snow pack
//                                 (depth, density, age of the canopy+ground and ground-only layers carried from hour to
//                                 hour; snowoneB :3835-3972, radoneB :3773-3833, canopysnowintCpp :3713-3739)
//   k_snowmicro   gridmicrosnow1  src/microclimfCpp.cpp:4894-5057   microclimate above / below the snow surface for
//                                 cell-hours with snow water equivalent > 0 (snowabovepoint :4739-4866,
//                                 belowpointsnow :4868-4892, snowdayan :4679, meanDsnow :4713), overwriting runmicro's
//                                 outputs in place
//
// One thread per cell, cells fastest (R layout), so every [rows, cols, hours] access of a warp is one contiguous
// segment; the time loop runs inside the thread with the snow state in registers — the only true hour-to-hour
// recurrence of the package.  Per-hour quantities that the reference recomputes per cell (solar position, snow albedo,
// daily radiation extremes) come from a per-hour table built once by k_snow_prep.
//
// Unlike the grid solver's hot loops (mcf_physics.cuh), vegetation height and plant area change with the snow depth
// every hour, so nothing can be hoisted per cell: the scalar physics below is a direct restatement of the reference's
// functions, standard double-precision libm, same evaluation order.  It is bound by FP64 latency like k_grid.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "mcf_kernels.cuh"

namespace mcf {
namespace snow {

constexpr double kPiS = 3.14159265358979323846;
constexpr double kToRadS = 3.14159265358979323846 / 180.0;
constexpr double kSbS = 5.67e-8;
constexpr double kKaS = 0.4;
constexpr double kOmdyS = (2.0 * 3.14159265358979323846) / (24.0 * 3600.0);

__device__ __forceinline__ double radem(double tc) { return pow(tc + 273.15, 4.0); }
__device__ __forceinline__ double satvapS(double tc) { // ref :480-490
    return (tc > 0) ? 0.61078 * exp(17.27 * tc / (tc + 237.3)) : 0.61078 * exp(21.875 * tc / (tc + 265.5));
}
__device__ __forceinline__ double dewpointC(double ea) { // ref dewpointCpp :493-496
    return 243.5 * log(ea / 0.6112) / (17.67 - log(ea / 0.6112));
}
// round(x) % n as the reference forms its horizon / wind-shelter sector; a negative angle (outside what checkinputs
// admits; the reference would index out of bounds) is folded into range instead
__device__ __forceinline__ int sector(double x, int n) {
    const int s = ((int)round(x)) % n;
    return s < 0 ? s + n : s;
}
__device__ __forceinline__ double na_realS() { return __longlong_as_double(0x7FF00000000007A2LL); }

// ref solarindexCpp :85-102
__device__ double solarindex(double slope, double aspect, double zend, double azid, bool shadowmask) {
    double si;
    if (zend > 90.0 && !shadowmask) {
        si = 0;
    } else {
        if (slope == 0.0) si = cos(zend * kToRadS);
        else si = cos(zend * kToRadS) * cos(slope * kToRadS) + sin(zend * kToRadS) * sin(slope * kToRadS) * cos((azid - aspect) * kToRadS);
    }
    if (si < 0.0) si = 0.0;
    return si;
}
struct KS { double k, kd, Kc; };
// ref cankCpp :104-132
__device__ KS cank(double zenr, double x, double si) {
    double k;
    if (zenr > (kPiS / 2.0)) zenr = kPiS / 2.0;
    if (si < 0.0) si = 0.0;
    if (x == 1.0) k = 1.0 / (2.0 * cos(zenr));
    else if (isinf(x)) k = 1.0;
    else if (x == 0.0) k = tan(zenr);
    else k = sqrt(x * x + (tan(zenr) * tan(zenr))) / (x + 1.774 * pow((x + 1.182), -0.733));
    if (k > 6000.0) k = 6000.0;
    KS o;
    o.k = k;
    o.kd = k * cos(zenr) / si;
    if (si == 0) o.kd = 1.0;
    o.Kc = 1.0 / si;
    if (si == 0.0) o.Kc = 600.0;
    return o;
}
struct TsDif { double om, a, gma, J, del, h, u1, S1, D1, D2, p1, p2, p3, p4; };
// ref twostreamdifCpp :134-162
__device__ TsDif tsdif(double pait, double x, double lref, double ltra, double gref) {
    TsDif p;
    p.om = lref + ltra;
    p.a = 1.0 - p.om;
    p.del = lref - ltra;
    p.J = 1.0 / 3.0;
    if (x != 1.0) {
        double mla = 9.65 * pow((3.0 + x), -1.65);
        if (mla > kPiS / 2.0) mla = kPiS / 2.0;
        p.J = cos(mla) * cos(mla);
    }
    p.gma = 0.5 * (p.om + p.J * p.del);
    p.h = sqrt(p.a * p.a + 2.0 * p.a * p.gma);
    p.S1 = exp(-p.h * pait);
    p.u1 = p.a + p.gma * (1.0 - 1.0 / gref);
    const double u2 = p.a + p.gma * (1.0 - gref);
    p.D1 = (p.a + p.gma + p.h) * (p.u1 - p.h) * 1.0 / p.S1 - (p.a + p.gma - p.h) * (p.u1 + p.h) * p.S1;
    p.D2 = (u2 + p.h) * 1.0 / p.S1 - (u2 - p.h) * p.S1;
    p.p1 = (p.gma / (p.D1 * p.S1)) * (p.u1 - p.h);
    p.p2 = (-p.gma * p.S1 / p.D1) * (p.u1 + p.h);
    p.p3 = (1.0 / (p.D2 * p.S1)) * (u2 + p.h);
    p.p4 = (-p.S1 / p.D2) * (u2 - p.h);
    return p;
}
struct TsDir { double p5, p6, p7, p8, p9, p10, sig; };
// ref twostreamdirCpp :164-185
__device__ TsDir tsdir(double pait, const TsDif& d, double gref, double kd) {
    TsDir p;
    const double a = d.a, gma = d.gma;
    const double sig = kd * kd + gma * gma - pow((a + gma), 2.0);
    const double ss = 0.5 * (d.om + d.J * d.del / kd) * kd;
    const double sstr = d.om * kd - ss;
    const double S2 = exp(-kd * pait);
    const double u2 = a + gma * (1.0 - gref);
    p.p5 = -ss * (a + gma - kd) - gma * sstr;
    const double v1 = ss - (p.p5 * (a + gma + kd)) / sig;
    const double v2 = ss - gma - (p.p5 / sig) * (d.u1 + kd);
    p.p6 = (1.0 / d.D1) * ((v1 / d.S1) * (d.u1 - d.h) - (a + gma - d.h) * S2 * v2);
    p.p7 = (-1.0 / d.D1) * ((v1 * d.S1) * (d.u1 + d.h) - (a + gma + d.h) * S2 * v2);
    p.sig = -sig;
    p.p8 = sstr * (a + gma + kd) - gma * ss;
    const double v3 = (sstr + gma * gref - (p.p8 / p.sig) * (u2 - kd)) * S2;
    p.p9 = (-1 / d.D2) * ((p.p8 / (p.sig * d.S1)) * (u2 + d.h) + v3);
    p.p10 = (1 / d.D2) * (((p.p8 * d.S1) / p.sig) * (u2 - d.h) + v3);
    return p;
}
// ref zeroplanedisCpp :294, roughlengthCpp :302
__device__ double zeroplanedisS(double h, double pai) {
    if (pai < 0.001) pai = 0.001;
    return (1.0 - (1.0 - exp(-sqrt(7.5 * pai))) / sqrt(7.5 * pai)) * h;
}
__device__ double roughlengthS(double h, double pai, double d, double psi_h) {
    const double Be = sqrt(0.003 + (0.2 * pai) / 2);
    double zm = (h - d) * exp(-kKaS / Be) * exp(kKaS * psi_h);
    if (zm > (0.9 * (h - d))) zm = 0.9 * (h - d);
    if (zm < 0.0005) zm = 0.0005;
    return zm;
}
// ref gturbCpp :373-380
__device__ double gturbS(double uf, double d, double zm, double zref, double ph, double psi_h, double gmin) {
    const double z0 = 0.2 * zm + d;
    const double ln = log((zref - d) / (z0 - d));
    double g = (kKaS * ph * uf) / (ln + psi_h);
    if (g < gmin) g = gmin;
    return g;
}
// ref PenmanMonteithCpp :498-514
__device__ double penmanmonteith(double Rabs, double gHa, double gV, double tc, double te, double pk, double ea, double em,
                                 double G, double erh) {
    const double Rema = em * kSbS * radem(tc);
    double la;
    if (te >= 0) la = 45068.7 - 42.8428 * te;
    else la = 51078.69 - 4.338 * te - 0.06367 * te * te;
    const double cp = 2e-05 * pow(te, 2.0) + 0.0002 * te + 29.119; // cpairCpp :287
    const double Da = satvapS(tc) - ea;
    const double gR = (4.0 * em * kSbS * pow(te + 273.15, 3.0)) / cp;
    const double De = satvapS(te + 0.5) - satvapS(te - 0.5);
    return tc + ((Rabs - Rema - la * (gV / pk) * Da * erh - G) / (cp * (gHa + gR) + la * (gV / pk) * De * erh));
}
// ref canopysnowintCpp :3713-3739
__device__ double canopysnowint(double hgt, double pai, double uf, double prec, double tc, double Li) {
    const double Sh = 6.2;
    if (hgt < 0.001) hgt = 0.001;
    if (pai < 0.001) pai = 0.001;
    const double Be = sqrt(0.003 + (0.2 * pai) / 2.0);
    const double uh = uf / Be;
    const double a = pai / hgt;
    const double Lc = pow(0.25 * a, -1.0);
    const double Lm = 2.0 * pow(Be, 3.0) * Lc;
    const double k1 = Be / Lm;
    double uzm = (uh / (hgt * k1)) * (1 - exp(-k1 * hgt));
    if (uzm < uf) uzm = uf;
    const double rhos = 67.92 + 51.25 * exp(tc / 2.59);
    const double S = Sh * (0.26 + 46 / rhos);
    const double Lstr = S * pai;
    const double Z = atan(uzm / 0.8);
    const double kc = 1.0 / (2.0 * cos(Z));
    const double Cp = 1.0 - exp(-kc * pai);
    const double k2 = Cp / Lstr;
    const double I1 = (Lstr - Li) * (1.0 - exp(-k2 * prec));
    double cis = I1 * 0.678;
    if (cis > prec) cis = prec;
    return cis;
}

// ---------------------------------------------------------------------------------------------
// per-hour table
// ---------------------------------------------------------------------------------------------
struct SnowHour {
    double tc, ea, pk, u2, Rsw, Rdif, Rlw, prec, Tcp, te;
    double Gp, umu, salb;
    double Rmx, Rmn, Rswmx, Rlwmx, Rswmn, Rlwmn, Gmx; // daily extremes replicated per hour (ref :4231-4283)
    double zend, zenr, azid, cosz;
    int32_t sindex, windex;
    double rh; // relative humidity (snow microclimate)
};

struct SnowPrepArgs {
    int tsteps;
    const int32_t *year, *month, *day;
    const double* hour;
    const double *temp, *relhum, *pres, *swdown, *difrad, *lwdown, *windspeed, *winddir, *precip;
    const double *Gp, *Tcp, *RswabsG, *RlwabsG, *umu; // may be NULL (snow microclimate uses only climate + umu)
    double lat, lon;
    SnowHour* hours;
    double* scal; // [0] series maximum of tc
};

// one block; thread 0 runs the sequential snow-age scan (ref snowalbCpp :3752-3771)
__global__ void __launch_bounds__(256) k_snow_prep(const __grid_constant__ SnowPrepArgs a) {
    const int T = a.tsteps;
    for (int k = threadIdx.x; k < T; k += blockDim.x) {
        SnowHour h;
        h.tc = a.temp[k];
        h.rh = a.relhum[k];
        h.ea = satvapS(h.tc) * h.rh / 100.0;
        h.pk = a.pres[k];
        h.u2 = a.windspeed[k];
        h.Rsw = a.swdown[k];
        h.Rdif = a.difrad[k];
        h.Rlw = a.lwdown[k];
        h.prec = a.precip[k];
        h.Tcp = a.Tcp ? a.Tcp[k] : 0.0;
        h.te = (h.Tcp + h.tc) / 2.0;
        h.Gp = a.Gp ? a.Gp[k] : 0.0;
        h.umu = a.umu ? a.umu[k] : 1.0;
        const SolPos s = solposition(a.lat, a.lon, a.year[k], a.month[k], a.day[k], a.hour[k]);
        h.zend = s.zend;
        h.zenr = s.zenr;
        h.azid = s.azid;
        h.cosz = cos(s.zenr);
        h.sindex = sector(s.azid / 15, 24);
        h.windex = sector(a.winddir[k] / 45, 8);
        h.salb = 0.0;
        h.Rmx = h.Rmn = h.Rswmx = h.Rlwmx = h.Rswmn = h.Rlwmn = h.Gmx = 0.0;
        a.hours[k] = h;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int hs = 0;
        double mx = -273.15;
        for (int k = 0; k < T; ++k) {
            if (k > 0) hs = (a.precip[k] > 0) ? 0 : hs + 1;
            // `hs[i] / 24` is an INTEGER division in the reference: log(0) = -inf for the first day after snowfall
            double alb = (-9.8740 * log((double)(hs / 24)) + 78.3434) / 100.0;
            if (alb > 0.95) alb = 0.95;
            if (alb < 0.1) alb = 0.1;
            a.hours[k].salb = alb;
            if (a.temp[k] > mx) mx = a.temp[k];
        }
        a.scal[0] = mx;
    }
    // daily extremes of the point model's net radiation (ref :4222-4283); only whole days of `tr`
    if (a.RswabsG) {
        const int ndays = T / 24;
        for (int d = threadIdx.x; d < ndays; d += blockDim.x) {
            double Rmxd = -1352.0, Rmnd = 1352.0, swmx = 0, lwmx = 0, swmn = 0, lwmn = 0, Gmxd = 0.0;
            for (int hh = 0; hh < 24; ++hh) {
                const int k = d * 24 + hh;
                const double Rem = 0.97 * kSbS * radem(a.temp[k]);
                const double Rnet = a.RswabsG[k] + a.RlwabsG[k] - Rem;
                if (Rmxd < Rnet) { Rmxd = Rnet; swmx = a.swdown[k]; lwmx = a.lwdown[k]; }
                if (Rmnd > Rnet) { Rmnd = Rnet; swmn = a.swdown[k]; lwmn = a.lwdown[k]; }
                if (fabs(Rnet) > Gmxd) Gmxd = fabs(Rnet);
            }
            for (int hh = 0; hh < 24; ++hh) {
                SnowHour& h = a.hours[d * 24 + hh];
                h.Rmx = Rmxd; h.Rmn = Rmnd; h.Rswmx = swmx; h.Rlwmx = lwmx; h.Rswmn = swmn; h.Rlwmn = lwmn; h.Gmx = Gmxd;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// gridmodelsnow1
// ---------------------------------------------------------------------------------------------
struct SnowModelArgs {
    int rows, cols, tsteps;
    const SnowHour* hours;
    const double *pai, *hgt, *ltra, *clump;
    const double *slope, *aspect, *skyview, *wsa, *hor;
    const double *isnowdc, *isnowdg;
    const int32_t *isnowac, *isnowag;
    double zref;
    double sdp[4];
    double *Tc, *Tg, *sdepc, *sdepg, *sden;   // [rows, cols, tsteps]
    double *agec, *ageg, *meltc, *meltg;      // [rows, cols]
};

struct SnowRad { double RabsC, RswabsG, RlwabsG, tr; };
// ref radoneB :3773-3833 (solar position from the hour table)
__device__ SnowRad radone(const SnowHour& h, double Rsw, double Rdif, double Rlw, double pai, double hgt, double ltra,
                          double clump, double alb, double slope, double aspect) {
    SnowRad out;
    const double RlwabsC = 0.97 * Rlw;
    out.RlwabsG = RlwabsC;
    const double cld = clump * clump;
    const double pait = pai / (1.0 - clump);
    out.tr = (1.0 - cld) * exp(-pait) + cld;
    if (hgt > 0.0) {
        const double Rsky = out.tr * Rlw;
        const double Rcan = (1.0 - out.tr) * 0.97 * kSbS * radem(h.Tcp);
        out.RlwabsG = 0.97 * (Rsky + Rcan);
    }
    out.RabsC = RlwabsC;
    out.RswabsG = 0.0;
    if (Rsw > 0.0) {
        double si = solarindex(slope, aspect, h.zend, h.azid, false);
        if (si < 0.0) si = 0.0;
        const double cosz = h.cosz; // cos(zenr) of the unclamped zenith, as the reference (:3796)
        double Rbeam = (Rsw - Rdif) / cosz;
        if (Rbeam > 1352.2) Rbeam = 1352.2;
        const double RswabsC = (1.0 - alb) * (Rdif + Rbeam * cosz);
        out.RabsC = RswabsC + RlwabsC;
        out.RswabsG = RswabsC;
        if (hgt > 0.0) {
            if ((alb + ltra) > 0.999) ltra = 0.999 - alb;
            const TsDif d = tsdif(pait, 1.0, alb, ltra, alb);
            const KS kp = cank(h.zenr, 1.0, si);
            const TsDir r = tsdir(pait, d, alb, kp.kd);
            const double clb = pow(clump, kp.Kc);
            double Rddm = (1.0 - cld) * (d.p3 * exp(-d.h * pait) + d.p4 * exp(d.h * pait)) + cld;
            if (Rddm > 1.0) Rddm = 1.0;
            if (Rddm < 0.0) Rddm = 0.0;
            double Rdbm = (1.0 - clb) * ((r.p8 / r.sig) * exp(-kp.kd * pait) + r.p9 * exp(-d.h * pait) + r.p10 * exp(d.h * pait));
            if (Rdbm > 1.0) Rdbm = 1.0;
            if (Rdbm < 0.0) Rdbm = 0.0;
            double Rbgm = (1.0 - clb) * exp(-kp.kd * pait) + clb;
            if (Rbgm > 1.0) Rbgm = 1.0;
            if (Rbgm < 0.0) Rbgm = 0.0;
            const double RdifG = (1.0 - alb) * (Rdbm * Rbeam * cosz) + Rddm * Rdif;
            const double RdirG = (1.0 - alb) * (Rbgm * Rbeam * 0.5);
            out.RswabsG = RdifG + RdirG;
        }
    }
    return out;
}

struct SnowStep { double Tc, Tg, sdepc, sdepg, sdenc, sdeng, agec, ageg, melc, melg; };
// one hour of one cell: snowoneB (ref :3835-3972) with umu = 1, psim = psih = 0
__device__ SnowStep snow_step(const SnowHour& h, double Rswp, double Rdifp, double Rlwp, double u2p, double G, double hgt0,
                              double pai0, double ltra0, double clump, double slope, double aspect, double zref,
                              const double* sdp, double sdepcp, double sdepgp, double sdencp, double sdengp, int snowagec,
                              int snowageg) {
    // ---- snowoneB (ref :3835-3972), umu = 1
    double pai = 0.0;
    if (hgt0 > sdepgp) pai = pai0 * (hgt0 - sdepgp) / hgt0;
    double hgt = hgt0 - sdepgp;
    if (hgt < 0.0) hgt = 0.0;
    double zi = 0.0;
    if (sdepgp > 0.0 && hgt > 0.0) zi = ((sdepcp - sdepgp) * sdencp) / (hgt * 1000.0);
    const double ltra = ltra0 * exp(-10.1 * zi);
    const SnowRad rad = radone(h, Rswp, Rdifp, Rlwp, pai, hgt, ltra, clump, h.salb, slope, aspect);
    const double RabsG = rad.RswabsG + rad.RlwabsG;
    double d = 0.0, zm = 0.005;
    if (hgt > 0.0) {
        d = zeroplanedisS(hgt, pai);
        zm = roughlengthS(hgt, pai, d, 0.0);
    }
    if (zm < 0.0009) zm = 0.0009;
    const double uf = (kKaS * u2p) / (log((zref - d) / zm) + 0.0);
    const double ph = 44.6 * (h.pk / 101.3) * (273.15 / (h.tc + 273.15)); // phairCpp :280
    const double gHa = gturbS(uf, d, zm, zref, ph, 0.0, 0.03);
    double Tc = penmanmonteith(rad.RabsC, gHa, gHa, h.tc, h.te, h.pk, h.ea, 0.97, G, 1.0);
    double Tg = penmanmonteith(RabsG, gHa, gHa, h.tc, h.te, h.pk, h.ea, 0.97, G, 1.0);
    const double tdew = dewpointC(h.ea);
    if (Tc < tdew) Tc = tdew;
    if (Tg < tdew) Tg = tdew;
    // canopy + ground pack
    double la;
    if (Tc < 0.0) la = 51078.69 - 4.338 * Tc - 0.06367 * Tc * Tc;
    else la = 45068.7 - 42.8428 * Tc;
    double L = la * (gHa / h.pk) * (satvapS(Tc) - h.ea);
    la = la / 0.018015;
    const double mSc = (L / la) * 3.6;
    double mMc = 0.0;
    if (Tc > 0.0) {
        const double S = sdepcp * (sdencp / 1000);
        const double Fm = 583.3 * Tc * S;
        mMc = (Fm / 334000.0) * 3.6;
        if (sdepcp > 0.0) Tc = 0.0;
    }
    double mRc = 0.0;
    if (h.tc > 0.0) mRc = 0.0125 * h.tc * h.prec / 1000;
    // ground-only pack
    if (Tg < 0.0) la = 51078.69 - 4.338 * Tg - 0.06367 * Tg * Tg;
    else la = 45068.7 - 42.8428 * Tg;
    double mu = exp(-pai);
    if (mu > 1.0) mu = 1.0;
    L = la * (gHa / h.pk) * (satvapS(Tg) - h.ea) * mu;
    la = la / 0.018015;
    const double mSg = (L / la) * 3.6;
    double mMg = 0.0;
    if (Tg > 0.0) {
        const double S = sdepgp * (sdengp / 1000.0);
        const double Fm = 583.3 * Tg * S;
        mMg = (Fm / 334000.0) * 3.6;
        if (sdepgp > 0.0) Tg = 0.0;
    }
    double Li = 0.0;
    if (sdepcp > 0.0) {
        double wgtg = sdepgp / sdepcp;
        if (wgtg < 0.0) wgtg = 0.0;
        if (wgtg > 1.0) wgtg = 1.0;
        const double sdencc = wgtg * sdengp + (1.0 - wgtg) * sdencp;
        Li = (sdepcp - sdepgp) * sdencc;
    }
    if (Li < 0.0) Li = 0.0;
    double cis = canopysnowint(hgt, pai, uf, h.prec, h.tc, Li);
    if (cis > h.prec) cis = h.prec;
    double mRg = 0.0;
    if (h.tc > 0.0) mRg = 0.0125 * h.tc * (h.prec - cis) / 1000.0;
    double snowc = h.prec, snowg = h.prec - cis;
    if (h.tc > 2.0) { snowc = 0.0; snowg = 0.0; }
    const double swec = snowc / 1000.0 - mSc - mMc - mRc;
    const double sweg = snowg / 1000.0 - mSg - mMg - mRg;
    double agec_n = (double)snowagec + 1.0, ageg_n = (double)snowageg + 1.0;
    const double sdenc_n = ((sdp[0] - sdp[1]) * (1.0 - exp(-sdp[2] * sdepcp / 100.0 - sdp[3] * agec_n / 24.0)) + sdp[1]) * 1000.0;
    const double sdeng_n = ((sdp[0] - sdp[1]) * (1.0 - exp(-sdp[2] * sdepgp / 100.0 - sdp[3] * ageg_n / 24.0)) + sdp[1]) * 1000.0;
    double sdepc_n = sdepcp + (swec * 1000.0) / sdenc_n;
    double sdepg_n = sdepgp + (sweg * 1000.0) / sdeng_n;
    if (sdepc_n < 0.0) { sdepc_n = 0.0; agec_n = 0.0; }
    if (sdepg_n < 0.0) { sdepg_n = 0.0; ageg_n = 0.0; }
    SnowStep o;
    o.Tc = Tc; o.Tg = Tg; o.sdepc = sdepc_n; o.sdepg = sdepg_n; o.sdenc = sdenc_n; o.sdeng = sdeng_n;
    o.agec = agec_n; o.ageg = ageg_n;
    o.melc = mSc + mMc + mRc;
    o.melg = mSg + mMg + mRg;
    return o;
}

// ref belowpointsnow :4868-4892
__device__ double below_snow(double reqhgts, double meanD, double stg, double Tzd, double mat, int hiy) {
    const double nb = -118.35 * reqhgts / meanD;
    double Tz = stg;
    if (nb > 1.0) {
        if (nb <= 24.0) {
            const double w1 = 1.0 / nb, w2 = nb / 24.0, wgt = w1 / (w1 + w2);
            Tz = wgt * stg + (1 - wgt) * Tzd;
        } else if (nb <= (double)hiy) {
            const double w1 = 24.0 / nb, w2 = nb / (double)hiy, wgt = w1 / (w1 + w2);
            Tz = wgt * Tzd + (1 - wgt) * mat;
        } else {
            Tz = mat;
        }
    }
    return Tz;
}

__global__ void __launch_bounds__(128) k_snowmodel(const __grid_constant__ SnowModelArgs a) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    const int nc = a.rows * a.cols;
    if (cell >= nc) return;
    const double hgt0 = a.hgt[cell];
    const double NA = na_realS();
    if (isnan(hgt0)) { // untouched cells keep the reference's NA prefill (:4296-4308, bioclimfill)
        for (int k = 0; k < a.tsteps; ++k) {
            const size_t idx = (size_t)k * nc + cell;
            a.Tc[idx] = NA; a.Tg[idx] = NA; a.sdepc[idx] = NA; a.sdepg[idx] = NA; a.sden[idx] = NA;
        }
        a.agec[cell] = NA; a.ageg[cell] = NA; a.meltc[cell] = NA; a.meltg[cell] = NA;
        return;
    }
    const double pai0 = a.pai[cell], clump = a.clump[cell], ltra0 = a.ltra[cell];
    const double slope = a.slope[cell], aspect = a.aspect[cell], svf = a.skyview[cell];
    const double* sdp = a.sdp;
    int snowagec = a.isnowac[cell], snowageg = a.isnowag[cell];
    double sdepcp = a.isnowdc[cell], sdepgp = a.isnowdg[cell];
    double sdencp = ((sdp[0] - sdp[1]) * (1 - exp(-sdp[2] * sdepcp / 100.0 - sdp[3] * snowagec / 24.0)) + sdp[1]) * 1000.0;
    double sdengp = ((sdp[0] - sdp[1]) * (1 - exp(-sdp[2] * sdepgp * 0.5 / 100.0 - sdp[3] * snowageg / 24.0)) + sdp[1]) * 1000.0;
    double meltc = 0.0;
    double meltg = NA; // bioclimfill leaves NA and the loop only ever adds to it (:4401): reproduced
#pragma unroll 1
    for (int k = 0; k < a.tsteps; ++k) {
        const SnowHour h = a.hours[k];
        const size_t idx = (size_t)k * nc + cell;
        int snowtest = 0;
        if (sdepcp > 0.0) snowtest = 1;
        if (h.tc < 2.0 && h.prec > 0.0) snowtest = 1;
        if (snowtest > 0) {
            double paip = pai0;
            if (hgt0 > sdepgp) paip = paip * (hgt0 - sdepgp) / hgt0;
            // ground heat flux (ref :4338-4347)
            const double dtR = h.Rmx - h.Rmn;
            const double trS = svf * exp(-paip);
            const double Rem = 0.97 * kSbS * radem(h.tc);
            const double dmxS = trS * h.Rswmx + trS * h.Rlwmx + (1 - trS) * Rem - Rem;
            const double dmnS = trS * h.Rswmn + trS * h.Rlwmn + (1 - trS) * Rem - Rem;
            const double Gmu = (dmxS - dmnS) / dtR;
            double G = h.Gp * Gmu;
            if (G > h.Gmx) G = h.Gmx;
            if (G < -h.Gmx) G = -h.Gmx;
            const double ha = a.hor[(size_t)h.sindex * nc + cell];
            const double sa = 90 - h.zend;
            double smu = 1.0;
            if (ha > tan(sa * kToRadS)) smu = 0.0;
            const double ws = a.wsa[(size_t)h.windex * nc + cell];
            const double u2p = h.umu * ws * h.u2;
            const double Rdifp = h.Rdif * svf;
            const double Rdirp = (h.Rsw - h.Rdif) * smu;
            const double Rswp = Rdirp + Rdifp;
            const double Rlwp = h.Rlw * svf;
            const SnowStep st = snow_step(h, Rswp, Rdifp, Rlwp, u2p, G, hgt0, pai0, ltra0, clump, slope, aspect, a.zref, sdp,
                                          sdepcp, sdepgp, sdencp, sdengp, snowagec, snowageg);
            const double Tc = st.Tc, Tg = st.Tg, sdepc_n = st.sdepc, sdepg_n = st.sdepg, sdenc_n = st.sdenc, sdeng_n = st.sdeng;
            const double agec_n = st.agec, ageg_n = st.ageg;
            // ---- outputs and state update (ref :4382-4402)
            a.Tc[idx] = Tc;
            a.Tg[idx] = Tg;
            a.sdepc[idx] = sdepc_n;
            a.sdepg[idx] = sdepg_n;
            a.sden[idx] = sdenc_n;
            sdencp = sdenc_n;
            sdengp = sdeng_n;
            sdepcp = sdepc_n;
            sdepgp = sdepg_n;
            snowagec = (int)agec_n; // double -> int, as the reference's assignment (:4393-4394)
            snowageg = (int)ageg_n;
            const double melc = st.melc, melg = st.melg;
            meltc = meltc + melc;
            meltc = meltc + (melc * 1000.0) / sdenc_n;
            meltg = meltg + (melg * 1000.0) / sdeng_n;
        } else {
            a.Tc[idx] = 0.0;
            a.Tg[idx] = 0.0;
            a.sdepc[idx] = 0.0;
            a.sdepg[idx] = 0.0;
            a.sden[idx] = sdp[1] * 1000.0;
        }
    }
    a.agec[cell] = (double)snowagec;
    a.ageg[cell] = (double)snowageg;
    a.meltc[cell] = meltc;
    a.meltg[cell] = meltg;
}

// ---------------------------------------------------------------------------------------------
// gridmicrosnow1
// ---------------------------------------------------------------------------------------------
struct SnowMicroArgs {
    int rows, cols, tsteps;
    const SnowHour* hours;
    const double* scal; // [0] mxtc
    double reqhgt, zref, mat;
    int hiy;
    const double *pai, *paia, *hgt, *ltra, *clump, *leafd, *leafden;
    const double *slope, *aspect, *skyview, *wsa, *hor, *Smax;
    const double *snowtempc, *snowtempg, *swe, *sdepg, *sden; // [rows, cols, tsteps]
    double* out[10];                                            // in/out (runmicro's arrays), NULL = absent
};

struct WindS { double uf, uz, gHa; };
// ref windCpp :1189-1218
__device__ WindS windS(double reqhgt, double zref, double h, double uref, double umu, double ws, double d, double zm, double aw) {
    WindS o;
    if (isnan(ws)) ws = 1.0;
    if (ws < 0.05) ws = 0.05;
    const double ufs = (kKaS * uref) / log((zref - d) / zm);
    o.uf = ufs * umu * ws;
    if (o.uf < 0.001) o.uf = 0.001;
    o.uz = o.uf;
    if (reqhgt > 0) {
        if (reqhgt >= h) {
            o.uz = (o.uf / kKaS) * log((reqhgt - d) / zm);
        } else {
            double uh = (o.uf / kKaS) * log((h - d) / zm);
            if (uh < o.uf) uh = o.uf;
            double Be = o.uf / uh;
            if (Be < 0.001) Be = 0.001;
            const double Lc = pow(0.25 * aw, -1.0);
            const double Lm = 2 * pow(Be, 3.0) * Lc;
            o.uz = uh * exp(Be * (reqhgt - h) / Lm);
        }
        if (o.uz > uref) o.uz = uref;
    }
    o.gHa = gturbS(o.uf, d, zm, zref, 43, 0, 0.0001);
    return o;
}
// ref TVabove :1298-1313
__device__ void tvabove(double reqhgt, double zref, double d, double zm, double T0, double tc, double ea, double surfwet,
                        double& Tz, double& ez) {
    const double zh = 0.2 * zm;
    const double estl = satvapS(T0);
    if (reqhgt > (d + zh)) {
        const double lnr = log((reqhgt - d) / zh) / log((zref - d) / zh);
        Tz = tc + (T0 - tc) * (1 - lnr);
        ez = ea + (estl - ea) * surfwet * (1 - lnr);
    } else {
        Tz = T0;
        ez = ea + (estl - ea) * surfwet;
    }
}
// ref rhcanopy :1365-1380, TVbelow :1381-1409
__device__ double rhcanopyS(double uf, double h, double d, double z) {
    const double a2 = 0.4 * (1.0 - (d / h)) / pow(1.25, 2);
    double inth = 4.293251 * h;
    if (z != h) {
        const double s = sin((kPiS * z) / h), c = cos((kPiS * z) / h);
        inth = (2.0 * h * ((48 * atan((sqrt(5.0) * s) / (c + 1))) / pow(5.0, 1.5) +
                           (32.0 * s) / ((c + 1) * ((25.0 * pow(s, 2.0)) / pow((c + 1.0), 2.0) + 5.0)))) / kPiS;
    }
    const double mu = uf / (a2 * h) * 1.0 / (uf * uf);
    double rHa = inth * mu;
    if (rHa < 0.001) rHa = 0.001;
    return rHa;
}
__device__ double tvbelow(double z, double d, double h, double pai, double uf, double leafden, double Flux, double Fluxz,
                          double SH, double SG, double mxnear) {
    const double Rc = rhcanopyS(uf, h, d, h);
    const double Kc = h / Rc;
    double Kg = 1.0 / rhcanopyS(uf, h, d, z);
    double Kh = 1.0 / (Rc - rhcanopyS(uf, h, d, z));
    Kg = Kg / z;
    Kh = Kh / (h - z);
    const double SC = SH + Flux / Kc;
    const double farg = (Kg * SG + Kh * SH + Kc * SC) / (Kg + Kh + Kc);
    const double SN = Fluxz * leafden;
    double near = (3.047519 + 0.128642 * log(pai)) * SN;
    if (fabs(near) > mxnear) near = (near > 0.0) ? mxnear : -mxnear;
    if (isnan(near)) near = 0;
    return near + farg;
}
struct PM2 { double Ts, H, L, mu; };
// ref PenmanMonteith2Cpp :1220-1247
__device__ PM2 penmon2(double Rabs, double gHa, double gV, double tc, double mxtc, double pk, double ea, double es, double G,
                       double surfwet, double tdew) {
    const double De = satvapS(tc + 0.5) - satvapS(tc - 0.5);
    const double gHr = gHa + (4 * 0.97 * kSbS * pow(tc + 273.15, 3.0)) / 29.3;
    const double Rem = 0.97 * kSbS * radem(tc);
    double la;
    if (tc >= 0) la = 45068.7 - 42.8428 * tc;
    else la = 51078.69 - 4.338 * tc - 0.06367 * tc * tc;
    const double m = la * (gV / pk);
    const double L = m * (es - ea) * surfwet;
    double dT = (Rabs - Rem - L - G) / (29.3 * gHr + m * De);
    const double dTmx = -0.6273 * mxtc + 49.79;
    if (dT > dTmx) dT = dTmx;
    if (dT > 80.0) dT = 80.0;
    PM2 o;
    o.Ts = dT + tc;
    if (o.Ts < tdew) o.Ts = tdew;
    o.H = 29.3 * gHa * (o.Ts - tc);
    o.L = m * (satvapS(o.Ts) - ea) * surfwet;
    o.mu = la * (43.0 / pk);
    return o;
}
// ref mincondCpp :1316-1331
__device__ double mincondS(double leafabs, double gs, double tc, double leafd) {
    const double Rnet = leafabs - 0.97 * kSbS * radem(tc);
    double rs = 500.0;
    if (gs > 0.0) rs = 1 / gs;
    if (rs > 500.0) rs = 500.0;
    const double Hlf = 1.09767 * pow(rs, 0.2672778);
    const double Hf = -1.0 / (1.0 + exp(2.0 - Hlf));
    const double H = Hf * Rnet;
    double gmin = 0.0463 * pow(fabs(H) / leafd, 0.2);
    if (gmin < 0.05) gmin = 0.05;
    return gmin;
}

struct SnowMicroOut { double Tz, tleaf, rh, uz, Rbdown, Rddown, Rlwdn, Rdup, Rlwup; };

// ref snowabovepoint :4739-4866
__device__ SnowMicroOut snowabove(double reqhgt, double zref, const SnowHour& h, double hgt, double pai, double paia,
                                  double leafd, double clump, double ltra, double leafden, double si, double svfa,
                                  int shadowmask, double ws, double mxtc, double snowtempg, double snowtempc, double sdepc,
                                  double sdepg, double sdenc, double albc, double albg) {
    if (reqhgt == 0.0) reqhgt = 0.001;
    SnowMicroOut out;
    const double tc = h.tc, pk = h.pk, Rsw = h.Rsw, Rdif = h.Rdif, Rlw = h.Rlw;
    const double es = satvapS(tc);
    const double ea = es * h.rh / 100.0;
    const double tdew = dewpointC(ea);
    double hgts = hgt - sdepg;
    if (hgts < 0.0) hgts = 0.0;
    double pais = 0.0, d = 0.0, zm = 1e-5, aw = 0.0;
    if (hgts > 0.0) {
        pais = pai * hgts / hgt;
        d = zeroplanedisS(hgts, pais); // windtiCpp :1179-1188
        zm = roughlengthS(hgts, pais, d, 0.0);
        if (zm < 1e-6) zm = 1e-6;
        aw = pais / hgts;
    }
    // tiw.a is left uninitialised by the reference when hgts == 0; windCpp only reads it for reqhgt < h (never then)
    const WindS wind = windS(reqhgt, zref, hgts, h.u2, h.umu, ws, d, zm, aw);
    out.uz = wind.uz;
    double ez;
    if (reqhgt >= hgts) {
        if (Rsw > 0.0) {
            out.Rddown = Rdif * svfa;
            if (si > 0.0) {
                if (shadowmask > 0) {
                    out.Rbdown = (Rsw - Rdif) / si;
                    if (out.Rbdown > 1352.0) out.Rbdown = 1352.0;
                    out.Rdup = albc * Rsw * svfa;
                } else {
                    out.Rbdown = 0.0;
                    out.Rdup = albc * Rdif * svfa;
                }
            } else {
                out.Rbdown = 0.0;
                out.Rdup = albc * Rdif * svfa;
            }
        } else {
            out.Rbdown = 0.0;
            out.Rddown = 0.0;
            out.Rdup = 0.0;
        }
        out.Rlwdn = svfa * Rlw;
        out.Rlwup = svfa * 0.97 * kSbS * radem(snowtempc);
        tvabove(reqhgt, zref, d, zm, snowtempc, tc, ea, 1.0, out.Tz, ez);
        out.tleaf = snowtempc;
    } else {
        double paias = 0.0;
        if (hgts > 0.0) paias = paia * hgts / hgt;
        double zi = 0.0;
        if (sdepg > 0.0) zi = ((sdepc - sdepg) * sdenc) / (hgts * 1000.0);
        double ltras = ltra * exp(-10.1 * zi);
        const double sm = ltras + albc;
        if (sm > 0.999) ltras = 0.999 - albc;
        double clumps = clump;
        if (clump > 0.0) clumps = pow(clump, pais / pai);
        double pait = pais;
        if (clump > 0.0) pait = pais / (1.0 - clumps);
        const TsDif dif1 = tsdif(pait, 1.0, albc, ltras, albg);
        // twostreamdif (ref :1034-1084) with x = 1, lref = albc, gref = albg
        const double pait2 = pais / (1.0 - clumps);
        const TsDif dif2 = tsdif(pait2, 1.0, albc, ltras, albg);
        double gi = 0.0;
        if (clumps > 0.0) gi = pow(clumps, paias / pais);
        if (gi > 0.99) gi = 0.99;
        double giu = 0.0;
        if (clumps > 0.0) giu = pow(clumps, (pais - paias) / pais);
        if (giu > 0.99) giu = 0.99;
        const double trd = gi * gi;
        const double trdn = pow(clumps, 2.0);
        const double trdu = giu * giu;
        const double paiaa = paias / (1.0 - gi);
        double amx = albg;
        if (amx < albc) amx = albc;
        double albd = (1.0 - trdn * trdn) * (dif2.p1 + dif2.p2) + trdn * trdn * albg;
        if (albd > amx) albd = amx;
        if (albd < 0.01) albd = 0.01;
        double Rddn_z = (1.0 - trd) * (dif2.p3 * exp(-dif2.h * paiaa) + dif2.p4 * exp(dif2.h * paiaa)) + trd;
        if (Rddn_z > 1.0) Rddn_z = 1.0;
        if (Rddn_z < 0.0) Rddn_z = 0.0;
        double Rdup_z = (1.0 - trdu * trdn) * (dif2.p1 * exp(-dif2.h * paiaa) + dif2.p2 * exp(dif2.h * paiaa)) + trdu * trdn * albg;
        if (Rdup_z > 1.0) Rdup_z = 1.0;
        if (Rdup_z < 0.0) Rdup_z = 0.0;
        const KS kp = cank(h.zenr, 1.0, si);
        const TsDir dir = tsdir(pait, dif1, albg, kp.kd);
        // twostreamCpp (ref :1086-1178): the streams snowabovepoint uses
        double Rbdown = 0.0, Rddown = 0.0, Rdup = 0.0, radLsw = 0.0;
        if (Rsw > 0.0) {
            const double cosz = cos(h.zenr);
            if (pais > 0.0) {
                double trbn = pow(clumps, kp.Kc);
                if (trbn > 0.999) trbn = 0.999;
                if (trbn < 0.0) trbn = 0.0;
                double trb = pow(gi, kp.Kc);
                if (trb > 0.999) trb = 0.999;
                if (trb < 0.0) trb = 0.0;
                double Rdbup_z = (1.0 - trdu * trbn) * ((dir.p5 / -dir.sig) * exp(-kp.kd * paiaa) + dir.p6 * exp(-dif2.h * paiaa) +
                                                       dir.p7 * exp(dif2.h * paiaa)) + trdu * trbn * albg;
                if (Rdbup_z > amx) Rdbup_z = amx;
                if (Rdbup_z < 0.0) Rdbup_z = 0.0;
                double Rdbdn_z = (1.0 - trb) * ((dir.p8 / dir.sig) * exp(-kp.kd * paiaa) + dir.p9 * exp(-dif2.h * paiaa) +
                                                dir.p10 * exp(dif2.h * paiaa));
                if (Rdbdn_z > amx) Rdbdn_z = amx;
                if (Rdbdn_z < 0.0) Rdbdn_z = 0.0;
                double Rbeam = (Rsw - Rdif) / cosz;
                if (Rbeam > 1352.0) Rbeam = 1352.0;
                const double Rb = Rbeam * cosz;
                Rbdown = (trb + (1.0 - trb) * exp(-kp.kd * paiaa)) * Rbeam;
                Rddown = Rddn_z * Rdif * svfa + Rdbdn_z * Rb;
                Rdup = Rdup_z * Rdif * svfa + Rdbup_z * Rb;
                radLsw = 0.5 * (1.0 - dif2.om) * (Rddown + Rdup + kp.k * cosz * Rbdown);
            } else {
                Rbdown = (Rsw - Rdif) / cosz;
                Rddown = Rdif * svfa;
                Rdup = albg * (Rdif * svfa + (Rsw - Rdif));
            }
        }
        (void)albd;
        if (shadowmask == 0) Rbdown = 0.0;
        // leaftemp (ref :1333-1364) with gsmax = 999.999: the stomatal branch is skipped
        const double lwcan = 0.97 * kSbS * radem(snowtempc);
        const double lwgro = 0.97 * kSbS * radem(snowtempg);
        const double paig = pais - paias;
        const double lwup = exp(-paig) * lwgro + (1 - exp(-paig)) * lwcan;
        const double lwdn = exp(-paias) * Rlw + (1 - exp(-paias)) * lwcan;
        const double lwabs = 0.97 * 0.5 * (lwup + lwdn);
        const double leafabs = radLsw + lwabs;
        double gh = 0.135 * sqrt(wind.uz / leafd) * 1.4;
        const double gmin = mincondS(leafabs, 999.99, snowtempc, leafd);
        if (gh < gmin) gh = gmin;
        const PM2 pm = penmon2(leafabs, gh, gh, tc, mxtc, pk, ea, es, 0.0, 1.0, tdew);
        out.tleaf = pm.Ts;
        const double H = 29.3 * wind.gHa * (snowtempc - tc);
        double Flux = H * (1.0 - exp(-pais));
        double Fluxz = pm.H;
        double Th, eh;
        tvabove(hgts, zref, d, zm, snowtempc, tc, ea, 1.0, Th, eh);
        double SH = Th * 29.3 * 43.0;
        double SG = snowtempg * 29.3 * 43.0;
        double mxnear = fabs(out.tleaf - Th) * 29.3 * 43.0;
        out.Tz = tvbelow(reqhgt, d, hgts, pais, wind.uf, leafden, Flux, Fluxz, SH, SG, mxnear) / (29.3 * 43);
        double la;
        if (tc < 0) la = 51078.69 - 4.338 * tc - 0.06367 * tc * tc;
        else la = 45068.7 - 42.8428 * tc;
        const double m = la * (wind.gHa / pk);
        const double L = m * (es - ea);
        Flux = L * (1.0 - exp(-pais));
        Fluxz = pm.L;
        const double mu = la * (43 / pk);
        SH = eh * mu;
        SG = satvapS(snowtempg) * mu;
        mxnear = fabs(satvapS(out.tleaf) - eh) * mu;
        ez = tvbelow(reqhgt, d, hgts, pais, wind.uf, leafden, Flux, Fluxz, SH, SG, mxnear) / mu;
        out.Rbdown = Rbdown;
        out.Rddown = Rddown;
        out.Rdup = Rdup;
        out.Rlwdn = lwdn;
        out.Rlwup = lwup;
    }
    out.rh = (ez / satvapS(out.Tz)) * 100.0;
    if (out.rh > 100.0) out.rh = 100.0;
    // std::max / std::min over {tleaf, tc, snowtempg, snowtempc} in the reference's fold order
    double tmx = out.tleaf;
    if (tmx < tc) tmx = tc;
    if (tmx < snowtempg) tmx = snowtempg;
    if (tmx < snowtempc) tmx = snowtempc;
    double tmn = out.tleaf;
    if (tc < tmn) tmn = tc;
    if (snowtempg < tmn) tmn = snowtempg;
    if (snowtempc < tmn) tmn = snowtempc;
    tmx += 2.0;
    tmn -= 2.0;
    if (out.Tz > tmx) out.Tz = tmx;
    if (out.Tz < tmn) out.Tz = tmn;
    return out;
}

__global__ void __launch_bounds__(128) k_snowmicro(const __grid_constant__ SnowMicroArgs a) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    const int nc = a.rows * a.cols;
    if (cell >= nc) return;
    const double hgt = a.hgt[cell];
    if (isnan(hgt)) return;
    const int T = a.tsteps;
    // meanDsnow (ref :4713-4737); NA when the first hour's density is NA
    double meanD = na_realS();
    if (!isnan(a.sden[cell])) {
        double sumD = 0.0;
        for (int k = 0; k < T; ++k) {
            const double sd = a.sden[(size_t)k * nc + cell];
            const double co = 0.0442 * exp(5.181 * sd / 1000.0);
            const double kap = co / (sd * 2090.0);
            sumD += sqrt(2.0 * kap / kOmdyS);
        }
        meanD = sumD / (double)T;
    }
    const bool tzd_ok = !isnan(a.snowtempg[cell]); // snowdayan (ref :4679-4711): NA series when hour 0 is NA
    const double pai = a.pai[cell], paia = a.paia[cell], ltra = a.ltra[cell], clump = a.clump[cell];
    const double leafd = a.leafd[cell], leafden = a.leafden[cell];
    const double slope = a.slope[cell], aspect = a.aspect[cell], svf = a.skyview[cell], Smax = a.Smax[cell];
    const double mxtc = a.scal[0];
    const int ndays = T / 24;
    double Tzd = na_realS();
#pragma unroll 1
    for (int k = 0; k < T; ++k) {
        const size_t idx = (size_t)k * nc + cell;
        if ((k % 24) == 0) { // daily mean of the ground snow temperature
            Tzd = na_realS();
            if (tzd_ok && k / 24 < ndays) {
                double s = 0.0;
                for (int hh = 0; hh < 24; ++hh) s += a.snowtempg[(size_t)(k + hh) * nc + cell];
                Tzd = s / 24.0;
            }
        }
        const double swe = a.swe[idx];
        if (!(swe > 0.0)) continue;
        const SnowHour h = a.hours[k];
        const double sdepg = a.sdepg[idx];
        const double reqhgts = a.reqhgt - sdepg;
        if (reqhgts >= 0.0) {
            int shadowmask = 1;
            const double ha = a.hor[(size_t)h.sindex * nc + cell];
            const double sa = (kPiS / 2.0) - h.zenr;
            double si = solarindex(slope, aspect, h.zend, h.azid, true);
            if (isnan(si)) si = cos(h.zend * kToRadS);
            if (ha > tan(sa)) shadowmask = 0;
            const double ws = a.wsa[(size_t)h.windex * nc + cell];
            const double sden = a.sden[idx];
            const double sdepc = swe / sden;
            const SnowMicroOut o = snowabove(reqhgts, a.zref, h, hgt, pai, paia, leafd, clump, ltra, leafden, si, svf, shadowmask,
                                             ws, mxtc, a.snowtempg[idx], a.snowtempc[idx], sdepc, sdepg, sden, h.salb, h.salb);
            if (a.out[0]) a.out[0][idx] = o.Tz;
            if (a.out[1]) a.out[1][idx] = o.tleaf;
            if (a.out[2]) a.out[2][idx] = o.rh;
            if (a.out[4]) a.out[4][idx] = o.uz;
            if (a.out[5]) a.out[5][idx] = o.Rbdown;
            if (a.out[6]) a.out[6][idx] = o.Rddown;
            if (a.out[7]) a.out[7][idx] = o.Rlwdn;
            if (a.out[8]) a.out[8][idx] = o.Rdup;
            if (a.out[9]) a.out[9][idx] = o.Rlwup;
        } else {
            const double Tz = below_snow(reqhgts, meanD, a.snowtempg[idx], Tzd, a.mat, a.hiy);
            if (a.out[0]) a.out[0][idx] = Tz;
            if (a.out[1]) a.out[1][idx] = Tz;
            if (a.out[2]) a.out[2][idx] = 100.0;
            if (a.out[4]) a.out[4][idx] = 0.0;
            if (a.out[5]) a.out[5][idx] = 0.0;
            if (a.out[6]) a.out[6][idx] = 0.0;
            if (a.out[7]) a.out[7][idx] = 0.0;
            if (a.out[8]) a.out[8][idx] = 0.0;
            if (a.out[9]) a.out[9][idx] = 0.0;
        }
        if (a.out[3]) a.out[3][idx] = Smax;
    }
}

// ---------------------------------------------------------------------------------------------
// array climate: gridmodelsnow2 (ref :4426-4673) and gridmicrosnow2 (ref :5059-5214)
// ---------------------------------------------------------------------------------------------
// Climate and point-model series are [rows, cols, tsteps] arrays (coalesced per hour); winddir stays a per-hour
// vector.  What the data.frame kernels take from the hour table is formed per cell here: the snow-albedo age scan
// runs along the cell's own precipitation series, the daily radiation extremes are gathered at each day start, the
// solar position comes from the cell's latitude / longitude.
struct SnowArr {
    const int32_t *year, *month, *day;
    const double* hour;
    const double *temp, *relhum, *pres, *swdown, *difrad, *lwdown, *windspeed, *precip; // [nc * T]
    const double* winddir;                                                              // [T]
    const double *Gp, *Tcp, *RswabsG, *RlwabsG, *umu;                                  // [nc * T]
    const double *lats, *lons;                                                          // [nc]
};

__device__ __forceinline__ double snow_albedo(int hs) { // ref snowalbCpp :3765-3769 (integer hs / 24)
    double alb = (-9.8740 * log((double)(hs / 24)) + 78.3434) / 100.0;
    if (alb > 0.95) alb = 0.95;
    if (alb < 0.1) alb = 0.1;
    return alb;
}

__global__ void __launch_bounds__(128) k_snowmodel_arr(const __grid_constant__ SnowModelArgs a, const __grid_constant__ SnowArr c) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    const int nc = a.rows * a.cols;
    if (cell >= nc) return;
    const double hgt0 = a.hgt[cell];
    const double NA = na_realS();
    if (isnan(hgt0)) {
        for (int k = 0; k < a.tsteps; ++k) {
            const size_t idx = (size_t)k * nc + cell;
            a.Tc[idx] = NA; a.Tg[idx] = NA; a.sdepc[idx] = NA; a.sdepg[idx] = NA; a.sden[idx] = NA;
        }
        a.agec[cell] = NA; a.ageg[cell] = NA; a.meltc[cell] = NA; a.meltg[cell] = NA;
        return;
    }
    const double pai0 = a.pai[cell], clump = a.clump[cell], ltra0 = a.ltra[cell];
    const double slope = a.slope[cell], aspect = a.aspect[cell], svf = a.skyview[cell];
    const double lat = c.lats[cell], lon = c.lons[cell];
    const double* sdp = a.sdp;
    int snowagec = a.isnowac[cell], snowageg = a.isnowag[cell];
    double sdepcp = a.isnowdc[cell], sdepgp = a.isnowdg[cell];
    double sdencp = ((sdp[0] - sdp[1]) * (1 - exp(-sdp[2] * sdepcp / 100.0 - sdp[3] * snowagec / 24.0)) + sdp[1]) * 1000.0;
    double sdengp = ((sdp[0] - sdp[1]) * (1 - exp(-sdp[2] * sdepgp * 0.5 / 100.0 - sdp[3] * snowageg / 24.0)) + sdp[1]) * 1000.0;
    double meltc = NA, meltg = NA; // neither is initialised in the array-climate driver (:4487-4488)
    const int ndays = a.tsteps / 24;
    int hs = 0;
    double Rmx = 0, Rmn = 0, Rswmx = 0, Rlwmx = 0, Rswmn = 0, Rlwmn = 0, Gmx = 0;
#pragma unroll 1
    for (int k = 0; k < a.tsteps; ++k) {
        const size_t idx = (size_t)k * nc + cell;
        const double tc = c.temp[idx], prec = c.precip[idx];
        if (k > 0) hs = (prec > 0) ? 0 : hs + 1;
        if ((k % 24) == 0) { // daily extremes of the cell's net radiation (:4503-4566); zero beyond the whole days
            Rmx = Rmn = Rswmx = Rlwmx = Rswmn = Rlwmn = Gmx = 0.0;
            if (k / 24 < ndays) {
                double mx = -1352.0, mn = 1352.0;
                for (int hh = 0; hh < 24; ++hh) {
                    const size_t i2 = (size_t)(k + hh) * nc + cell;
                    const double Rnet = c.RswabsG[i2] + c.RlwabsG[i2] - 0.97 * kSbS * radem(c.temp[i2]);
                    if (mx < Rnet) { mx = Rnet; Rswmx = c.swdown[i2]; Rlwmx = c.lwdown[i2]; }
                    if (mn > Rnet) { mn = Rnet; Rswmn = c.swdown[i2]; Rlwmn = c.lwdown[i2]; }
                    if (fabs(Rnet) > Gmx) Gmx = fabs(Rnet);
                }
                Rmx = mx;
                Rmn = mn;
            }
        }
        int snowtest = 0;
        if (sdepcp > 0.0) snowtest = 1;
        if (tc < 2.0 && prec > 0.0) snowtest = 1;
        if (snowtest > 0) {
            SnowHour h;
            h.tc = tc;
            h.rh = c.relhum[idx];
            h.ea = satvapS(tc) * h.rh / 100.0;
            h.pk = c.pres[idx];
            h.u2 = c.windspeed[idx];
            h.Rsw = c.swdown[idx];
            h.Rdif = c.difrad[idx];
            h.Rlw = c.lwdown[idx];
            h.prec = prec;
            h.Tcp = c.Tcp[idx];
            h.te = (h.Tcp + tc) / 2.0;
            h.Gp = c.Gp[idx];
            h.umu = c.umu[idx];
            h.salb = snow_albedo(hs);
            const SolPos sp = solposition(lat, lon, c.year[k], c.month[k], c.day[k], c.hour[k]);
            h.zend = sp.zend; h.zenr = sp.zenr; h.azid = sp.azid; h.cosz = cos(sp.zenr);
            h.sindex = sector(sp.azid / 15, 24);
            h.windex = sector(c.winddir[k] / 45, 8);
            double paip = pai0;
            if (hgt0 > sdepgp) paip = paip * (hgt0 - sdepgp) / hgt0;
            const double dtR = Rmx - Rmn;
            const double trS = svf * exp(-paip);
            const double Rem = 0.97 * kSbS * radem(tc);
            const double dmxS = trS * Rswmx + trS * Rlwmx + (1 - trS) * Rem - Rem;
            const double dmnS = trS * Rswmn + trS * Rlwmn + (1 - trS) * Rem - Rem;
            const double Gmu = (dmxS - dmnS) / dtR;
            double G = h.Gp * Gmu;
            if (G > Gmx) G = Gmx;
            if (G < -Gmx) G = -Gmx;
            const double ha = a.hor[(size_t)h.sindex * nc + cell];
            const double sa = kPiS / 2.0 - h.zenr; // radians here, degrees in the data.frame driver
            double smu = 1.0;
            if (ha > tan(sa)) smu = 0.0;
            const double ws = a.wsa[(size_t)h.windex * nc + cell];
            const double u2p = h.umu * ws * h.u2;
            const double Rdifp = h.Rdif * svf;
            const double Rdirp = (h.Rsw - h.Rdif) * smu;
            const double Rswp = Rdirp + Rdifp;
            const double Rlwp = h.Rlw * svf;
            const SnowStep st = snow_step(h, Rswp, Rdifp, Rlwp, u2p, G, hgt0, pai0, ltra0, clump, slope, aspect, a.zref, sdp,
                                          sdepcp, sdepgp, sdencp, sdengp, snowagec, snowageg);
            a.Tc[idx] = st.Tc; a.Tg[idx] = st.Tg; a.sdepc[idx] = st.sdepc; a.sdepg[idx] = st.sdepg; a.sden[idx] = st.sdenc;
            sdencp = st.sdenc; sdengp = st.sdeng; sdepcp = st.sdepc; sdepgp = st.sdepg;
            snowagec = (int)st.agec; snowageg = (int)st.ageg;
            meltc = meltc + st.melc;
            meltc = meltc + (st.melc * 1000.0) / st.sdenc;
            meltg = meltg + (st.melg * 1000.0) / st.sdeng;
        } else {
            a.Tc[idx] = 0.0; a.Tg[idx] = 0.0; a.sdepc[idx] = 0.0; a.sdepg[idx] = 0.0; a.sden[idx] = sdp[1] * 1000.0;
        }
    }
    a.agec[cell] = (double)snowagec;
    a.ageg[cell] = (double)snowageg;
    a.meltc[cell] = meltc;
    a.meltg[cell] = meltg;
}

__global__ void __launch_bounds__(128) k_snowmicro_arr(const __grid_constant__ SnowMicroArgs a, const __grid_constant__ SnowArr c) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    const int nc = a.rows * a.cols;
    if (cell >= nc) return;
    const double hgt = a.hgt[cell];
    if (isnan(hgt)) return;
    const int T = a.tsteps;
    double meanD = na_realS();
    if (!isnan(a.sden[cell])) {
        double sumD = 0.0;
        for (int k = 0; k < T; ++k) {
            const double sd = a.sden[(size_t)k * nc + cell];
            const double co = 0.0442 * exp(5.181 * sd / 1000.0);
            sumD += sqrt(2.0 * (co / (sd * 2090.0)) / kOmdyS);
        }
        meanD = sumD / (double)T;
    }
    double mxtc = -273.15; // per cell here (:5138-5143)
    for (int k = 0; k < T; ++k) {
        const double t = c.temp[(size_t)k * nc + cell];
        if (t > mxtc) mxtc = t;
    }
    const bool tzd_ok = !isnan(a.snowtempg[cell]);
    const double pai = a.pai[cell], paia = a.paia[cell], ltra = a.ltra[cell], clump = a.clump[cell];
    const double leafd = a.leafd[cell], leafden = a.leafden[cell];
    const double slope = a.slope[cell], aspect = a.aspect[cell], svf = a.skyview[cell], Smax = a.Smax[cell];
    const double lat = c.lats[cell], lon = c.lons[cell];
    const int ndays = T / 24;
    double Tzd = na_realS();
    int hs = 0;
#pragma unroll 1
    for (int k = 0; k < T; ++k) {
        const size_t idx = (size_t)k * nc + cell;
        if (k > 0) hs = (c.precip[idx] > 0) ? 0 : hs + 1;
        if ((k % 24) == 0) {
            Tzd = na_realS();
            if (tzd_ok && k / 24 < ndays) {
                double s = 0.0;
                for (int hh = 0; hh < 24; ++hh) s += a.snowtempg[(size_t)(k + hh) * nc + cell];
                Tzd = s / 24.0;
            }
        }
        const double swe = a.swe[idx];
        if (!(swe > 0.0)) continue;
        const double sdepg = a.sdepg[idx];
        const double reqhgts = a.reqhgt - sdepg;
        if (reqhgts >= 0.0) {
            SnowHour h;
            h.tc = c.temp[idx]; h.rh = c.relhum[idx]; h.pk = c.pres[idx]; h.u2 = c.windspeed[idx];
            h.Rsw = c.swdown[idx]; h.Rdif = c.difrad[idx]; h.Rlw = c.lwdown[idx]; h.umu = c.umu[idx];
            const SolPos sp = solposition(lat, lon, c.year[k], c.month[k], c.day[k], c.hour[k]);
            h.zend = sp.zend; h.zenr = sp.zenr; h.azid = sp.azid;
            const int sindex = sector(sp.azid / 15, 24);
            const int windex = sector(c.winddir[k] / 45, 8);
            int shadowmask = 1;
            const double ha = a.hor[(size_t)sindex * nc + cell];
            const double sa = (kPiS / 2.0) - sp.zenr;
            double si = solarindex(slope, aspect, sp.zend, sp.azid, true);
            if (isnan(si)) si = cos(sp.zenr);
            if (ha > tan(sa)) shadowmask = 0;
            const double ws = a.wsa[(size_t)windex * nc + cell];
            const double sden = a.sden[idx];
            const double alb = snow_albedo(hs);
            const SnowMicroOut o = snowabove(reqhgts, a.zref, h, hgt, pai, paia, leafd, clump, ltra, leafden, si, svf, shadowmask,
                                             ws, mxtc, a.snowtempg[idx], a.snowtempc[idx], swe / sden, sdepg, sden, alb, alb);
            if (a.out[0]) a.out[0][idx] = o.Tz;
            if (a.out[1]) a.out[1][idx] = o.tleaf;
            if (a.out[2]) a.out[2][idx] = o.rh;
            if (a.out[4]) a.out[4][idx] = o.uz;
            if (a.out[5]) a.out[5][idx] = o.Rbdown;
            if (a.out[6]) a.out[6][idx] = o.Rddown;
            if (a.out[7]) a.out[7][idx] = o.Rlwdn;
            if (a.out[8]) a.out[8][idx] = o.Rdup;
            if (a.out[9]) a.out[9][idx] = o.Rlwup;
        } else {
            const double Tz = below_snow(reqhgts, meanD, a.snowtempg[idx], Tzd, a.mat, a.hiy);
            if (a.out[0]) a.out[0][idx] = Tz;
            if (a.out[1]) a.out[1][idx] = Tz;
            if (a.out[2]) a.out[2][idx] = 100.0;
            if (a.out[4]) a.out[4][idx] = 0.0;
            if (a.out[5]) a.out[5][idx] = 0.0;
            if (a.out[6]) a.out[6][idx] = 0.0;
            if (a.out[7]) a.out[7][idx] = 0.0;
            if (a.out[8]) a.out[8][idx] = 0.0;
            if (a.out[9]) a.out[9][idx] = 0.0;
        }
        if (a.out[3]) a.out[3][idx] = Smax;
    }
}

} // namespace snow
} // namespace mcf

// ---------------------------------------------------------------------------------------------
// host side of the two C-ABI entry points (declared in include/microclimf_b200.h)
// ---------------------------------------------------------------------------------------------
#include <cstdio>
#include <cstring>
#include <vector>

#include "mcf_host.h"
#include "microclimf_b200.h"

namespace {
using namespace mcf::snow;

struct DevBuf { // RAII device allocations of one call, and its queue of host->device uploads
    // Allocations are stream-ordered on the default stream, out of the device's default memory pool (kept between
    // calls: host_prepare_device sets the release threshold).  Uploads are queued and sent in one batch by flush()
    // through the copy pool of mcf_api.cu — R hands us pageable arrays as large as the outputs.
    std::vector<void*> ptrs;
    std::vector<mcf::HostXfer> pend;
    ~DevBuf() {
        for (void* p : ptrs) cudaFreeAsync(p, 0);
    }
    template <class T> cudaError_t up(const T* h, size_t n, const T** d) {
        *d = nullptr;
        if (!h) return cudaSuccess;
        void* q = nullptr;
        cudaError_t e = cudaMallocAsync(&q, n * sizeof(T), 0);
        if (e != cudaSuccess) return e;
        ptrs.push_back(q);
        *d = (const T*)q;
        pend.push_back(mcf::HostXfer{q, h, n * sizeof(T)});
        return cudaSuccess;
    }
    template <class T> cudaError_t alloc(T** d, size_t n) {
        void* q = nullptr;
        cudaError_t e = cudaMallocAsync(&q, n * sizeof(T), 0);
        if (e != cudaSuccess) return e;
        ptrs.push_back(q);
        *d = (T*)q;
        return cudaSuccess;
    }
    // sends the queued uploads; on return they have landed (the copy pool synchronises its streams)
    int flush(char* err, size_t errlen) {
        if (pend.empty()) return MCF_OK;
        cudaError_t e = cudaStreamSynchronize(0); // the allocations are ordered on the default stream
        if (e != cudaSuccess) {
            if (err && errlen) snprintf(err, errlen, "%s", cudaGetErrorString(e));
            return MCF_ERR_CUDA;
        }
        const int rc = mcf::host_transfer(pend.data(), (int)pend.size(), true, err, errlen);
        pend.clear();
        return rc;
    }
};

int fail(int code, const char* msg, char* err, size_t errlen) {
    if (err && errlen) snprintf(err, errlen, "%s", msg);
    return code;
}
#define SCU(x)                                                                           \
    do {                                                                                 \
        cudaError_t e_ = (x);                                                            \
        if (e_ != cudaSuccess) return fail(MCF_ERR_CUDA, cudaGetErrorString(e_), err, errlen); \
    } while (0)

int prep_hours(DevBuf& db, const mcf_snow_climate* c, const double* Gp, const double* Tcp, const double* RswabsG,
               const double* RlwabsG, const double* umu, double lat, double lon, SnowHour** hours, double** scal, char* err,
               size_t errlen) {
    const size_t T = (size_t)c->tsteps;
    SnowPrepArgs pa;
    std::memset(&pa, 0, sizeof pa);
    pa.tsteps = c->tsteps;
    SCU(db.up(c->year, T, &pa.year));
    SCU(db.up(c->month, T, &pa.month));
    SCU(db.up(c->day, T, &pa.day));
    SCU(db.up(c->hour, T, &pa.hour));
    SCU(db.up(c->temp, T, &pa.temp));
    SCU(db.up(c->relhum, T, &pa.relhum));
    SCU(db.up(c->pres, T, &pa.pres));
    SCU(db.up(c->swdown, T, &pa.swdown));
    SCU(db.up(c->difrad, T, &pa.difrad));
    SCU(db.up(c->lwdown, T, &pa.lwdown));
    SCU(db.up(c->windspeed, T, &pa.windspeed));
    SCU(db.up(c->winddir, T, &pa.winddir));
    SCU(db.up(c->precip, T, &pa.precip));
    SCU(db.up(Gp, T, &pa.Gp));
    SCU(db.up(Tcp, T, &pa.Tcp));
    SCU(db.up(RswabsG, T, &pa.RswabsG));
    SCU(db.up(RlwabsG, T, &pa.RlwabsG));
    SCU(db.up(umu, T, &pa.umu));
    pa.lat = lat;
    pa.lon = lon;
    SCU(db.alloc(hours, T));
    SCU(db.alloc(scal, 4));
    pa.hours = *hours;
    pa.scal = *scal;
    {
        const int rc = db.flush(err, errlen);
        if (rc != MCF_OK) return rc;
    }
    k_snow_prep<<<1, 256>>>(pa);
    SCU(cudaGetLastError());
    return MCF_OK;
}

int upload_arr(DevBuf& db, const mcf_snow_climate* c, const double* Gp, const double* Tcp, const double* RswabsG,
               const double* RlwabsG, const double* umu, const mcf_snow_static* st, SnowArr* ca, char* err, size_t errlen) {
    const size_t T = (size_t)c->tsteps, nc = (size_t)st->rows * st->cols, n = nc * T;
    SCU(db.up(c->year, T, &ca->year));
    SCU(db.up(c->month, T, &ca->month));
    SCU(db.up(c->day, T, &ca->day));
    SCU(db.up(c->hour, T, &ca->hour));
    SCU(db.up(c->temp, n, &ca->temp));
    SCU(db.up(c->relhum, n, &ca->relhum));
    SCU(db.up(c->pres, n, &ca->pres));
    SCU(db.up(c->swdown, n, &ca->swdown));
    SCU(db.up(c->difrad, n, &ca->difrad));
    SCU(db.up(c->lwdown, n, &ca->lwdown));
    SCU(db.up(c->windspeed, n, &ca->windspeed));
    SCU(db.up(c->precip, n, &ca->precip));
    SCU(db.up(c->winddir, T, &ca->winddir));
    SCU(db.up(Gp, n, &ca->Gp));
    SCU(db.up(Tcp, n, &ca->Tcp));
    SCU(db.up(RswabsG, n, &ca->RswabsG));
    SCU(db.up(RlwabsG, n, &ca->RlwabsG));
    SCU(db.up(umu, n, &ca->umu));
    SCU(db.up(st->lats, nc, &ca->lats));
    SCU(db.up(st->lons, nc, &ca->lons));
    return MCF_OK;
}
} // namespace

static int gridmodelsnow_impl(bool arr, const mcf_snow_climate* clim, const mcf_snow_point* pt, const mcf_snow_static* st,
                              int32_t snowenv, double* const out3d[5], double* const out2d[4], char* err, size_t errlen) {
    if (!clim || !pt || !st || !out3d || !out2d) return fail(MCF_ERR_ARG, "NULL argument", err, errlen);
    if (st->rows <= 0 || st->cols <= 0 || clim->tsteps <= 0) return fail(MCF_ERR_ARG, "rows, cols, tsteps must be > 0", err, errlen);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(MCF_ERR_CUDA, "no CUDA device (there is no CPU fallback)", err, errlen);
    {
        const int rc0 = mcf::host_prepare_device(err, errlen);
        if (rc0 != MCF_OK) return rc0;
    }
    DevBuf db;
    SnowHour* hours = nullptr;
    double* scal = nullptr;
    const size_t nc = (size_t)st->rows * st->cols, T = (size_t)clim->tsteps;
    SnowArr ca;
    std::memset(&ca, 0, sizeof ca);
    if (arr) {
        if (!st->lats || !st->lons) return fail(MCF_ERR_ARG, "array-climate snow needs lats and lons", err, errlen);
        int rc2 = upload_arr(db, clim, pt->Gp, pt->Tc, pt->RswabsG, pt->RlwabsG, pt->umu, st, &ca, err, errlen);
        if (rc2 != MCF_OK) return rc2;
    } else {
        int rc = prep_hours(db, clim, pt->Gp, pt->Tc, pt->RswabsG, pt->RlwabsG, pt->umu, st->lat, st->lon, &hours, &scal, err, errlen);
        if (rc != MCF_OK) return rc;
    }
    SnowModelArgs a;
    std::memset(&a, 0, sizeof a);
    a.rows = st->rows; a.cols = st->cols; a.tsteps = clim->tsteps; a.hours = hours; a.zref = st->zref;
    // snowdenp (ref :3741-3750)
    static const double kSdp[5][4] = {{0.5975, 0.2237, 0.0012, 0.0038}, {0.5979, 0.2578, 0.001, 0.0038},
                                      {0.594, 0.2332, 0.0016, 0.0031},  {0.363, 0.2425, 0.0029, 0.0049},
                                      {0.217, 0.217, 0.0, 0.0}};
    if (snowenv < 0 || snowenv > 4) return fail(MCF_ERR_ARG, "snowenv must be 0 (Alpine) .. 4 (Taiga)", err, errlen);
    for (int i = 0; i < 4; ++i) a.sdp[i] = kSdp[snowenv][i];
    SCU(db.up(st->pai, nc, &a.pai));
    SCU(db.up(st->hgt, nc, &a.hgt));
    SCU(db.up(st->leaft, nc, &a.ltra));
    SCU(db.up(st->clump, nc, &a.clump));
    SCU(db.up(st->slope, nc, &a.slope));
    SCU(db.up(st->aspect, nc, &a.aspect));
    SCU(db.up(st->skyview, nc, &a.skyview));
    SCU(db.up(st->wsa, nc * 8, &a.wsa));
    SCU(db.up(st->hor, nc * 24, &a.hor));
    SCU(db.up(st->isnowdc, nc, &a.isnowdc));
    SCU(db.up(st->isnowdg, nc, &a.isnowdg));
    SCU(db.up(st->isnowac, nc, &a.isnowac));
    SCU(db.up(st->isnowag, nc, &a.isnowag));
    double* d3[5];
    for (int v = 0; v < 5; ++v) SCU(db.alloc(&d3[v], nc * T));
    double* d2[4];
    for (int v = 0; v < 4; ++v) SCU(db.alloc(&d2[v], nc));
    a.Tc = d3[0]; a.Tg = d3[1]; a.sdepc = d3[2]; a.sdepg = d3[3]; a.sden = d3[4];
    a.agec = d2[0]; a.ageg = d2[1]; a.meltc = d2[2]; a.meltg = d2[3];
    {
        const int rc = db.flush(err, errlen);
        if (rc != MCF_OK) return rc;
    }
    if (arr) k_snowmodel_arr<<<(unsigned)((nc + 127) / 128), 128>>>(a, ca);
    else k_snowmodel<<<(unsigned)((nc + 127) / 128), 128>>>(a);
    SCU(cudaGetLastError());
    SCU(cudaDeviceSynchronize());
    std::vector<mcf::HostXfer> back;
    for (int v = 0; v < 5; ++v)
        if (out3d[v]) back.push_back(mcf::HostXfer{out3d[v], d3[v], nc * T * sizeof(double)});
    for (int v = 0; v < 4; ++v)
        if (out2d[v]) back.push_back(mcf::HostXfer{out2d[v], d2[v], nc * sizeof(double)});
    return mcf::host_transfer(back.data(), (int)back.size(), false, err, errlen);
}

extern "C" int mcf_gridmodelsnow(const mcf_snow_climate* clim, const mcf_snow_point* pt, const mcf_snow_static* st,
                                 int32_t snowenv, double* const out3d[5], double* const out2d[4], char* err, size_t errlen) {
    return gridmodelsnow_impl(false, clim, pt, st, snowenv, out3d, out2d, err, errlen);
}
extern "C" int mcf_gridmodelsnow2(const mcf_snow_climate* clim, const mcf_snow_point* pt, const mcf_snow_static* st,
                                  int32_t snowenv, double* const out3d[5], double* const out2d[4], char* err, size_t errlen) {
    return gridmodelsnow_impl(true, clim, pt, st, snowenv, out3d, out2d, err, errlen);
}

static int gridmicrosnow_impl(bool arr, double reqhgt, const mcf_snow_climate* clim, const double* umu, const mcf_snow_state* sm,
                              const mcf_snow_static* st, double mat, double* const micro[MCF_NOUT], char* err,
                              size_t errlen) {
    if (!clim || !sm || !st || !micro) return fail(MCF_ERR_ARG, "NULL argument", err, errlen);
    if (st->rows <= 0 || st->cols <= 0 || clim->tsteps <= 0) return fail(MCF_ERR_ARG, "rows, cols, tsteps must be > 0", err, errlen);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(MCF_ERR_CUDA, "no CUDA device (there is no CPU fallback)", err, errlen);
    {
        const int rc0 = mcf::host_prepare_device(err, errlen);
        if (rc0 != MCF_OK) return rc0;
    }
    DevBuf db;
    SnowHour* hours = nullptr;
    double* scal = nullptr;
    const size_t nc = (size_t)st->rows * st->cols, T = (size_t)clim->tsteps;
    SnowArr ca;
    std::memset(&ca, 0, sizeof ca);
    if (arr) {
        if (!st->lats || !st->lons) return fail(MCF_ERR_ARG, "array-climate snow needs lats and lons", err, errlen);
        int rc2 = upload_arr(db, clim, nullptr, nullptr, nullptr, nullptr, umu, st, &ca, err, errlen);
        if (rc2 != MCF_OK) return rc2;
    } else {
        int rc = prep_hours(db, clim, nullptr, nullptr, nullptr, nullptr, umu, st->lat, st->lon, &hours, &scal, err, errlen);
        if (rc != MCF_OK) return rc;
    }
    SnowMicroArgs a;
    std::memset(&a, 0, sizeof a);
    a.rows = st->rows; a.cols = st->cols; a.tsteps = clim->tsteps; a.hours = hours; a.scal = scal;
    a.reqhgt = reqhgt; a.zref = st->zref; a.mat = mat;
    const int y0 = clim->year[0];
    a.hiy = (y0 % 4 == 0 && (y0 % 100 != 0 || y0 % 400 == 0)) ? 366 * 24 : 365 * 24;
    SCU(db.up(st->pai, nc, &a.pai));
    SCU(db.up(st->paia, nc, &a.paia));
    SCU(db.up(st->hgt, nc, &a.hgt));
    SCU(db.up(st->leaft, nc, &a.ltra));
    SCU(db.up(st->clump, nc, &a.clump));
    SCU(db.up(st->leafd, nc, &a.leafd));
    SCU(db.up(st->leafden, nc, &a.leafden));
    SCU(db.up(st->slope, nc, &a.slope));
    SCU(db.up(st->aspect, nc, &a.aspect));
    SCU(db.up(st->skyview, nc, &a.skyview));
    SCU(db.up(st->wsa, nc * 8, &a.wsa));
    SCU(db.up(st->hor, nc * 24, &a.hor));
    SCU(db.up(st->Smax, nc, &a.Smax));
    SCU(db.up(sm->Tc, nc * T, &a.snowtempc));
    SCU(db.up(sm->Tg, nc * T, &a.snowtempg));
    SCU(db.up(sm->totalSWE, nc * T, &a.swe));
    SCU(db.up(sm->groundsnowdepth, nc * T, &a.sdepg));
    SCU(db.up(sm->snowden, nc * T, &a.sden));
    for (int v = 0; v < MCF_NOUT; ++v) {
        a.out[v] = nullptr;
        if (micro[v]) {
            const double* d = nullptr;
            SCU(db.up((const double*)micro[v], nc * T, &d));
            a.out[v] = const_cast<double*>(d);
        }
    }
    {
        const int rc = db.flush(err, errlen);
        if (rc != MCF_OK) return rc;
    }
    if (arr) k_snowmicro_arr<<<(unsigned)((nc + 127) / 128), 128>>>(a, ca);
    else k_snowmicro<<<(unsigned)((nc + 127) / 128), 128>>>(a);
    SCU(cudaGetLastError());
    SCU(cudaDeviceSynchronize());
    std::vector<mcf::HostXfer> back;
    for (int v = 0; v < MCF_NOUT; ++v)
        if (micro[v]) back.push_back(mcf::HostXfer{micro[v], a.out[v], nc * T * sizeof(double)});
    return mcf::host_transfer(back.data(), (int)back.size(), false, err, errlen);
}

extern "C" int mcf_gridmicrosnow(double reqhgt, const mcf_snow_climate* clim, const double* umu, const mcf_snow_state* sm,
                                 const mcf_snow_static* st, double mat, double* const micro[MCF_NOUT], char* err,
                                 size_t errlen) {
    return gridmicrosnow_impl(false, reqhgt, clim, umu, sm, st, mat, micro, err, errlen);
}
extern "C" int mcf_gridmicrosnow2(double reqhgt, const mcf_snow_climate* clim, const double* umu, const mcf_snow_state* sm,
                                  const mcf_snow_static* st, double mat, double* const micro[MCF_NOUT], char* err,
                                  size_t errlen) {
    return gridmicrosnow_impl(true, reqhgt, clim, umu, sm, st, mat, micro, err, errlen);
}
