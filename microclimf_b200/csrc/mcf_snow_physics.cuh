// microclimf_b200 — snow physics, re-derived for one-thread-per-cell execution (SURVEY.md NEXT-3).
//
// What the reference computes (src/microclimfCpp.cpp): gridmodelsnow1/2 (:4172-4673) carry a canopy+ground and a
// ground-only snow pack through the hourly series, calling snowoneB (:3835-3972) with radoneB (:3773-3833),
// twostreamdifCpp / twostreamdirCpp / cankCpp (:104-185), PenmanMonteithCpp (:498-514), canopysnowintCpp (:3713-3739)
// and the wind profile helpers every snow hour; gridmicrosnow1/2 (:4894-5214) recompute the microclimate of
// snow-covered cell-hours with snowabovepoint (:4739-4866) / belowpointsnow (:4868-4892).
//
// How it is organised here (this is NOT the reference's call tree):
//   * SnowHr — everything that depends on the hour only is folded once per hour by snow_hour(): the Penman-Monteith
//     air terms (the reference evaluates 3 satvap, 2 pow and the latent-heat polynomial per CELL-hour and call), the
//     dew point, the molar conductance factor, solar geometry as sines / cosines (no per-cell trigonometry), the
//     x = 1 extinction coefficient, the interception capacity of the hour's air temperature, the ground-heat-flux
//     scaling of the day (its radiation extremes collapse to ONE ratio: the emitted-radiation terms of :4341-4344 cancel
//     algebraically), rain-on-snow melt per mm of precipitation.
//   * SnowCell — slope / aspect as the three products the solar index needs, clump^2, log(clump), 1 / (1 - clump).
//   * snow_geometry() — canopy height, plant area, displacement height, roughness and both profile logarithms above the
//     CURRENT ground pack, once per snow hour (the reference derives the same quantities three times per hour; the second
//     logarithm is the first plus log 5).
//   * snow_hour_step() — the energy and mass balance of both packs.  The two Penman-Monteith solves share one
//     denominator; sublimation is (gHa / pk)(es(T) - ea) x 0.018015 x 3.6 — the latent heat cancels between :3898 and :3900;
//     the two-stream solution is specialised to what radoneB uses of it (x = 1, leaf reflectance = ground reflectance =
//     snow albedo: p3, p4, p8..p10 only, exp(+h pait) = 1 / exp(-h pait)); 1 / (2 cos(atan u)) of the interception
//     model is sqrt(1 + u^2) / 2; hours without precipitation skip the interception model (it returns 0 for them).
//   * elementary functions through SM_EXP / SM_LOG / SM_RCP / SM_SQRT: the branch-free MUFU-seeded forms of mcf_math.cuh
//     on the device.  The header also compiles for the HOST (std:: functions) — used only by tests/hostcheck, which
//     checks this algebra against the compiled reference without a GPU; the product library never builds a host path.
// Parity bar 1e-6 (tests/test_snow_gpu.py); the restructuring changes results at rounding level only.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#include "mcf_math.cuh"
#define SNOW_HD __host__ __device__ __forceinline__
#else
#define SNOW_HD inline
#endif

namespace mcf {
namespace snowphys {

#if defined(__CUDA_ARCH__)
#define SM_EXP(x) ::mcf::mexp(x)
#define SM_LOG(x) ::mcf::mlog(x)
#define SM_RCP(x) ::mcf::mrcp(x)
#define SM_SQRT(x) ::mcf::msqrt(x)
#define SM_SINCOS(x, s, c) ::mcf::msincos(x, s, c)
#else
#define SM_SINCOS(x, s, c) (*(s) = sin(x), *(c) = cos(x))
#define SM_EXP(x) exp(x)
#define SM_LOG(x) log(x)
#define SM_RCP(x) (1.0 / (x))
#define SM_SQRT(x) sqrt(x)
#endif

constexpr double kPi = 3.14159265358979323846;
constexpr double kRad = kPi / 180.0;
constexpr double kEmSb = 0.97 * 5.67e-8; // emissivity x Stefan-Boltzmann
constexpr double kKa = 0.4;
constexpr double kLog5 = 1.6094379124341003; // log((zref - d) / (0.2 zm)) - log((zref - d) / zm)

SNOW_HD double pow4(double x) { const double y = x * x; return y * y; }
SNOW_HD double clamp01(double x) { x = (x > 1.0) ? 1.0 : x; return (x < 0.0) ? 0.0 : x; }
// saturated vapour pressure over water / ice (ref satvapCpp :480-490)
SNOW_HD double satvap(double tc) {
    const bool w = tc > 0;
    return 0.61078 * SM_EXP((w ? 17.27 : 21.875) * tc * SM_RCP(tc + (w ? 237.3 : 265.5)));
}
SNOW_HD double satvap_libm(double tc) { // once per hour: plain libm
    return (tc > 0) ? 0.61078 * exp(17.27 * tc / (tc + 237.3)) : 0.61078 * exp(21.875 * tc / (tc + 265.5));
}
SNOW_HD int sector(double x, int n) { // round(x) % n, folded into range for angles outside what checkinputs admits
    const int s = ((int)round(x)) % n;
    return s < 0 ? s + n : s;
}
SNOW_HD double snow_albedo(int hours_since_snow) { // ref snowalbCpp :3765-3769 (the hs / 24 is an INTEGER division)
    double alb = (-9.8740 * log((double)(hours_since_snow / 24)) + 78.3434) / 100.0;
    alb = (alb > 0.95) ? 0.95 : alb;
    return (alb < 0.1) ? 0.1 : alb;
}

// ---------------------------------------------------------------------------------------------------------------------
// per hour
// ---------------------------------------------------------------------------------------------------------------------
struct SnowHr {
    // forcing
    double tc, ea, es, pk, u2, Rsw, Rdif, Rlw, prec, umu, rh;
    double alb;          // snow albedo of the hour (age scan)
    // Penman-Monteith air terms of PenmanMonteithCpp (:498-514) at te = (Tc_point + tc) / 2
    double Rema;         // 0.97 sb (tc + 273.15)^4
    double la_pk;        // latent heat(te) / pk
    double Da;           // es(tc) - ea
    double De;           // es(te + 0.5) - es(te - 0.5)
    double cp;           // molar heat capacity(te)
    double gR;           // radiative conductance
    double tdew;         // dewpointCpp(ea) :493-496
    double gcoef;        // 0.4 x phairCpp(tc, pk) :280: gHa = gcoef uf / log(...)
    double subl;         // 0.018015 x 3.6 / pk: sublimation (m water equivalent per hour) per unit (gHa (es(T) - ea))
    double emTcp;        // 0.97 sb (Tc_point + 273.15)^4: long-wave emitted by the canopy above the pack
    double icap;         // interception capacity per unit plant area, 6.2 (0.26 + 46 / rhos(tc)) (:3727-3729)
    double rainmelt;     // melt per mm of rain on snow, 0.0125 tc / 1000 for tc > 0 (:3921, :3963)
    // ground heat flux: G = clamp(gflux x tr, +-Gmx), tr = sky view x canopy gap (ref :4338-4347)
    double gflux, Gmx;
    // solar geometry
    double cosz, sinz;   // cos / sin of the (unclamped) zenith
    double cazi, sazi;   // cos / sin of the azimuth
    double kcz;          // min(1 / (2 cos zc), 6000) cos zc, zc = min(zenith, pi/2): cankCpp's kd x si for x = 1
    double k1;           // min(1 / (2 cos zc), 6000): cankCpp's k for x = 1
    double tan_alt_deg;  // tan((90 - zenith in degrees) pi / 180): horizon test of gridmodelsnow1 (:4350-4352)
    double tan_alt_rad;  // tan(pi/2 - zenith in radians): the same test as the other three drivers write it
    double zend;         // zenith, degrees
    int32_t sindex, windex;
    int32_t snowing;     // tc < 2 and prec > 0 (:4328)
    int32_t warm;        // tc > 2: precipitation falls as rain (:3965)
    // Penman-Monteith air terms of PenmanMonteith2Cpp (:1220-1247) at tc (snow microclimate)
    double De_tc;        // es(tc + 0.5) - es(tc - 0.5)
    double gr4;          // 4 x 0.97 sb (tc + 273.15)^3 / 29.3
    double la_tc;        // latent heat(tc)
    double pmmu;         // la_tc x 43 / pk
};

struct SolarPos { double zend, zenr, azid; };
// ref solpositionCpp :48-83 with juldayCpp :28-37 and soltimeCpp :39-46 (libm: once per hour, or per cell-hour in the
// array-climate drivers, where it is the reference's own per-cell-hour cost as well)
SNOW_HD SolarPos solar_position(double lat, double lon, int year, int month, int day, double lt) {
    const double dd = day + 0.5;
    const int madj = month + (month < 3) * 12, yadj = year + (month < 3) * -1;
    const double j0 = trunc(365.25 * (yadj + 4716)) + trunc(30.6001 * (madj + 1)) + dd - 1524.5;
    const int b = (int)(2 - trunc((double)(yadj / 100)) + trunc(trunc((double)(yadj / 100)) / 4));
    const int jd = (int)(j0 + (j0 > 2299160) * b);
    const double m = 6.24004077 + 0.01720197 * (jd - 2451545.0);
    const double eot = -7.659 * sin(m) + 9.863 * sin(2 * m + 3.5932);
    const double st = lt + (4.0 * lon + eot) / 60.0;
    const double latr = lat * kPi / 180.0, tt = 0.261799 * (st - 12);
    const double dec = (kPi * 23.5 / 180) * cos(2 * kPi * ((jd - 159.5) / 365.25));
    const double sd = sin(dec), cd = cos(dec), sl = sin(latr), cl = cos(latr), stt = sin(tt), ctt = cos(tt);
    const double coh = sd * sl + cd * cl * ctt;
    SolarPos s;
    s.zend = acos(coh) * (180 / kPi);
    const double hh = atan(coh / sqrt(1 - coh * coh));
    const double sazi = cd * stt / cos(hh);
    const double num = sl * cd * ctt - cl * sd;
    const double cazi = num / sqrt((cd * stt) * (cd * stt) + num * num);
    double sqt = 1 - sazi * sazi;
    sqt = (sqt < 0) ? 0 : sqt;
    double azi = 180 + (180 * atan(sazi / sqrt(sqt))) / kPi;
    if (cazi < 0) azi = (sazi < 0) ? 180 - azi : 540 - azi;
    s.zenr = s.zend * kRad;
    s.azid = azi;
    return s;
}

// The hour's record from its forcing.  `day`: radiation extremes of the hour's day (gridmodelsnow only, else zeros).
struct DayExtremes { double Rmx, Rmn, Rswmx, Rlwmx, Rswmn, Rlwmn, Gmx; };
SNOW_HD void snow_hour(SnowHr& h, double tc, double rh, double pk, double u2, double Rsw, double Rdif, double Rlw, double prec,
                       double Tcp, double Gp, double umu, double winddir, const SolarPos& sp, int hours_since_snow,
                       const DayExtremes& day) {
    h.tc = tc; h.rh = rh; h.pk = pk; h.u2 = u2; h.Rsw = Rsw; h.Rdif = Rdif; h.Rlw = Rlw; h.prec = prec; h.umu = umu;
    h.es = satvap_libm(tc);
    h.ea = h.es * rh / 100.0;
    h.alb = snow_albedo(hours_since_snow);
    const double te = (Tcp + tc) / 2.0;
    h.Rema = kEmSb * pow4(tc + 273.15);
    const double la = (te >= 0) ? 45068.7 - 42.8428 * te : 51078.69 - 4.338 * te - 0.06367 * te * te;
    h.la_pk = la / pk;
    h.Da = h.es - h.ea;
    h.De = satvap_libm(te + 0.5) - satvap_libm(te - 0.5);
    h.cp = 2e-05 * te * te + 0.0002 * te + 29.119; // cpairCpp :287
    const double tk = te + 273.15;
    h.gR = (4.0 * kEmSb * tk * tk * tk) / h.cp;
    const double lea = log(h.ea / 0.6112);
    h.tdew = 243.5 * lea / (17.67 - lea);
    h.gcoef = kKa * (44.6 * (pk / 101.3) * (273.15 / (tc + 273.15)));
    h.subl = 0.018015 * 3.6 / pk;
    h.emTcp = kEmSb * pow4(Tcp + 273.15);
    h.icap = 6.2 * (0.26 + 46 / (67.92 + 51.25 * exp(tc / 2.59)));
    h.rainmelt = (tc > 0.0) ? 0.0125 * tc / 1000 : 0.0;
    // (dmxS - dmnS) / dtR of :4341-4346 = tr x ((Rswmx + Rlwmx) - (Rswmn + Rlwmn)) / (Rmx - Rmn): the (1 - tr) Rem - Rem
    // terms are identical in both and cancel.  0 / 0 (a day of constant net radiation, or hours beyond the whole days)
    // stays NaN as in the reference.
    h.gflux = Gp * (((day.Rswmx + day.Rlwmx) - (day.Rswmn + day.Rlwmn)) / (day.Rmx - day.Rmn));
    h.Gmx = day.Gmx;
    h.zend = sp.zend;
    h.cosz = cos(sp.zenr);
    h.sinz = sin(sp.zenr);
    h.cazi = cos(sp.azid * kRad);
    h.sazi = sin(sp.azid * kRad);
    const double zc = (sp.zenr > kPi / 2.0) ? kPi / 2.0 : sp.zenr;
    double k1 = 1.0 / (2.0 * cos(zc));
    k1 = (k1 > 6000.0) ? 6000.0 : k1;
    h.k1 = k1;
    h.kcz = k1 * cos(zc);
    h.tan_alt_deg = tan((90 - sp.zend) * kRad);
    h.tan_alt_rad = tan(kPi / 2.0 - sp.zenr);
    h.sindex = sector(sp.azid / 15, 24);
    h.windex = sector(winddir / 45, 8);
    h.snowing = (tc < 2.0 && prec > 0.0) ? 1 : 0;
    h.warm = (tc > 2.0) ? 1 : 0;
    h.De_tc = satvap_libm(tc + 0.5) - satvap_libm(tc - 0.5);
    const double tkc = tc + 273.15;
    h.gr4 = (4 * kEmSb * tkc * tkc * tkc) / 29.3;
    h.la_tc = (tc >= 0) ? 45068.7 - 42.8428 * tc : 51078.69 - 4.338 * tc - 0.06367 * tc * tc;
    h.pmmu = h.la_tc * (43.0 / pk);
}

// ---------------------------------------------------------------------------------------------------------------------
// per cell
// ---------------------------------------------------------------------------------------------------------------------
struct SnowCell {
    double hgt0, pai0, inv_hgt0, ltra0;
    double clump, cld, logclump, inv_1mclump;
    double cs, ssca, sssa; // solar index = cosz cs + sinz (cazi ssca + sazi sssa)   (ref solarindexCpp :85-102)
    double svf;
    int flat;              // slope == 0: the index is exactly cos(zenith)
};
SNOW_HD void snow_cell(SnowCell& c, double hgt, double pai, double ltra, double clump, double slope, double aspect, double svf) {
    c.hgt0 = hgt; c.pai0 = pai; c.inv_hgt0 = 1.0 / hgt; c.ltra0 = ltra;
    c.clump = clump; c.cld = clump * clump; c.logclump = log(clump); c.inv_1mclump = 1.0 / (1.0 - clump);
    const double ss = sin(slope * kRad);
    c.cs = cos(slope * kRad);
    c.ssca = ss * cos(aspect * kRad);
    c.sssa = ss * sin(aspect * kRad);
    c.flat = (slope == 0.0) ? 1 : 0;
    c.svf = svf;
}
// cos of the solar incidence angle on the cell's surface; shadowmask = false: 0 once the sun is below the horizon
SNOW_HD double solar_index(const SnowCell& c, const SnowHr& h, bool shadowmask) {
    if (!shadowmask && h.zend > 90.0) return 0.0;
    // cos(azi - aspect) expanded: the reference's single cosine and this sum differ at rounding level
    double si = c.flat ? h.cosz : h.cosz * c.cs + h.sinz * (h.cazi * c.ssca + h.sazi * c.sssa);
    return (si < 0.0) ? 0.0 : si;
}

// ---------------------------------------------------------------------------------------------------------------------
// canopy above the ground pack
// ---------------------------------------------------------------------------------------------------------------------
struct SnowGeom {
    double hgt, pai;   // canopy height and plant area index above the ground pack (0, 0 once buried)
    double paip;       // plant area used by the ground-heat-flux gap fraction: pai0 when the canopy is buried (:4334-4336)
    double d, zm;      // zero-plane displacement, roughness length
    double ln1;        // log((zref - d) / zm)
    double Be;         // sqrt(0.003 + 0.1 pai)
    int above;         // the canopy top is above the ground pack
};
SNOW_HD void snow_geometry(SnowGeom& g, const SnowCell& c, double sdepg, double zref) {
    const bool above = c.hgt0 > sdepg;
    const double frac = (c.hgt0 - sdepg) * c.inv_hgt0;
    g.above = above ? 1 : 0;
    g.pai = above ? c.pai0 * frac : 0.0;
    g.paip = above ? g.pai : c.pai0;
    double hgt = c.hgt0 - sdepg;
    g.hgt = (hgt < 0.0) ? 0.0 : hgt;
    g.Be = SM_SQRT(0.003 + 0.1 * g.pai);
    double d = 0.0, zm = 0.005;
    if (g.hgt > 0.0) {
        // zeroplanedisCpp :294 (plant area floored at 0.001), roughlengthCpp :302 with psi_h = 0
        const double p = (g.pai < 0.001) ? 0.001 : g.pai;
        const double r = SM_SQRT(7.5 * p);
        d = (1.0 - (1.0 - SM_EXP(-r)) * SM_RCP(r)) * g.hgt;
        zm = (g.hgt - d) * SM_EXP(-kKa * SM_RCP(g.Be));
        const double cap = 0.9 * (g.hgt - d);
        zm = (zm > cap) ? cap : zm;
        zm = (zm < 0.0005) ? 0.0005 : zm;
    }
    zm = (zm < 0.0009) ? 0.0009 : zm;
    g.d = d;
    g.zm = zm;
    g.ln1 = SM_LOG((zref - d) * SM_RCP(zm));
}

// ---------------------------------------------------------------------------------------------------------------------
// one hour of both snow packs (ref snowoneB :3835-3972 with umu = 1, psi_m = psi_h = 0)
// ---------------------------------------------------------------------------------------------------------------------
struct SnowState {
    double sdepc, sdepg; // depth of the canopy + ground pack, of the ground-only pack (m)
    double sdenc, sdeng; // their densities (kg / m^3)
    int agec, ageg;      // hours since they were last empty
};
struct SnowStepOut { double Tc, Tg, melc, melg; };

// density of a pack of depth `dep` (m) and age `age` hours in snow environment sdp (ref :3947-3950)
SNOW_HD double pack_density(const double* sdp, double dep, double age) {
    return ((sdp[0] - sdp[1]) * (1.0 - SM_EXP(-sdp[2] * dep / 100.0 - sdp[3] * age / 24.0)) + sdp[1]) * 1000.0;
}

// short-wave absorbed by the ground pack under the canopy: the two-stream solution with x = 1 and leaf reflectance =
// ground reflectance = snow albedo, reduced to the three transmissions radoneB (:3812-3830) takes from it
SNOW_HD double ground_shortwave(const SnowCell& c, const SnowHr& h, double pait, double ltra, double si, double Rbeam, double Rdifp) {
    const double alb = h.alb;
    const double lt = ((alb + ltra) > 0.999) ? 0.999 - alb : ltra;
    const double om = alb + lt, a = 1.0 - om, del3 = (alb - lt) * (1.0 / 3.0);
    const double gma = 0.5 * (om + del3);
    const double hh = SM_SQRT(a * a + 2.0 * a * gma);
    const double S1 = SM_EXP(-hh * pait), iS1 = SM_RCP(S1);
    const double u2 = a + gma * (1.0 - alb);
    const double iD2 = SM_RCP((u2 + hh) * iS1 - (u2 - hh) * S1);
    // diffuse: p3 exp(-h pait) + p4 exp(h pait) with p3 = (u2 + h) / (D2 S1), p4 = -S1 (u2 - h) / D2
    const double Rddm = clamp01((1.0 - c.cld) * (((u2 + hh) * iD2 * iS1) * S1 + (-S1 * (u2 - hh) * iD2) * iS1) + c.cld);
    // direct beam
    const bool dark = (si == 0.0);
    const double isi = SM_RCP(si);
    const double kd = dark ? 1.0 : h.kcz * isi;
    const double Kc = dark ? 600.0 : isi;
    const double apg = a + gma;
    const double sig = kd * kd + gma * gma - apg * apg;
    const double ss = 0.5 * (om + del3 * SM_RCP(kd)) * kd;
    const double sstr = om * kd - ss;
    const double S2 = SM_EXP(-kd * pait);
    const double q = (sstr * (apg + kd) - gma * ss) * SM_RCP(-sig); // p8 / (-sig)
    const double v3 = (sstr + gma * alb - q * (u2 - kd)) * S2;
    const double p9 = -iD2 * ((q * iS1) * (u2 + hh) + v3);
    const double p10 = iD2 * ((q * S1) * (u2 - hh) + v3);
    const double clb = SM_EXP(Kc * c.logclump); // clump^Kc; clump = 0: exp(-inf) = 0
    const double Rdbm = clamp01((1.0 - clb) * (q * S2 + p9 * S1 + p10 * iS1));
    const double Rbgm = clamp01((1.0 - clb) * S2 + clb);
    return (1.0 - alb) * (Rdbm * Rbeam * h.cosz) + Rddm * Rdifp + (1.0 - alb) * (Rbgm * Rbeam * 0.5);
}

// snow intercepted by the canopy in an hour with precipitation (ref canopysnowintCpp :3713-3739)
SNOW_HD double intercepted(const SnowHr& h, double hgt, double pai, double uf, double Li) {
    hgt = (hgt < 0.001) ? 0.001 : hgt;
    pai = (pai < 0.001) ? 0.001 : pai;
    const double Be = SM_SQRT(0.003 + 0.1 * pai);
    const double uh = uf * SM_RCP(Be);
    const double Lm = 2.0 * (Be * Be * Be) * (4.0 * hgt * SM_RCP(pai)); // 2 Be^3 / (0.25 pai / hgt)
    const double k1 = Be * SM_RCP(Lm);
    double uzm = (uh * SM_RCP(hgt * k1)) * (1 - SM_EXP(-k1 * hgt));
    uzm = (uzm < uf) ? uf : uzm;
    const double Lstr = h.icap * pai;
    const double u = uzm * (1.0 / 0.8);
    const double kc = 0.5 * SM_SQRT(1.0 + u * u); // 1 / (2 cos(atan u))
    const double Cp = 1.0 - SM_EXP(-kc * pai);
    const double I1 = (Lstr - Li) * (1.0 - SM_EXP(-(Cp * SM_RCP(Lstr)) * h.prec));
    const double cis = I1 * 0.678;
    return (cis > h.prec) ? h.prec : cis;
}

// `ws`: wind-shelter coefficient of the hour's sector, `ha`: horizon tangent of the hour's solar sector,
// `tan_alt`: the hour's tangent of the solar altitude as the calling driver forms it
SNOW_HD SnowStepOut snow_hour_step(const SnowCell& c, const SnowHr& h, const double* sdp, double zref, double ws, double ha,
                                   double tan_alt, SnowState& s) {
    SnowGeom g;
    snow_geometry(g, c, s.sdepg, zref);
    // ---- what reaches the cell (ref :4348-4360)
    const double smu = (ha > tan_alt) ? 0.0 : 1.0;
    const double u2p = h.umu * ws * h.u2;
    const double Rdifp = h.Rdif * c.svf;
    const double Rswp = (h.Rsw - h.Rdif) * smu + Rdifp;
    const double Rlwp = h.Rlw * c.svf;
    // ---- ground heat flux from the point model's, scaled by the gap fraction (ref :4338-4347)
    const double egap = SM_EXP(-g.paip);
    double G = h.gflux * (c.svf * egap);
    G = (G > h.Gmx) ? h.Gmx : G;
    G = (G < -h.Gmx) ? -h.Gmx : G;
    // ---- absorbed radiation (ref radoneB :3773-3833)
    double zi = 0.0; // water equivalent held per unit canopy height thins the leaves' transmittance (:3843-3845)
    if (s.sdepg > 0.0 && g.hgt > 0.0) zi = ((s.sdepc - s.sdepg) * s.sdenc) * SM_RCP(g.hgt * 1000.0);
    const double ltra = c.ltra0 * SM_EXP(-10.1 * zi);
    const double pait = g.pai * c.inv_1mclump;
    const double tr = (1.0 - c.cld) * SM_EXP(-pait) + c.cld;
    const double RlwabsC = 0.97 * Rlwp;
    double RlwabsG = RlwabsC;
    if (g.hgt > 0.0) RlwabsG = 0.97 * (tr * Rlwp + (1.0 - tr) * h.emTcp);
    double RabsC = RlwabsC, RswabsG = 0.0;
    if (Rswp > 0.0) {
        const double si = solar_index(c, h, false);
        double Rbeam = (Rswp - Rdifp) * SM_RCP(h.cosz);
        Rbeam = (Rbeam > 1352.2) ? 1352.2 : Rbeam;
        const double RswabsC = (1.0 - h.alb) * (Rdifp + Rbeam * h.cosz);
        RabsC = RswabsC + RlwabsC;
        RswabsG = (g.hgt > 0.0) ? ground_shortwave(c, h, pait, ltra, si, Rbeam, Rdifp) : RswabsC;
    }
    const double RabsG = RswabsG + RlwabsG;
    // ---- conductance and the two surface temperatures (ref :3866-3883; PenmanMonteithCpp :498-514 with gV = gHa, erh = 1)
    const double uf = (kKa * u2p) * SM_RCP(g.ln1);
    double gHa = (h.gcoef * uf) * SM_RCP(g.ln1 + kLog5);
    gHa = (gHa < 0.03) ? 0.03 : gHa;
    const double m = h.la_pk * gHa;
    const double iden = SM_RCP(h.cp * (gHa + h.gR) + m * h.De);
    const double sink = h.Rema + m * h.Da + G;
    double Tc = h.tc + (RabsC - sink) * iden;
    double Tg = h.tc + (RabsG - sink) * iden;
    Tc = (Tc < h.tdew) ? h.tdew : Tc;
    Tg = (Tg < h.tdew) ? h.tdew : Tg;
    // ---- mass balance of the canopy + ground pack (ref :3885-3922): sublimation, temperature melt, rain melt
    const double gsub = gHa * h.subl;
    const double mSc = gsub * (satvap(Tc) - h.ea);
    double mMc = 0.0;
    if (Tc > 0.0) {
        mMc = ((583.3 * Tc * (s.sdepc * (s.sdenc / 1000))) / 334000.0) * 3.6;
        if (s.sdepc > 0.0) Tc = 0.0;
    }
    const double mRc = h.rainmelt * h.prec;
    // ---- the ground-only pack (ref :3923-3964): vapour exchange damped by the canopy above it
    const double mu = g.above ? ((egap > 1.0) ? 1.0 : egap) : 1.0; // exp(-pai): the gap fraction again, 1 once buried
    const double mSg = gsub * (satvap(Tg) - h.ea) * mu;
    double mMg = 0.0;
    if (Tg > 0.0) {
        mMg = ((583.3 * Tg * (s.sdepg * (s.sdeng / 1000.0))) / 334000.0) * 3.6;
        if (s.sdepg > 0.0) Tg = 0.0;
    }
    double cis = 0.0;
    if (h.prec > 0.0) { // without precipitation the interception model returns 0 (:3735-3737)
        double Li = 0.0; // snow already held by the canopy, kg / m^2 (:3940-3946)
        if (s.sdepc > 0.0) {
            double w = s.sdepg * SM_RCP(s.sdepc);
            w = clamp01(w);
            Li = (s.sdepc - s.sdepg) * (w * s.sdeng + (1.0 - w) * s.sdenc);
        }
        Li = (Li < 0.0) ? 0.0 : Li;
        cis = intercepted(h, g.hgt, g.pai, uf, Li);
    }
    const double mRg = h.rainmelt * (h.prec - cis);
    const double snowc = h.warm ? 0.0 : h.prec, snowg = h.warm ? 0.0 : h.prec - cis;
    const double melc = mSc + mMc + mRc, melg = mSg + mMg + mRg;
    const double swec = snowc / 1000.0 - melc, sweg = snowg / 1000.0 - melg;
    // ---- age, density and depth (ref :3947-3962)
    double agec = (double)s.agec + 1.0, ageg = (double)s.ageg + 1.0;
    const double denc = pack_density(sdp, s.sdepc, agec), deng = pack_density(sdp, s.sdepg, ageg);
    double depc = s.sdepc + (swec * 1000.0) * SM_RCP(denc);
    double depg = s.sdepg + (sweg * 1000.0) * SM_RCP(deng);
    if (depc < 0.0) { depc = 0.0; agec = 0.0; }
    if (depg < 0.0) { depg = 0.0; ageg = 0.0; }
    s.sdepc = depc; s.sdepg = depg; s.sdenc = denc; s.sdeng = deng;
    s.agec = (int)agec; s.ageg = (int)ageg; // the reference keeps the ages in ints between hours (:4393-4394)
    SnowStepOut o;
    o.Tc = Tc; o.Tg = Tg; o.melc = melc; o.melg = melg;
    return o;
}

// ---------------------------------------------------------------------------------------------------------------------
// microclimate of a snow-covered cell-hour (ref snowabovepoint :4739-4866, belowpointsnow :4868-4892)
// ---------------------------------------------------------------------------------------------------------------------
struct MicroCell {
    double hgt, pai, paia, leafd, inv_leafd, leafden, ltra, clump, inv_hgt, inv_pai, svf, Smax;
    double logclump;   // log(clump): clump^(pais / pai) = exp((pais / pai) log clump)
    double Hf0;        // mincondCpp's Hf for gs = 999.99 (the only conductance leaftemp sees over snow, :4818)
    SnowCell sc;       // slope / aspect products
};
SNOW_HD void micro_cell(MicroCell& m, double hgt, double pai, double paia, double leafd, double leafden, double ltra, double clump,
                        double slope, double aspect, double svf, double Smax) {
    m.hgt = hgt; m.pai = pai; m.paia = paia; m.leafd = leafd; m.inv_leafd = 1.0 / leafd; m.leafden = leafden; m.ltra = ltra;
    m.clump = clump; m.inv_hgt = 1.0 / hgt; m.inv_pai = 1.0 / pai; m.svf = svf; m.Smax = Smax;
    m.logclump = log(clump);
    m.Hf0 = -1.0 / (1.0 + exp(2.0 - 1.09767 * pow(1 / 999.99, 0.2672778)));
    snow_cell(m.sc, hgt, pai, ltra, clump, slope, aspect, svf);
}
struct MicroOut { double Tz, tleaf, rh, uz, Rbdown, Rddown, Rlwdn, Rdup, Rlwup; };

// the z-dependent integral of rhcanopy (:1365-1380) without its uf factor: s / (1 + c) = tan(theta / 2)
SNOW_HD double canopy_integral(double h, double z) {
    if (z == h) return 4.293251 * h;
    double sn, cs;
    SM_SINCOS((kPi * z) * SM_RCP(h), &sn, &cs);
    const double t = sn * SM_RCP(cs + 1.0);
    return (2.0 * h * ((48 * atan(2.23606797749979 * t)) / 11.180339887498949 + (32.0 * t) * SM_RCP(25.0 * t * t + 5.0))) / kPi;
}

// `shadow`: the horizon hides the sun (shadowmask == 0 of :4997-5003).  `si`: solar index WITH shadow mask.
SNOW_HD MicroOut snow_micro_above(const MicroCell& c, const SnowHr& h, double reqhgt, double zref, double si, bool shadow,
                                  double ws, double dTmx, double Tg, double Tc, double sdepc, double sdepg, double sden) {
    if (reqhgt == 0.0) reqhgt = 0.001;
    MicroOut o;
    const double tc = h.tc, ea = h.ea;
    // ---- canopy above the ground pack and the wind profile through it (windtiCpp :1179-1188, windCpp :1189-1218)
    double hgts = c.hgt - sdepg;
    hgts = (hgts < 0.0) ? 0.0 : hgts;
    double pais = 0.0, d = 0.0, zm = 1e-5;
    if (hgts > 0.0) {
        pais = c.pai * hgts * c.inv_hgt;
        const double p = (pais < 0.001) ? 0.001 : pais;
        const double r = SM_SQRT(7.5 * p);
        d = (1.0 - (1.0 - SM_EXP(-r)) * SM_RCP(r)) * hgts;
        zm = (hgts - d) * SM_EXP(-kKa * SM_RCP(SM_SQRT(0.003 + 0.1 * pais)));
        const double cap = 0.9 * (hgts - d);
        zm = (zm > cap) ? cap : zm;
        zm = (zm < 0.0005) ? 0.0005 : zm;
        zm = (zm < 1e-6) ? 1e-6 : zm;
    }
    const double izm = SM_RCP(zm);
    const double ln1 = SM_LOG((zref - d) * izm);          // log((zref - d) / zm)
    const double lnh = ln1 + kLog5;                        // log((zref - d) / zh), zh = 0.2 zm
    ws = (ws != ws) ? 1.0 : ws;
    ws = (ws < 0.05) ? 0.05 : ws;
    double uf = ((kKa * h.u2) * SM_RCP(ln1)) * h.umu * ws;
    uf = (uf < 0.001) ? 0.001 : uf;
    double uz = uf;
    if (reqhgt > 0) {
        if (reqhgt >= hgts) {
            uz = (uf * (1.0 / kKa)) * SM_LOG((reqhgt - d) * izm);
        } else {
            double uh = (uf * (1.0 / kKa)) * SM_LOG((hgts - d) * izm);
            uh = (uh < uf) ? uf : uh;
            double Be = uf * SM_RCP(uh);
            Be = (Be < 0.001) ? 0.001 : Be;
            const double Lm = 2 * (Be * Be * Be) * (4.0 * hgts * SM_RCP(pais)); // 2 Be^3 / (0.25 pais / hgts)
            uz = uh * SM_EXP(Be * (reqhgt - hgts) * SM_RCP(Lm));
        }
        uz = (uz > h.u2) ? h.u2 : uz;
    }
    double gHa = (kKa * 43 * uf) * SM_RCP(lnh);
    gHa = (gHa < 0.0001) ? 0.0001 : gHa;
    o.uz = uz;
    const double zh = 0.2 * zm;
    const double esTc = satvap(Tc);
    const double lwcan = kEmSb * pow4(Tc + 273.15);
    double ez;
    if (reqhgt >= hgts) {
        // ---- above the canopy (or the canopy is buried): the snow surface is all there is below
        o.Rbdown = 0.0; o.Rddown = 0.0; o.Rdup = 0.0;
        if (h.Rsw > 0.0) {
            o.Rddown = h.Rdif * c.svf;
            const bool lit = (si > 0.0) && !shadow;
            double rb = (h.Rsw - h.Rdif) * SM_RCP(si);
            rb = (rb > 1352.0) ? 1352.0 : rb;
            o.Rbdown = lit ? rb : 0.0;
            o.Rdup = h.alb * (lit ? h.Rsw : h.Rdif) * c.svf;
        }
        o.Rlwdn = c.svf * h.Rlw;
        o.Rlwup = c.svf * lwcan;
        // TVabove (:1298-1313) from the canopy + ground pack's temperature, surface wetness 1
        double w = 1.0;
        if (reqhgt > (d + zh)) w = 1 - SM_LOG((reqhgt - d) * SM_RCP(zh)) * SM_RCP(lnh);
        o.Tz = (reqhgt > (d + zh)) ? tc + (Tc - tc) * w : Tc;
        ez = ea + (esTc - ea) * w;
        o.tleaf = Tc;
    } else {
        // ---- inside the canopy above the pack
        const double paias = c.paia * hgts * c.inv_hgt;
        double zi = 0.0;
        if (sdepg > 0.0) zi = ((sdepc - sdepg) * sden) * SM_RCP(hgts * 1000.0);
        const double alb = h.alb; // canopy and ground snow share the hour's albedo (:5007-5008)
        double lt = c.ltra * SM_EXP(-10.1 * zi);
        lt = ((lt + alb) > 0.999) ? 0.999 - alb : lt;
        // clumping rescaled to the exposed part of the canopy; gap fractions above / below the height (:4759-4790)
        const bool clumped = c.clump > 0.0;
        const double lcl = (pais * c.inv_pai) * c.logclump; // log(clumps)
        const double clumps = clumped ? SM_EXP(lcl) : c.clump;
        const double pait = clumped ? pais * SM_RCP(1.0 - clumps) : pais;
        const double ipais = SM_RCP(pais);
        double gi = (clumps > 0.0) ? SM_EXP((paias * ipais) * lcl) : 0.0;
        gi = (gi > 0.99) ? 0.99 : gi;
        double giu = (clumps > 0.0) ? SM_EXP(((pais - paias) * ipais) * lcl) : 0.0;
        giu = (giu > 0.99) ? 0.99 : giu;
        const double trd = gi * gi, trdn = clumps * clumps, trdu = giu * giu;
        const double paiaa = paias * SM_RCP(1.0 - gi);
        const double amx = alb; // max(albg, albc), equal here
        // two-stream, diffuse (twostreamdifCpp :134-162 with x = 1, leaf reflectance = ground reflectance = albedo); the
        // reference solves it twice with identical arguments (:4765-4768)
        const double om = alb + lt, a = 1.0 - om, del3 = (alb - lt) * (1.0 / 3.0);
        const double gma = 0.5 * (om + del3);
        const double hh = SM_SQRT(a * a + 2.0 * a * gma);
        const double S1 = SM_EXP(-hh * pait), iS1 = SM_RCP(S1);
        const double u1 = a + gma * (1.0 - SM_RCP(alb)), u2 = a + gma * (1.0 - alb);
        const double apg = a + gma;
        const double iD1 = SM_RCP((apg + hh) * (u1 - hh) * iS1 - (apg - hh) * (u1 + hh) * S1);
        const double iD2 = SM_RCP((u2 + hh) * iS1 - (u2 - hh) * S1);
        const double p1 = (gma * iD1 * iS1) * (u1 - hh), p2 = (-gma * S1 * iD1) * (u1 + hh);
        const double p3 = (iD2 * iS1) * (u2 + hh), p4 = (-S1 * iD2) * (u2 - hh);
        const double Eh = SM_EXP(-hh * paiaa), iEh = SM_RCP(Eh);
        const double Rddn_z = clamp01((1.0 - trd) * (p3 * Eh + p4 * iEh) + trd);
        const double Rdup_z = clamp01((1.0 - trdu * trdn) * (p1 * Eh + p2 * iEh) + trdu * trdn * alb);
        double Rbdown = 0.0, Rddown = 0.0, Rdup = 0.0, radLsw = 0.0;
        if (h.Rsw > 0.0) {
            const double cosz = h.cosz;
            if (pais > 0.0) {
                // direct beam (cankCpp :104-132, twostreamdirCpp :164-185)
                const bool dark = (si == 0.0);
                const double isi = SM_RCP(si);
                const double kd = dark ? 1.0 : h.kcz * isi;
                const double Kc = dark ? 600.0 : isi;
                const double sig = kd * kd + gma * gma - apg * apg, isig = SM_RCP(sig);
                const double ss = 0.5 * (om + del3 * SM_RCP(kd)) * kd;
                const double sstr = om * kd - ss;
                const double S2 = SM_EXP(-kd * pait);
                const double p5 = -ss * (apg - kd) - gma * sstr;
                const double p5s = p5 * isig;
                const double v1 = ss - p5s * (apg + kd);
                const double v2 = ss - gma - p5s * (u1 + kd);
                const double p6 = iD1 * ((v1 * iS1) * (u1 - hh) - (apg - hh) * S2 * v2);
                const double p7 = -iD1 * ((v1 * S1) * (u1 + hh) - (apg + hh) * S2 * v2);
                const double q = -(sstr * (apg + kd) - gma * ss) * isig; // p8 / (-sig)
                const double v3 = (sstr + gma * alb - q * (u2 - kd)) * S2;
                const double p9 = -iD2 * ((q * iS1) * (u2 + hh) + v3);
                const double p10 = iD2 * ((q * S1) * (u2 - hh) + v3);
                double trbn = SM_EXP(Kc * lcl); // clumps^Kc
                trbn = (trbn > 0.999) ? 0.999 : trbn;
                double trb = (gi > 0.0) ? SM_EXP(Kc * SM_LOG(gi)) : 0.0; // gi^Kc
                trb = (trb > 0.999) ? 0.999 : trb;
                const double Ek = SM_EXP(-kd * paiaa);
                double Rdbup_z = (1.0 - trdu * trbn) * (p5s * Ek + p6 * Eh + p7 * iEh) + trdu * trbn * alb;
                Rdbup_z = (Rdbup_z > amx) ? amx : Rdbup_z;
                Rdbup_z = (Rdbup_z < 0.0) ? 0.0 : Rdbup_z;
                double Rdbdn_z = (1.0 - trb) * (q * Ek + p9 * Eh + p10 * iEh);
                Rdbdn_z = (Rdbdn_z > amx) ? amx : Rdbdn_z;
                Rdbdn_z = (Rdbdn_z < 0.0) ? 0.0 : Rdbdn_z;
                double Rbeam = (h.Rsw - h.Rdif) * SM_RCP(cosz);
                Rbeam = (Rbeam > 1352.0) ? 1352.0 : Rbeam;
                const double Rb = Rbeam * cosz;
                Rbdown = (trb + (1.0 - trb) * Ek) * Rbeam;
                Rddown = Rddn_z * h.Rdif * c.svf + Rdbdn_z * Rb;
                Rdup = Rdup_z * h.Rdif * c.svf + Rdbup_z * Rb;
                radLsw = 0.5 * (1.0 - om) * (Rddown + Rdup + h.k1 * cosz * Rbdown);
            } else {
                Rbdown = (h.Rsw - h.Rdif) / cosz;
                Rddown = h.Rdif * c.svf;
                Rdup = alb * (h.Rdif * c.svf + (h.Rsw - h.Rdif));
            }
        }
        if (shadow) Rbdown = 0.0;
        // ---- leaf temperature (leaftemp :1333-1364 with gsmax = 999.99: no stomatal branch over snow)
        const double lwgro = kEmSb * pow4(Tg + 273.15);
        const double Eg = SM_EXP(-(pais - paias)), Ea = SM_EXP(-paias);
        const double lwup = Eg * lwgro + (1 - Eg) * lwcan;
        const double lwdn = Ea * h.Rlw + (1 - Ea) * lwcan;
        const double leafabs = radLsw + 0.97 * 0.5 * (lwup + lwdn);
        double gh = 0.135 * SM_SQRT(uz * c.inv_leafd) * 1.4;
        // mincondCpp (:1316-1331): 0.0463 (|Hf (leafabs - lwcan)| / leafd)^0.2, floored at 0.05
        const double hmag = fabs(c.Hf0 * (leafabs - lwcan)) * c.inv_leafd;
        double gmin = 0.0463 * SM_EXP(0.2 * SM_LOG(hmag)); // hmag = 0: exp(0.2 x -709) ~ 1e-62, floored below as pow's 0 is
        gmin = (gmin < 0.05) ? 0.05 : gmin;
        gh = (gh < gmin) ? gmin : gh;
        // PenmanMonteith2Cpp (:1220-1247) with gV = gHa = gh, G = 0, surface wetness 1
        const double ml = h.la_tc * (gh / h.pk);
        double dT = (leafabs - h.Rema - ml * (h.es - ea)) * SM_RCP(29.3 * (gh + h.gr4) + ml * h.De_tc);
        dT = (dT > dTmx) ? dTmx : dT;
        dT = (dT > 80.0) ? 80.0 : dT;
        double tleaf = dT + tc;
        tleaf = (tleaf < h.tdew) ? h.tdew : tleaf;
        const double esTl = satvap(tleaf);
        const double lfH = 29.3 * gh * (tleaf - tc), lfL = ml * (esTl - ea);
        o.tleaf = tleaf;
        // ---- state at the canopy top (TVabove at hgts) and the diffusivities below it (TVbelow :1381-1409)
        double w = 1.0;
        if (hgts > (d + zh)) w = 1 - SM_LOG((hgts - d) * SM_RCP(zh)) * SM_RCP(lnh);
        const double Th = (hgts > (d + zh)) ? tc + (Tc - tc) * w : Tc;
        const double eh = ea + (esTc - ea) * w;
        const double a2h = (0.4 * (1.0 - d * SM_RCP(hgts)) / 1.5625) * hgts;
        const double mu_r = SM_RCP(a2h * uf); // (uf / (a2 h)) / uf^2
        double Rc = canopy_integral(hgts, hgts) * mu_r;
        Rc = (Rc < 0.001) ? 0.001 : Rc;
        double Rz = canopy_integral(hgts, reqhgt) * mu_r;
        Rz = (Rz < 0.001) ? 0.001 : Rz;
        const double iKc = Rc * SM_RCP(hgts), Kc_ = SM_RCP(iKc);
        const double Kg = SM_RCP(Rz * reqhgt), Kh = SM_RCP((Rc - Rz) * (hgts - reqhgt));
        const double iK = SM_RCP(Kg + Kh + Kc_);
        const double efac = 1.0 - SM_EXP(-pais);
        const double nearc = 3.047519 + 0.128642 * SM_LOG(pais);
        {
            const double cp = 29.3 * 43.0;
            const double SH = Th * cp, SG = Tg * cp, mxnear = fabs(tleaf - Th) * cp;
            const double SC = SH + ((29.3 * gHa * (Tc - tc)) * efac) * iKc;
            double nearf = nearc * (lfH * c.leafden);
            if (fabs(nearf) > mxnear) nearf = (nearf > 0.0) ? mxnear : -mxnear;
            nearf = (nearf != nearf) ? 0.0 : nearf;
            o.Tz = (nearf + (Kg * SG + Kh * SH + Kc_ * SC) * iK) * (1.0 / cp);
        }
        {
            const double mu = h.pmmu;
            const double SH = eh * mu, SG = satvap(Tg) * mu, mxnear = fabs(esTl - eh) * mu;
            const double SC = SH + (((h.la_tc * (gHa / h.pk)) * (h.es - ea)) * efac) * iKc;
            double nearf = nearc * (lfL * c.leafden);
            if (fabs(nearf) > mxnear) nearf = (nearf > 0.0) ? mxnear : -mxnear;
            nearf = (nearf != nearf) ? 0.0 : nearf;
            ez = (nearf + (Kg * SG + Kh * SH + Kc_ * SC) * iK) * SM_RCP(mu);
        }
        o.Rbdown = Rbdown; o.Rddown = Rddown; o.Rdup = Rdup; o.Rlwdn = lwdn; o.Rlwup = lwup;
    }
    o.rh = (ez * SM_RCP(satvap(o.Tz))) * 100.0;
    o.rh = (o.rh > 100.0) ? 100.0 : o.rh;
    // limits: within 2 K of the extremes of {tleaf, tc, Tg, Tc}, std::max / std::min fold order (:4858-4863)
    double tmx = o.tleaf, tmn = o.tleaf;
    tmx = (tmx < tc) ? tc : tmx; tmx = (tmx < Tg) ? Tg : tmx; tmx = (tmx < Tc) ? Tc : tmx;
    tmn = (tc < tmn) ? tc : tmn; tmn = (Tg < tmn) ? Tg : tmn; tmn = (Tc < tmn) ? Tc : tmn;
    o.Tz = (o.Tz > tmx + 2.0) ? tmx + 2.0 : o.Tz;
    o.Tz = (o.Tz < tmn - 2.0) ? tmn - 2.0 : o.Tz;
    return o;
}

// temperature inside the pack at depth -reqhgts below its surface (ref belowpointsnow :4868-4892): the pack's surface
// temperature, its daily mean and the mean annual temperature blended by the damping depth's time scale
SNOW_HD double snow_micro_below(double reqhgts, double meanD, double Tg, double Tg_daymean, double mat, int hiy) {
    const double nb = -118.35 * reqhgts / meanD;
    if (!(nb > 1.0)) return Tg;
    if (nb <= 24.0) {
        const double w1 = 1.0 / nb, w2 = nb / 24.0, wgt = w1 / (w1 + w2);
        return wgt * Tg + (1 - wgt) * Tg_daymean;
    }
    if (nb <= (double)hiy) {
        const double w1 = 24.0 / nb, w2 = nb / (double)hiy, wgt = w1 / (w1 + w2);
        return wgt * Tg_daymean + (1 - wgt) * mat;
    }
    return mat;
}

} // namespace snowphys
} // namespace mcf
