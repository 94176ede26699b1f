// TEST INFRASTRUCTURE ONLY — not part of the product path.
//
// extern "C" driver over the UNMODIFIED reference translation unit (/root/reference/src/
// microclimfCpp.cpp, compiled in place against oracle/rcpp_shim/Rcpp.h).  It takes the same
// `mcf_problem` description as the product C ABI (include/microclimf_b200.h), rebuilds the
// DataFrame / List arguments the reference drivers expect (names per src/microclimfCpp.cpp:2056-2111
// for modes 1/3 and :2344-2399 for modes 2/4) and copies the returned arrays out.
//
// Used by: tests/ (parity), tests/golden/make_golden.py (fixture generation), bench.py's
// cpu_baseline / --impl reference legs.  Never by the product.
//
// Compiled a second time with -DMCF_GLUE_DRIVER (rglue/Makefile) it drives, with the very same DataFrame / List
// arguments, the twelve functions of rglue/microclimf_glue.cpp — the Rcpp-typed binding a maintainer puts in the
// reference's src/ — instead of the reference's own; the entry points are then called glue_* (tests/test_glue_gpu.py).
#include <Rcpp.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "microclimf_b200.h"
#ifdef MCF_GLUE_DRIVER
#define DRV(name) glue_##name
#else
#define DRV(name) ref_##name
#include "microclimfheaders.h" // reference POD structs, included in place from /root/reference/src
#endif

using namespace Rcpp;

// prototypes of the reference entry points (defined in the reference .cpp; no header exports them)
List runmicro1Cpp(DataFrame obstime, DataFrame climdata, DataFrame pointm, List vegp, List soilc, double reqhgt,
                  double zref, double lat, double lon, double Sminp, double Smaxp, double tfact, bool complete,
                  double mat, std::vector<bool> out);
List runmicro2Cpp(DataFrame obstime, List climdata, List pointm, List vegp, List soilc, double reqhgt, double zref,
                  NumericMatrix lats, NumericMatrix lons, double Sminp, double Smaxp, double tfact, bool complete,
                  double mat, std::vector<bool> out);
List runmicro3Cpp(DataFrame dfsel, DataFrame obstime, DataFrame climdata, DataFrame pointm, List vegp, List soilc,
                  double reqhgt, double zref, double lat, double lon, double Sminp, double Smaxp, double tfact,
                  bool complete, double mat, std::vector<bool> out);
List runmicro4Cpp(DataFrame dfsel, DataFrame obstime, List climdata, List pointm, List vegp, List soilc,
                  double reqhgt, double zref, NumericMatrix lats, NumericMatrix lons, double Sminp, double Smaxp,
                  double tfact, bool complete, double mat, std::vector<bool> out);
List runbioclim1Cpp(DataFrame obstime, DataFrame climdata, DataFrame pointm, List vegp, List soilc, double reqhgt,
                    double zref, double lat, double lon, double Sminp, double Smaxp, double tfact, double mat,
                    std::vector<bool> out, IntegerVector wetq, IntegerVector dryq, IntegerVector hotq,
                    IntegerVector colq, bool air);
List runbioclim2Cpp(DataFrame obstime, List climdata, List pointm, List vegp, List soilc, double reqhgt, double zref,
                    NumericMatrix lats, NumericMatrix lons, double Sminp, double Smaxp, double tfact, double mat,
                    std::vector<bool> out, IntegerVector wetq, IntegerVector dryq, IntegerVector hotq,
                    IntegerVector colq, bool air);
List runbioclim3Cpp(DataFrame obstime, DataFrame climdata, DataFrame pointm, List vegp, List soilc, double reqhgt,
                    double zref, double lat, double lon, double Sminp, double Smaxp, double tfact, double mat,
                    std::vector<bool> out, IntegerVector wetq, IntegerVector dryq, IntegerVector hotq,
                    IntegerVector colq, bool air);
List runbioclim4Cpp(DataFrame obstime, List climdata, List pointm, List vegp, List soilc, double reqhgt, double zref,
                    NumericMatrix lats, NumericMatrix lons, double Sminp, double Smaxp, double tfact, double mat,
                    std::vector<bool> out, IntegerVector wetq, IntegerVector dryq, IntegerVector hotq,
                    IntegerVector colq, bool air);
#ifndef MCF_GLUE_DRIVER
solmodel solpositionCpp(double lat, double lon, int year, int month, int day, double lt);
NumericVector clearskyradCpp(IntegerVector year, IntegerVector month, IntegerVector day, NumericVector lt, double lat,
                             double lon, NumericVector tc, NumericVector rh, NumericVector pk);
std::vector<double> manCpp(std::vector<double> x, int n);
double satvapCpp(double tc);
#endif

namespace {

NumericVector vec(const double* p, size_t n) {
    NumericVector v(n);
    if (p) std::memcpy(v.raw(), p, n * sizeof(double));
    return v;
}
IntegerVector ivec(const int32_t* p, size_t n) {
    IntegerVector v(n);
    for (size_t i = 0; i < n; ++i) v[i] = p[i];
    return v;
}
NumericMatrix mat2(const double* p, int r, int c) {
    NumericMatrix m(r, c);
    if (p) std::memcpy(m.raw(), p, (size_t)r * c * sizeof(double));
    return m;
}
NumericVector arr3(const double* p, int r, int c, int n) {
    NumericVector v((size_t)r * c * n);
    if (p) std::memcpy(v.raw(), p, (size_t)r * c * n * sizeof(double));
    v.set_dims({r, c, n});
    return v;
}

struct Args {
    DataFrame dfsel, obstime;
    List climdata, pointm, vegp, soilc;
    NumericMatrix lats, lons;
};

// swap_types: store the integral-valued columns with the OTHER storage type than the reference reads them with (year /
// month / day as doubles, hour and winddir as integers), as R data.frames often hold them: Rcpp's typed vectors coerce
// on extraction, and so must any binding (tests/test_glue_gpu.py).
Args build(const mcf_problem* p, bool swap_types = false) {
    Args a;
    const bool arrclim = (p->mode == 2 || p->mode == 4);
    const bool layered = (p->mode == 3 || p->mode == 4);
    const int R = p->rows, C = p->cols, T = p->tsteps;
    const size_t nc = (size_t)R * C;
    const size_t nclim = arrclim ? nc * T : (size_t)T;
    a.obstime["year"] = ivec(p->year, T);
    a.obstime["month"] = ivec(p->month, T);
    a.obstime["day"] = ivec(p->day, T);
    a.obstime["hour"] = vec(p->hour, T);
    a.climdata[arrclim ? "tc" : "temp"] = vec(p->temp, nclim);
    a.climdata["es"] = vec(p->es, nclim);
    a.climdata["ea"] = vec(p->ea, nclim);
    a.climdata["tdew"] = vec(p->tdew, nclim);
    a.climdata[arrclim ? "pk" : "pres"] = vec(p->pres, nclim);
    a.climdata["swdown"] = vec(p->swdown, nclim);
    a.climdata["difrad"] = vec(p->difrad, nclim);
    a.climdata["lwdown"] = vec(p->lwdown, nclim);
    a.climdata["windspeed"] = vec(p->windspeed, nclim);
    a.climdata["winddir"] = vec(p->winddir, T);
    if (swap_types) {
        NumericVector y(T), m(T), d(T);
        IntegerVector hr(T), wd(T);
        for (int k = 0; k < T; ++k) {
            y[k] = p->year[k]; m[k] = p->month[k]; d[k] = p->day[k];
            hr[k] = (int)p->hour[k]; wd[k] = (int)p->winddir[k];
        }
        a.obstime["year"] = y; a.obstime["month"] = m; a.obstime["day"] = d;
        a.obstime["hour"] = hr;
        a.climdata["winddir"] = wd;
    }
    a.pointm["soilm"] = vec(p->p_soilm, nclim);
    a.pointm["Tg"] = vec(p->p_Tg, nclim);
    a.pointm["Tbp"] = vec(p->p_Tbp, nclim);
    a.pointm[arrclim ? "Gp" : "G"] = vec(p->p_G, nclim);
    a.pointm["umu"] = vec(p->p_umu, nclim);
    a.pointm["kp"] = vec(p->p_kp, nclim);
    a.pointm["muGp"] = vec(p->p_muGp, nclim);
    a.pointm["dtrp"] = vec(p->p_dtrp, nclim);
    // read by the reference into locals that are never used (src/microclimfCpp.cpp:2075, 2078)
    a.pointm["T0p"] = NumericVector(nclim);
    a.pointm["DDp"] = NumericVector(nclim);
    const int L = layered ? p->nlyr : 1;
    auto veg = [&](const double* q) -> NumericVector {
        if (layered) return arr3(q, R, C, L);
        return mat2(q, R, C);
    };
    a.vegp["hgt"] = veg(p->hgt);
    a.vegp["pai"] = veg(p->pai);
    a.vegp["x"] = veg(p->x);
    a.vegp["gsmax"] = veg(p->gsmax);
    a.vegp["leafr"] = veg(p->leafr);
    a.vegp["leaft"] = veg(p->leaft);
    a.vegp["clump"] = veg(p->clump);
    a.vegp["leafd"] = veg(p->leafd);
    a.vegp["paia"] = veg(p->paia);
    a.vegp["leafden"] = veg(p->leafden);
    a.soilc["Smin"] = mat2(p->Smin, R, C);
    a.soilc["Smax"] = mat2(p->Smax, R, C);
    a.soilc["gref"] = mat2(p->gref, R, C);
    a.soilc["soilb"] = mat2(p->soilb, R, C);
    a.soilc["Psie"] = mat2(p->Psie, R, C);
    a.soilc["Vq"] = mat2(p->Vq, R, C);
    a.soilc["Vm"] = mat2(p->Vm, R, C);
    a.soilc["Mc"] = mat2(p->Mc, R, C);
    a.soilc["rho"] = mat2(p->rho, R, C);
    a.soilc["slope"] = mat2(p->slope, R, C);
    a.soilc["aspect"] = mat2(p->aspect, R, C);
    a.soilc["twi"] = mat2(p->twi, R, C);
    a.soilc["svfa"] = mat2(p->svfa, R, C);
    a.soilc["wsa"] = arr3(p->wsa, R, C, 8);
    a.soilc["hor"] = arr3(p->hor, R, C, 24);
    if (arrclim) {
        a.lats = mat2(p->lats, R, C);
        a.lons = mat2(p->lons, R, C);
    }
    if (layered) {
        IntegerVector lyr(L);
        for (int i = 0; i < L; ++i) lyr[i] = i + 1;
        a.dfsel["lyr"] = lyr;
        a.dfsel["st"] = ivec(p->lyr_st, L);
        a.dfsel["ed"] = ivec(p->lyr_ed, L);
    }
    return a;
}

int fail(char* err, size_t errlen, const std::string& msg, int code) {
    if (err && errlen) std::snprintf(err, errlen, "%s", msg.c_str());
    return code;
}

const char* const kOutNames[MCF_NOUT] = {"Tz",       "tleaf",    "relhum",  "soilm", "windspeed",
                                         "Rdirdown", "Rdifdown", "Rlwdown", "Rswup", "Rlwup"};

} // namespace

static bool g_swap_types = false;
extern "C" void DRV(set_swap_types)(int on) { g_swap_types = on != 0; }

extern "C" int DRV(runmicro)(const mcf_problem* p, double* const out[MCF_NOUT], char* err, size_t errlen) {
    try {
        Args a = build(p, g_swap_types);
        std::vector<bool> o(MCF_NOUT);
        for (int v = 0; v < MCF_NOUT; ++v) o[v] = out[v] != nullptr;
        List res;
        switch (p->mode) {
        case 1:
            res = runmicro1Cpp(a.obstime, a.climdata, a.pointm, a.vegp, a.soilc, p->reqhgt, p->zref, p->lat, p->lon,
                               p->Sminp, p->Smaxp, p->tfact, p->complete != 0, p->mat, o);
            break;
        case 2:
            res = runmicro2Cpp(a.obstime, a.climdata, a.pointm, a.vegp, a.soilc, p->reqhgt, p->zref, a.lats, a.lons,
                               p->Sminp, p->Smaxp, p->tfact, p->complete != 0, p->mat, o);
            break;
        case 3:
            res = runmicro3Cpp(a.dfsel, a.obstime, a.climdata, a.pointm, a.vegp, a.soilc, p->reqhgt, p->zref, p->lat,
                               p->lon, p->Sminp, p->Smaxp, p->tfact, p->complete != 0, p->mat, o);
            break;
        case 4:
            res = runmicro4Cpp(a.dfsel, a.obstime, a.climdata, a.pointm, a.vegp, a.soilc, p->reqhgt, p->zref, a.lats,
                               a.lons, p->Sminp, p->Smaxp, p->tfact, p->complete != 0, p->mat, o);
            break;
        default:
            return fail(err, errlen, "ref_runmicro: mode must be 1..4", MCF_ERR_ARG);
        }
        const size_t n = (size_t)p->rows * p->cols * p->tsteps;
        for (int v = 0; v < MCF_NOUT; ++v) {
            if (!out[v]) continue;
            NumericVector r = res[kOutNames[v]];
            std::memcpy(out[v], r.raw(), n * sizeof(double));
        }
    } catch (const std::exception& e) {
        return fail(err, errlen, e.what(), MCF_ERR_ARG);
    }
    return MCF_OK;
}

extern "C" int DRV(runbioclim)(const mcf_problem* p, const int32_t* wetq, int32_t nwetq, const int32_t* dryq,
                              int32_t ndryq, const int32_t* hotq, int32_t nhotq, const int32_t* colq, int32_t ncolq,
                              int32_t air, double* const bio[MCF_NBIO], char* err, size_t errlen) {
    try {
        Args a = build(p);
        std::vector<bool> o(MCF_NBIO);
        for (int v = 0; v < MCF_NBIO; ++v) o[v] = bio[v] != nullptr;
        IntegerVector wq = ivec(wetq, nwetq), dq = ivec(dryq, ndryq), hq = ivec(hotq, nhotq), cq = ivec(colq, ncolq);
        List res;
        switch (p->mode) {
        case 1:
            res = runbioclim1Cpp(a.obstime, a.climdata, a.pointm, a.vegp, a.soilc, p->reqhgt, p->zref, p->lat, p->lon,
                                 p->Sminp, p->Smaxp, p->tfact, p->mat, o, wq, dq, hq, cq, air != 0);
            break;
        case 2:
            res = runbioclim2Cpp(a.obstime, a.climdata, a.pointm, a.vegp, a.soilc, p->reqhgt, p->zref, a.lats, a.lons,
                                 p->Sminp, p->Smaxp, p->tfact, p->mat, o, wq, dq, hq, cq, air != 0);
            break;
        case 3:
            res = runbioclim3Cpp(a.obstime, a.climdata, a.pointm, a.vegp, a.soilc, p->reqhgt, p->zref, p->lat, p->lon,
                                 p->Sminp, p->Smaxp, p->tfact, p->mat, o, wq, dq, hq, cq, air != 0);
            break;
        case 4:
            res = runbioclim4Cpp(a.obstime, a.climdata, a.pointm, a.vegp, a.soilc, p->reqhgt, p->zref, a.lats, a.lons,
                                 p->Sminp, p->Smaxp, p->tfact, p->mat, o, wq, dq, hq, cq, air != 0);
            break;
        default:
            return fail(err, errlen, "ref_runbioclim: mode must be 1..4", MCF_ERR_ARG);
        }
        const size_t n = (size_t)p->rows * p->cols;
        for (int v = 0; v < MCF_NBIO; ++v) {
            if (!bio[v]) continue;
            NumericMatrix r = res["bio" + std::to_string(v + 1)];
            std::memcpy(bio[v], r.raw(), n * sizeof(double));
        }
    } catch (const std::exception& e) {
        return fail(err, errlen, e.what(), MCF_ERR_ARG);
    }
    return MCF_OK;
}

#ifndef MCF_GLUE_DRIVER
// Small helpers of the reference exposed for fixture generation and unit checks of the restatement.
extern "C" void ref_clearskyrad(int32_t n, const int32_t* year, const int32_t* month, const int32_t* day,
                                const double* lt, double lat, double lon, const double* tc, const double* rh,
                                const double* pk, double* out) {
    NumericVector r = clearskyradCpp(ivec(year, n), ivec(month, n), ivec(day, n), vec(lt, n), lat, lon, vec(tc, n),
                                     vec(rh, n), vec(pk, n));
    std::memcpy(out, r.raw(), (size_t)n * sizeof(double));
}

extern "C" void ref_solposition(double lat, double lon, int32_t year, int32_t month, int32_t day, double lt,
                                double* zend_azid /* [2] */) {
    solmodel s = solpositionCpp(lat, lon, year, month, day, lt);
    zend_azid[0] = s.zend;
    zend_azid[1] = s.azid;
}

extern "C" void ref_man(const double* x, int32_t m, int32_t n, double* out) {
    std::vector<double> r = manCpp(std::vector<double>(x, x + m), n);
    std::memcpy(out, r.data(), (size_t)m * sizeof(double));
}

extern "C" double ref_satvap(double tc) { return satvapCpp(tc); }

// ---------------------------------------------------------------------------------------------
// Point model and terrain helpers of the reference (upstream of the grid solver), exposed so that
// tools/make_bundled_fixtures.py can build a `micropoint` for the bundled example data the way
// runpointmodel does (R/Cppwrappers.R:59-148), and so that mcf_flowacc has a compiled-reference checker.
// ---------------------------------------------------------------------------------------------
DataFrame weatherhgtCpp(DataFrame obstime, DataFrame climdata, double zin, double uzin, double zout, double lat,
                        double lon);
std::vector<double> soilmCpp(DataFrame climdata, double rmu, double mult, double pwr, double Smax, double Smin,
                             double Ksat, double a);
Rcpp::List BigLeafCpp(DataFrame obstime, DataFrame climdata, std::vector<double> vegp, std::vector<double> groundp,
                      std::vector<double> soilm, double lat, double lon, double dTmx, double zref, int maxiter,
                      double bwgt, double tol, double gmn, bool yearG);
DataFrame pointmprocess(DataFrame pointvars, double zref, double h, double pai, double rho, double Vm, double Vq,
                        double Mc);
NumericMatrix flowaccCpp(NumericMatrix dm);

#endif // !MCF_GLUE_DRIVER
namespace {
DataFrame obstime_df(int32_t n, const int32_t* year, const int32_t* month, const int32_t* day, const double* hour) {
    DataFrame o;
    o["year"] = ivec(year, n);
    o["month"] = ivec(month, n);
    o["day"] = ivec(day, n);
    o["hour"] = vec(hour, n);
    return o;
}
#ifndef MCF_GLUE_DRIVER
// weather: 9 columns of length n in the order temp, relhum, pres, swdown, difrad, lwdown, windspeed, winddir, precip
const char* const kWeatherCols[9] = {"temp", "relhum", "pres", "swdown", "difrad", "lwdown", "windspeed", "winddir", "precip"};
DataFrame weather_df(int32_t n, const double* w) {
    DataFrame c;
    for (int k = 0; k < 9; ++k) c[kWeatherCols[k]] = vec(w + (size_t)k * n, n);
    return c;
}
#endif
} // namespace
#ifndef MCF_GLUE_DRIVER

// weatherhgtCpp: writes the adjusted temp / relhum / windspeed columns (3 x n)
extern "C" int ref_weatherhgt(int32_t n, const int32_t* year, const int32_t* month, const int32_t* day,
                              const double* hour, const double* weather, double zin, double uzin, double zout,
                              double lat, double lon, double* out3) {
    try {
        DataFrame r = weatherhgtCpp(obstime_df(n, year, month, day, hour), weather_df(n, weather), zin, uzin, zout, lat, lon);
        const char* cols[3] = {"temp", "relhum", "windspeed"};
        for (int k = 0; k < 3; ++k) {
            std::vector<double> v = r[cols[k]];
            std::memcpy(out3 + (size_t)k * n, v.data(), (size_t)n * sizeof(double));
        }
        return 0;
    } catch (...) {
        return 1;
    }
}

// soilmCpp: daily soil moisture, n / 24 values
extern "C" int ref_soilm(int32_t n, const double* weather, double rmu, double mult, double pwr, double Smax, double Smin,
                         double Ksat, double a, double* out, int32_t* nout) {
    try {
        std::vector<double> v = soilmCpp(weather_df(n, weather), rmu, mult, pwr, Smax, Smin, Ksat, a);
        *nout = (int32_t)v.size();
        std::memcpy(out, v.data(), v.size() * sizeof(double));
        return 0;
    } catch (...) {
        return 1;
    }
}

// BigLeafCpp: out6 = Tc, Tg, G, uf, RabsG, psih (6 x n)
extern "C" int ref_bigleaf(int32_t n, const int32_t* year, const int32_t* month, const int32_t* day, const double* hour,
                           const double* weather, const double* vegp, int32_t nvegp, const double* groundp,
                           int32_t ngroundp, const double* soilm, double lat, double lon, double dTmx, double zref,
                           int32_t maxiter, double bwgt, double tol, double gmn, int32_t yearG, double* out6) {
    try {
        Rcpp::List r = BigLeafCpp(obstime_df(n, year, month, day, hour), weather_df(n, weather),
                                  std::vector<double>(vegp, vegp + nvegp), std::vector<double>(groundp, groundp + ngroundp),
                                  std::vector<double>(soilm, soilm + n), lat, lon, dTmx, zref, maxiter, bwgt, tol, gmn,
                                  yearG != 0);
        const char* cols[6] = {"Tc", "Tg", "G", "uf", "RabsG", "psih"};
        for (int k = 0; k < 6; ++k) {
            std::vector<double> v = Rcpp::as<std::vector<double>>(r[cols[k]]);
            std::memcpy(out6 + (size_t)k * n, v.data(), (size_t)n * sizeof(double));
        }
        return 0;
    } catch (...) {
        return 1;
    }
}

// pointmprocess: in7 = windspeed, tc, rh, pk, uf, soilm, RabsG (7 x n); out6 = umu, kp, muGp, DDp, T0p, dtrp (6 x n)
extern "C" int ref_pointmprocess(int32_t n, const double* in7, double zref, double h, double pai, double rho, double Vm,
                                 double Vq, double Mc, double* out6) {
    try {
        const char* icols[7] = {"windspeed", "tc", "rh", "pk", "uf", "soilm", "RabsG"};
        DataFrame p;
        for (int k = 0; k < 7; ++k) p[icols[k]] = vec(in7 + (size_t)k * n, n);
        DataFrame r = pointmprocess(p, zref, h, pai, rho, Vm, Vq, Mc);
        const char* cols[6] = {"umu", "kp", "muGp", "DDp", "T0p", "dtrp"};
        for (int k = 0; k < 6; ++k) {
            std::vector<double> v = r[cols[k]];
            std::memcpy(out6 + (size_t)k * n, v.data(), (size_t)n * sizeof(double));
        }
        return 0;
    } catch (...) {
        return 1;
    }
}

// flowaccCpp on a column-major [rows, cols] matrix
extern "C" int ref_flowacc(const double* dtm, int32_t rows, int32_t cols, double* fa) {
    try {
        NumericMatrix m(rows, cols);
        for (size_t i = 0; i < (size_t)rows * cols; ++i) m[i] = dtm[i];
        NumericMatrix r = flowaccCpp(m);
        for (size_t i = 0; i < (size_t)rows * cols; ++i) fa[i] = r[i];
        return 0;
    } catch (...) {
        return 1;
    }
}

// canintfrac (src/microclimfCpp.cpp:5417) and meltmu (:5454) of the quick snow model, on column-major matrices
NumericMatrix canintfrac(NumericMatrix hgt, NumericMatrix pai, double uf, double prec, double tc, double Li);
NumericMatrix meltmu(NumericMatrix skyview, NumericVector stemp, NumericVector tc);
extern "C" int ref_canintfrac(const double* hgt, const double* pai, int32_t rows, int32_t cols, double uf, double prec, double tc,
                              double Li, double* out) {
    try {
        NumericMatrix h(rows, cols), p(rows, cols);
        for (size_t i = 0; i < (size_t)rows * cols; ++i) { h[i] = hgt[i]; p[i] = pai[i]; }
        NumericMatrix r = canintfrac(h, p, uf, prec, tc, Li);
        for (size_t i = 0; i < (size_t)rows * cols; ++i) out[i] = r[i];
        return 0;
    } catch (...) {
        return 1;
    }
}
extern "C" int ref_meltmu(const double* skyview, int32_t rows, int32_t cols, const double* stemp, const double* tc, int32_t n,
                          double* out) {
    try {
        NumericMatrix sv(rows, cols);
        for (size_t i = 0; i < (size_t)rows * cols; ++i) sv[i] = skyview[i];
        NumericVector st(n), t(n);
        for (int i = 0; i < n; ++i) { st[i] = stemp[i]; t[i] = tc[i]; }
        NumericMatrix r = meltmu(sv, st, t);
        for (size_t i = 0; i < (size_t)rows * cols; ++i) out[i] = r[i];
        return 0;
    } catch (...) {
        return 1;
    }
}

NumericMatrix meltmu2(NumericMatrix mu, NumericVector stemp, NumericVector tc);
// meltmu2 (src/microclimfCpp.cpp:5495): stemp / tc are column-major [rows, cols, n] arrays
extern "C" int ref_meltmu2(const double* mu, int32_t rows, int32_t cols, const double* stemp, const double* tc, int32_t n,
                           double* out) {
    try {
        NumericMatrix m(rows, cols);
        for (size_t i = 0; i < (size_t)rows * cols; ++i) m[i] = mu[i];
        const size_t len = (size_t)rows * cols * n;
        NumericVector st(len), t(len);
        for (size_t i = 0; i < len; ++i) { st[i] = stemp[i]; t[i] = tc[i]; }
        IntegerVector dim = {rows, cols, n};
        st.attr("dim") = dim;
        t.attr("dim") = dim;
        NumericMatrix r = meltmu2(m, st, t);
        for (size_t i = 0; i < (size_t)rows * cols; ++i) out[i] = r[i];
        return 0;
    } catch (...) {
        return 1;
    }
}

#endif // !MCF_GLUE_DRIVER
// ---------------------------------------------------------------------------------------------
// Snow (SURVEY.md NEXT-3): the reference's gridmodelsnow1 / gridmicrosnow1 behind the product's structs
// ---------------------------------------------------------------------------------------------
List gridmodelsnow1(DataFrame obstime, DataFrame climdata, DataFrame pointm, List vegp, List other, std::string snowenv);
List gridmicrosnow1(double reqhgt, DataFrame obstime, DataFrame climdata, List snowm, List micro, List vegp, List other,
                    double mat, std::vector<bool> out);

namespace {
IntegerMatrix imat2(const int32_t* p, int rows, int cols) {
    IntegerMatrix m(rows, cols);
    for (size_t i = 0; i < (size_t)rows * cols; ++i) m[i] = p[i];
    return m;
}
DataFrame snow_clim_df(const mcf_snow_climate* c) {
    const int n = c->tsteps;
    DataFrame d;
    d["temp"] = vec(c->temp, n); d["relhum"] = vec(c->relhum, n); d["pres"] = vec(c->pres, n);
    d["swdown"] = vec(c->swdown, n); d["difrad"] = vec(c->difrad, n); d["lwdown"] = vec(c->lwdown, n);
    d["windspeed"] = vec(c->windspeed, n); d["winddir"] = vec(c->winddir, n); d["precip"] = vec(c->precip, n);
    return d;
}
const char* const kSnowEnv[5] = {"Alpine", "Maritime", "Prairie", "Tundra", "Taiga"};
} // namespace

extern "C" int DRV(gridmodelsnow)(const mcf_snow_climate* c, const mcf_snow_point* pt, const mcf_snow_static* st,
                                 int32_t snowenv, double* const out3d[5], double* const out2d[4], char* err, size_t errlen) {
    try {
        const int n = c->tsteps, R = st->rows, C = st->cols;
        DataFrame pm;
        pm["Gp"] = vec(pt->Gp, n); pm["Tc"] = vec(pt->Tc, n); pm["RswabsG"] = vec(pt->RswabsG, n);
        pm["RlwabsG"] = vec(pt->RlwabsG, n); pm["umu"] = vec(pt->umu, n); pm["tr"] = vec(pt->umu, n);
        List vegp;
        vegp["pai"] = mat2(st->pai, R, C); vegp["hgt"] = mat2(st->hgt, R, C); vegp["leaft"] = mat2(st->leaft, R, C);
        vegp["clump"] = mat2(st->clump, R, C);
        List other;
        other["slope"] = mat2(st->slope, R, C); other["aspect"] = mat2(st->aspect, R, C);
        other["skyview"] = mat2(st->skyview, R, C); other["wsa"] = arr3(st->wsa, R, C, 8); other["hor"] = arr3(st->hor, R, C, 24);
        other["lat"] = st->lat; other["lon"] = st->lon; other["zref"] = st->zref;
        other["isnowdc"] = mat2(st->isnowdc, R, C); other["isnowdg"] = mat2(st->isnowdg, R, C);
        other["isnowac"] = imat2(st->isnowac, R, C); other["isnowag"] = imat2(st->isnowag, R, C);
        List r = gridmodelsnow1(obstime_df(n, c->year, c->month, c->day, c->hour), snow_clim_df(c), pm, vegp, other,
                                kSnowEnv[snowenv]);
        const char* n3[5] = {"Tc", "Tg", "sdepc", "sdepg", "sden"};
        for (int v = 0; v < 5; ++v)
            if (out3d[v]) {
                NumericVector a = r[n3[v]];
                std::memcpy(out3d[v], a.raw(), (size_t)R * C * n * sizeof(double));
            }
        const char* n2[4] = {"agec", "ageg", "meltc", "meltg"};
        for (int v = 0; v < 4; ++v)
            if (out2d[v]) {
                NumericMatrix a = r[n2[v]];
                for (size_t i = 0; i < (size_t)R * C; ++i) out2d[v][i] = a[i];
            }
        return MCF_OK;
    } catch (const std::exception& e) {
        std::snprintf(err, errlen, "%s", e.what());
        return MCF_ERR_ARG;
    }
}

extern "C" int DRV(gridmicrosnow)(double reqhgt, const mcf_snow_climate* c, const double* umu, const mcf_snow_state* sm,
                                 const mcf_snow_static* st, double mat, double* const micro[MCF_NOUT], char* err,
                                 size_t errlen) {
    try {
        const int n = c->tsteps, R = st->rows, C = st->cols;
        DataFrame clim = snow_clim_df(c);
        clim["umu"] = vec(umu, n);
        List snowm;
        snowm["Tc"] = arr3(sm->Tc, R, C, n); snowm["Tg"] = arr3(sm->Tg, R, C, n); snowm["totalSWE"] = arr3(sm->totalSWE, R, C, n);
        snowm["groundsnowdepth"] = arr3(sm->groundsnowdepth, R, C, n); snowm["snowden"] = arr3(sm->snowden, R, C, n);
        static const char* nm[MCF_NOUT] = {"Tz", "tleaf", "relhum", "soilm", "windspeed", "Rdirdown", "Rdifdown", "Rlwdown",
                                           "Rswup", "Rlwup"};
        List mic;
        std::vector<bool> o(MCF_NOUT);
        for (int v = 0; v < MCF_NOUT; ++v) {
            o[v] = micro[v] != nullptr;
            if (micro[v]) mic[nm[v]] = arr3(micro[v], R, C, n);
        }
        List vegp;
        vegp["pai"] = mat2(st->pai, R, C); vegp["paia"] = mat2(st->paia, R, C); vegp["hgt"] = mat2(st->hgt, R, C);
        vegp["leaft"] = mat2(st->leaft, R, C); vegp["clump"] = mat2(st->clump, R, C); vegp["leafd"] = mat2(st->leafd, R, C);
        vegp["leafden"] = mat2(st->leafden, R, C);
        List other;
        other["slope"] = mat2(st->slope, R, C); other["aspect"] = mat2(st->aspect, R, C);
        other["skyview"] = mat2(st->skyview, R, C); other["wsa"] = arr3(st->wsa, R, C, 8); other["hor"] = arr3(st->hor, R, C, 24);
        other["lat"] = st->lat; other["lon"] = st->lon; other["zref"] = st->zref; other["Smax"] = mat2(st->Smax, R, C);
        List r = gridmicrosnow1(reqhgt, obstime_df(n, c->year, c->month, c->day, c->hour), clim, snowm, mic, vegp, other, mat, o);
        for (int v = 0; v < MCF_NOUT; ++v)
            if (micro[v]) {
                NumericVector a = r[nm[v]];
                std::memcpy(micro[v], a.raw(), (size_t)R * C * n * sizeof(double));
            }
        return MCF_OK;
    } catch (const std::exception& e) {
        std::snprintf(err, errlen, "%s", e.what());
        return MCF_ERR_ARG;
    }
}

// array-climate snow drivers
List gridmodelsnow2(DataFrame obstime, List climdata, List pointm, List vegp, List other, std::string snowenv);
List gridmicrosnow2(double reqhgt, DataFrame obstime, List climdata, List snowm, List micro, List vegp, List other, double mat,
                    std::vector<bool> out);

extern "C" int DRV(gridmodelsnow2)(const mcf_snow_climate* c, const mcf_snow_point* pt, const mcf_snow_static* st,
                                  int32_t snowenv, double* const out3d[5], double* const out2d[4], char* err, size_t errlen) {
    try {
        const int n = c->tsteps, R = st->rows, C = st->cols;
        List clim;
        clim["temp"] = arr3(c->temp, R, C, n); clim["relhum"] = arr3(c->relhum, R, C, n); clim["pres"] = arr3(c->pres, R, C, n);
        clim["swdown"] = arr3(c->swdown, R, C, n); clim["difrad"] = arr3(c->difrad, R, C, n); clim["lwdown"] = arr3(c->lwdown, R, C, n);
        clim["windspeed"] = arr3(c->windspeed, R, C, n); clim["winddir"] = vec(c->winddir, n); clim["precip"] = arr3(c->precip, R, C, n);
        List pm;
        pm["Gp"] = arr3(pt->Gp, R, C, n); pm["Tc"] = arr3(pt->Tc, R, C, n); pm["RswabsG"] = arr3(pt->RswabsG, R, C, n);
        pm["RlwabsG"] = arr3(pt->RlwabsG, R, C, n); pm["umu"] = arr3(pt->umu, R, C, n); pm["tr"] = arr3(pt->umu, R, C, n);
        List vegp;
        vegp["pai"] = mat2(st->pai, R, C); vegp["hgt"] = mat2(st->hgt, R, C); vegp["leaft"] = mat2(st->leaft, R, C);
        vegp["clump"] = mat2(st->clump, R, C);
        List other;
        other["slope"] = mat2(st->slope, R, C); other["aspect"] = mat2(st->aspect, R, C);
        other["skyview"] = mat2(st->skyview, R, C); other["wsa"] = arr3(st->wsa, R, C, 8); other["hor"] = arr3(st->hor, R, C, 24);
        other["lats"] = mat2(st->lats, R, C); other["lons"] = mat2(st->lons, R, C); other["zref"] = st->zref;
        other["isnowdc"] = mat2(st->isnowdc, R, C); other["isnowdg"] = mat2(st->isnowdg, R, C);
        other["isnowac"] = imat2(st->isnowac, R, C); other["isnowag"] = imat2(st->isnowag, R, C);
        List r = gridmodelsnow2(obstime_df(n, c->year, c->month, c->day, c->hour), clim, pm, vegp, other, kSnowEnv[snowenv]);
        const char* n3[5] = {"Tc", "Tg", "sdepc", "sdepg", "sden"};
        for (int v = 0; v < 5; ++v)
            if (out3d[v]) {
                NumericVector a = r[n3[v]];
                std::memcpy(out3d[v], a.raw(), (size_t)R * C * n * sizeof(double));
            }
        const char* n2[4] = {"agec", "ageg", "meltc", "meltg"};
        for (int v = 0; v < 4; ++v)
            if (out2d[v]) {
                NumericMatrix a = r[n2[v]];
                for (size_t i = 0; i < (size_t)R * C; ++i) out2d[v][i] = a[i];
            }
        return MCF_OK;
    } catch (const std::exception& e) {
        std::snprintf(err, errlen, "%s", e.what());
        return MCF_ERR_ARG;
    }
}

extern "C" int DRV(gridmicrosnow2)(double reqhgt, const mcf_snow_climate* c, const double* umu, const mcf_snow_state* sm,
                                  const mcf_snow_static* st, double mat, double* const micro[MCF_NOUT], char* err,
                                  size_t errlen) {
    try {
        const int n = c->tsteps, R = st->rows, C = st->cols;
        List clim;
        clim["temp"] = arr3(c->temp, R, C, n); clim["relhum"] = arr3(c->relhum, R, C, n); clim["pres"] = arr3(c->pres, R, C, n);
        clim["swdown"] = arr3(c->swdown, R, C, n); clim["difrad"] = arr3(c->difrad, R, C, n); clim["lwdown"] = arr3(c->lwdown, R, C, n);
        clim["windspeed"] = arr3(c->windspeed, R, C, n); clim["winddir"] = vec(c->winddir, n); clim["prec"] = arr3(c->precip, R, C, n);
        clim["umu"] = arr3(umu, R, C, n);
        List snowm;
        snowm["Tc"] = arr3(sm->Tc, R, C, n); snowm["Tg"] = arr3(sm->Tg, R, C, n); snowm["totalSWE"] = arr3(sm->totalSWE, R, C, n);
        snowm["groundsnowdepth"] = arr3(sm->groundsnowdepth, R, C, n); snowm["snowden"] = arr3(sm->snowden, R, C, n);
        static const char* nm[MCF_NOUT] = {"Tz", "tleaf", "relhum", "soilm", "windspeed", "Rdirdown", "Rdifdown", "Rlwdown",
                                           "Rswup", "Rlwup"};
        List mic;
        std::vector<bool> o(MCF_NOUT);
        for (int v = 0; v < MCF_NOUT; ++v) {
            o[v] = micro[v] != nullptr;
            if (micro[v]) mic[nm[v]] = arr3(micro[v], R, C, n);
        }
        List vegp;
        vegp["pai"] = mat2(st->pai, R, C); vegp["paia"] = mat2(st->paia, R, C); vegp["hgt"] = mat2(st->hgt, R, C);
        vegp["leaft"] = mat2(st->leaft, R, C); vegp["clump"] = mat2(st->clump, R, C); vegp["leafd"] = mat2(st->leafd, R, C);
        vegp["leafden"] = mat2(st->leafden, R, C);
        List other;
        other["slope"] = mat2(st->slope, R, C); other["aspect"] = mat2(st->aspect, R, C);
        other["skyview"] = mat2(st->skyview, R, C); other["wsa"] = arr3(st->wsa, R, C, 8); other["hor"] = arr3(st->hor, R, C, 24);
        other["lat"] = mat2(st->lats, R, C); other["lon"] = mat2(st->lons, R, C); other["zref"] = st->zref;
        other["Smax"] = mat2(st->Smax, R, C);
        List r = gridmicrosnow2(reqhgt, obstime_df(n, c->year, c->month, c->day, c->hour), clim, snowm, mic, vegp, other, mat, o);
        for (int v = 0; v < MCF_NOUT; ++v)
            if (micro[v]) {
                NumericVector a = r[nm[v]];
                std::memcpy(micro[v], a.raw(), (size_t)R * C * n * sizeof(double));
            }
        return MCF_OK;
    } catch (const std::exception& e) {
        std::snprintf(err, errlen, "%s", e.what());
        return MCF_ERR_ARG;
    }
}
