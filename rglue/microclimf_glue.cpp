// microclimf_glue.cpp — the reference-side binding of the B200 grid solver.
//
// A maintainer of ilyamaclean/microclimf drops this file into src/, removes the bodies of the twelve functions of the
// same names from src/microclimfCpp.cpp (rglue/apply_glue.sh does it by name) and links libmicroclimf_b200 (rglue/Makevars).
// Nothing else changes: the twelve functions keep the reference's names, argument lists and `// [[Rcpp::export]]` markers
// (src/microclimfCpp.cpp:2050-2052, 2338-2340, 2622-2624, 2924-2926, 3561-3563, 3592-3594, 3622-3624, 3661-3663,
// 4170-4172, 4424-4426, 4892-4894, 5057-5059), so `Rcpp::compileAttributes()` regenerates src/RcppExports.cpp:248-465
// and R/RcppExports.R:72-102 byte for byte, `.runmodel1Cpp` ... (R/internal.R:1168, 1342, 1458, 1640, 1877, 2064, 2181,
// 2370) and the snow drivers (R/internal.R:2567-2616, 2952-3010, 3580-3745) call them as before, and modelin(), runmicro(),
// runmicro_big() and runbioclim() keep their signatures and return values.
//
// Written against Rcpp's PUBLIC API only (NumericVector / IntegerVector / NumericMatrix coercing constructors from list
// elements, attr("dim"), Rcpp::stop): an integer column handed where the reference reads a NumericVector is coerced
// exactly as the reference coerces it, a missing column raises the same R condition.  The file therefore also compiles
// against the repository's Rcpp stand-in (oracle/rcpp_shim/Rcpp.h), which is how tests/test_glue_gpu.py drives it with
// the DataFrame / List arguments the reference drivers take and compares with the compiled reference.
//
// There is NO CPU fallback: without a usable sm_100 device every function raises an R error carrying the library's
// message.
#include <Rcpp.h>

#include <cstdint>
#include <string>
#include <vector>

#include "microclimf_b200.h"

using namespace Rcpp;

namespace {

// ---- pointers into R-owned (or shim-owned) storage.  The vectors are kept alive by the holder below for the duration of
// the call; R's GC cannot move them.
inline const double* ptr(NumericVector& v) { return v.size() > 0 ? &v[0] : nullptr; }
inline const int32_t* ptr(IntegerVector& v) {
    static_assert(sizeof(int) == sizeof(int32_t), "R integers are 32-bit");
    return v.size() > 0 ? reinterpret_cast<const int32_t*>(&v[0]) : nullptr;
}

struct Holder { // keeps every coerced vector alive until the C call has returned
    std::vector<NumericVector> num;
    std::vector<IntegerVector> ints;
    const double* d(const NumericVector& v) {
        num.push_back(v);
        return ptr(num.back());
    }
    const int32_t* i(const IntegerVector& v) {
        ints.push_back(v);
        return ptr(ints.back());
    }
};

void fail_if(int rc, const char* err) {
    if (rc != MCF_OK) Rcpp::stop(std::string(err));
}

// obstime, vegp, soilc and the scalars: common to all eight grid drivers
void pack_common(mcf_problem& p, Holder& h, int mode, DataFrame& obstime, List& vegp, List& soilc, double reqhgt,
                 double zref, double Sminp, double Smaxp, double tfact, bool complete, double mat) {
    p = mcf_problem();
    p.mode = mode;
    p.reqhgt = reqhgt;
    p.zref = zref;
    p.Sminp = Sminp;
    p.Smaxp = Smaxp;
    p.tfact = tfact;
    p.complete = complete ? 1 : 0;
    p.mat = mat;
    IntegerVector year = obstime["year"], month = obstime["month"], day = obstime["day"];
    NumericVector hour = obstime["hour"];
    p.tsteps = (int32_t)year.size();
    p.year = h.i(year);
    p.month = h.i(month);
    p.day = h.i(day);
    p.hour = h.d(hour);
    // vegetation: [rows, cols] matrices (modes 1/2) or [rows, cols, nlyr] arrays (modes 3/4); same element order
    NumericVector hgt = vegp["hgt"];
    IntegerVector dims = hgt.attr("dim");
    if (dims.size() < 2) Rcpp::stop("vegp$hgt must be a matrix or a 3-D array");
    p.rows = dims[0];
    p.cols = dims[1];
    p.nlyr = 1;
    p.hgt = h.d(hgt);
    p.pai = h.d(vegp["pai"]);
    p.x = h.d(vegp["x"]);
    p.gsmax = h.d(vegp["gsmax"]);
    p.leafr = h.d(vegp["leafr"]);
    p.leaft = h.d(vegp["leaft"]);
    p.clump = h.d(vegp["clump"]);
    p.leafd = h.d(vegp["leafd"]);
    p.paia = h.d(vegp["paia"]);
    p.leafden = h.d(vegp["leafden"]);
    p.Smin = h.d(soilc["Smin"]);
    p.Smax = h.d(soilc["Smax"]);
    p.gref = h.d(soilc["gref"]);
    p.soilb = h.d(soilc["soilb"]);
    p.Psie = h.d(soilc["Psie"]);
    p.Vq = h.d(soilc["Vq"]);
    p.Vm = h.d(soilc["Vm"]);
    p.Mc = h.d(soilc["Mc"]);
    p.rho = h.d(soilc["rho"]);
    p.slope = h.d(soilc["slope"]);
    p.aspect = h.d(soilc["aspect"]);
    p.twi = h.d(soilc["twi"]);
    p.svfa = h.d(soilc["svfa"]);
    p.wsa = h.d(soilc["wsa"]);
    p.hor = h.d(soilc["hor"]);
}

// climdata / pointm of the data.frame drivers (column names of src/microclimfCpp.cpp:2062-2082)
void pack_series_df(mcf_problem& p, Holder& h, DataFrame& climdata, DataFrame& pointm, double lat, double lon) {
    p.lat = lat;
    p.lon = lon;
    p.temp = h.d(climdata["temp"]);
    p.es = h.d(climdata["es"]);
    p.ea = h.d(climdata["ea"]);
    p.tdew = h.d(climdata["tdew"]);
    p.pres = h.d(climdata["pres"]);
    p.swdown = h.d(climdata["swdown"]);
    p.difrad = h.d(climdata["difrad"]);
    p.lwdown = h.d(climdata["lwdown"]);
    p.windspeed = h.d(climdata["windspeed"]);
    p.winddir = h.d(climdata["winddir"]);
    p.p_soilm = h.d(pointm["soilm"]);
    p.p_Tg = h.d(pointm["Tg"]);
    p.p_Tbp = h.d(pointm["Tbp"]);
    p.p_G = h.d(pointm["G"]);
    p.p_umu = h.d(pointm["umu"]);
    p.p_kp = h.d(pointm["kp"]);
    p.p_muGp = h.d(pointm["muGp"]);
    p.p_dtrp = h.d(pointm["dtrp"]);
}

// climdata / pointm of the array drivers (names of src/microclimfCpp.cpp:2350-2370: tc, pk, Gp)
void pack_series_arr(mcf_problem& p, Holder& h, List& climdata, List& pointm, NumericMatrix& lats, NumericMatrix& lons) {
    p.temp = h.d(climdata["tc"]);
    p.es = h.d(climdata["es"]);
    p.ea = h.d(climdata["ea"]);
    p.tdew = h.d(climdata["tdew"]);
    p.pres = h.d(climdata["pk"]);
    p.swdown = h.d(climdata["swdown"]);
    p.difrad = h.d(climdata["difrad"]);
    p.lwdown = h.d(climdata["lwdown"]);
    p.windspeed = h.d(climdata["windspeed"]);
    p.winddir = h.d(climdata["winddir"]);
    p.p_soilm = h.d(pointm["soilm"]);
    p.p_Tg = h.d(pointm["Tg"]);
    p.p_Tbp = h.d(pointm["Tbp"]);
    p.p_G = h.d(pointm["Gp"]);
    p.p_umu = h.d(pointm["umu"]);
    p.p_kp = h.d(pointm["kp"]);
    p.p_muGp = h.d(pointm["muGp"]);
    p.p_dtrp = h.d(pointm["dtrp"]);
    p.lats = h.d(NumericVector(lats));
    p.lons = h.d(NumericVector(lons));
}

// dfsel of the layered drivers: 0-based inclusive hour spans per layer (R/internal.R:1391-1399).  The drivers index
// layers 0 .. nrow(dfsel) - 1 of the vegetation arrays (src/microclimfCpp.cpp:2770-2778).
void pack_layers(mcf_problem& p, Holder& h, DataFrame& dfsel) {
    IntegerVector st = dfsel["st"], ed = dfsel["ed"];
    p.nlyr = (int32_t)st.size();
    p.lyr_st = h.i(st);
    p.lyr_ed = h.i(ed);
}

const char* const kOut[MCF_NOUT] = {"Tz", "tleaf", "relhum", "soilm", "windspeed", "Rdirdown", "Rdifdown", "Rlwdown",
                                    "Rswup", "Rlwup"};

// the named list of [rows, cols, tsteps] arrays every runmicroNCpp returns (src/microclimfCpp.cpp:2325-2336)
List solve(const mcf_problem& p, const std::vector<bool>& out) {
    if ((int)out.size() != MCF_NOUT) Rcpp::stop("out must have 10 elements");
    const R_xlen_t n = (R_xlen_t)p.rows * p.cols * p.tsteps;
    std::vector<NumericVector> res(MCF_NOUT);
    double* bufs[MCF_NOUT];
    for (int v = 0; v < MCF_NOUT; ++v) {
        bufs[v] = nullptr;
        if (!out[v]) continue;
        res[v] = NumericVector(n);
        res[v].attr("dim") = IntegerVector::create(p.rows, p.cols, p.tsteps);
        bufs[v] = n > 0 ? &res[v][0] : nullptr;
    }
    char err[512] = {0};
    fail_if(mcf_runmicro(&p, bufs, err, sizeof err), err);
    List mout;
    for (int v = 0; v < MCF_NOUT; ++v)
        if (out[v]) mout[kOut[v]] = res[v];
    return mout;
}

// the named list of [rows, cols] matrices every runbioclimNCpp returns (src/microclimfCpp.cpp:3538-3559)
List solve_bioclim(const mcf_problem& p, const std::vector<bool>& out, IntegerVector& wetq, IntegerVector& dryq,
                   IntegerVector& hotq, IntegerVector& colq, bool air) {
    if ((int)out.size() != MCF_NBIO) Rcpp::stop("out must have 19 elements");
    std::vector<NumericMatrix> res(MCF_NBIO);
    double* bufs[MCF_NBIO];
    for (int v = 0; v < MCF_NBIO; ++v) {
        bufs[v] = nullptr;
        if (!out[v]) continue;
        res[v] = NumericMatrix(p.rows, p.cols);
        bufs[v] = (p.rows > 0 && p.cols > 0) ? &res[v][0] : nullptr;
    }
    char err[512] = {0};
    fail_if(mcf_runbioclim(&p, ptr(wetq), (int32_t)wetq.size(), ptr(dryq), (int32_t)dryq.size(), ptr(hotq),
                           (int32_t)hotq.size(), ptr(colq), (int32_t)colq.size(), air ? 1 : 0, bufs, err, sizeof err),
            err);
    List bout;
    for (int v = 0; v < MCF_NBIO; ++v)
        if (out[v]) bout["bio" + std::to_string(v + 1)] = res[v];
    return bout;
}

} // namespace

// ------------------------------------------------------------------------------------------------------------------
// grid microclimate drivers (replace src/microclimfCpp.cpp:2052-3223)
// ------------------------------------------------------------------------------------------------------------------
// [[Rcpp::export]]
List runmicro1Cpp(DataFrame obstime, DataFrame climdata, DataFrame pointm, List vegp, List soilc, double reqhgt,
                  double zref, double lat, double lon, double Sminp, double Smaxp, double tfact, bool complete,
                  double mat, std::vector<bool> out) {
    mcf_problem p;
    Holder h;
    pack_common(p, h, 1, obstime, vegp, soilc, reqhgt, zref, Sminp, Smaxp, tfact, complete, mat);
    pack_series_df(p, h, climdata, pointm, lat, lon);
    return solve(p, out);
}

// [[Rcpp::export]]
List runmicro2Cpp(DataFrame obstime, List climdata, List pointm, List vegp, List soilc, double reqhgt, double zref,
                  NumericMatrix lats, NumericMatrix lons, double Sminp, double Smaxp, double tfact, bool complete,
                  double mat, std::vector<bool> out) {
    mcf_problem p;
    Holder h;
    pack_common(p, h, 2, obstime, vegp, soilc, reqhgt, zref, Sminp, Smaxp, tfact, complete, mat);
    pack_series_arr(p, h, climdata, pointm, lats, lons);
    return solve(p, out);
}

// [[Rcpp::export]]
List runmicro3Cpp(DataFrame dfsel, DataFrame obstime, DataFrame climdata, DataFrame pointm, List vegp, List soilc,
                  double reqhgt, double zref, double lat, double lon, double Sminp, double Smaxp, double tfact,
                  bool complete, double mat, std::vector<bool> out) {
    mcf_problem p;
    Holder h;
    pack_common(p, h, 3, obstime, vegp, soilc, reqhgt, zref, Sminp, Smaxp, tfact, complete, mat);
    pack_series_df(p, h, climdata, pointm, lat, lon);
    pack_layers(p, h, dfsel);
    return solve(p, out);
}

// [[Rcpp::export]]
List runmicro4Cpp(DataFrame dfsel, DataFrame obstime, List climdata, List pointm, List vegp, List soilc,
                  double reqhgt, double zref, NumericMatrix lats, NumericMatrix lons, double Sminp, double Smaxp,
                  double tfact, bool complete, double mat, std::vector<bool> out) {
    mcf_problem p;
    Holder h;
    pack_common(p, h, 4, obstime, vegp, soilc, reqhgt, zref, Sminp, Smaxp, tfact, complete, mat);
    pack_series_arr(p, h, climdata, pointm, lats, lons);
    pack_layers(p, h, dfsel);
    return solve(p, out);
}

// ------------------------------------------------------------------------------------------------------------------
// bioclim drivers (replace src/microclimfCpp.cpp:3563-3700; complete = true and, for the layered variants, the 14
// one-day layers of :3635-3646 are set by the library)
// ------------------------------------------------------------------------------------------------------------------
// [[Rcpp::export]]
List runbioclim1Cpp(DataFrame obstime, DataFrame climdata, DataFrame pointm, List vegp, List soilc, double reqhgt,
                    double zref, double lat, double lon, double Sminp, double Smaxp, double tfact, double mat,
                    std::vector<bool> out, IntegerVector wetq, IntegerVector dryq, IntegerVector hotq,
                    IntegerVector colq, bool air) {
    mcf_problem p;
    Holder h;
    pack_common(p, h, 1, obstime, vegp, soilc, reqhgt, zref, Sminp, Smaxp, tfact, true, mat);
    pack_series_df(p, h, climdata, pointm, lat, lon);
    return solve_bioclim(p, out, wetq, dryq, hotq, colq, air);
}

// [[Rcpp::export]]
List runbioclim2Cpp(DataFrame obstime, List climdata, List pointm, List vegp, List soilc, double reqhgt, double zref,
                    NumericMatrix lats, NumericMatrix lons, double Sminp, double Smaxp, double tfact, double mat,
                    std::vector<bool> out, IntegerVector wetq, IntegerVector dryq, IntegerVector hotq,
                    IntegerVector colq, bool air) {
    mcf_problem p;
    Holder h;
    pack_common(p, h, 2, obstime, vegp, soilc, reqhgt, zref, Sminp, Smaxp, tfact, true, mat);
    pack_series_arr(p, h, climdata, pointm, lats, lons);
    return solve_bioclim(p, out, wetq, dryq, hotq, colq, air);
}

// [[Rcpp::export]]
List runbioclim3Cpp(DataFrame obstime, DataFrame climdata, DataFrame pointm, List vegp, List soilc, double reqhgt,
                    double zref, double lat, double lon, double Sminp, double Smaxp, double tfact, double mat,
                    std::vector<bool> out, IntegerVector wetq, IntegerVector dryq, IntegerVector hotq,
                    IntegerVector colq, bool air) {
    mcf_problem p;
    Holder h;
    pack_common(p, h, 3, obstime, vegp, soilc, reqhgt, zref, Sminp, Smaxp, tfact, true, mat);
    pack_series_df(p, h, climdata, pointm, lat, lon);
    IntegerVector dims = NumericVector(vegp["hgt"]).attr("dim");
    p.nlyr = dims.size() >= 3 ? dims[2] : 1; // the library checks for the 14 layers and builds their day spans
    return solve_bioclim(p, out, wetq, dryq, hotq, colq, air);
}

// [[Rcpp::export]]
List runbioclim4Cpp(DataFrame obstime, List climdata, List pointm, List vegp, List soilc, double reqhgt, double zref,
                    NumericMatrix lats, NumericMatrix lons, double Sminp, double Smaxp, double tfact, double mat,
                    std::vector<bool> out, IntegerVector wetq, IntegerVector dryq, IntegerVector hotq,
                    IntegerVector colq, bool air) {
    mcf_problem p;
    Holder h;
    pack_common(p, h, 4, obstime, vegp, soilc, reqhgt, zref, Sminp, Smaxp, tfact, true, mat);
    pack_series_arr(p, h, climdata, pointm, lats, lons);
    IntegerVector dims = NumericVector(vegp["hgt"]).attr("dim");
    p.nlyr = dims.size() >= 3 ? dims[2] : 1;
    return solve_bioclim(p, out, wetq, dryq, hotq, colq, air);
}

// ------------------------------------------------------------------------------------------------------------------
// snow operators (replace src/microclimfCpp.cpp:4172-4424, 4426-4673, 4894-5057, 5059-5214)
// ------------------------------------------------------------------------------------------------------------------
namespace {

int snowenv_index(const std::string& s) { // snowdenp, src/microclimfCpp.cpp:3741-3750: anything else is Alpine
    if (s == "Maritime") return 1;
    if (s == "Prairie") return 2;
    if (s == "Tundra") return 3;
    if (s == "Taiga") return 4;
    return 0;
}

void pack_snow_time(mcf_snow_climate& c, Holder& h, DataFrame& obstime) {
    c = mcf_snow_climate();
    IntegerVector year = obstime["year"], month = obstime["month"], day = obstime["day"];
    NumericVector hour = obstime["hour"];
    c.tsteps = (int32_t)year.size();
    c.year = h.i(year);
    c.month = h.i(month);
    c.day = h.i(day);
    c.hour = h.d(hour);
}
void pack_snow_clim(mcf_snow_climate& c, Holder& h, List& climdata, const char* precip_name) {
    c.temp = h.d(climdata["temp"]);
    c.relhum = h.d(climdata["relhum"]);
    c.pres = h.d(climdata["pres"]);
    c.swdown = h.d(climdata["swdown"]);
    c.difrad = h.d(climdata["difrad"]);
    c.lwdown = h.d(climdata["lwdown"]);
    c.windspeed = h.d(climdata["windspeed"]);
    c.winddir = h.d(climdata["winddir"]);
    c.precip = h.d(climdata[precip_name]);
}
void pack_snow_static(mcf_snow_static& s, Holder& h, List& vegp, List& other, bool array_climate, bool micro,
                      const char* lat_name, const char* lon_name) {
    s = mcf_snow_static();
    NumericMatrix pai = vegp["pai"];
    s.rows = pai.nrow();
    s.cols = pai.ncol();
    s.pai = h.d(NumericVector(pai));
    s.hgt = h.d(vegp["hgt"]);
    s.leaft = h.d(vegp["leaft"]);
    s.clump = h.d(vegp["clump"]);
    s.slope = h.d(other["slope"]);
    s.aspect = h.d(other["aspect"]);
    s.skyview = h.d(other["skyview"]);
    s.wsa = h.d(other["wsa"]);
    s.hor = h.d(other["hor"]);
    s.zref = other["zref"];
    if (array_climate) {
        s.lats = h.d(other[lat_name]);
        s.lons = h.d(other[lon_name]);
    } else {
        s.lat = other["lat"];
        s.lon = other["lon"];
    }
    if (micro) {
        s.paia = h.d(vegp["paia"]);
        s.leafd = h.d(vegp["leafd"]);
        s.leafden = h.d(vegp["leafden"]);
        s.Smax = h.d(other["Smax"]);
    } else {
        s.isnowdc = h.d(other["isnowdc"]);
        s.isnowdg = h.d(other["isnowdg"]);
        s.isnowac = h.i(other["isnowac"]);
        s.isnowag = h.i(other["isnowag"]);
    }
}

// gridmodelsnow1/2 return value (src/microclimfCpp.cpp:4412-4423)
List snow_model(const mcf_snow_climate& c, const mcf_snow_point& pt, const mcf_snow_static& s, const std::string& snowenv,
                bool array_climate) {
    const R_xlen_t nc = (R_xlen_t)s.rows * s.cols, n = nc * c.tsteps;
    const char* n3[5] = {"Tc", "Tg", "sdepc", "sdepg", "sden"};
    const char* n2[4] = {"agec", "ageg", "meltc", "meltg"};
    std::vector<NumericVector> a3(5);
    std::vector<NumericMatrix> a2(4);
    double *o3[5], *o2[4];
    for (int v = 0; v < 5; ++v) {
        a3[v] = NumericVector(n);
        a3[v].attr("dim") = IntegerVector::create(s.rows, s.cols, c.tsteps);
        o3[v] = n > 0 ? &a3[v][0] : nullptr;
    }
    for (int v = 0; v < 4; ++v) {
        a2[v] = NumericMatrix(s.rows, s.cols);
        o2[v] = nc > 0 ? &a2[v][0] : nullptr;
    }
    char err[512] = {0};
    fail_if((array_climate ? mcf_gridmodelsnow2 : mcf_gridmodelsnow)(&c, &pt, &s, snowenv_index(snowenv), o3, o2, err,
                                                                     sizeof err),
            err);
    List out;
    for (int v = 0; v < 5; ++v) out[n3[v]] = a3[v];
    // ages and melts are NumericMatrix in the reference too (bioclimfill, src/microclimfCpp.cpp:4306-4309): skipped cells NA
    for (int v = 0; v < 4; ++v) out[n2[v]] = a2[v];
    return out;
}

// gridmicrosnow1/2: `micro`'s arrays are updated where SWE > 0 and returned (src/microclimfCpp.cpp:4954-5056)
List snow_micro(double reqhgt, const mcf_snow_climate& c, const double* umu, List& snowm, List& micro,
                const mcf_snow_static& s, double mat, const std::vector<bool>& out, Holder& h, bool array_climate) {
    if ((int)out.size() != MCF_NOUT) Rcpp::stop("out must have 10 elements");
    mcf_snow_state st;
    st.Tc = h.d(snowm["Tc"]);
    st.Tg = h.d(snowm["Tg"]);
    st.totalSWE = h.d(snowm["totalSWE"]);
    st.groundsnowdepth = h.d(snowm["groundsnowdepth"]);
    st.snowden = h.d(snowm["snowden"]);
    std::vector<NumericVector> arr(MCF_NOUT);
    double* bufs[MCF_NOUT];
    for (int v = 0; v < MCF_NOUT; ++v) {
        bufs[v] = nullptr;
        if (!out[v]) continue;
        arr[v] = NumericVector(micro[kOut[v]]); // the reference writes through the handle it was given, in place
        bufs[v] = arr[v].size() > 0 ? &arr[v][0] : nullptr;
    }
    char err[512] = {0};
    fail_if((array_climate ? mcf_gridmicrosnow2 : mcf_gridmicrosnow)(reqhgt, &c, umu, &st, &s, mat, bufs, err, sizeof err),
            err);
    List outp;
    for (int v = 0; v < MCF_NOUT; ++v)
        if (out[v]) outp[kOut[v]] = arr[v];
    return outp;
}

} // namespace

// [[Rcpp::export]]
List gridmodelsnow1(DataFrame obstime, DataFrame climdata, DataFrame pointm, List vegp, List other, std::string snowenv) {
    Holder h;
    mcf_snow_climate c;
    pack_snow_time(c, h, obstime);
    pack_snow_clim(c, h, climdata, "precip");
    mcf_snow_point pt;
    pt.Gp = h.d(pointm["Gp"]);
    pt.Tc = h.d(pointm["Tc"]);
    pt.RswabsG = h.d(pointm["RswabsG"]);
    pt.RlwabsG = h.d(pointm["RlwabsG"]);
    pt.umu = h.d(pointm["umu"]);
    mcf_snow_static s;
    pack_snow_static(s, h, vegp, other, false, false, "lat", "lon");
    return snow_model(c, pt, s, snowenv, false);
}

// [[Rcpp::export]]
List gridmodelsnow2(DataFrame obstime, List climdata, List pointm, List vegp, List other, std::string snowenv) {
    Holder h;
    mcf_snow_climate c;
    pack_snow_time(c, h, obstime);
    pack_snow_clim(c, h, climdata, "precip");
    mcf_snow_point pt;
    pt.Gp = h.d(pointm["Gp"]);
    pt.Tc = h.d(pointm["Tc"]);
    pt.RswabsG = h.d(pointm["RswabsG"]);
    pt.RlwabsG = h.d(pointm["RlwabsG"]);
    pt.umu = h.d(pointm["umu"]);
    mcf_snow_static s;
    pack_snow_static(s, h, vegp, other, true, false, "lats", "lons"); // other$lats / other$lons (:4470-4471)
    return snow_model(c, pt, s, snowenv, true);
}

// [[Rcpp::export]]
List gridmicrosnow1(double reqhgt, DataFrame obstime, DataFrame climdata, List snowm, List micro, List vegp, List other,
                    double mat, std::vector<bool> out) {
    Holder h;
    mcf_snow_climate c;
    pack_snow_time(c, h, obstime);
    pack_snow_clim(c, h, climdata, "precip");
    const double* umu = h.d(climdata["umu"]);
    mcf_snow_static s;
    pack_snow_static(s, h, vegp, other, false, true, "lat", "lon");
    return snow_micro(reqhgt, c, umu, snowm, micro, s, mat, out, h, false);
}

// [[Rcpp::export]]
List gridmicrosnow2(double reqhgt, DataFrame obstime, List climdata, List snowm, List micro, List vegp, List other,
                    double mat, std::vector<bool> out) {
    Holder h;
    mcf_snow_climate c;
    pack_snow_time(c, h, obstime);
    pack_snow_clim(c, h, climdata, "prec"); // the array variant reads climdata$prec (:5085)
    const double* umu = h.d(climdata["umu"]);
    mcf_snow_static s;
    pack_snow_static(s, h, vegp, other, true, true, "lat", "lon"); // other$lat / other$lon are matrices here (:5101-5102)
    return snow_micro(reqhgt, c, umu, snowm, micro, s, mat, out, h, true);
}
