// TEST INFRASTRUCTURE ONLY — never part of the product library.
//
// The re-derived snow physics (microclimf_b200/csrc/mcf_snow_physics.cuh, mcf_snow_drivers.cuh) compiled for the HOST and
// looped over the cells, behind the C signatures of mcf_gridmodelsnow[2] / mcf_gridmicrosnow[2].  It lets the CPU test
// suite check the ALGEBRA of the re-derivation (hoisted hour terms, the specialised two-stream solution, the cancelled
// latent heat ...) against the compiled reference without a GPU; the elementary functions are libm here and the
// MUFU-seeded ones of mcf_math.cuh on the device (covered by tests/test_snow_gpu.py).  tests/test_snow_hostcheck_cpu.py
// builds it with g++ into tests/hostcheck/_build/.
#include <cstring>
#include <vector>

#include "microclimf_b200.h"
#include "../../microclimf_b200/csrc/mcf_snow_drivers.cuh"

using namespace mcf::snowphys;

namespace {
const double kSdp[5][4] = {{0.5975, 0.2237, 0.0012, 0.0038}, {0.5979, 0.2578, 0.001, 0.0038}, {0.594, 0.2332, 0.0016, 0.0031},
                           {0.363, 0.2425, 0.0029, 0.0049}, {0.217, 0.217, 0.0, 0.0}}; // snowdenp, ref :3741-3750

SnowSeries series(const mcf_snow_climate* c, const mcf_snow_point* pt, const double* umu, double lat, double lon) {
    SnowSeries s;
    std::memset(&s, 0, sizeof s);
    s.tsteps = c->tsteps;
    s.year = c->year; s.month = c->month; s.day = c->day; s.hour = c->hour;
    s.temp = c->temp; s.relhum = c->relhum; s.pres = c->pres; s.swdown = c->swdown; s.difrad = c->difrad; s.lwdown = c->lwdown;
    s.windspeed = c->windspeed; s.winddir = c->winddir; s.precip = c->precip;
    if (pt) { s.Gp = pt->Gp; s.Tcp = pt->Tc; s.RswabsG = pt->RswabsG; s.RlwabsG = pt->RlwabsG; s.umu = pt->umu; }
    if (umu) s.umu = umu;
    s.lat = lat; s.lon = lon;
    return s;
}
std::vector<SnowHr> table(const SnowSeries& s, double* mxtc) {
    const int T = s.tsteps, ndays = T / 24;
    std::vector<SnowHr> hours(T);
    std::vector<DayExtremes> days(ndays + 1);
    if (s.RswabsG)
        for (int d = 0; d < ndays; ++d) day_extremes(days[d], s.RswabsG, s.RlwabsG, s.temp, s.swdown, s.lwdown, (size_t)d * 24, 1);
    int hs = 0;
    double mx = -273.15;
    for (int k = 0; k < T; ++k) {
        if (k > 0) hs = (s.precip[k] > 0) ? 0 : hs + 1;
        if (s.temp[k] > mx) mx = s.temp[k];
        prep_hour(s, k, hs, (s.RswabsG && k / 24 < ndays) ? days[k / 24] : no_extremes(), hours[k]);
    }
    if (mxtc) *mxtc = mx;
    return hours;
}
SnowArr arrays(const mcf_snow_climate* c, const mcf_snow_point* pt, const double* umu, const mcf_snow_static* st) {
    SnowArr a;
    std::memset(&a, 0, sizeof a);
    a.year = c->year; a.month = c->month; a.day = c->day; a.hour = c->hour;
    a.temp = c->temp; a.relhum = c->relhum; a.pres = c->pres; a.swdown = c->swdown; a.difrad = c->difrad; a.lwdown = c->lwdown;
    a.windspeed = c->windspeed; a.precip = c->precip; a.winddir = c->winddir;
    if (pt) { a.Gp = pt->Gp; a.Tcp = pt->Tc; a.RswabsG = pt->RswabsG; a.RlwabsG = pt->RlwabsG; a.umu = pt->umu; }
    if (umu) a.umu = umu;
    a.lats = st->lats; a.lons = st->lons;
    return a;
}
int model(bool arr, const mcf_snow_climate* c, const mcf_snow_point* pt, const mcf_snow_static* st, int32_t snowenv,
          double* const out3d[5], double* const out2d[4]) {
    const size_t nc = (size_t)st->rows * st->cols, T = (size_t)c->tsteps;
    std::vector<std::vector<double>> b3(5, std::vector<double>(nc * T)), b2(4, std::vector<double>(nc));
    SnowModelArgs a;
    std::memset(&a, 0, sizeof a);
    a.rows = st->rows; a.cols = st->cols; a.tsteps = c->tsteps; a.zref = st->zref;
    for (int i = 0; i < 4; ++i) a.sdp[i] = kSdp[snowenv][i];
    a.pai = st->pai; a.hgt = st->hgt; a.ltra = st->leaft; a.clump = st->clump; a.slope = st->slope; a.aspect = st->aspect;
    a.skyview = st->skyview; a.wsa = st->wsa; a.hor = st->hor; a.isnowdc = st->isnowdc; a.isnowdg = st->isnowdg;
    a.isnowac = st->isnowac; a.isnowag = st->isnowag;
    a.Tc = b3[0].data(); a.Tg = b3[1].data(); a.sdepc = b3[2].data(); a.sdepg = b3[3].data(); a.sden = b3[4].data();
    a.agec = b2[0].data(); a.ageg = b2[1].data(); a.meltc = b2[2].data(); a.meltg = b2[3].data();
    std::vector<SnowHr> hours;
    if (arr) {
        const SnowArr ca = arrays(c, pt, nullptr, st);
        for (size_t cell = 0; cell < nc; ++cell) snowmodel_cell_arr(a, ca, (int)cell);
    } else {
        hours = table(series(c, pt, nullptr, st->lat, st->lon), nullptr);
        a.hours = hours.data();
        for (size_t cell = 0; cell < nc; ++cell) snowmodel_cell(a, (int)cell);
    }
    for (int v = 0; v < 5; ++v)
        if (out3d[v]) std::memcpy(out3d[v], b3[v].data(), nc * T * sizeof(double));
    for (int v = 0; v < 4; ++v)
        if (out2d[v]) std::memcpy(out2d[v], b2[v].data(), nc * sizeof(double));
    return MCF_OK;
}
int micro_(bool arr, double reqhgt, const mcf_snow_climate* c, const double* umu, const mcf_snow_state* sm, const mcf_snow_static* st,
           double mat, double* const micro[MCF_NOUT]) {
    SnowMicroArgs a;
    std::memset(&a, 0, sizeof a);
    a.rows = st->rows; a.cols = st->cols; a.tsteps = c->tsteps; a.reqhgt = reqhgt; a.zref = st->zref; a.mat = mat;
    const int y0 = c->year[0];
    a.hiy = (y0 % 4 == 0 && (y0 % 100 != 0 || y0 % 400 == 0)) ? 366 * 24 : 365 * 24;
    a.pai = st->pai; a.paia = st->paia; a.hgt = st->hgt; a.ltra = st->leaft; a.clump = st->clump; a.leafd = st->leafd;
    a.leafden = st->leafden; a.slope = st->slope; a.aspect = st->aspect; a.skyview = st->skyview; a.wsa = st->wsa; a.hor = st->hor;
    a.Smax = st->Smax; a.snowtempc = sm->Tc; a.snowtempg = sm->Tg; a.swe = sm->totalSWE; a.sdepg = sm->groundsnowdepth;
    a.sden = sm->snowden;
    for (int v = 0; v < MCF_NOUT; ++v) a.out[v] = micro[v];
    const size_t nc = (size_t)st->rows * st->cols;
    double mxtc = 0.0;
    std::vector<SnowHr> hours;
    if (arr) {
        const SnowArr ca = arrays(c, nullptr, umu, st);
        for (size_t cell = 0; cell < nc; ++cell) snowmicro_cell_t<true>(a, &ca, (int)cell);
    } else {
        hours = table(series(c, nullptr, umu, st->lat, st->lon), &mxtc);
        a.hours = hours.data();
        a.scal = &mxtc;
        for (size_t cell = 0; cell < nc; ++cell) snowmicro_cell_t<false>(a, nullptr, (int)cell);
    }
    return MCF_OK;
}
} // namespace

extern "C" int host_gridmodelsnow(const mcf_snow_climate* c, const mcf_snow_point* pt, const mcf_snow_static* st, int32_t snowenv,
                                  double* const out3d[5], double* const out2d[4], char*, size_t) {
    return model(false, c, pt, st, snowenv, out3d, out2d);
}
extern "C" int host_gridmodelsnow2(const mcf_snow_climate* c, const mcf_snow_point* pt, const mcf_snow_static* st, int32_t snowenv,
                                   double* const out3d[5], double* const out2d[4], char*, size_t) {
    return model(true, c, pt, st, snowenv, out3d, out2d);
}
extern "C" int host_gridmicrosnow(double reqhgt, const mcf_snow_climate* c, const double* umu, const mcf_snow_state* sm,
                                  const mcf_snow_static* st, double mat, double* const micro[MCF_NOUT], char*, size_t) {
    return micro_(false, reqhgt, c, umu, sm, st, mat, micro);
}
extern "C" int host_gridmicrosnow2(double reqhgt, const mcf_snow_climate* c, const double* umu, const mcf_snow_state* sm,
                                   const mcf_snow_static* st, double mat, double* const micro[MCF_NOUT], char*, size_t) {
    return micro_(true, reqhgt, c, umu, sm, st, mat, micro);
}
