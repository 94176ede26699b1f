// microclimf_b200 — per-cell drivers of the snow operators (one thread = one cell walks the hourly series).
//
//   snowmodel_cell / snowmodel_cell_arr : gridmodelsnow1 / gridmodelsnow2 (ref src/microclimfCpp.cpp:4172-4424, 4426-4673)
//   snowmicro_cell / snowmicro_cell_arr : gridmicrosnow1 / gridmicrosnow2 (ref :4894-5057, 5059-5214)
//
// The functions are __host__ __device__: the kernels of mcf_snow.cu call them with cell = the thread's index; the
// no-GPU check of the re-derived algebra (tests/hostcheck) loops them over the cells on the host.  Arrays are R layout:
// element (cell, hour k) of a [rows, cols, tsteps] array is k * ncells + cell, so a warp's accesses of one hour are one
// contiguous 256-byte segment.
#pragma once
#include "mcf_snow_physics.cuh"

namespace mcf {
namespace snowphys {

SNOW_HD double na_real() {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(0x7FF00000000007A2LL);
#else
    const uint64_t b = 0x7FF00000000007A2ULL;
    double d;
    memcpy(&d, &b, sizeof d);
    return d;
#endif
}

// per-hour series of the data.frame drivers
struct SnowSeries {
    int tsteps;
    const int32_t *year, *month, *day;
    const double* hour;
    const double *temp, *relhum, *pres, *swdown, *difrad, *lwdown, *windspeed, *winddir, *precip;
    const double *Gp, *Tcp, *RswabsG, *RlwabsG, *umu; // NULL for the snow microclimate (climate + umu only)
    double lat, lon;
};
// [rows, cols, tsteps] series of the array-climate drivers (winddir stays per hour)
struct SnowArr {
    const int32_t *year, *month, *day;
    const double* hour;
    const double *temp, *relhum, *pres, *swdown, *difrad, *lwdown, *windspeed, *precip;
    const double* winddir;
    const double *Gp, *Tcp, *RswabsG, *RlwabsG, *umu;
    const double *lats, *lons;
};

// extremes of the point model's net radiation over day `d` of a series with stride `st` starting at `base` (ref
// :4222-4283 / :4503-4566): the short- and long-wave forcing at the hour of the largest / smallest net radiation
SNOW_HD void day_extremes(DayExtremes& e, const double* RswabsG, const double* RlwabsG, const double* temp, const double* swdown,
                          const double* lwdown, size_t base, size_t st) {
    double mx = -1352.0, mn = 1352.0;
    e.Rswmx = e.Rlwmx = e.Rswmn = e.Rlwmn = e.Gmx = 0.0;
    for (int hh = 0; hh < 24; ++hh) {
        const size_t i = base + (size_t)hh * st;
        const double Rnet = RswabsG[i] + RlwabsG[i] - kEmSb * pow4(temp[i] + 273.15);
        if (mx < Rnet) { mx = Rnet; e.Rswmx = swdown[i]; e.Rlwmx = lwdown[i]; }
        if (mn > Rnet) { mn = Rnet; e.Rswmn = swdown[i]; e.Rlwmn = lwdown[i]; }
        if (fabs(Rnet) > e.Gmx) e.Gmx = fabs(Rnet);
    }
    e.Rmx = mx;
    e.Rmn = mn;
}
SNOW_HD DayExtremes no_extremes() {
    DayExtremes e;
    e.Rmx = e.Rmn = e.Rswmx = e.Rlwmx = e.Rswmn = e.Rlwmn = e.Gmx = 0.0;
    return e;
}
// hour k of the per-hour table (data.frame drivers)
SNOW_HD void prep_hour(const SnowSeries& a, int k, int hours_since_snow, const DayExtremes& day, SnowHr& h) {
    const SolarPos sp = solar_position(a.lat, a.lon, a.year[k], a.month[k], a.day[k], a.hour[k]);
    snow_hour(h, a.temp[k], a.relhum[k], a.pres[k], a.windspeed[k], a.swdown[k], a.difrad[k], a.lwdown[k], a.precip[k],
              a.Tcp ? a.Tcp[k] : 0.0, a.Gp ? a.Gp[k] : 0.0, a.umu ? a.umu[k] : 1.0, a.winddir[k], sp, hours_since_snow, day);
}

// ---------------------------------------------------------------------------------------------------------------------
// snow-pack model
// ---------------------------------------------------------------------------------------------------------------------
struct SnowModelArgs {
    int rows, cols, tsteps;
    const SnowHr* hours; // data.frame climate: the per-hour table
    const double *pai, *hgt, *ltra, *clump;
    const double *slope, *aspect, *skyview, *wsa, *hor;
    const double *isnowdc, *isnowdg;
    const int32_t *isnowac, *isnowag;
    double zref;
    double sdp[4];
    double *Tc, *Tg, *sdepc, *sdepg, *sden; // [rows, cols, tsteps]
    double *agec, *ageg, *meltc, *meltg;    // [rows, cols]
};

SNOW_HD void snowmodel_skip(const SnowModelArgs& a, int cell) { // cells the reference leaves at its NA prefill (:4296-4309)
    const size_t nc = (size_t)a.rows * a.cols;
    const double NA = na_real();
    for (int k = 0; k < a.tsteps; ++k) {
        const size_t idx = (size_t)k * nc + cell;
        a.Tc[idx] = NA; a.Tg[idx] = NA; a.sdepc[idx] = NA; a.sdepg[idx] = NA; a.sden[idx] = NA;
    }
    a.agec[cell] = NA; a.ageg[cell] = NA; a.meltc[cell] = NA; a.meltg[cell] = NA;
}
SNOW_HD void snowmodel_init(const SnowModelArgs& a, int cell, SnowCell& c, SnowState& s) {
    snow_cell(c, a.hgt[cell], a.pai[cell], a.ltra[cell], a.clump[cell], a.slope[cell], a.aspect[cell], a.skyview[cell]);
    s.agec = a.isnowac[cell];
    s.ageg = a.isnowag[cell];
    s.sdepc = a.isnowdc[cell];
    s.sdepg = a.isnowdg[cell];
    s.sdenc = pack_density(a.sdp, s.sdepc, (double)s.agec);       // ref :4311-4314
    s.sdeng = pack_density(a.sdp, s.sdepg * 0.5, (double)s.ageg); // (the ground pack's start density uses half its depth)
}
// one hour's outputs and melt totals; `meltc` receives both the water equivalent and the depth equivalent (ref :4399-4401)
SNOW_HD void snowmodel_store(const SnowModelArgs& a, size_t idx, const SnowStepOut& o, const SnowState& s, double& meltc, double& meltg) {
    a.Tc[idx] = o.Tc; a.Tg[idx] = o.Tg; a.sdepc[idx] = s.sdepc; a.sdepg[idx] = s.sdepg; a.sden[idx] = s.sdenc;
    meltc = meltc + o.melc;
    meltc = meltc + (o.melc * 1000.0) / s.sdenc;
    meltg = meltg + (o.melg * 1000.0) / s.sdeng;
}
SNOW_HD void snowmodel_bare(const SnowModelArgs& a, size_t idx) { // no pack and no snowfall (:4403-4409)
    a.Tc[idx] = 0.0; a.Tg[idx] = 0.0; a.sdepc[idx] = 0.0; a.sdepg[idx] = 0.0; a.sden[idx] = a.sdp[1] * 1000.0;
}

SNOW_HD void snowmodel_cell(const SnowModelArgs& a, int cell) {
    const size_t nc = (size_t)a.rows * a.cols;
    if (a.hgt[cell] != a.hgt[cell]) { snowmodel_skip(a, cell); return; }
    SnowCell c;
    SnowState s;
    snowmodel_init(a, cell, c, s);
    double meltc = 0.0;
    double meltg = na_real(); // bioclimfill leaves NA and the loop only ever adds to it (:4401): reproduced
    for (int k = 0; k < a.tsteps; ++k) {
        const SnowHr& h = a.hours[k];
        const size_t idx = (size_t)k * nc + cell;
        if (s.sdepc > 0.0 || h.snowing) {
            const double ws = a.wsa[(size_t)h.windex * nc + cell], ha = a.hor[(size_t)h.sindex * nc + cell];
            const SnowStepOut o = snow_hour_step(c, h, a.sdp, a.zref, ws, ha, h.tan_alt_deg, s);
            snowmodel_store(a, idx, o, s, meltc, meltg);
        } else {
            snowmodel_bare(a, idx);
        }
    }
    a.agec[cell] = (double)s.agec;
    a.ageg[cell] = (double)s.ageg;
    a.meltc[cell] = meltc;
    a.meltg[cell] = meltg;
}

// array climate: the hour record is formed per cell-hour (only for hours with a pack or snowfall), the albedo age scan
// runs along the cell's own precipitation, the radiation extremes are gathered at each day start
SNOW_HD void snowmodel_cell_arr(const SnowModelArgs& a, const SnowArr& c4, int cell) {
    const size_t nc = (size_t)a.rows * a.cols;
    if (a.hgt[cell] != a.hgt[cell]) { snowmodel_skip(a, cell); return; }
    SnowCell c;
    SnowState s;
    snowmodel_init(a, cell, c, s);
    const double lat = c4.lats[cell], lon = c4.lons[cell];
    double meltc = na_real(), meltg = na_real(); // neither is initialised in the array-climate driver (:4487-4488)
    const int ndays = a.tsteps / 24;
    int hs = 0;
    DayExtremes day = no_extremes();
    for (int k = 0; k < a.tsteps; ++k) {
        const size_t idx = (size_t)k * nc + cell;
        const double tc = c4.temp[idx], prec = c4.precip[idx];
        if (k > 0) hs = (prec > 0) ? 0 : hs + 1;
        if ((k % 24) == 0) {
            day = no_extremes(); // zero beyond the whole days (:4503-4566)
            if (k / 24 < ndays) day_extremes(day, c4.RswabsG, c4.RlwabsG, c4.temp, c4.swdown, c4.lwdown, idx, nc);
        }
        if (s.sdepc > 0.0 || (tc < 2.0 && prec > 0.0)) {
            SnowHr h;
            const SolarPos sp = solar_position(lat, lon, c4.year[k], c4.month[k], c4.day[k], c4.hour[k]);
            snow_hour(h, tc, c4.relhum[idx], c4.pres[idx], c4.windspeed[idx], c4.swdown[idx], c4.difrad[idx], c4.lwdown[idx], prec,
                      c4.Tcp[idx], c4.Gp[idx], c4.umu[idx], c4.winddir[k], sp, hs, day);
            const double ws = a.wsa[(size_t)h.windex * nc + cell], ha = a.hor[(size_t)h.sindex * nc + cell];
            const SnowStepOut o = snow_hour_step(c, h, a.sdp, a.zref, ws, ha, h.tan_alt_rad, s);
            snowmodel_store(a, idx, o, s, meltc, meltg);
        } else {
            snowmodel_bare(a, idx);
        }
    }
    a.agec[cell] = (double)s.agec;
    a.ageg[cell] = (double)s.ageg;
    a.meltc[cell] = meltc;
    a.meltg[cell] = meltg;
}

// ---------------------------------------------------------------------------------------------------------------------
// snow microclimate
// ---------------------------------------------------------------------------------------------------------------------
struct SnowMicroArgs {
    int rows, cols, tsteps;
    const SnowHr* hours;
    const double* scal; // [0] series maximum of air temperature
    double reqhgt, zref, mat;
    int hiy;
    const double *pai, *paia, *hgt, *ltra, *clump, *leafd, *leafden;
    const double *slope, *aspect, *skyview, *wsa, *hor, *Smax;
    const double *snowtempc, *snowtempg, *swe, *sdepg, *sden; // [rows, cols, tsteps]
    double* out[10];                                            // in / out (runmicro's arrays), NULL = absent
};

// mean damping depth of the cell's pack over the series (ref meanDsnow :4713-4737); NA when the first hour's density is
SNOW_HD double mean_damping_depth(const double* sden, size_t nc, int cell, int T) {
    if (sden[cell] != sden[cell]) return na_real();
    double sum = 0.0;
    for (int k = 0; k < T; ++k) {
        const double sd = sden[(size_t)k * nc + cell];
        const double kap = (0.0442 * SM_EXP(5.181 * sd / 1000.0)) / (sd * 2090.0);
        sum += SM_SQRT(2.0 * kap / ((2.0 * kPi) / (24.0 * 3600.0)));
    }
    return sum / (double)T;
}
SNOW_HD void micro_store(const SnowMicroArgs& a, size_t idx, const MicroOut& o) {
    if (a.out[0]) a.out[0][idx] = o.Tz;
    if (a.out[1]) a.out[1][idx] = o.tleaf;
    if (a.out[2]) a.out[2][idx] = o.rh;
    if (a.out[4]) a.out[4][idx] = o.uz;
    if (a.out[5]) a.out[5][idx] = o.Rbdown;
    if (a.out[6]) a.out[6][idx] = o.Rddown;
    if (a.out[7]) a.out[7][idx] = o.Rlwdn;
    if (a.out[8]) a.out[8][idx] = o.Rdup;
    if (a.out[9]) a.out[9][idx] = o.Rlwup;
}
SNOW_HD void micro_store_buried(const SnowMicroArgs& a, size_t idx, double Tz) { // inside the pack (:5025-5036)
    if (a.out[0]) a.out[0][idx] = Tz;
    if (a.out[1]) a.out[1][idx] = Tz;
    if (a.out[2]) a.out[2][idx] = 100.0;
    for (int v = 4; v < 10; ++v)
        if (a.out[v]) a.out[v][idx] = 0.0;
}

// ARRC = false: `hours` table; true: the record is formed per snow-covered cell-hour from the arrays of `c4`
template <bool ARRC>
SNOW_HD void snowmicro_cell_t(const SnowMicroArgs& a, const SnowArr* c4, int cell) {
    const size_t nc = (size_t)a.rows * a.cols;
    const double hgt = a.hgt[cell];
    if (hgt != hgt) return;
    const int T = a.tsteps;
    const double meanD = mean_damping_depth(a.sden, nc, cell, T);
    double mxtc = ARRC ? -273.15 : a.scal[0];
    if (ARRC) // per cell here (:5138-5143)
        for (int k = 0; k < T; ++k) {
            const double t = c4->temp[(size_t)k * nc + cell];
            if (t > mxtc) mxtc = t;
        }
    const double dTmx = -0.6273 * mxtc + 49.79;
    const bool tzd_ok = !(a.snowtempg[cell] != a.snowtempg[cell]); // snowdayan (:4679-4711): NA series when hour 0 is NA
    MicroCell c;
    micro_cell(c, hgt, a.pai[cell], a.paia[cell], a.leafd[cell], a.leafden[cell], a.ltra[cell], a.clump[cell], a.slope[cell],
               a.aspect[cell], a.skyview[cell], a.Smax[cell]);
    const double lat = ARRC ? c4->lats[cell] : 0.0, lon = ARRC ? c4->lons[cell] : 0.0;
    const int ndays = T / 24;
    double Tzd = na_real();
    int hs = 0;
    for (int k = 0; k < T; ++k) {
        const size_t idx = (size_t)k * nc + cell;
        if (ARRC && k > 0) hs = (c4->precip[idx] > 0) ? 0 : hs + 1;
        if ((k % 24) == 0) { // daily mean of the ground pack's temperature
            Tzd = na_real();
            if (tzd_ok && k / 24 < ndays) {
                double sum = 0.0;
                for (int hh = 0; hh < 24; ++hh) sum += a.snowtempg[(size_t)(k + hh) * nc + cell];
                Tzd = sum / 24.0;
            }
        }
        const double swe = a.swe[idx];
        if (!(swe > 0.0)) continue;
        const double sdepg = a.sdepg[idx];
        const double reqhgts = a.reqhgt - sdepg;
        if (reqhgts >= 0.0) {
            SnowHr hloc;
            if (ARRC) {
                const SolarPos sp = solar_position(lat, lon, c4->year[k], c4->month[k], c4->day[k], c4->hour[k]);
                snow_hour(hloc, c4->temp[idx], c4->relhum[idx], c4->pres[idx], c4->windspeed[idx], c4->swdown[idx], c4->difrad[idx],
                          c4->lwdown[idx], 0.0, 0.0, 0.0, c4->umu[idx], c4->winddir[k], sp, hs, no_extremes());
            }
            const SnowHr& h = ARRC ? hloc : a.hours[k];
            double si = solar_index(c.sc, h, true);
            if (si != si) si = h.cosz;
            const bool shadow = a.hor[(size_t)h.sindex * nc + cell] > h.tan_alt_rad;
            const double ws = a.wsa[(size_t)h.windex * nc + cell];
            const double sden = a.sden[idx];
            const MicroOut o = snow_micro_above(c, h, reqhgts, a.zref, si, shadow, ws, dTmx, a.snowtempg[idx], a.snowtempc[idx],
                                                swe / sden, sdepg, sden);
            micro_store(a, idx, o);
        } else {
            micro_store_buried(a, idx, snow_micro_below(reqhgts, meanD, a.snowtempg[idx], Tzd, a.mat, a.hiy));
        }
        if (a.out[3]) a.out[3][idx] = c.Smax;
    }
}

} // namespace snowphys
} // namespace mcf
