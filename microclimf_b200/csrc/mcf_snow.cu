// microclimf_b200 — snow kernels (SURVEY.md §8f NEXT-3).
//
//   k_snowmodel[_arr]   gridmodelsnow1 / 2  src/microclimfCpp.cpp:4172-4424, 4426-4673   per-cell HOURLY RECURRENCE of the
//                       snow pack: depth, density and age of the canopy+ground and ground-only layers carried from hour
//                       to hour in registers — the only true hour-to-hour recurrence of the package
//   k_snowmicro[_arr]   gridmicrosnow1 / 2  src/microclimfCpp.cpp:4894-5057, 5059-5214   microclimate above / inside the
//                       snow pack for cell-hours with snow water equivalent > 0, overwriting runmicro's outputs in place
//   k_snow_prep         the per-hour table of the data.frame drivers: everything the physics needs that depends on the
//                       hour alone (mcf_snow_physics.cuh: SnowHr), built once per call
//
// One thread per cell, cells fastest (R layout): every [rows, cols, hours] access of a warp is one contiguous 256-byte
// segment.  The physics (mcf_snow_physics.cuh) and the per-cell walks over the series (mcf_snow_drivers.cuh) are
// re-derived for this execution model — hour-only terms hoisted into the table, the canopy geometry above the pack formed
// once per snow hour, the two-stream solution specialised to the snow case, transcendentals through the branch-free
// MUFU-seeded functions of mcf_math.cuh; see the header of mcf_snow_physics.cuh for the list.  Bound by FP64 latency
// like k_grid: the recurrence leaves one dependent chain per thread, so occupancy (256-thread CTAs, no shared memory)
// is what hides it.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "mcf_kernels.cuh"
#include "mcf_snow_drivers.cuh"

namespace mcf {
namespace snow {

using namespace ::mcf::snowphys;

struct SnowPrepArgs {
    SnowSeries s;
    SnowHr* hours;
    double* scal;      // [0] series maximum of air temperature
    int32_t* hs;       // scratch [tsteps]: hours since the last precipitation (sequential scan, ref snowalbCpp :3752-3771)
    DayExtremes* days; // scratch [tsteps / 24]
};

// one block: thread 0 runs the sequential age scan while the others gather the daily extremes; then one record per thread
__global__ void __launch_bounds__(256) k_snow_prep(const __grid_constant__ SnowPrepArgs a) {
    const int T = a.s.tsteps, ndays = T / 24;
    if (threadIdx.x == 0) {
        int hs = 0;
        double mx = -273.15;
        for (int k = 0; k < T; ++k) {
            if (k > 0) hs = (a.s.precip[k] > 0) ? 0 : hs + 1;
            a.hs[k] = hs;
            if (a.s.temp[k] > mx) mx = a.s.temp[k];
        }
        a.scal[0] = mx;
    } else if (a.s.RswabsG) {
        for (int d = threadIdx.x - 1; d < ndays; d += blockDim.x - 1)
            day_extremes(a.days[d], a.s.RswabsG, a.s.RlwabsG, a.s.temp, a.s.swdown, a.s.lwdown, (size_t)d * 24, 1);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < T; k += blockDim.x) {
        const bool whole = a.s.RswabsG && (k / 24 < ndays); // only whole days have extremes (ref :4222-4283)
        prep_hour(a.s, k, a.hs[k], whole ? a.days[k / 24] : no_extremes(), a.hours[k]);
    }
}

__global__ void __launch_bounds__(256) k_snowmodel(const __grid_constant__ SnowModelArgs a) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell < a.rows * a.cols) snowmodel_cell(a, cell);
}
__global__ void __launch_bounds__(256) k_snowmodel_arr(const __grid_constant__ SnowModelArgs a, const __grid_constant__ SnowArr c) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell < a.rows * a.cols) snowmodel_cell_arr(a, c, cell);
}
__global__ void __launch_bounds__(256) k_snowmicro(const __grid_constant__ SnowMicroArgs a) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell < a.rows * a.cols) snowmicro_cell_t<false>(a, nullptr, cell);
}
__global__ void __launch_bounds__(256) k_snowmicro_arr(const __grid_constant__ SnowMicroArgs a, const __grid_constant__ SnowArr c) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell < a.rows * a.cols) snowmicro_cell_t<true>(a, &c, cell);
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
using namespace mcf::snowphys;

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
               const double* RlwabsG, const double* umu, double lat, double lon, SnowHr** hours, double** scal, char* err,
               size_t errlen) {
    const size_t T = (size_t)c->tsteps;
    SnowPrepArgs pa;
    std::memset(&pa, 0, sizeof pa);
    pa.s.tsteps = c->tsteps;
    SCU(db.up(c->year, T, &pa.s.year));
    SCU(db.up(c->month, T, &pa.s.month));
    SCU(db.up(c->day, T, &pa.s.day));
    SCU(db.up(c->hour, T, &pa.s.hour));
    SCU(db.up(c->temp, T, &pa.s.temp));
    SCU(db.up(c->relhum, T, &pa.s.relhum));
    SCU(db.up(c->pres, T, &pa.s.pres));
    SCU(db.up(c->swdown, T, &pa.s.swdown));
    SCU(db.up(c->difrad, T, &pa.s.difrad));
    SCU(db.up(c->lwdown, T, &pa.s.lwdown));
    SCU(db.up(c->windspeed, T, &pa.s.windspeed));
    SCU(db.up(c->winddir, T, &pa.s.winddir));
    SCU(db.up(c->precip, T, &pa.s.precip));
    SCU(db.up(Gp, T, &pa.s.Gp));
    SCU(db.up(Tcp, T, &pa.s.Tcp));
    SCU(db.up(RswabsG, T, &pa.s.RswabsG));
    SCU(db.up(RlwabsG, T, &pa.s.RlwabsG));
    SCU(db.up(umu, T, &pa.s.umu));
    pa.s.lat = lat;
    pa.s.lon = lon;
    SCU(db.alloc(hours, T));
    SCU(db.alloc(scal, 4));
    SCU(db.alloc(&pa.hs, T));
    SCU(db.alloc(&pa.days, T / 24 + 1));
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
    SnowHr* hours = nullptr;
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
    if (arr) k_snowmodel_arr<<<(unsigned)((nc + 255) / 256), 256>>>(a, ca);
    else k_snowmodel<<<(unsigned)((nc + 255) / 256), 256>>>(a);
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
    SnowHr* hours = nullptr;
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
    if (arr) k_snowmicro_arr<<<(unsigned)((nc + 255) / 256), 256>>>(a, ca);
    else k_snowmicro<<<(unsigned)((nc + 255) / 256), 256>>>(a);
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
