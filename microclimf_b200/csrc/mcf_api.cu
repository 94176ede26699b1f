// microclimf_b200 — C ABI (include/microclimf_b200.h): host orchestration around the kernels.
//
//   mcf_runmicro_dev : prepare (hour table / calendar, twi mean, day-block list) -> grid kernel over the
//                      requested window -> (reqhgt < 0) below-ground kernel, all on one stream.
//   mcf_runmicro     : host buffers.  Statics and forcing are uploaded once; the time axis is streamed
//                      through two device output chunks so the device->host copy of chunk i overlaps the
//                      kernels of chunk i+1 (outputs of large rasters do not fit HBM).
//   mcf_runbioclim*  : runmicro with outm = {Tz|tleaf, soilm} into device scratch, then the 19 reductions.
//
// No CPU compute path exists here: without a usable CUDA device every entry point fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "mcf_kernels.cuh"
#include "mcf_host.h"
#include "microclimf_b200.h"

using namespace mcf;

namespace {

std::atomic<int64_t> g_launches{0};
std::mutex g_time_mu;
bool g_timing = false;
std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_events;
double g_time_ms = 0.0;
int64_t g_time_n = 0;

struct Err {
    int code = MCF_OK;
    std::string msg;
};

int report(const Err& e, char* err, size_t errlen) {
    if (err && errlen) std::snprintf(err, errlen, "%s", e.msg.c_str());
    return e.code;
}
Err make_err(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    std::vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    Err e;
    e.code = code;
    e.msg = buf;
    return e;
}
#define CU(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t _e = (call);                                                                              \
        if (_e != cudaSuccess)                                                                                \
            return make_err(_e == cudaErrorMemoryAllocation ? MCF_ERR_NOMEM : MCF_ERR_CUDA, "%s failed: %s",  \
                            #call, cudaGetErrorString(_e));                                                   \
    } while (0)
#define TRY(expr)                       \
    do {                                \
        Err _r = (expr);                \
        if (_r.code != MCF_OK) return _r; \
    } while (0)

// stream-ordered scratch allocations, released when the holder dies
struct Scratch {
    cudaStream_t stream;
    std::vector<void*> ptrs;
    explicit Scratch(cudaStream_t s) : stream(s) {}
    ~Scratch() {
        for (void* p : ptrs) cudaFreeAsync(p, stream);
    }
    template <class T> cudaError_t alloc(T** p, size_t n) {
        void* q = nullptr;
        cudaError_t e = cudaMallocAsync(&q, std::max<size_t>(n, 1) * sizeof(T), stream);
        if (e == cudaSuccess) ptrs.push_back(q);
        *p = (T*)q;
        return e;
    }
};

const double* const* clim_ptrs(const mcf_problem* p, const double* out[10]) {
    out[0] = p->temp; out[1] = p->es; out[2] = p->ea; out[3] = p->tdew; out[4] = p->pres;
    out[5] = p->swdown; out[6] = p->difrad; out[7] = p->lwdown; out[8] = p->windspeed; out[9] = p->winddir;
    return out;
}

Err validate(const mcf_problem* p) {
    if (!p) return make_err(MCF_ERR_ARG, "problem is NULL");
    if (p->mode < 1 || p->mode > 4) return make_err(MCF_ERR_ARG, "mode must be 1..4 (got %d)", p->mode);
    if (p->rows <= 0 || p->cols <= 0 || p->tsteps <= 0) return make_err(MCF_ERR_ARG, "rows, cols, tsteps must be > 0");
    if ((int64_t)p->rows * p->cols > INT32_MAX - 256) return make_err(MCF_ERR_ARG, "rows*cols exceeds 2^31");
    const bool layered = p->mode >= 3;
    const bool arr = (p->mode == 2 || p->mode == 4);
    const bool coarse = arr && p->clim_rows > 0;
    if (p->clim_rows != 0 && !arr) return make_err(MCF_ERR_ARG, "coarse-grid climate (clim_rows > 0) needs mode 2 or 4");
    if (coarse) {
        if (p->clim_cols <= 0) return make_err(MCF_ERR_ARG, "clim_cols must be > 0");
        if (!p->relhum || !p->wu || !p->wv) return make_err(MCF_ERR_ARG, "coarse-grid climate needs relhum, wu and wv");
        if (p->altcorrect < 0 || p->altcorrect > 2) return make_err(MCF_ERR_ARG, "altcorrect must be 0, 1 or 2");
        if (p->altcorrect && (!p->elevd || !p->pfac)) return make_err(MCF_ERR_ARG, "altcorrect needs elevd and pfac");
    }
    // es / ea / tdew / windspeed are derived in the kernel for coarse-grid climate
    const double* const need_es = coarse ? p->temp : p->es;
    const double* const need_ea = coarse ? p->temp : p->ea;
    const double* const need_td = coarse ? p->temp : p->tdew;
    const double* const need_ws = coarse ? p->temp : p->windspeed;
    const void* req[] = {p->year, p->month, p->day, p->hour, p->temp, need_es, need_ea, need_td, p->pres, p->swdown,
                         p->difrad, p->lwdown, need_ws, p->winddir, p->p_soilm, p->p_G, p->p_umu, p->p_kp,
                         p->p_muGp, p->p_dtrp, p->hgt, p->pai, p->x, p->gsmax, p->leafr, p->leaft, p->clump,
                         p->leafd, p->paia, p->leafden, p->Smin, p->Smax, p->gref, p->soilb, p->Psie, p->Vq, p->Vm,
                         p->Mc, p->rho, p->slope, p->aspect, p->twi, p->svfa, p->wsa, p->hor};
    for (const void* q : req)
        if (!q) return make_err(MCF_ERR_ARG, "a required input pointer is NULL");
    if (p->reqhgt < 0 && (!p->p_Tg || !p->p_Tbp) && !p->complete)
        return make_err(MCF_ERR_ARG, "reqhgt < 0 with an incomplete series needs pointm Tg and Tbp");
    if (arr && (!p->lats || !p->lons)) return make_err(MCF_ERR_ARG, "modes 2/4 need lats and lons");
    if (layered) {
        if (p->nlyr < 1 || !p->lyr_st || !p->lyr_ed) return make_err(MCF_ERR_ARG, "modes 3/4 need nlyr >= 1 and dfsel");
    }
    return Err();
}

// day-block list (ref :2194 for modes 1/2; :2629-2639 + :2770-2800 for modes 3/4)
Err build_blocks(const mcf_problem* p, std::vector<DayBlock>& blocks) {
    blocks.clear();
    if (p->mode <= 2) {
        const int nd = p->tsteps / 24;
        for (int d = 0; d < nd; ++d) blocks.push_back(DayBlock{24 * d, 0});
        return Err();
    }
    int prev_end = 0;
    for (int l = 0; l < p->nlyr; ++l) {
        const int span = p->lyr_ed[l] - p->lyr_st[l] + 1;
        if (span < 24) // the reference's Rcpp::stop (src/microclimfCpp.cpp:2636-2637)
            return make_err(MCF_ERR_ARG, "Too many layers in vegp. Max layers must be <= max days");
        const int nd = span / 24;
        if (p->lyr_st[l] < prev_end)
            return make_err(MCF_ERR_ARG, "dfsel layers must be ascending and non-overlapping (layer %d)", l + 1);
        if (p->lyr_st[l] + 24 * nd > p->tsteps)
            return make_err(MCF_ERR_ARG, "dfsel layer %d runs past the end of the series", l + 1);
        for (int d = 0; d < nd; ++d) blocks.push_back(DayBlock{p->lyr_st[l] + 24 * d, l});
        prev_end = p->lyr_st[l] + 24 * nd;
    }
    return Err();
}

int rq_of(double reqhgt) { return reqhgt > 0.0 ? RQ_ABOVE : (reqhgt == 0.0 ? RQ_SURFACE : RQ_BELOW); }

bool kernel_writes(int rq, int v) {
    switch (v) {
    case MCF_OUT_TZ: return true;
    case MCF_OUT_TLEAF:
    case MCF_OUT_RELHUM: return rq == RQ_ABOVE;
    case MCF_OUT_RLWDOWN:
    case MCF_OUT_RLWUP: return rq != RQ_BELOW;
    default: return true;
    }
}

int g_sm_count = 0;
Err device_info() {
    if (g_sm_count) return Err();
    int dev = 0;
    CU(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10)
        return make_err(MCF_ERR_CUDA, "device %s is sm_%d%d; this library is built for sm_100a only", prop.name,
                        prop.major, prop.minor);
    g_sm_count = prop.multiProcessorCount;
    // Scratch is stream-ordered (cudaMallocAsync): keep freed blocks in the pool instead of returning them to the
    // driver at every synchronisation — the bioclim path alone re-allocates 2 x rows*cols*336 doubles per call
    // (22.6 GB for config 3), the windowed drivers a stash and a few tables per launch.
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t keep = UINT64_MAX;
        (void)cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    // The day stash of k_grid is written and read back with the L2 evict_last ("persisting") policy, which only has
    // somewhere to persist if part of L2 is set aside for it (the default set-aside is 0).  MCF_L2_PERSIST_MB overrides
    // the size (0 = leave the device as it is).
    // 48 MB holds the live stash of 148 CTAs; 64 MB and more starve the write stream of the outputs of L2 ways and
    // cost 4-20 %.
    size_t persist = std::min((size_t)prop.persistingL2CacheMaxSize, (size_t)48 << 20);
    if (const char* e = std::getenv("MCF_L2_PERSIST_MB"))
        persist = std::min((size_t)prop.persistingL2CacheMaxSize, (size_t)std::atoll(e) << 20);
    if (persist > 0) (void)cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist);
    (void)cudaGetLastError();
    return Err();
}

// MCF_NO_PAIR=1 in the environment keeps every launch on k_grid (A/B measurements of the two builds)
const bool g_use_pair = [] { const char* e = std::getenv("MCF_NO_PAIR"); return !(e && e[0] == '1'); }();

void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// A prepared problem on the device: everything a window launch needs.
struct Plan {
    const mcf_problem* p = nullptr;
    int arr = 0; // 0: per-hour table (modes 1/3), 1: fine [rows, cols, T] arrays, 2: coarse arrays (modes 2/4)
    int rq = RQ_ABOVE;
    int ncells = 0;
    std::vector<DayBlock> blocks;
    DayBlock* d_blocks = nullptr;
    HourRec* d_hours = nullptr;
    HourCal* d_cal = nullptr;
    double* d_scal = nullptr;      // [0] mxtc, [1] twi sum, [2] twi count, then reduction scratch
    double* d_mxtc_cell = nullptr;
    double* d_stash = nullptr;
    double* d_cpack = nullptr;     // coarse-grid climate: one 128-byte record per (hour, coarse node), see k_pack_coarse
    int grid = 0;
    // element type of the outputs: 0 FP64; 1 int16 (the packed integer sink); 2 FP32 (the FP32 build: k_grid_f32).
    // out[v] pointers are int16_t* / float* in disguise for 1 / 2.
    int pack = 0;
    char* d_hoursf = nullptr;  // FP32 build, modes 1/3: the narrowed hour table
    float* d_stashf = nullptr; // FP32 build: its day stash
    // below ground the FP32 build runs the FP64 hour loops (see mcf_kernels_f32.inl) and narrows its outputs
    bool f32_kernel() const { return pack == 2 && rq != RQ_BELOW; }
    int tile() const { return f32_kernel() ? f32_tile() : kTile; }
};

void fill_common(const Plan& pl, GridArgs& a);

Err plan_prepare(Plan& pl, const mcf_problem* p, Scratch& sc, cudaStream_t st) {
    TRY(validate(p));
    TRY(device_info());
    pl.p = p;
    pl.arr = (p->mode == 2 || p->mode == 4) ? (p->clim_rows > 0 ? 2 : 1) : 0;
    pl.rq = rq_of(p->reqhgt);
    pl.ncells = p->rows * p->cols;
    TRY(build_blocks(p, pl.blocks));
    const int T = p->tsteps;
    CU(sc.alloc(&pl.d_blocks, pl.blocks.size()));
    if (!pl.blocks.empty())
        CU(cudaMemcpyAsync(pl.d_blocks, pl.blocks.data(), pl.blocks.size() * sizeof(DayBlock), cudaMemcpyHostToDevice, st));
    CU(sc.alloc(&pl.d_scal, 4 + 2 * 256 + 2));
    // calendar ints are host arrays by contract: stage them
    int32_t* d_cal3 = nullptr;
    CU(sc.alloc(&d_cal3, (size_t)3 * T));
    CU(cudaMemcpyAsync(d_cal3, p->year, T * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_cal3 + T, p->month, T * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_cal3 + 2 * T, p->day, T * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    if (pl.arr) CU(sc.alloc(&pl.d_cal, T));
    else CU(sc.alloc(&pl.d_hours, T));
    const double* clim[10];
    clim_ptrs(p, clim);
    const double* pnt[6] = {p->p_soilm, p->p_G, p->p_umu, p->p_kp, p->p_muGp, p->p_dtrp};
    CU(launch_prep_hours(d_cal3, d_cal3 + T, d_cal3 + 2 * T, p->hour, clim, pnt, p->lat, p->lon, T, pl.arr != 0, pl.d_hours,
                         pl.d_cal, pl.d_scal, st));
    count_launch();
    if (pl.arr) {
        CU(sc.alloc(&pl.d_mxtc_cell, pl.ncells));
        if (pl.arr == 2) {
            GridArgs ga;
            fill_common(pl, ga);
            CU(launch_mxtc_cell_coarse(ga, pl.d_mxtc_cell, st));
            if (!pl.f32_kernel()) { // the FP64 grid kernel reads the coarse series as packed node records
                CU(sc.alloc(&pl.d_cpack, coarse_pack_doubles(ga)));
                CU(launch_pack_coarse(ga, pl.d_cpack, st));
                count_launch();
            }
        } else {
            CU(launch_mxtc_cell(p->temp, pl.ncells, T, pl.d_mxtc_cell, st));
        }
        count_launch();
    }
    if (!p->has_twi_mean) {
        CU(launch_twi_sum(p->twi, pl.ncells, p->tfact, pl.d_scal + 1, st));
        count_launch(2);
    }
    if (pl.f32_kernel()) { // FP32 build: its own CTA shape, stash and (modes 1/3) narrowed hour table
        pl.grid = g_sm_count * f32_blocks_per_sm();
        CU(sc.alloc(&pl.d_stashf, (size_t)pl.grid * 24 * kStashVars * f32_tile()));
        if (!pl.arr) {
            CU(sc.alloc(&pl.d_hoursf, (size_t)T * hourrec_f32_bytes()));
            CU(launch_narrow_hours(pl.d_hours, T, pl.d_hoursf, st));
            count_launch();
        }
        return Err();
    }
    pl.grid = g_sm_count * grid_blocks_per_sm(pl.arr != 0, pl.rq);
    // one scratch serves both builds of the FP64 kernel (k_grid's day stash; k_grid_pair's stash + reduction exchange)
    CU(sc.alloc(&pl.d_stash, (size_t)pl.grid * std::max((size_t)24 * kStashVars * kTile, pair_scratch_doubles())));
    return Err();
}

void fill_common(const Plan& pl, GridArgs& a) {
    const mcf_problem* p = pl.p;
    std::memset(&a, 0, sizeof a);
    a.ncells = pl.ncells;
    a.tsteps = p->tsteps;
    a.nlyr = p->mode >= 3 ? p->nlyr : 1;
    a.reqhgt2 = p->reqhgt < 0.00001 ? 0.00001 : p->reqhgt;
    a.zref = p->zref;
    a.lat = p->lat;
    a.tfact = p->tfact;
    a.has_tadd_mean = p->has_twi_mean;
    a.tadd_mean = p->twi_mean;
    a.dscal = pl.d_scal;
    a.hours = pl.d_hours;
    const double* clim[10];
    clim_ptrs(p, clim);
    for (int i = 0; i < 9; ++i) a.clim[i] = clim[i];
    a.pnt[0] = p->p_soilm; a.pnt[1] = p->p_G; a.pnt[2] = p->p_umu; a.pnt[3] = p->p_kp; a.pnt[4] = p->p_muGp;
    a.pnt[5] = p->p_dtrp;
    a.lats = p->lats;
    a.lons = p->lons;
    a.mxtc_cell = pl.d_mxtc_cell;
    a.cal = pl.d_cal;
    const double* veg[10] = {p->hgt, p->pai, p->x, p->gsmax, p->leafr, p->leaft, p->clump, p->leafd, p->paia, p->leafden};
    for (int i = 0; i < 10; ++i) a.veg[i] = veg[i];
    const double* soil[13] = {p->Smin, p->Smax, p->gref, p->soilb, p->Psie, p->Vq, p->Vm, p->Mc, p->rho, p->slope,
                              p->aspect, p->twi, p->svfa};
    for (int i = 0; i < 13; ++i) a.soil[i] = soil[i];
    a.wsa = p->wsa;
    a.hor = p->hor;
    a.blocks = pl.d_blocks;
    a.stash = pl.d_stash;
    a.pack = pl.pack;
    a.rows = p->rows;
    a.out_stride = pl.ncells;
    a.out_cell0 = 0;
    if (pl.arr == 2) {
        a.clim_rows = p->clim_rows;
        a.clim_cols = p->clim_cols;
        a.altcorrect = p->altcorrect;
        a.clim_row0 = p->clim_row0;
        a.clim_drow = p->clim_drow;
        a.clim_col0 = p->clim_col0;
        a.clim_dcol = p->clim_dcol;
        a.relhum = p->relhum;
        a.wu = p->wu;
        a.wv = p->wv;
        a.elevd = p->elevd;
        a.pfac = p->pfac;
        a.cpack = pl.d_cpack;
    }
}

// one launch of the grid kernel of the plan's build over cells [a.cell_begin, a.cell_end), timed if timing is on
Err timed_grid_launch(const Plan& pl, const GridArgs& a, cudaStream_t st, int sink = -1) {
    // the pair build (two threads per cell, invariants in shared memory) serves the per-hour-table drivers above ground
    const int eff_sink = sink >= 0 ? sink : (a.pack == 1 ? SINK_PACK : SINK_F64);
    const bool pair = !pl.f32_kernel() && g_use_pair && pair_eligible(pl.arr, pl.rq, eff_sink);
    const int tl = pair ? pair_tile() : pl.tile();
    const int ntiles = (a.cell_end - a.cell_begin + tl - 1) / tl;
    const int grid = std::min(pair ? g_sm_count : pl.grid, ntiles);
    if (grid <= 0) return Err();
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (g_timing) {
        CU(cudaEventCreate(&e0));
        CU(cudaEventCreate(&e1));
        CU(cudaEventRecord(e0, st));
    }
    if (pl.f32_kernel()) {
        if (sink >= 0) return make_err(MCF_ERR_ARG, "the reducing sinks run in the FP64 build");
        CU(launch_grid_f32(a, pl.d_hoursf, reinterpret_cast<float* const*>(a.out), pl.d_stashf, pl.arr, pl.rq, grid, st));
    } else if (pair) {
        CU(launch_grid_pair(a, pl.arr, pl.rq, grid, st, sink));
    } else {
        CU(launch_grid(a, pl.arr, pl.rq, grid, st, sink));
    }
    count_launch();
    if (g_timing) {
        CU(cudaEventRecord(e1, st));
        std::lock_guard<std::mutex> lk(g_time_mu);
        g_events.emplace_back(e0, e1);
    }
    return Err();
}

// Run day-blocks [b0, b0+nb) into out[] (device).  rq != RQ_BELOW.
Err plan_run_window(const Plan& pl, double* const out[MCF_NOUT], int b0, int nb, long long hour0, long long ring,
                    Scratch& sc, cudaStream_t st) {
    if (nb <= 0) return Err();
    GridArgs a;
    fill_common(pl, a);
    a.cell_begin = 0;
    a.cell_end = pl.ncells;
    a.block0 = b0;
    a.nblocks = nb;
    a.hour0 = hour0;
    a.ring_hours = ring;
    a.outmask = 0;
    for (int v = 0; v < MCF_NOUT; ++v) {
        a.out[v] = out[v];
        if (out[v] && kernel_writes(pl.rq, v)) a.outmask |= 1u << v;
    }
    unsigned int* ctr = nullptr;
    CU(sc.alloc(&ctr, 1));
    CU(cudaMemsetAsync(ctr, 0, sizeof(unsigned int), st));
    a.tile_counter = ctr;
    return timed_grid_launch(pl, a, st);
}

// Whole series for reqhgt < 0: grid kernel writes Tg into scratch per cell chunk, then the time-axis pass.
// `bio` != NULL (runbioclim below ground): nothing hourly leaves the chunk — its below-ground Tz and soil moisture land
// in two chunk-sized series ([tsteps][W], allocated here) and are reduced to the chunk's cells of the 19 summaries
// before the next chunk reuses them, so the scratch is bounded whatever the raster size.
Err plan_run_below(const Plan& pl, double* const out[MCF_NOUT], Scratch& sc, cudaStream_t st, BioArgs* bio = nullptr) {
    const mcf_problem* p = pl.p;
    const int T = p->tsteps;
    const int numDays = T / 24;
    size_t freeb = 0, totalb = 0;
    CU(cudaMemGetInfo(&freeb, &totalb));
    size_t budget = std::max<size_t>(freeb / 4, (size_t)64 << 20);
    if (bio) budget = std::min<size_t>(budget, (size_t)1 << 30) / 3;
    long long wmax = (long long)(budget / ((size_t)T * sizeof(double)));
    const long long tl = pl.tile();
    wmax = std::max<long long>(tl, (wmax / tl) * tl);
    const int W = (int)std::min<long long>(wmax, ((pl.ncells + tl - 1) / tl) * tl);
    double *tg = nullptr, *dds = nullptr, *daily = nullptr, *ctz = nullptr, *csm = nullptr;
    CU(sc.alloc(&tg, (size_t)T * W));
    CU(sc.alloc(&dds, W));
    CU(sc.alloc(&daily, (size_t)2 * std::max(numDays, 1) * W));
    if (bio) {
        CU(sc.alloc(&ctz, (size_t)T * W));
        CU(sc.alloc(&csm, (size_t)T * W));
    }
    const int nchunks = (pl.ncells + W - 1) / W;
    unsigned int* ctr = nullptr;
    CU(sc.alloc(&ctr, nchunks));
    CU(cudaMemsetAsync(ctr, 0, nchunks * sizeof(unsigned int), st));
    int hiy = 365 * 24;
    if (p->year[0] % 4 == 0) hiy = 366 * 24; // ref :2171-2172
    // packed sink: the time-axis pass produces FP64; it lands in a scratch series and is packed afterwards
    double* tz64 = nullptr;
    if (pl.pack && out[MCF_OUT_TZ] && !bio) CU(sc.alloc(&tz64, (size_t)T * pl.ncells));
    // FP32 build: the hour loops run in FP64 below ground (mcf_kernels_f32.inl); what pass 1 writes (soil moisture, wind,
    // shortwave streams) lands in FP64 scratch series, NA where no day-block covers the hour, and is narrowed afterwards
    double* f64s[MCF_NOUT] = {};
    if (pl.pack == 2 && !bio)
        for (int v = 0; v < MCF_NOUT; ++v)
            if (out[v] && v != MCF_OUT_TZ && kernel_writes(pl.rq, v)) {
                CU(sc.alloc(&f64s[v], (size_t)T * pl.ncells));
                CU(launch_fill_na(f64s[v], (int64_t)T * pl.ncells, st));
                count_launch();
            }
    // coarse-grid climate: the time-axis pass reads the point model's Tg / Tbz per cell-hour; expand them once
    const double *tgp = p->p_Tg, *tbp = p->p_Tbp;
    if (pl.arr == 2 && out[MCF_OUT_TZ] && tgp && tbp) {
        GridArgs ga;
        fill_common(pl, ga);
        double *f1 = nullptr, *f2 = nullptr;
        CU(sc.alloc(&f1, (size_t)T * pl.ncells));
        CU(sc.alloc(&f2, (size_t)T * pl.ncells));
        CU(launch_interp_coarse(ga, tgp, f1, st));
        CU(launch_interp_coarse(ga, tbp, f2, st));
        count_launch(2);
        tgp = f1;
        tbp = f2;
    }
    for (int ch = 0; ch < nchunks; ++ch) {
        const int c0 = ch * W, c1 = std::min(pl.ncells, c0 + W);
        // uncovered hours keep Tg = 0, DD = 0, as the reference's zero-initialised vectors (:2192-2193)
        CU(cudaMemsetAsync(tg, 0, (size_t)T * (c1 - c0) * sizeof(double), st));
        CU(cudaMemsetAsync(dds, 0, (size_t)(c1 - c0) * sizeof(double), st));
        GridArgs a;
        fill_common(pl, a);
        a.cell_begin = c0;
        a.cell_end = c1;
        a.block0 = 0;
        a.nblocks = (int)pl.blocks.size();
        a.hour0 = 0;
        a.ring_hours = T;
        a.outmask = 0;
        for (int v = 0; v < MCF_NOUT; ++v) {
            a.out[v] = f64s[v] ? f64s[v] : out[v];
            if (out[v] && v != MCF_OUT_TZ && kernel_writes(pl.rq, v)) a.outmask |= 1u << v;
        }
        if (bio) { // the chunk's soil moisture into its compact series; hours no day-block covers stay NA
            CU(launch_fill_na(csm, (int64_t)T * W, st));
            count_launch();
            for (int v = 0; v < MCF_NOUT; ++v) a.out[v] = nullptr;
            a.out[MCF_OUT_SOILM] = csm;
            a.outmask = 1u << MCF_OUT_SOILM;
            a.out_stride = W;
            a.out_cell0 = c0;
        }
        a.tile_counter = ctr + ch;
        a.tg_scratch = tg;
        a.dd_sum = dds;
        if (a.nblocks > 0) TRY(timed_grid_launch(pl, a, st));
        if (out[MCF_OUT_TZ] || bio) {
            BelowArgs b;
            std::memset(&b, 0, sizeof b);
            b.width = c1 - c0;
            b.ncells = pl.ncells;
            b.tz_stride = bio ? W : pl.ncells;
            b.tz_cell0 = bio ? c0 : 0;
            b.cell_begin = c0;
            b.tsteps = T;
            b.arr = pl.arr ? 1 : 0;
            b.complete = p->complete;
            b.hiy = hiy;
            b.reqhgt = p->reqhgt;
            b.mat = p->mat;
            b.tg = tg;
            b.dd_sum = dds;
            b.Tgp = tgp;
            b.Tbp = tbp;
            b.hgt = p->hgt;
            b.daily = daily;
            b.Tz = bio ? ctz : (tz64 ? tz64 : out[MCF_OUT_TZ]);
            CU(launch_below(b, st));
            count_launch();
        }
        if (bio) {
            BioArgs bb = *bio;
            bb.width = c1 - c0;
            bb.stride = W;
            bb.Tz = ctz;
            bb.soilm = csm;
            bb.cell_begin = c0;
            CU(launch_bioclim(bb, st));
            count_launch();
        }
    }
    if (tz64) {
        if (pl.pack == 2) CU(launch_narrow32(tz64, reinterpret_cast<float*>(out[MCF_OUT_TZ]), (int64_t)T * pl.ncells, st));
        else CU(launch_pack16(tz64, reinterpret_cast<int16_t*>(out[MCF_OUT_TZ]), (int64_t)T * pl.ncells, 100.0, st));
        count_launch();
    }
    for (int v = 0; v < MCF_NOUT; ++v)
        if (f64s[v]) {
            CU(launch_narrow32(f64s[v], reinterpret_cast<float*>(out[v]), (int64_t)T * pl.ncells, st));
            count_launch();
        }
    return Err();
}

// NA prefill of whatever the kernels will not write in a whole-series run
Err prefill_whole(const Plan& pl, double* const out[MCF_NOUT], cudaStream_t st) {
    const int T = pl.p->tsteps;
    std::vector<char> covered(T, 0);
    for (const DayBlock& b : pl.blocks)
        for (int h = 0; h < 24; ++h) covered[b.k0 + h] = 1;
    bool gaps = false;
    for (int k = 0; k < T; ++k) gaps |= !covered[k];
    for (int v = 0; v < MCF_NOUT; ++v) {
        if (!out[v]) continue;
        const bool written = kernel_writes(pl.rq, v);
        const bool all_hours = (pl.rq == RQ_BELOW && v == MCF_OUT_TZ); // the time-axis pass writes every hour
        if (!written || (gaps && !all_hours)) {
            if (pl.pack == 1) CU(launch_fill16(reinterpret_cast<int16_t*>(out[v]), (int64_t)T * pl.ncells, (int16_t)-9999, st));
            else if (pl.pack == 2) CU(launch_fill32(reinterpret_cast<float*>(out[v]), (int64_t)T * pl.ncells, st));
            else CU(launch_fill_na(out[v], (int64_t)T * pl.ncells, st));
            count_launch();
        }
    }
    return Err();
}

// Resolve and validate a time window against the day-block list.  The kernels place hour k of the window in ring slot
// (k - hour0) mod ring_hours with C++ '%', so hour0 must not lie beyond the window's first hour (a negative remainder
// would index in front of the output buffers).
Err resolve_window(const mcf_window* win, const std::vector<DayBlock>& blocks, int tsteps, int& b0, int& nb,
                   long long& hour0, long long& ring) {
    const int nblk = (int)blocks.size();
    b0 = 0;
    nb = nblk;
    hour0 = 0;
    ring = tsteps;
    if (!win) return Err();
    b0 = win->block0;
    nb = win->nblocks < 0 ? nblk - b0 : win->nblocks;
    hour0 = win->hour0;
    ring = win->ring_hours;
    if (b0 < 0 || nb < 0 || b0 + nb > nblk) return make_err(MCF_ERR_ARG, "window outside the %d day-blocks", nblk);
    if (ring < 24) return make_err(MCF_ERR_ARG, "ring_hours must be >= 24");
    if (hour0 < 0) return make_err(MCF_ERR_ARG, "window hour0 must be >= 0 (got %lld)", hour0);
    if (nb > 0 && hour0 > blocks[b0].k0)
        return make_err(MCF_ERR_ARG, "window hour0 (%lld) lies beyond the window's first hour (%d)", hour0, blocks[b0].k0);
    return Err();
}

Err run_dev(const mcf_problem* p, double* const out[MCF_NOUT], const mcf_window* win, cudaStream_t st, int pack = 0) {
    Scratch sc(st);
    Plan pl;
    pl.pack = pack;
    {   // argument errors are reported before any device work is queued
        TRY(validate(p));
        std::vector<DayBlock> blocks;
        TRY(build_blocks(p, blocks));
        int b0_, nb_;
        long long h0_, ring_;
        TRY(resolve_window(win, blocks, p->tsteps, b0_, nb_, h0_, ring_));
    }
    TRY(plan_prepare(pl, p, sc, st));
    const int nblk = (int)pl.blocks.size();
    int b0 = 0, nb = nblk;
    long long hour0 = 0, ring = p->tsteps;
    TRY(resolve_window(win, pl.blocks, p->tsteps, b0, nb, hour0, ring));
    const bool whole = !win || (b0 == 0 && nb == nblk && hour0 == 0 && ring >= p->tsteps);
    if (pl.rq == RQ_BELOW) {
        if (!whole) return make_err(MCF_ERR_ARG, "reqhgt < 0 needs the whole series in one window");
        TRY(prefill_whole(pl, out, st));
        return plan_run_below(pl, out, sc, st);
    }
    if (whole) TRY(prefill_whole(pl, out, st));
    return plan_run_window(pl, out, b0, nb, hour0, ring, sc, st);
}

// ------------------------------------------------------------------------------------------------
// host-buffer path
// ------------------------------------------------------------------------------------------------
// Device workspace of the host-buffer entry points: one grow-only allocation kept across calls (a
// cudaMalloc / cudaFree pair per array and call cost more than the solve itself), carved by a bump
// pointer.  Calls are serialised by g_ws_mu (the R caller is single-threaded anyway).
struct Workspace {
    void* base = nullptr;
    size_t cap = 0;
};
Workspace g_ws;
std::mutex g_ws_mu;

struct DevCopy { // device mirror of a host problem, carved out of the workspace
    size_t off = 0;
    size_t need = 0;      // sizing pass: bytes that would have been carved
    bool sizing = false;
    cudaStream_t stream = nullptr;
    static size_t pad(size_t bytes) { return (std::max<size_t>(bytes, 8) + 255) & ~(size_t)255; }
    Err reserve(size_t bytes) {
        if (bytes > g_ws.cap) {
            if (g_ws.base) CU(cudaFree(g_ws.base));
            g_ws.base = nullptr;
            g_ws.cap = 0;
            CU(cudaMalloc(&g_ws.base, bytes));
            g_ws.cap = bytes;
        }
        off = 0;
        return Err();
    }
    Err carve(void** q, size_t bytes) {
        const size_t b = pad(bytes);
        if (sizing) {
            need += b;
            *q = nullptr;
            return Err();
        }
        if (off + b > g_ws.cap) return make_err(MCF_ERR_NOMEM, "internal: device workspace exhausted");
        *q = (char*)g_ws.base + off;
        off += b;
        return Err();
    }
    Err up(const double* h, size_t n, const double** d) {
        *d = nullptr;
        if (!h) return Err();
        void* q = nullptr;
        TRY(carve(&q, n * sizeof(double)));
        if (!sizing) CU(cudaMemcpyAsync(q, h, n * sizeof(double), cudaMemcpyHostToDevice, stream));
        *d = (const double*)q;
        return Err();
    }
    Err dalloc(double** d, size_t n, size_t esz = sizeof(double)) { // n elements of esz bytes
        void* q = nullptr;
        TRY(carve(&q, n * esz));
        *d = (double*)q;
        return Err();
    }
};

Err upload_problem(const mcf_problem* h, mcf_problem* d, DevCopy& dc) {
    *d = *h;
    const bool arr = (h->mode == 2 || h->mode == 4);
    const size_t nc = (size_t)h->rows * h->cols, T = h->tsteps;
    const bool coarse = arr && h->clim_rows > 0;
    const size_t ns = coarse ? (size_t)h->clim_rows * h->clim_cols * T : (arr ? nc * T : T);
    const size_t nv = nc * (h->mode >= 3 ? h->nlyr : 1);
    TRY(dc.up(h->hour, T, &d->hour));
    TRY(dc.up(h->temp, ns, &d->temp));
    TRY(dc.up(coarse ? nullptr : h->es, ns, &d->es));
    TRY(dc.up(coarse ? nullptr : h->ea, ns, &d->ea));
    TRY(dc.up(coarse ? nullptr : h->tdew, ns, &d->tdew));
    TRY(dc.up(h->pres, ns, &d->pres));
    TRY(dc.up(h->swdown, ns, &d->swdown));
    TRY(dc.up(h->difrad, ns, &d->difrad));
    TRY(dc.up(h->lwdown, ns, &d->lwdown));
    TRY(dc.up(coarse ? nullptr : h->windspeed, ns, &d->windspeed));
    TRY(dc.up(coarse ? h->relhum : nullptr, ns, &d->relhum));
    TRY(dc.up(coarse ? h->wu : nullptr, ns, &d->wu));
    TRY(dc.up(coarse ? h->wv : nullptr, ns, &d->wv));
    TRY(dc.up(coarse && h->altcorrect ? h->elevd : nullptr, nc, &d->elevd));
    TRY(dc.up(coarse && h->altcorrect ? h->pfac : nullptr, nc, &d->pfac));
    TRY(dc.up(h->winddir, T, &d->winddir));
    TRY(dc.up(h->p_soilm, ns, &d->p_soilm));
    TRY(dc.up(h->reqhgt < 0 ? h->p_Tg : nullptr, ns, &d->p_Tg));
    TRY(dc.up(h->reqhgt < 0 ? h->p_Tbp : nullptr, ns, &d->p_Tbp));
    TRY(dc.up(h->p_G, ns, &d->p_G));
    TRY(dc.up(h->p_umu, ns, &d->p_umu));
    TRY(dc.up(h->p_kp, ns, &d->p_kp));
    TRY(dc.up(h->p_muGp, ns, &d->p_muGp));
    TRY(dc.up(h->p_dtrp, ns, &d->p_dtrp));
    TRY(dc.up(h->hgt, nv, &d->hgt));
    TRY(dc.up(h->pai, nv, &d->pai));
    TRY(dc.up(h->x, nv, &d->x));
    TRY(dc.up(h->gsmax, nv, &d->gsmax));
    TRY(dc.up(h->leafr, nv, &d->leafr));
    TRY(dc.up(h->leaft, nv, &d->leaft));
    TRY(dc.up(h->clump, nv, &d->clump));
    TRY(dc.up(h->leafd, nv, &d->leafd));
    TRY(dc.up(h->paia, nv, &d->paia));
    TRY(dc.up(h->leafden, nv, &d->leafden));
    TRY(dc.up(h->Smin, nc, &d->Smin));
    TRY(dc.up(h->Smax, nc, &d->Smax));
    TRY(dc.up(h->gref, nc, &d->gref));
    TRY(dc.up(h->soilb, nc, &d->soilb));
    TRY(dc.up(h->Psie, nc, &d->Psie));
    TRY(dc.up(h->Vq, nc, &d->Vq));
    TRY(dc.up(h->Vm, nc, &d->Vm));
    TRY(dc.up(h->Mc, nc, &d->Mc));
    TRY(dc.up(h->rho, nc, &d->rho));
    TRY(dc.up(h->slope, nc, &d->slope));
    TRY(dc.up(h->aspect, nc, &d->aspect));
    TRY(dc.up(h->twi, nc, &d->twi));
    TRY(dc.up(h->svfa, nc, &d->svfa));
    TRY(dc.up(h->wsa, nc * 8, &d->wsa));
    TRY(dc.up(h->hor, nc * 24, &d->hor));
    TRY(dc.up(arr ? h->lats : nullptr, nc, &d->lats));
    TRY(dc.up(arr ? h->lons : nullptr, nc, &d->lons));
    return Err();
}

struct StreamGuard {
    cudaStream_t s = nullptr;
    ~StreamGuard() {
        if (!s) return;
        // every exit path, errors included: nothing queued on the stream may still be using the workspace (or the
        // caller's host buffers) once the call returns and g_ws_mu is released
        (void)cudaStreamSynchronize(s);
        (void)cudaStreamDestroy(s);
        (void)cudaGetLastError();
    }
};
struct EventGuard {
    cudaEvent_t e = nullptr;
    ~EventGuard() {
        if (e) cudaEventDestroy(e);
    }
};
// ------------------------------------------------------------------------------------------------
// device -> host result copies
// ------------------------------------------------------------------------------------------------
// A caller such as R hands us ordinary pageable memory.  cudaMemcpyAsync into pageable memory is staged by
// the driver through one small bounce buffer at ~10 GB/s, a fifth of what PCIe gen5 delivers, and it is the
// largest term of the host-buffer path (80 B per cell-hour).  Pageable destinations are therefore served by
// a pool of copy threads: each owns a stream and two pinned slots, DMAs a chunk into one slot while it
// memcpy's the previous chunk out of the other, so the DMA engine and several cores' worth of memcpy
// bandwidth run concurrently.  Pinned (or registered) destinations are copied directly.
struct CopyJob {
    char* dst;
    const char* src;
    size_t bytes;
    cudaEvent_t ready; // the producing kernels have completed (may be null)
};

constexpr size_t kSlotBytes = (size_t)32 << 20;
constexpr int kMaxCopyThreads = 8;
void* g_stage = nullptr; // kMaxCopyThreads * 2 slots of pinned memory, kept between calls (guarded by g_ws_mu)

bool is_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

// `to_device` reverses the roles: the copy threads memcpy a chunk of the pageable source into a pinned slot and DMA it
// to the device while they fill the other slot (used by the snow operators, whose inputs are [rows, cols, hours]
// arrays as large as their outputs).
Err stage_copies(const std::vector<CopyJob>& jobs, cudaStream_t direct_stream, bool to_device) {
    const cudaMemcpyKind kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    // split into pinned host buffers (direct) and pageable ones (staged, chunked)
    struct Chunk { char* dst; const char* src; size_t bytes; cudaEvent_t ready; };
    std::vector<Chunk> chunks;
    for (const CopyJob& j : jobs) {
        if (!j.bytes) continue;
        if (is_pinned(to_device ? (const void*)j.src : (const void*)j.dst)) {
            if (j.ready) CU(cudaStreamWaitEvent(direct_stream, j.ready, 0));
            CU(cudaMemcpyAsync(j.dst, j.src, j.bytes, kind, direct_stream));
        } else {
            for (size_t off = 0; off < j.bytes; off += kSlotBytes)
                chunks.push_back(Chunk{j.dst + off, j.src + off, std::min(kSlotBytes, j.bytes - off), j.ready});
        }
    }
    if (!chunks.empty()) {
        int nthreads = (int)std::min<size_t>(kMaxCopyThreads, std::max(1u, std::thread::hardware_concurrency() / 2));
        nthreads = (int)std::min<size_t>(nthreads, chunks.size());
        if (!g_stage) CU(cudaMallocHost(&g_stage, kSlotBytes * 2 * kMaxCopyThreads));
        int dev = 0;
        CU(cudaGetDevice(&dev));
        std::vector<cudaError_t> status(nthreads, cudaSuccess);
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; ++t) {
            pool.emplace_back([&, t]() {
                cudaError_t e = cudaSetDevice(dev);
                cudaStream_t st = nullptr;
                cudaEvent_t ev[2] = {nullptr, nullptr};
                if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
                for (int k = 0; k < 2 && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming);
                char* slot[2] = {(char*)g_stage + kSlotBytes * (2 * t), (char*)g_stage + kSlotBytes * (2 * t + 1)};
                // chunks t, t + nthreads, ...
                std::vector<size_t> mine;
                for (size_t c = t; c < chunks.size(); c += nthreads) mine.push_back(c);
                if (to_device) {
                    // fill slot i&1 on the CPU while the DMA of the other slot is in flight
                    for (size_t i = 0; i < mine.size() && e == cudaSuccess; ++i) {
                        const Chunk& c = chunks[mine[i]];
                        if (i >= 2) e = cudaEventSynchronize(ev[i & 1]); // the slot's previous DMA has drained
                        if (e != cudaSuccess) break;
                        std::memcpy(slot[i & 1], c.src, c.bytes);
                        e = cudaMemcpyAsync(c.dst, slot[i & 1], c.bytes, kind, st);
                        if (e == cudaSuccess) e = cudaEventRecord(ev[i & 1], st);
                    }
                } else {
                    // DMA of chunk i+1 overlaps the memcpy of chunk i
                    auto issue = [&](size_t i) {
                        const Chunk& c = chunks[mine[i]];
                        if (c.ready) e = cudaStreamWaitEvent(st, c.ready, 0);
                        if (e == cudaSuccess) e = cudaMemcpyAsync(slot[i & 1], c.src, c.bytes, kind, st);
                        if (e == cudaSuccess) e = cudaEventRecord(ev[i & 1], st);
                    };
                    if (e == cudaSuccess && !mine.empty()) issue(0);
                    for (size_t i = 0; i < mine.size() && e == cudaSuccess; ++i) {
                        if (i + 1 < mine.size()) issue(i + 1);
                        if (e == cudaSuccess) e = cudaEventSynchronize(ev[i & 1]);
                        if (e == cudaSuccess) std::memcpy(chunks[mine[i]].dst, slot[i & 1], chunks[mine[i]].bytes);
                    }
                }
                if (st) {
                    const cudaError_t e2 = cudaStreamSynchronize(st);
                    if (e == cudaSuccess) e = e2;
                }
                for (int k = 0; k < 2; ++k)
                    if (ev[k]) cudaEventDestroy(ev[k]);
                if (st) cudaStreamDestroy(st);
                status[t] = e;
            });
        }
        for (auto& th : pool) th.join();
        for (cudaError_t e : status)
            if (e != cudaSuccess)
                return make_err(MCF_ERR_CUDA, "staged %s copy failed: %s", to_device ? "host->device" : "device->host",
                                cudaGetErrorString(e));
    }
    CU(cudaStreamSynchronize(direct_stream));
    return Err();
}
Err copy_back(const std::vector<CopyJob>& jobs, cudaStream_t direct_stream) { return stage_copies(jobs, direct_stream, false); }

void host_fill_na(double* p, size_t n, int pack) {
    if (pack == 1) {
        int16_t* q = reinterpret_cast<int16_t*>(p);
        for (size_t i = 0; i < n; ++i) q[i] = (int16_t)-9999;
        return;
    }
    if (pack == 2) {
        uint32_t* q = reinterpret_cast<uint32_t*>(p);
        for (size_t i = 0; i < n; ++i) q[i] = 0x7FC00000u; // quiet NaN: FP32 has no NA payload convention
        return;
    }
    const uint64_t bits = MCF_NA_REAL_BITS;
    uint64_t* q = reinterpret_cast<uint64_t*>(p);
    for (size_t i = 0; i < n; ++i) q[i] = bits;
}

// `pack`: out[v] are int16_t* (packed integer sink), else double*; esz = bytes per output element
Err run_host(const mcf_problem* hp, double* const out[MCF_NOUT], int pack = 0) {
    const size_t esz = pack == 1 ? sizeof(int16_t) : (pack == 2 ? sizeof(float) : sizeof(double));
    TRY(validate(hp));
    TRY(device_info());
    std::lock_guard<std::mutex> ws_lock(g_ws_mu);
    StreamGuard cs, xs;
    CU(cudaStreamCreateWithFlags(&cs.s, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&xs.s, cudaStreamNonBlocking));
    const size_t nc = (size_t)hp->rows * hp->cols;
    const int T = hp->tsteps;
    int nreq = 0;
    for (int v = 0; v < MCF_NOUT; ++v) nreq += out[v] != nullptr;
    if (nreq == 0) return Err();
    const int rq = rq_of(hp->reqhgt);
    const size_t per_hour = (size_t)nreq * nc * esz;
    std::vector<DayBlock> blocks;
    TRY(build_blocks(hp, blocks));
    const int nblk = (int)blocks.size();

    // ---- sizing pass: inputs + output buffers, then one workspace reservation
    DevCopy dc;
    dc.stream = cs.s;
    dc.sizing = true;
    mcf_problem dp;
    TRY(upload_problem(hp, &dp, dc));
    const size_t in_bytes = dc.need;
    size_t freeb = 0, totalb = 0;
    CU(cudaMemGetInfo(&freeb, &totalb));
    const size_t avail = freeb + g_ws.cap; // what a fresh reservation could use
    // scratch the solve allocates beside inputs and outputs: the day stash, and for reqhgt < 0 the FP64 series the
    // packed sink is packed from and the two expanded point-model series of coarse-grid climate (the ground-temperature
    // chunk of plan_run_below sizes itself from what is left)
    double scratch_bytes = 148.0 * 2 * 24 * kStashVars * kTile * sizeof(double);
    if (hp->clim_rows > 0) scratch_bytes += (double)T * hp->clim_rows * hp->clim_cols * 16 * sizeof(double); // packed node records
    if (rq == RQ_BELOW) {
        if (pack && out[MCF_OUT_TZ]) scratch_bytes += (double)nc * T * sizeof(double);
        if (pack == 2) // FP32 build below ground: FP64 series of whatever pass 1 writes, narrowed afterwards
            for (int v = 0; v < MCF_NOUT; ++v)
                if (out[v] && v != MCF_OUT_TZ && kernel_writes(rq, v)) scratch_bytes += (double)nc * T * sizeof(double);
        if (hp->clim_rows > 0 && hp->p_Tg && hp->p_Tbp) scratch_bytes += 2.0 * (double)nc * T * sizeof(double);
    }
    bool fits = (double)in_bytes + (double)per_hour * T + (double)nreq * 256 + scratch_bytes < 0.55 * (double)avail;
    long long chunk_blocks = 0;
    // test hook: MCF_FORCE_STREAM_BLOCKS=n forces the streaming path with n day-blocks per device chunk
    const char* force = std::getenv("MCF_FORCE_STREAM_BLOCKS");
    const long long forced = (force && rq != RQ_BELOW) ? std::atoll(force) : 0;
    if (forced > 0) fits = false;
    if (!fits) {
        if (rq == RQ_BELOW)
            return make_err(MCF_ERR_NOMEM, "reqhgt < 0 on a raster whose outputs exceed device memory: "
                                           "split the raster into column bands (has_twi_mean)");
        chunk_blocks = (long long)((0.45 * (double)avail - (double)in_bytes) / (2.0 * 24.0 * (double)per_hour));
        if (chunk_blocks < 1) return make_err(MCF_ERR_NOMEM, "one day of outputs does not fit device memory");
        if (forced > 0) chunk_blocks = forced;
        chunk_blocks = std::min<long long>(chunk_blocks, std::max(nblk, 1));
    }
    const long long chunk_hours = chunk_blocks * 24;
    const size_t out_elems = fits ? nc * (size_t)T : nc * (size_t)chunk_hours;
    const size_t out_bytes = (size_t)nreq * (fits ? 1 : 2) * DevCopy::pad(out_elems * esz);
    TRY(dc.reserve(in_bytes + out_bytes));
    dc.sizing = false;
    TRY(upload_problem(hp, &dp, dc)); // asynchronous H2D on the compute stream
    {
        Scratch sc(cs.s);
        Plan pl;
        pl.pack = pack;
        TRY(plan_prepare(pl, &dp, sc, cs.s));
        if (fits) {
            double* dout[MCF_NOUT] = {nullptr};
            for (int v = 0; v < MCF_NOUT; ++v)
                if (out[v]) TRY(dc.dalloc(&dout[v], out_elems, esz));
            TRY(prefill_whole(pl, dout, cs.s));
            // Time windows: the device->host copy of a window's hours overlaps the kernels of the next
            // window.  Window i owns the hour range from its first block to the next window's first block
            // (the first from hour 0, the last to T), so gaps the day-blocks do not cover travel with the NA
            // prefill.  reqhgt < 0 needs the whole series before its time-axis pass: one window.
            // The first window is a single day, so that the copy engine starts as early as possible; the rest of
            // the series is split into up to five windows.
            const int nrest = (rq == RQ_BELOW || nblk < 2) ? 0 : std::min(nblk - 1, 5);
            const int nwin = 1 + nrest;
            auto first_block = [&](int w) { // first day-block of window w; first_block(nwin) = nblk
                if (nrest == 0) return w == 0 ? 0 : nblk;
                return w == 0 ? 0 : 1 + (int)((long long)(nblk - 1) * (w - 1) / nrest);
            };
            std::vector<EventGuard> done_k(nwin);
            std::vector<CopyJob> jobs;
            for (int w = 0; w < nwin; ++w) {
                CU(cudaEventCreateWithFlags(&done_k[w].e, cudaEventDisableTiming));
                const int b0 = first_block(w), b1 = first_block(w + 1);
                if (rq == RQ_BELOW) TRY(plan_run_below(pl, dout, sc, cs.s));
                else TRY(plan_run_window(pl, dout, b0, b1 - b0, 0, T, sc, cs.s));
                CU(cudaEventRecord(done_k[w].e, cs.s));
                const size_t h0 = (w == 0) ? 0 : (size_t)pl.blocks[b0].k0;
                const size_t h1 = (w == nwin - 1) ? (size_t)T : (size_t)pl.blocks[b1].k0;
                for (int v = 0; v < MCF_NOUT; ++v)
                    if (out[v])
                        jobs.push_back(CopyJob{(char*)out[v] + h0 * nc * esz, (const char*)dout[v] + h0 * nc * esz,
                                               (h1 - h0) * nc * esz, done_k[w].e});
            }
            TRY(copy_back(jobs, xs.s));
            CU(cudaStreamSynchronize(cs.s));
        } else {
            // The outputs do not fit the device: stream the time axis through two device chunk buffers; the
            // copy-back of chunk i (blocking this thread) overlaps the kernels of chunk i+1 (already queued).
            double* dout[2][MCF_NOUT] = {{nullptr}, {nullptr}};
            for (int s = 0; s < 2; ++s)
                for (int v = 0; v < MCF_NOUT; ++v)
                    if (out[v]) TRY(dc.dalloc(&dout[s][v], out_elems, esz));
            EventGuard done_k[2];
            for (int s = 0; s < 2; ++s) CU(cudaEventCreateWithFlags(&done_k[s].e, cudaEventDisableTiming));
            // hours no day-block covers, and outputs this reqhgt never writes, are NA (host side)
            std::vector<char> covered(T, 0);
            for (const DayBlock& b : pl.blocks)
                for (int h = 0; h < 24; ++h) covered[b.k0 + h] = 1;
            for (int v = 0; v < MCF_NOUT; ++v) {
                if (!out[v]) continue;
                if (!kernel_writes(rq, v)) { host_fill_na(out[v], nc * T, pack); continue; }
                for (int k = 0; k < T; ++k)
                    if (!covered[k]) host_fill_na((double*)((char*)out[v] + (size_t)k * nc * esz), nc, pack);
            }
            std::vector<CopyJob> pending;
            int w = 0;
            for (int b0 = 0; b0 < nblk; ++w) {
                const int s = w & 1;
                int nb = (int)std::min<long long>(chunk_blocks, nblk - b0);
                // a chunk must cover a contiguous hour range to be copied back in one piece per output
                for (int i = 1; i < nb; ++i)
                    if (pl.blocks[b0 + i].k0 != pl.blocks[b0 + i - 1].k0 + 24) { nb = i; break; }
                const long long h0 = pl.blocks[b0].k0;
                TRY(plan_run_window(pl, dout[s], b0, nb, h0, chunk_hours, sc, cs.s)); // buffer s was drained below
                CU(cudaEventRecord(done_k[s].e, cs.s));
                if (!pending.empty()) TRY(copy_back(pending, xs.s)); // chunk w-1, overlapping the kernels of chunk w
                pending.clear();
                for (int v = 0; v < MCF_NOUT; ++v)
                    if (out[v] && kernel_writes(rq, v))
                        pending.push_back(CopyJob{(char*)out[v] + (size_t)h0 * nc * esz, (const char*)dout[s][v], nc * 24 * (size_t)nb * esz,
                                                  done_k[s].e});
                b0 += nb;
            }
            if (!pending.empty()) TRY(copy_back(pending, xs.s));
            CU(cudaStreamSynchronize(cs.s));
        }
    }
    CU(cudaStreamSynchronize(cs.s));
    return Err();
}

// ------------------------------------------------------------------------------------------------
// bioclim
// ------------------------------------------------------------------------------------------------
// runbioclim3/4Cpp hard-code 14 one-day vegetation layers (ref :3635-3646) and complete = true (:3576-3577)
struct BioPatch {
    mcf_problem pp;
    int32_t st14[14], ed14[14];
    Err apply(const mcf_problem* p) {
        if (!p) return make_err(MCF_ERR_ARG, "problem is NULL");
        pp = *p;
        pp.complete = 1;
        if (p->mode >= 3) {
            if (p->nlyr < 14) return make_err(MCF_ERR_ARG, "runbioclim3/4 need 14 vegetation layers, got %d", p->nlyr);
            for (int i = 0; i < 14; ++i) { st14[i] = i * 24; ed14[i] = i * 24 + 23; }
            pp.nlyr = 14;
            pp.lyr_st = st14;
            pp.lyr_ed = ed14;
        }
        return validate(&pp);
    }
};

// Launch the grid kernel with a reducing sink over day-blocks [b0, b0 + nb) of the whole raster (reqhgt >= 0).
Err plan_run_reduce(const Plan& pl, GridArgs& a, int sink, int b0, int nb, Scratch& sc, cudaStream_t st) {
    if (nb <= 0) return Err();
    a.cell_begin = 0;
    a.cell_end = pl.ncells;
    a.block0 = b0;
    a.nblocks = nb;
    a.hour0 = 0;
    a.ring_hours = std::max(pl.p->tsteps, 24);
    unsigned int* ctr = nullptr;
    CU(sc.alloc(&ctr, 1));
    CU(cudaMemsetAsync(ctr, 0, sizeof(unsigned int), st));
    a.tile_counter = ctr;
    return timed_grid_launch(pl, a, st, sink);
}

Err run_bioclim_dev(const mcf_problem* p, const int32_t* const q[4], const int32_t nq[4], int air,
                    double* const bio[MCF_NBIO], cudaStream_t st) {
    if (p->tsteps < 336) return make_err(MCF_ERR_ARG, "runbioclim needs the 14 selected days (336 hours), got %d", p->tsteps);
    for (int i = 0; i < 4; ++i) {
        if (nq[i] < 0 || (nq[i] > 0 && !q[i])) return make_err(MCF_ERR_ARG, "quarter index vector %d is invalid", i);
        for (int j = 0; j < nq[i]; ++j)
            if (q[i][j] < 0 || q[i][j] >= p->tsteps) return make_err(MCF_ERR_ARG, "quarter index out of range");
    }
    Scratch sc(st);
    const mcf_problem& pp = *p;
    Plan pl;
    TRY(plan_prepare(pl, &pp, sc, st));
    const int T = p->tsteps;
    const int nc = pl.ncells;
    uint32_t mask = 0;
    for (int v = 0; v < MCF_NBIO; ++v)
        if (bio[v]) mask |= 1u << v;
    if (!mask) return Err();
    // tleaf does not exist at or below the surface (ref :2300-2303): every cell keeps its NA fill (:3507-3508)
    if (!air && pl.rq != RQ_ABOVE) {
        for (int v = 0; v < MCF_NBIO; ++v)
            if (bio[v]) {
                CU(launch_fill_na(bio[v], nc, st));
                count_launch();
            }
        return Err();
    }
    if (pl.rq != RQ_BELOW) {
        // Fused: the 19 reductions are accumulated inside the grid kernel's day loop; no hourly series exists.
        std::vector<char> covered(T, 0);
        for (const DayBlock& b : pl.blocks)
            for (int h = 0; h < 24; ++h) covered[b.k0 + h] = 1;
        std::vector<uint32_t> qcnt(T, 0u);
        uint32_t q_na = 0;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < nq[i]; ++j) {
                const int k = q[i][j];
                if (!covered[k]) q_na |= 1u << i;
                if (((qcnt[k] >> (8 * i)) & 255u) == 255u)
                    return make_err(MCF_ERR_ARG, "quarter index vector %d repeats an hour more than 255 times", i);
                qcnt[k] += 1u << (8 * i);
            }
        int soil_gap = 0;
        for (int k = 0; k < T; ++k) soil_gap |= !covered[k];
        uint32_t* d_qcnt = nullptr;
        CU(sc.alloc(&d_qcnt, T));
        // (pageable source: the copy is staged by the driver before the call returns)
        CU(cudaMemcpyAsync(d_qcnt, qcnt.data(), (size_t)T * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        GridArgs a;
        fill_common(pl, a);
        for (int v = 0; v < MCF_NBIO; ++v) a.red[v] = bio[v];
        a.red_mask = mask;
        a.bio_air = air ? 1 : 0;
        a.bio_soil_gap = soil_gap;
        a.bio_q_na = q_na;
        a.bio_qcnt = d_qcnt;
        a.outmask = 0;
        TRY(plan_run_reduce(pl, a, SINK_BIO, 0, (int)pl.blocks.size(), sc, st));
        CU(cudaStreamSynchronize(st)); // qcnt lives on this frame
        return Err();
    }
    // Below ground the series only exists after the time-axis pass over the whole year: chunks of cells, each reduced
    // before the next one reuses the chunk's two series.
    int32_t* dq = nullptr;
    const int ntot = nq[0] + nq[1] + nq[2] + nq[3];
    CU(sc.alloc(&dq, ntot));
    BioArgs b;
    std::memset(&b, 0, sizeof b);
    int off = 0;
    for (int i = 0; i < 4; ++i) {
        if (nq[i]) CU(cudaMemcpyAsync(dq + off, q[i], nq[i] * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        b.q[i] = dq + off;
        b.nq[i] = nq[i];
        off += nq[i];
    }
    b.tsteps = T;
    b.mask = mask;
    for (int v = 0; v < MCF_NBIO; ++v) b.bio[v] = bio[v];
    double* out[MCF_NOUT] = {nullptr};
    TRY(plan_run_below(pl, out, sc, st, &b));
    // the quarter index vectors are read from pageable host memory by the async copies above
    CU(cudaStreamSynchronize(st));
    return Err();
}

// Summary sink: per-cell sum / min / max over the window's hours of each requested output (reqhgt >= 0).
Err run_summary_dev(const mcf_problem* p, double* const sum[MCF_NOUT], double* const mn[MCF_NOUT],
                    double* const mx[MCF_NOUT], const mcf_window* win, int accumulate, int64_t* hours_done,
                    cudaStream_t st) {
    Scratch sc(st);
    Plan pl;
    TRY(plan_prepare(pl, p, sc, st));
    if (pl.rq == RQ_BELOW)
        return make_err(MCF_ERR_ARG, "the summary sink covers reqhgt >= 0 (below ground the series needs its time-axis pass)");
    int b0 = 0, nb = 0;
    long long hour0 = 0, ring = p->tsteps;
    TRY(resolve_window(win, pl.blocks, p->tsteps, b0, nb, hour0, ring));
    GridArgs a;
    fill_common(pl, a);
    a.outmask = 0;
    for (int v = 0; v < MCF_NOUT; ++v) {
        const int n = (sum[v] != nullptr) + (mn[v] != nullptr) + (mx[v] != nullptr);
        if (n == 0) continue;
        if (n != 3) return make_err(MCF_ERR_ARG, "summary output %d needs its sum, min and max buffers together", v);
        if (!kernel_writes(pl.rq, v)) {
            // never produced at this height (ref :2300-2303): NA, as the reference's array would be
            if (!accumulate) {
                CU(launch_fill_na(sum[v], pl.ncells, st));
                CU(launch_fill_na(mn[v], pl.ncells, st));
                CU(launch_fill_na(mx[v], pl.ncells, st));
                count_launch(3);
            }
            continue;
        }
        a.outmask |= 1u << v;
        a.red[v] = sum[v];
        a.red[10 + v] = mn[v];
        a.red[20 + v] = mx[v];
    }
    a.red_accumulate = accumulate ? 1 : 0;
    if (hours_done) *hours_done = (int64_t)nb * 24;
    if (!a.outmask) return Err();
    return plan_run_reduce(pl, a, SINK_SUMMARY, b0, nb, sc, st);
}

} // namespace

// Internal services for the other translation units (mcf_snow.cu): device check + memory-pool retention, and the
// staged host<->device copy pool above.  Declared in mcf_kernels.cuh.
namespace mcf {
int host_prepare_device(char* err, size_t errlen) { return report(device_info(), err, errlen); }
int host_transfer(const HostXfer* jobs, int n, bool to_device, char* err, size_t errlen) {
    std::lock_guard<std::mutex> ws_lock(g_ws_mu); // the pinned slots are shared with the grid solver's copy-back
    std::vector<CopyJob> cj;
    for (int i = 0; i < n; ++i) cj.push_back(CopyJob{(char*)jobs[i].dst, (const char*)jobs[i].src, jobs[i].bytes, nullptr});
    return report(stage_copies(cj, nullptr, to_device), err, errlen);
}
} // namespace mcf

// ------------------------------------------------------------------------------------------------
// exported C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int mcf_abi_version(void) { return MCF_ABI_VERSION; }

int mcf_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

int mcf_set_device(int device) {
    // The grow-only workspace and the pinned staging slots belong to the device that was current when they were
    // allocated: give them back while that device is still current, so that the next host-buffer call carves its
    // inputs and outputs from memory of the device it launches on.
    int cur = -1;
    if (cudaGetDevice(&cur) == cudaSuccess && cur != device) mcf_release_workspace();
    (void)cudaGetLastError();
    g_sm_count = 0;
    return cudaSetDevice(device) == cudaSuccess ? MCF_OK : MCF_ERR_CUDA;
}

int64_t mcf_launch_count(void) { return g_launches.load(); }
void mcf_launch_count_reset(void) { g_launches.store(0); }

void mcf_kernel_timing_enable(int on) { g_timing = on != 0; }
void mcf_kernel_time_reset(void) {
    std::lock_guard<std::mutex> lk(g_time_mu);
    for (auto& pr : g_events) {
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    g_events.clear();
    g_time_ms = 0.0;
    g_time_n = 0;
}
int mcf_kernel_time(double* total_ms, int64_t* launches) {
    std::lock_guard<std::mutex> lk(g_time_mu);
    for (auto& pr : g_events) {
        if (cudaEventSynchronize(pr.second) != cudaSuccess) return MCF_ERR_CUDA;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, pr.first, pr.second) != cudaSuccess) return MCF_ERR_CUDA;
        g_time_ms += ms;
        g_time_n += 1;
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    g_events.clear();
    if (total_ms) *total_ms = g_time_ms;
    if (launches) *launches = g_time_n;
    return MCF_OK;
}

int mcf_runmicro_dev(const mcf_problem* prob, double* const out[MCF_NOUT], const mcf_window* win, void* stream,
                     char* err, size_t errlen) {
    if (!out) return report(make_err(MCF_ERR_ARG, "out is NULL"), err, errlen);
    return report(run_dev(prob, out, win, (cudaStream_t)stream), err, errlen);
}

int mcf_runmicro(const mcf_problem* prob, double* const out[MCF_NOUT], char* err, size_t errlen) {
    if (!out) return report(make_err(MCF_ERR_ARG, "out is NULL"), err, errlen);
    return report(run_host(prob, out), err, errlen);
}

// FP32 build (north_star: optional, within 0.05 degC / 0.5 % radiation): every mode and height, device or host buffers.
// Inputs are the FP64 problem; the hour loops, the day stash and the outputs are FP32 (Plan::pack == 2).
int mcf_runmicro_f32_dev(const mcf_problem* prob, float* const out[MCF_NOUT], const mcf_window* win, void* stream, char* err,
                         size_t errlen) {
    if (!out) return report(make_err(MCF_ERR_ARG, "out is NULL"), err, errlen);
    return report(run_dev(prob, reinterpret_cast<double* const*>(out), win, (cudaStream_t)stream, 2), err, errlen);
}

int mcf_runmicro_f32(const mcf_problem* prob, float* const out[MCF_NOUT], char* err, size_t errlen) {
    if (!out) return report(make_err(MCF_ERR_ARG, "out is NULL"), err, errlen);
    return report(run_host(prob, reinterpret_cast<double* const*>(out), 2), err, errlen);
}

// Packed integer sink (SURVEY.md NEXT-4): the same solve, results stored as writetonc stores them
// (R/dataprep.R:1064-1069, 1164-1173).  The int16_t* buffers travel through the double* plumbing.
int mcf_runmicro_packed_dev(const mcf_problem* prob, int16_t* const out[MCF_NOUT], const mcf_window* win, void* stream,
                            char* err, size_t errlen) {
    if (!out) return report(make_err(MCF_ERR_ARG, "out is NULL"), err, errlen);
    return report(run_dev(prob, reinterpret_cast<double* const*>(out), win, (cudaStream_t)stream, 1), err, errlen);
}

int mcf_runmicro_packed(const mcf_problem* prob, int16_t* const out[MCF_NOUT], char* err, size_t errlen) {
    if (!out) return report(make_err(MCF_ERR_ARG, "out is NULL"), err, errlen);
    return report(run_host(prob, reinterpret_cast<double* const*>(out), 1), err, errlen);
}

int mcf_runbioclim_dev(const mcf_problem* prob, const int32_t* wetq, int32_t nwetq, const int32_t* dryq, int32_t ndryq,
                       const int32_t* hotq, int32_t nhotq, const int32_t* colq, int32_t ncolq, int32_t air,
                       double* const bio[MCF_NBIO], void* stream, char* err, size_t errlen) {
    BioPatch bp;
    Err e = bp.apply(prob);
    if (e.code == MCF_OK) e = device_info();
    if (e.code == MCF_OK) {
        const int32_t* q[4] = {wetq, dryq, hotq, colq};
        const int32_t nq[4] = {nwetq, ndryq, nhotq, ncolq};
        e = run_bioclim_dev(&bp.pp, q, nq, air, bio, (cudaStream_t)stream);
    }
    return report(e, err, errlen);
}

int mcf_runbioclim(const mcf_problem* prob, const int32_t* wetq, int32_t nwetq, const int32_t* dryq, int32_t ndryq,
                   const int32_t* hotq, int32_t nhotq, const int32_t* colq, int32_t ncolq, int32_t air,
                   double* const bio[MCF_NBIO], char* err, size_t errlen) {
    auto body = [&]() -> Err {
        BioPatch bp;
        TRY(bp.apply(prob));
        TRY(device_info());
        std::lock_guard<std::mutex> ws_lock(g_ws_mu);
        StreamGuard sg;
        CU(cudaStreamCreateWithFlags(&sg.s, cudaStreamNonBlocking));
        const size_t nc = (size_t)prob->rows * prob->cols;
        DevCopy dc;
        dc.stream = sg.s;
        mcf_problem dp;
        double* dbio[MCF_NBIO] = {nullptr};
        for (int pass = 0; pass < 2; ++pass) { // sizing pass, then the real one
            dc.sizing = (pass == 0);
            TRY(upload_problem(&bp.pp, &dp, dc));
            for (int v = 0; v < MCF_NBIO; ++v)
                if (bio[v]) TRY(dc.dalloc(&dbio[v], nc));
            if (pass == 0) TRY(dc.reserve(dc.need));
        }
        const int32_t* q[4] = {wetq, dryq, hotq, colq};
        const int32_t nq[4] = {nwetq, ndryq, nhotq, ncolq};
        TRY(run_bioclim_dev(&dp, q, nq, air, dbio, sg.s));
        for (int v = 0; v < MCF_NBIO; ++v)
            if (bio[v]) CU(cudaMemcpyAsync(bio[v], dbio[v], nc * sizeof(double), cudaMemcpyDeviceToHost, sg.s));
        CU(cudaStreamSynchronize(sg.s));
        return Err();
    };
    return report(body(), err, errlen);
}

int mcf_runmicro_summary_dev(const mcf_problem* prob, double* const sum[MCF_NOUT], double* const mn[MCF_NOUT],
                             double* const mx[MCF_NOUT], const mcf_window* win, int32_t accumulate, int64_t* hours_done,
                             void* stream, char* err, size_t errlen) {
    if (!sum || !mn || !mx) return report(make_err(MCF_ERR_ARG, "sum / min / max is NULL"), err, errlen);
    return report(run_summary_dev(prob, sum, mn, mx, win, accumulate, hours_done, (cudaStream_t)stream), err, errlen);
}

int mcf_runmicro_summary(const mcf_problem* prob, double* const mean[MCF_NOUT], double* const mn[MCF_NOUT],
                         double* const mx[MCF_NOUT], int64_t* hours_done, char* err, size_t errlen) {
    auto body = [&]() -> Err {
        if (!mean || !mn || !mx) return make_err(MCF_ERR_ARG, "mean / min / max is NULL");
        TRY(validate(prob));
        TRY(device_info());
        std::lock_guard<std::mutex> ws_lock(g_ws_mu);
        StreamGuard sg;
        CU(cudaStreamCreateWithFlags(&sg.s, cudaStreamNonBlocking));
        const size_t nc = (size_t)prob->rows * prob->cols;
        DevCopy dc;
        dc.stream = sg.s;
        mcf_problem dp;
        double* d[3][MCF_NOUT] = {{nullptr}};
        double* const* h[3] = {mean, mn, mx};
        for (int pass = 0; pass < 2; ++pass) { // sizing pass, then the real one
            dc.sizing = (pass == 0);
            TRY(upload_problem(prob, &dp, dc));
            for (int v = 0; v < MCF_NOUT; ++v)
                for (int k = 0; k < 3; ++k)
                    if (h[k][v]) TRY(dc.dalloc(&d[k][v], nc));
            if (pass == 0) TRY(dc.reserve(dc.need));
        }
        int64_t hours = 0;
        TRY(run_summary_dev(&dp, d[0], d[1], d[2], nullptr, 0, &hours, sg.s));
        std::vector<CopyJob> jobs;
        for (int v = 0; v < MCF_NOUT; ++v)
            for (int k = 0; k < 3; ++k)
                if (h[k][v]) jobs.push_back(CopyJob{(char*)h[k][v], (const char*)d[k][v], nc * sizeof(double), nullptr});
        CU(cudaStreamSynchronize(sg.s));
        TRY(copy_back(jobs, sg.s));
        if (hours_done) *hours_done = hours;
        if (hours > 0)
            for (int v = 0; v < MCF_NOUT; ++v)
                if (mean[v])
                    for (size_t i = 0; i < nc; ++i)
                        if (!std::isnan(mean[v][i])) mean[v][i] /= (double)hours; // NA / NaN cells keep their bits
        return Err();
    };
    return report(body(), err, errlen);
}

int mcf_twi_partial(const double* twi, int64_t n, double tfact, double* sum, int64_t* count, char* err,
                    size_t errlen) {
    auto body = [&]() -> Err {
        if (!twi || !sum || !count || n < 0) return make_err(MCF_ERR_ARG, "bad argument");
        TRY(device_info());
        double* d = nullptr;
        double* sc = nullptr;
        CU(cudaMalloc((void**)&d, std::max<int64_t>(n, 1) * sizeof(double)));
        cudaError_t e = cudaMalloc((void**)&sc, (2 + 2 * 256) * sizeof(double));
        if (e != cudaSuccess) { cudaFree(d); return make_err(MCF_ERR_CUDA, "cudaMalloc failed"); }
        double res[2] = {0, 0};
        e = cudaMemcpy(d, twi, n * sizeof(double), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = launch_twi_sum(d, n, tfact, sc, nullptr);
        count_launch(2);
        if (e == cudaSuccess) e = cudaMemcpy(res, sc, sizeof res, cudaMemcpyDeviceToHost);
        cudaFree(d);
        cudaFree(sc);
        if (e != cudaSuccess) return make_err(MCF_ERR_CUDA, "twi reduction failed: %s", cudaGetErrorString(e));
        *sum = res[0];
        *count = (int64_t)res[1];
        return Err();
    };
    return report(body(), err, errlen);
}

// terrain preparation on host buffers (SURVEY.md NEXT-2).  windcoef: thr_hgt = wind height (m), else horizon.
static Err terrain_stencil(const double* dtm, int32_t rows, int32_t cols, double reso, int32_t ndir,
                           const double* dir_deg, bool windcoef, double hgt, double* out, double* svfa, double* blend8) {
    if (!dtm || !dir_deg || rows <= 0 || cols <= 0 || ndir <= 0 || !(reso > 0)) return make_err(MCF_ERR_ARG, "bad argument");
    if (blend8 && ndir != 16) return make_err(MCF_ERR_ARG, "the 16 -> 8 blend needs 16 directions");
    TRY(device_info());
    std::lock_guard<std::mutex> ws_lock(g_ws_mu);
    const size_t nc = (size_t)rows * cols;
    std::vector<double> offs((size_t)ndir * 20);
    for (int a = 0; a < ndir; ++a) {
        const double azi = dir_deg[a] * (3.14159265358979323846 / 180);
        for (int s = 1; s <= 10; ++s) { // R/internal.R:921-922: 101 -/+ cos/sin(azi) * step^2
            offs[((size_t)a * 10 + s - 1) * 2 + 0] = 101 - std::cos(azi) * (double)(s * s);
            offs[((size_t)a * 10 + s - 1) * 2 + 1] = 101 + std::sin(azi) * (double)(s * s);
        }
    }
    DevCopy dc;
    const double *d_dtm = nullptr, *d_offs = nullptr;
    double *d_scaled = nullptr, *d_out = nullptr, *d_svf = nullptr, *d_b8 = nullptr;
    for (int pass = 0; pass < 2; ++pass) {
        dc.sizing = (pass == 0);
        TRY(dc.up(dtm, nc, &d_dtm));
        TRY(dc.up(offs.data(), offs.size(), &d_offs));
        TRY(dc.dalloc(&d_scaled, nc));
        TRY(dc.dalloc(&d_out, nc * ndir));
        if (svfa) TRY(dc.dalloc(&d_svf, nc));
        if (blend8) TRY(dc.dalloc(&d_b8, nc * 8));
        if (pass == 0) TRY(dc.reserve(dc.need));
    }
    CU(launch_scale_dtm(d_dtm, (int64_t)nc, reso, d_scaled, nullptr));
    CU(launch_horizon(d_scaled, rows, cols, ndir, d_offs, hgt / reso, windcoef, d_out, nullptr));
    count_launch(2);
    if (svfa) {
        CU(launch_skyview(d_out, (int64_t)nc, ndir, d_svf, nullptr));
        count_launch();
    }
    if (blend8) {
        CU(launch_blend16to8(d_out, (int64_t)nc, d_b8, nullptr));
        count_launch();
    }
    if (out) CU(cudaMemcpyAsync(out, d_out, nc * ndir * sizeof(double), cudaMemcpyDeviceToHost, nullptr));
    if (svfa) CU(cudaMemcpyAsync(svfa, d_svf, nc * sizeof(double), cudaMemcpyDeviceToHost, nullptr));
    if (blend8) CU(cudaMemcpyAsync(blend8, d_b8, nc * 8 * sizeof(double), cudaMemcpyDeviceToHost, nullptr));
    CU(cudaStreamSynchronize(nullptr));
    return Err();
}

int mcf_horizon(const double* dtm, int32_t rows, int32_t cols, double reso, int32_t nazi, const double* azimuth_deg,
                double* hor, double* svfa, char* err, size_t errlen) {
    return report(terrain_stencil(dtm, rows, cols, reso, nazi, azimuth_deg, false, 0.0, hor, svfa, nullptr), err, errlen);
}

int mcf_windcoef(const double* dsm, int32_t rows, int32_t cols, double reso, double hgt, int32_t ndir,
                 const double* direction_deg, double* index, double* blend8, char* err, size_t errlen) {
    return report(terrain_stencil(dsm, rows, cols, reso, ndir, direction_deg, true, hgt, index, nullptr, blend8), err, errlen);
}

// .windsheltera (R/internal.R:970-991) end to end on the device: .windcoef in 16 directions, each smoothed by
// terra::aggregate(fact = s, mean) + terra::resample(bilinear) (:979-981; s <= 1 skips the smoothing), blended to the 8
// sectors the solver indexes (:983-989).  Only the [rows, cols, 8] result crosses PCIe.
int mcf_windshelter(const double* dsm, int32_t rows, int32_t cols, double reso, double hgt, int32_t s, double* wsa8,
                    char* err, size_t errlen) {
    auto body = [&]() -> Err {
        if (!dsm || !wsa8 || rows <= 0 || cols <= 0 || !(reso > 0)) return make_err(MCF_ERR_ARG, "bad argument");
        TRY(device_info());
        std::lock_guard<std::mutex> ws_lock(g_ws_mu);
        const size_t nc = (size_t)rows * cols;
        std::vector<double> offs((size_t)16 * 20);
        for (int a = 0; a < 16; ++a) {
            const double azi = (a * 22.5) * (3.14159265358979323846 / 180);
            for (int st = 1; st <= 10; ++st) {
                offs[((size_t)a * 10 + st - 1) * 2 + 0] = 101 - std::cos(azi) * (double)(st * st);
                offs[((size_t)a * 10 + st - 1) * 2 + 1] = 101 + std::sin(azi) * (double)(st * st);
            }
        }
        const int fact = s > 1 ? s : 1;
        const size_t onc = (size_t)((rows + fact - 1) / fact) * ((cols + fact - 1) / fact);
        DevCopy dc;
        const double *d_dsm = nullptr, *d_offs = nullptr;
        double *d_scaled = nullptr, *d_idx = nullptr, *d_coarse = nullptr, *d_smooth = nullptr, *d_b8 = nullptr;
        for (int pass = 0; pass < 2; ++pass) {
            dc.sizing = (pass == 0);
            TRY(dc.up(dsm, nc, &d_dsm));
            TRY(dc.up(offs.data(), offs.size(), &d_offs));
            TRY(dc.dalloc(&d_scaled, nc));
            TRY(dc.dalloc(&d_idx, nc * 16));
            if (fact > 1) {
                TRY(dc.dalloc(&d_coarse, onc * 16));
                TRY(dc.dalloc(&d_smooth, nc * 16));
            }
            TRY(dc.dalloc(&d_b8, nc * 8));
            if (pass == 0) TRY(dc.reserve(dc.need));
        }
        CU(launch_scale_dtm(d_dsm, (int64_t)nc, reso, d_scaled, nullptr));
        CU(launch_horizon(d_scaled, rows, cols, 16, d_offs, hgt / reso, true, d_idx, nullptr));
        count_launch(2);
        const double* blend_in = d_idx;
        if (fact > 1) {
            CU(launch_smooth(d_idx, rows, cols, 16, fact, d_coarse, d_smooth, nullptr));
            count_launch(2);
            blend_in = d_smooth;
        }
        CU(launch_blend16to8(blend_in, (int64_t)nc, d_b8, nullptr));
        count_launch();
        CU(cudaMemcpyAsync(wsa8, d_b8, nc * 8 * sizeof(double), cudaMemcpyDeviceToHost, nullptr));
        CU(cudaStreamSynchronize(nullptr));
        return Err();
    };
    return report(body(), err, errlen);
}

// terra::terrain(v = "slope") and (v = "aspect") of the DTM (R/internal.R:1124-1129; R/Cppwrappers.R:483-484), degrees,
// NaN on the raster's edge and beside missing cells as terra leaves NA there.  Either output may be NULL.
int mcf_slope_aspect(const double* dtm, int32_t rows, int32_t cols, double xres, double yres, double* slope, double* aspect,
                     char* err, size_t errlen) {
    auto body = [&]() -> Err {
        if (!dtm || rows <= 0 || cols <= 0 || !(xres > 0) || !(yres > 0)) return make_err(MCF_ERR_ARG, "bad argument");
        TRY(device_info());
        std::lock_guard<std::mutex> ws_lock(g_ws_mu);
        const size_t nc = (size_t)rows * cols;
        DevCopy dc;
        const double* d_dtm = nullptr;
        double *d_sl = nullptr, *d_as = nullptr;
        for (int pass = 0; pass < 2; ++pass) {
            dc.sizing = (pass == 0);
            TRY(dc.up(dtm, nc, &d_dtm));
            if (slope) TRY(dc.dalloc(&d_sl, nc));
            if (aspect) TRY(dc.dalloc(&d_as, nc));
            if (pass == 0) TRY(dc.reserve(dc.need));
        }
        CU(launch_horn(d_dtm, rows, cols, xres, yres, d_sl, d_as, nullptr));
        count_launch();
        if (slope) CU(cudaMemcpyAsync(slope, d_sl, nc * sizeof(double), cudaMemcpyDeviceToHost, nullptr));
        if (aspect) CU(cudaMemcpyAsync(aspect, d_as, nc * sizeof(double), cudaMemcpyDeviceToHost, nullptr));
        CU(cudaStreamSynchronize(nullptr));
        return Err();
    };
    return report(body(), err, errlen);
}

// .topidx (R/internal.R:861-874): topographic wetness index a / tan(B), B = Horn slope (device) floored at
// atan(0.02 / mean(res)) with missing slopes replaced by the median, a = (flow accumulation + 1) x cell area floored at 1
// (flowaccCpp: the sequential sweep of mcf_flowacc, host), masked by the DTM.
int mcf_topidx(const double* dtm, int32_t rows, int32_t cols, double xres, double yres, double* twi, char* err, size_t errlen) {
    auto body = [&]() -> Err {
        if (!dtm || !twi || rows <= 0 || cols <= 0 || !(xres > 0) || !(yres > 0)) return make_err(MCF_ERR_ARG, "bad argument");
        const size_t nc = (size_t)rows * cols;
        std::vector<double> B(nc), fa(nc);
        char e2[256];
        if (mcf_slope_aspect(dtm, rows, cols, xres, yres, B.data(), nullptr, e2, sizeof e2) != MCF_OK)
            return make_err(MCF_ERR_CUDA, "%s", e2);
        if (mcf_flowacc(dtm, rows, cols, fa.data(), e2, sizeof e2) != MCF_OK) return make_err(MCF_ERR_ARG, "%s", e2);
        const double minslope = std::atan(0.02 / (0.5 * (xres + yres)));
        const double torad = 3.14159265358979323846 / 180.0;
        std::vector<double> finite;
        finite.reserve(nc);
        for (size_t i = 0; i < nc; ++i) {
            double b = B[i] * torad; // terrain(unit = "radians")
            if (b < minslope) b = minslope;
            B[i] = b;
            if (!std::isnan(b)) finite.push_back(b);
        }
        double med = std::nan("");
        if (!finite.empty()) { // R's median: mean of the two middle values for an even count
            const size_t m = finite.size() / 2;
            std::nth_element(finite.begin(), finite.begin() + m, finite.end());
            med = finite[m];
            if (finite.size() % 2 == 0) {
                const double lo = *std::max_element(finite.begin(), finite.begin() + m);
                med = 0.5 * (lo + med);
            }
        }
        for (size_t i = 0; i < nc; ++i) {
            const double b = std::isnan(B[i]) ? med : B[i];
            double a = (fa[i] + 1) * xres * yres;
            if (a < 1) a = 1;
            twi[i] = std::isnan(dtm[i]) ? std::nan("") : a / std::tan(b);
        }
        return Err();
    };
    return report(body(), err, errlen);
}

// flowdirCpp + flowaccCpp (src/microclimfCpp.cpp:5326-5414): D8 flow direction to the lowest of the 3x3
// neighbourhood (first minimum in column-major scan order, the cell itself included) and accumulation by one
// sweep over the cells in decreasing (elevation, row * cols + col) order.  A sequential sorted sweep over the
// whole raster: HOST code, run once per model run before the solver ("TWI stays on one
// rank/host").  NA cells receive (double)INT_MIN as in the reference (fa = NA_INTEGER stored in a double).
// "NA" is R_IsNA, i.e. ONLY R's NA_real_ payload (low word 1954): the plain quiet NaN terra hands back for
// missing cells is NOT NA to the reference — such cells take part in every comparison (all false), drain
// into their lowest neighbour and are sorted by index among "equal" elevations.  Reproduced as written
// (same std::sort, same comparator), because the accumulation next to a masked region depends on it.
int mcf_flowacc(const double* dtm, int32_t rows, int32_t cols, double* fa, char* err, size_t errlen) {
    auto body = [&]() -> Err {
        if (!dtm || !fa || rows < 1 || cols < 1) return make_err(MCF_ERR_ARG, "bad argument");
        const int64_t n = (int64_t)rows * cols;
        const double na_int = (double)INT32_MIN;
        std::vector<int8_t> fd((size_t)n, 0);
        std::vector<std::pair<double, int32_t>> order;
        order.reserve((size_t)n);
        auto at = [&](int i, int j) { return dtm[(int64_t)j * rows + i]; };
        auto is_na = [](double v) {
            uint64_t b;
            std::memcpy(&b, &v, sizeof b);
            return std::isnan(v) && (uint32_t)(b & 0xFFFFFFFFu) == 1954u;
        };
        for (int i = 0; i < rows; ++i) {
            for (int j = 0; j < cols; ++j) {
                const double v = at(i, j);
                if (is_na(v)) {
                    fa[(int64_t)j * rows + i] = na_int;
                    continue;
                }
                fa[(int64_t)j * rows + i] = 1.0;
                order.push_back({v, (int32_t)(i * cols + j)});
                double minval = 9999.99;
                int indx = 1, best = 0;
                for (int jj = -1; jj <= 1; ++jj) {
                    for (int ii = -1; ii <= 1; ++ii, ++indx) {
                        const int y = i + ii, x = j + jj;
                        if (y < 0 || y >= rows || x < 0 || x >= cols) continue; // the NA border
                        const double v2 = at(y, x);
                        if (!is_na(v2) && v2 < minval) {
                            minval = v2;
                            best = indx;
                        }
                    }
                }
                fd[(size_t)((int64_t)j * rows + i)] = (int8_t)best;
            }
        }
        if (order.empty()) return Err();
        std::sort(order.begin(), order.end(), std::greater<std::pair<double, int32_t>>());
        for (size_t k = 0; k + 1 < order.size(); ++k) { // the lowest cell is not propagated (ref :5392)
            const int idx = order[k].second;
            const int y = idx / cols, x = idx % cols;
            const int f = fd[(size_t)((int64_t)x * rows + y)];
            if (f < 1 || f > 9) continue;
            const int y2 = y + (f - 1) % 3 - 1, x2 = x + (f - 1) / 3 - 1;
            if (x2 >= 0 && x2 < cols && y2 >= 0 && y2 < rows) {
                double& t = fa[(int64_t)x2 * rows + y2];
                if (t != na_int) t += fa[(int64_t)x * rows + y];
            }
        }
        return Err();
    };
    return report(body(), err, errlen);
}

void mcf_release_workspace(void) {
    std::lock_guard<std::mutex> ws_lock(g_ws_mu);
    if (g_ws.base) cudaFree(g_ws.base);
    g_ws.base = nullptr;
    g_ws.cap = 0;
    if (g_stage) cudaFreeHost(g_stage);
    g_stage = nullptr;
    // scratch blocks retained by the stream-ordered pool (see device_info) go back to the driver too
    int dev = 0;
    cudaMemPool_t pool = nullptr;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        (void)cudaDeviceSynchronize();
        (void)cudaMemPoolTrimTo(pool, 0);
    }
    (void)cudaCtxResetPersistingL2Cache(); // stash lines still marked persisting become ordinary lines
    (void)cudaGetLastError();
}

int mcf_math_eval(int fn, const double* x, const double* y, int64_t n, double* out, char* err, size_t errlen) {
    auto body = [&]() -> Err {
        if (!x || !out || n < 0 || fn < 0 || fn > 8) return make_err(MCF_ERR_ARG, "bad argument");
        if ((fn == 1 || fn == 6) && !y) return make_err(MCF_ERR_ARG, "fn %d needs a second operand", fn);
        TRY(device_info());
        std::lock_guard<std::mutex> ws_lock(g_ws_mu);
        DevCopy dc;
        const double *dx = nullptr, *dy = nullptr;
        double* dout = nullptr;
        for (int pass = 0; pass < 2; ++pass) {
            dc.sizing = (pass == 0);
            TRY(dc.up(x, (size_t)n, &dx));
            TRY(dc.up(y, (size_t)n, &dy));
            TRY(dc.dalloc(&dout, (size_t)n));
            if (pass == 0) TRY(dc.reserve(dc.need));
        }
        CU(launch_math_eval(fn, dx, dy, n, dout, nullptr));
        count_launch();
        CU(cudaMemcpyAsync(out, dout, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, nullptr));
        CU(cudaStreamSynchronize(nullptr));
        return Err();
    };
    return report(body(), err, errlen);
}

int mcf_fp64_peak(double* tflops, char* err, size_t errlen) {
    auto body = [&]() -> Err {
        if (!tflops) return make_err(MCF_ERR_ARG, "tflops is NULL");
        TRY(device_info());
        double* sink = nullptr;
        CU(cudaMalloc((void**)&sink, sizeof(double)));
        cudaEvent_t e0, e1;
        CU(cudaEventCreate(&e0));
        CU(cudaEventCreate(&e1));
        const int grid = g_sm_count * 8, iters = 4096;
        launch_fp64_peak(sink, grid, 256, nullptr); // warm-up
        double best = 0.0;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0, nullptr);
            launch_fp64_peak(sink, grid, iters, nullptr);
            cudaEventRecord(e1, nullptr);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            const double flops = (double)grid * 256.0 * 64.0 * iters * 2.0;
            best = std::max(best, flops / (ms * 1e-3) / 1e12);
        }
        count_launch(6);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        cudaFree(sink);
        CU(cudaGetLastError());
        *tflops = best;
        return Err();
    };
    return report(body(), err, errlen);
}

} // extern "C"
