/*
 * microclimf_b200 — C ABI of the B200-native grid solver (the drop-in boundary).
 *
 * This header is what the reference's Rcpp glue would bind for its hot path.  Each entry point
 * replaces one family of `.Call` symbols of the reference (ilyamaclean/microclimf v2.0.0):
 *
 *   mcf_runmicro    <- _microclimf_runmicro{1,2,3,4}Cpp    src/RcppExports.cpp:248-349,
 *                      R/RcppExports.R:72-86, drivers src/microclimfCpp.cpp:2052, 2340, 2624, 2926
 *   mcf_runbioclim  <- _microclimf_runbioclim{1,2,3,4}Cpp  src/RcppExports.cpp:350-465,
 *                      R/RcppExports.R:88-102, drivers src/microclimfCpp.cpp:3457-3700
 *
 * Plain pointers and sizes only.  All arrays are R layout (column-major, FP64):
 *   matrix [rows, cols]        : idx = i + rows*j
 *   array  [rows, cols, n]     : idx = i + rows*j + rows*cols*k     (cells contiguous per slice)
 * so for a fixed hour k the cells are contiguous ("cell-major"), which is exactly the layout the
 * kernels read and write coalesced.
 *
 * There is NO CPU fallback: every compute entry point returns MCF_ERR_CUDA when no sm_100 device
 * is usable.  INTEGRATION.md shows the Rcpp-side stub that forwards the SEXPs to these calls.
 */
#ifndef MICROCLIMF_B200_H
#define MICROCLIMF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCF_ABI_VERSION 2

/* status codes (the Rcpp stub maps non-zero to Rcpp::stop(err), cf. BEGIN_RCPP/END_RCPP
 * src/RcppExports.cpp:251-271) */
#define MCF_OK 0
#define MCF_ERR_ARG 1    /* bad argument (NULL where data is required, bad mode, bad dfsel span ...) */
#define MCF_ERR_CUDA 2   /* CUDA runtime failure / no usable device */
#define MCF_ERR_NOMEM 3  /* the problem does not fit the device even after chunking */

/* the 10 outputs of runmicroNCpp, in the order of its `out` logical vector
 * (src/microclimfCpp.cpp:2131-2140, 2325-2336) */
enum {
    MCF_OUT_TZ = 0,
    MCF_OUT_TLEAF = 1,
    MCF_OUT_RELHUM = 2,
    MCF_OUT_SOILM = 3,
    MCF_OUT_WINDSPEED = 4,
    MCF_OUT_RDIRDOWN = 5,
    MCF_OUT_RDIFDOWN = 6,
    MCF_OUT_RLWDOWN = 7,
    MCF_OUT_RSWUP = 8,
    MCF_OUT_RLWUP = 9,
    MCF_NOUT = 10
};
#define MCF_NBIO 19 /* bio1..bio19, src/microclimfCpp.cpp:3539-3559 */

/* R's NA_real_ bit pattern: outputs of skipped cells / never-computed hours carry it, as the
 * reference's NumericVector(n, NA_REAL) prefill does (src/microclimfCpp.cpp:2131). */
#define MCF_NA_REAL_BITS 0x7FF00000000007A2ULL

/*
 * One grid-model problem = the argument list of runmicroNCpp, flattened.
 *
 * mode 1: static vegetation, data.frame climate   (runmicro1Cpp, src/microclimfCpp.cpp:2052)
 * mode 2: static vegetation, array climate        (runmicro2Cpp, :2340)
 * mode 3: layered vegetation, data.frame climate  (runmicro3Cpp, :2624)
 * mode 4: layered vegetation, array climate       (runmicro4Cpp, :2926)
 *
 * Climate / point-model series have length `tsteps` in modes 1/3 and rows*cols*tsteps in modes 2/4
 * (`winddir` is always a length-tsteps vector, src/microclimfCpp.cpp:2359).  Vegetation fields have
 * rows*cols*nlyr elements (nlyr = 1 in modes 1/2).  Soil fields are [rows, cols]; `wsa` is
 * [rows, cols, 8] and `hor` is [rows, cols, 24].
 */
typedef struct mcf_problem {
    int32_t mode;     /* 1..4 */
    int32_t rows;     /* fast spatial axis */
    int32_t cols;     /* slow spatial axis: bands for multi-GPU sharding are column ranges */
    int32_t tsteps;   /* hours; only floor(tsteps/24) whole days are computed in modes 1/2 */
    int32_t nlyr;     /* vegetation layers (modes 3/4), else 1 */
    int32_t complete; /* `complete` flag: all hours of the year present (below-ground branch) */

    double reqhgt, zref;
    double lat, lon;      /* modes 1/3 */
    double Sminp, Smaxp;  /* accepted and ignored, as in soildCppm (src/microclimfCpp.cpp:975) */
    double tfact, mat;

    /* dfsel (modes 3/4): 0-based inclusive hour spans per layer, src/microclimfCpp.cpp:2629-2639 */
    const int32_t* lyr_st;
    const int32_t* lyr_ed;

    /* obstime */
    const int32_t* year;
    const int32_t* month;
    const int32_t* day;
    const double* hour;

    /* climdata: temp|tc, es, ea, tdew, pres|pk, swdown, difrad, lwdown, windspeed, winddir */
    const double* temp;
    const double* es;
    const double* ea;
    const double* tdew;
    const double* pres;
    const double* swdown;
    const double* difrad;
    const double* lwdown;
    const double* windspeed;
    const double* winddir;

    /* pointm: soilm, Tg, Tbp, G|Gp, umu, kp, muGp, dtrp  (T0p, DDp, Tc are never read by the grid
     * drivers; Tg/Tbp only when reqhgt < 0 and may be NULL otherwise) */
    const double* p_soilm;
    const double* p_Tg;
    const double* p_Tbp;
    const double* p_G;
    const double* p_umu;
    const double* p_kp;
    const double* p_muGp;
    const double* p_dtrp;

    /* vegp */
    const double* hgt;
    const double* pai;
    const double* x;
    const double* gsmax;
    const double* leafr;
    const double* leaft;
    const double* clump;
    const double* leafd;
    const double* paia;
    const double* leafden;

    /* soilc */
    const double* Smin;
    const double* Smax;
    const double* gref;
    const double* soilb;
    const double* Psie;
    const double* Vq;
    const double* Vm;
    const double* Mc;
    const double* rho;
    const double* slope;
    const double* aspect;
    const double* twi;
    const double* svfa;
    const double* wsa; /* [rows, cols, 8]  */
    const double* hor; /* [rows, cols, 24] */

    /* modes 2/4: per-cell latitude / longitude matrices */
    const double* lats;
    const double* lons;

    /* Band sharding (multi-GPU): when this process holds only a column band of the raster, the
     * whole-raster mean of log(twi)/tfact that soildCppm subtracts (src/microclimfCpp.cpp:993-1004)
     * must be supplied by the caller (all-reduce of mcf_twi_partial over the bands). */
    int32_t has_twi_mean;
    double twi_mean;

    /* Coarse-grid climate (modes 2/4, clim_rows > 0).  In the reference `.runmodel2Cpp` / `.runmodel4Cpp`
     * expand every climate and point-model variable of runpointmodela's coarse grid to the fine raster on the
     * host (`.cca` -> terra::resample, R/internal.R:523-542, 1219-1277) and hand [rows, cols, tsteps] arrays to
     * runmicro2Cpp: 15 arrays x 8 B per cell-hour that cannot exist for a large raster.  With
     * clim_rows > 0 the series below are [clim_rows, clim_cols, tsteps] arrays on the COARSE grid and the
     * kernels interpolate them bilinearly per cell-hour, deriving es / ea / tdew (R/internal.R:1222-1224), the
     * altitude correction (:1226-1245) and the wind speed from its interpolated components (:1250-1259) on the
     * fly.  Fields read in this layout: temp, relhum, pres, swdown, difrad, lwdown, wu, wv, winddir[tsteps],
     * p_soilm, p_G, p_umu, p_kp, p_muGp, p_dtrp (+ p_Tg, p_Tbp when reqhgt < 0); es, ea, tdew and windspeed are
     * ignored.  Fine row i / column j sits at fractional coarse row clim_row0 + clim_drow * i / column
     * clim_col0 + clim_dcol * j (clamped to the hull of the coarse cell centres).
     * altcorrect = 0: pres is the coarse pressure.  altcorrect = 1 / 2: pres is the coarse SEA-LEVEL pressure
     * pk / ((293 - 0.0065 dtmc) / 293)^5.26, pfac[rows, cols] = ((293 - 0.0065 dtm) / 293)^5.26 and
     * elevd[rows, cols] = resample(dtmc) - dtm; 1 = fixed lapse rate 5 K / km, 2 = humidity-dependent. */
    int32_t clim_rows, clim_cols;
    double clim_row0, clim_drow, clim_col0, clim_dcol;
    int32_t altcorrect;
    const double* relhum;
    const double* wu;
    const double* wv;
    const double* elevd;
    const double* pfac;
} mcf_problem;

/* ------------------------------------------------------------------------------------------- */
/* Host-buffer entry points: what the Rcpp stub calls.  Inputs and outputs are HOST memory; the
 * library uploads the static layers once, streams the time axis through the device in chunks and
 * copies the results back (pinned staging, copy/compute overlap).                               */
/* ------------------------------------------------------------------------------------------- */

/* out[v] == NULL  <=>  out[v] FALSE in the reference call.  Each non-NULL buffer holds
 * rows*cols*tsteps doubles and is completely overwritten (NA_REAL where the reference leaves its
 * prefill).  Returns MCF_OK or an error code with a message in err (NUL-terminated, truncated). */
int mcf_runmicro(const mcf_problem* prob, double* const out[MCF_NOUT], char* err, size_t errlen);

/* runbioclimNCpp: tsteps must be 336 (14 days).  wetq/dryq/hotq/colq are 0-based hour indices
 * (src/microclimfCpp.cpp:3317-3360); `air` selects Tz (1) or tleaf (0).  bio[b] == NULL <=> out[b]
 * FALSE; each non-NULL buffer holds rows*cols doubles.  bio3 and bio7 are derived from bio2/5/6
 * computed internally, whether or not those outputs are requested.
 * For reqhgt >= 0 the 19 reductions are accumulated inside the grid kernel while the 14 days are solved: the
 * two [rows, cols, 336] arrays the reference materialises (src/microclimfCpp.cpp:3509-3515) never exist, so the
 * raster size is bounded by the static layers alone.  Below ground the series needs the time-axis pass first
 * and is reduced in chunks of cells (bounded scratch).  tsteps > 336 is accepted as the reference accepts it
 * (soil statistics over all hours; hours no whole day covers make bio15 NA). */
int mcf_runbioclim(const mcf_problem* prob, const int32_t* wetq, int32_t nwetq, const int32_t* dryq,
                   int32_t ndryq, const int32_t* hotq, int32_t nhotq, const int32_t* colq,
                   int32_t ncolq, int32_t air, double* const bio[MCF_NBIO], char* err, size_t errlen);

/* ------------------------------------------------------------------------------------------- */
/* Device-buffer entry points: every pointer in `prob` and `out` is DEVICE memory on the current
 * device (scalars and the lyr_st/lyr_ed/year/month/day arrays stay on the HOST).  Used by the
 * throughput benchmark (inputs resident in HBM) and by callers that keep rasters on the GPU.    */
/* ------------------------------------------------------------------------------------------- */

/* A time window: compute day-blocks [block0, block0 + nblocks) of the problem (a day-block is one
 * 24-hour block; modes 1/2: block b = hours 24b..24b+23; modes 3/4: the blocks of layer 0, then
 * layer 1, ... which must be ascending and non-overlapping).  Hour k of the problem is written to
 * time slot ((k - hour0) mod ring_hours) of each output buffer, whose slot stride is rows*cols;
 * ring_hours >= 24 lets a caller reuse a small buffer as a ring (the output sink for rasters whose
 * full [rows, cols, tsteps] result exceeds HBM).  Window {0, -1, 0, tsteps} = the whole problem.
 * Preconditions (MCF_ERR_ARG otherwise): 0 <= hour0 <= first hour of block0, ring_hours >= 24.
 * A PARTIAL window writes the hours of its day-blocks and nothing else: hours no block covers and outputs
 * the requested height never produces are NA-filled only when the window is the whole problem. */
typedef struct mcf_window {
    int32_t block0;
    int32_t nblocks; /* -1 = all remaining blocks */
    int64_t hour0;
    int64_t ring_hours;
} mcf_window;

/* reqhgt >= 0 only when the window is partial (the below-ground pass needs every hour of a cell).
 * `stream` is a cudaStream_t (NULL = default stream); the call is asynchronous on that stream
 * unless `err` reporting requires otherwise (launch errors are reported synchronously). */
int mcf_runmicro_dev(const mcf_problem* prob, double* const out[MCF_NOUT], const mcf_window* win,
                     void* stream, char* err, size_t errlen);

int mcf_runbioclim_dev(const mcf_problem* prob, const int32_t* wetq, int32_t nwetq,
                       const int32_t* dryq, int32_t ndryq, const int32_t* hotq, int32_t nhotq,
                       const int32_t* colq, int32_t ncolq, int32_t air, double* const bio[MCF_NBIO],
                       void* stream, char* err, size_t errlen);

/* ------------------------------------------------------------------------------------------- */
/* Summary sink: the hourly arrays of a large raster exist nowhere (config 4 of BASELINE.json: 4.7 TB per   */
/* variable; the reference's runmicro_big, R/Cppwrappers.R:444-543, writes them tile by tile to disk).  These */
/* entry points run the same solve and keep, per cell and requested output, the SUM, MINIMUM and MAXIMUM over */
/* the computed hours of the window, accumulated inside the grid kernel; nothing hourly is stored.  NaN hours   */
/* poison the sum and are ignored by the extremes; cells the reference skips (hgt NA) hold NA_real_; outputs     */
/* the requested height never produces (tleaf / relhum at the surface) are NA.  reqhgt >= 0.                     */
/* sum[v], mn[v], mx[v]: all three NULL (output not summarised) or all three [rows * cols].  accumulate != 0     */
/* merges the window into what the buffers already hold (successive windows of one series); hours_done (may be   */
/* NULL) receives the number of hours the call added.                                                            */
/* ------------------------------------------------------------------------------------------- */
int mcf_runmicro_summary_dev(const mcf_problem* prob, double* const sum[MCF_NOUT], double* const mn[MCF_NOUT],
                             double* const mx[MCF_NOUT], const mcf_window* win, int32_t accumulate, int64_t* hours_done,
                             void* stream, char* err, size_t errlen);
/* HOST buffers, whole series: mean (= sum / hours_done), minimum and maximum per cell. */
int mcf_runmicro_summary(const mcf_problem* prob, double* const mean[MCF_NOUT], double* const mn[MCF_NOUT],
                         double* const mx[MCF_NOUT], int64_t* hours_done, char* err, size_t errlen);

/* ------------------------------------------------------------------------------------------- */
/* Packed integer sink (SURVEY.md NEXT-4).  The reference stores hourly grids on disk as integers: */
/* writetonc (R/dataprep.R:1063-1260) writes `as.integer(round(x * rd, 0))` with rd = 100 for Tz,   */
/* tleaf, soilm and windspeed and rd = 1 for relhum and the radiation streams (atonc :1064-1069,     */
/* :1164-1173), missing value -9999.  These entry points run the same solve and store exactly those  */
/* integers (round half to even, NA / NaN / Inf -> -9999) as int16 in the kernel's store, 2 bytes    */
/* instead of 8 per value on the PCIe link and in the sink.  Layout stays [rows, cols, tsteps] (R);  */
/* writetonc's aperm(c(2, 1, 3)) is file-format work for the writer.  |x * rd| > 32767 saturates     */
/* (not reachable for physical values; the reference's NC_INT would hold it).                        */
/* ------------------------------------------------------------------------------------------- */
#define MCF_PACKED_NA (-9999)
int mcf_runmicro_packed(const mcf_problem* prob, int16_t* const out[MCF_NOUT], char* err, size_t errlen);
int mcf_runmicro_packed_dev(const mcf_problem* prob, int16_t* const out[MCF_NOUT], const mcf_window* win,
                            void* stream, char* err, size_t errlen);

/* ------------------------------------------------------------------------------------------- */
/* FP32 build (BASELINE north_star: "an optional FP32 build must stay within 0.05 degC and 0.5 % radiation").  `prob` is
 * the FP64 problem of mcf_runmicro[_dev] — every mode (data.frame, fine-array and coarse-grid climate, layered
 * vegetation) and every height.  The hour loops run in FP32 (SFU transcendentals); per-cell invariants, the per-cell-hour
 * assembly of array climate (interpolation, altitude correction, solar position) and everything below ground stay FP64
 * (reqhgt < 0: the time-axis pass is discontinuous in its inputs — a rolling mean over round(-118.35 z / mean damping
 * depth) hours — so its ground temperatures and damping depths are computed in FP64 and only the results are narrowed).  Outputs are float arrays (quiet NaN where the FP64 entry points write NA_real_): 40 instead of 80 bytes per
 * cell-hour in HBM and over PCIe.  Window / ring as for mcf_runmicro_dev. */
int mcf_runmicro_f32_dev(const mcf_problem* prob, float* const out[MCF_NOUT], const mcf_window* win, void* stream,
                         char* err, size_t errlen);
/* HOST buffers: the same host path as mcf_runmicro (one upload, windowed copy-back overlapping the kernels, pageable
 * destinations served by the pinned-slot copy pool), moving half the bytes. */
int mcf_runmicro_f32(const mcf_problem* prob, float* const out[MCF_NOUT], char* err, size_t errlen);

/* sum and count of log(twi)/tfact over the non-NaN cells of a HOST twi buffer holding n cells
 * (the two numbers the bands all-reduce before calling with has_twi_mean = 1). */
int mcf_twi_partial(const double* twi, int64_t n, double tfact, double* sum, int64_t* count,
                    char* err, size_t errlen);

/* ------------------------------------------------------------------------------------------- */
/* Device management and instrumentation                                                          */
/* ------------------------------------------------------------------------------------------- */
int mcf_abi_version(void);
/* The host-buffer entry points keep one grow-only device workspace between calls (allocating and
 * freeing tens of arrays per call costs more than the solve); this releases it.  It belongs to the device
 * that was current when it was allocated: mcf_set_device() to another device releases it first. */
void mcf_release_workspace(void);
int mcf_device_count(void);
int mcf_set_device(int device);
/* kernels launched by this library in this process since the last reset (bench.py's gpu_launches) */
int64_t mcf_launch_count(void);
void mcf_launch_count_reset(void);
/* Device time (ms, CUDA events on the launching stream) and number of grid-kernel launches
 * accumulated since the last reset: the dominant kernel's average launch duration for the roofline. */
int mcf_kernel_time(double* total_ms, int64_t* launches);
void mcf_kernel_time_reset(void);
void mcf_kernel_timing_enable(int on);
/* DFMA micro-benchmark: measured non-tensor FP64 peak of the current device, TFLOP/s
 * (2 flop per DFMA); the FP64 roofline denominator that MEASURED_PEAKS.json does not carry. */
int mcf_fp64_peak(double* tflops, char* err, size_t errlen);

/* ------------------------------------------------------------------------------------------- */
/* Terrain preparation (SURVEY.md NEXT-2): the pure-R stencils that produce `hor`, `svfa` and the   */
/* wind-shelter coefficients from the DTM.  HOST buffers, R layout.                                  */
/* ------------------------------------------------------------------------------------------- */

/* .horizon (R/internal.R:909-925) for `nazi` azimuths (degrees; the grid model uses 0, 15, ..., 345,
 * R/internal.R:1142-1145): hor is [rows, cols, nazi].  svfa (may be NULL) receives the sky-view factor
 * 0.5 cos(2 tan(mean(atan(hor)))) + 0.5 over those azimuths (R/internal.R:1146-1148).  reso = cell size (m). */
int mcf_horizon(const double* dtm, int32_t rows, int32_t cols, double reso, int32_t nazi, const double* azimuth_deg,
                double* hor, double* svfa, char* err, size_t errlen);

/* .windcoef (R/internal.R:949-968) for `ndir` directions: index is [rows, cols, ndir] (may be NULL).
 * blend8 (may be NULL, needs ndir == 16) receives the 16 -> 8 direction blend of .windsheltera
 * (R/internal.R:983-989) WITHOUT its terra aggregate/resample smoothing, which is third-party arithmetic. */
int mcf_windcoef(const double* dsm, int32_t rows, int32_t cols, double reso, double hgt, int32_t ndir,
                 const double* direction_deg, double* index, double* blend8, char* err, size_t errlen);

/* .windsheltera (R/internal.R:970-991) as ONE call, everything on the device: .windcoef in 16 directions at height
 * `hgt`, each direction smoothed by terra::aggregate(fact = s, fun = "mean") + terra::resample (bilinear, :979-981;
 * s <= 1: no smoothing), blended to 8 sectors (:983-989).  wsa8 is [rows, cols, 8].  The terra steps are restated from
 * their published definitions (parity unpinned: no R here). */
int mcf_windshelter(const double* dsm, int32_t rows, int32_t cols, double reso, double hgt, int32_t s, double* wsa8,
                    char* err, size_t errlen);

/* terra::terrain(dtm, v = "slope" / "aspect") (R/internal.R:1124-1129, R/Cppwrappers.R:483-484): Horn's 8-neighbour
 * finite difference, degrees; NaN on the edge and beside missing cells (terra: NA); aspect = downslope bearing clockwise
 * from north, 90 on flat ground.  Either output may be NULL.  Row 0 of the matrix is the raster's northern edge. */
int mcf_slope_aspect(const double* dtm, int32_t rows, int32_t cols, double xres, double yres, double* slope, double* aspect,
                     char* err, size_t errlen);

/* .topidx (R/internal.R:861-874): a / tan(B) with B the Horn slope (device kernel) floored at atan(0.02 / mean(res)) and
 * missing slopes replaced by the median, a = (flowaccCpp + 1) x cell area floored at 1 (host sweep), NaN where dtm is. */
int mcf_topidx(const double* dtm, int32_t rows, int32_t cols, double xres, double yres, double* twi, char* err, size_t errlen);

/* flowaccCpp (src/microclimfCpp.cpp:5368-5414, with flowdirCpp :5326-5366): D8 flow accumulation of a
 * [rows, cols] elevation matrix, the input of .topidx (R/internal.R:861-874).  HOST code (one sequential
 * sweep over the cells sorted by elevation); NaN cells receive (double)INT_MIN as in the reference. */
int mcf_flowacc(const double* dtm, int32_t rows, int32_t cols, double* fa, char* err, size_t errlen);

/* ------------------------------------------------------------------------------------------- */
/* Snow (SURVEY.md NEXT-3), data.frame climate.                                                    */
/*   mcf_gridmodelsnow  <- _microclimf_gridmodelsnow1  (src/microclimfCpp.cpp:4172-4424): the hourly   */
/*                         snow-pack recurrence per cell (snowoneB :3835, radoneB :3773)             */
/*   mcf_gridmicrosnow  <- _microclimf_gridmicrosnow1  (src/microclimfCpp.cpp:4894-5057): microclimate  */
/*                         above / below the snow surface where SWE > 0, overwriting runmicro's outputs */
/* HOST buffers, R layout.  The R drivers around them (.snowmodel1, .runmicrosnow1: 5-day chunks with  */
/* terrain recomputed from DTM + snow depth, R/internal.R:2389-2779, 3367-3660) are host R code.       */
/* ------------------------------------------------------------------------------------------- */
typedef struct mcf_snow_climate { /* obstime + climdata, each of length tsteps */
    int32_t tsteps;
    const int32_t* year;
    const int32_t* month;
    const int32_t* day;
    const double* hour;
    const double *temp, *relhum, *pres, *swdown, *difrad, *lwdown, *windspeed, *winddir, *precip;
} mcf_snow_climate;

typedef struct mcf_snow_point { /* pointm of gridmodelsnow1 (:4181-4186), each of length tsteps */
    const double *Gp, *Tc, *RswabsG, *RlwabsG, *umu;
} mcf_snow_point;

typedef struct mcf_snow_static { /* vegp + other ([rows, cols] unless noted) */
    int32_t rows, cols;
    const double *pai, *hgt, *leaft, *clump;
    const double *paia, *leafd, *leafden, *Smax; /* gridmicrosnow only */
    const double *slope, *aspect, *skyview;
    const double* wsa;                          /* [rows, cols, 8]  */
    const double* hor;                          /* [rows, cols, 24] */
    double lat, lon, zref;
    const double *isnowdc, *isnowdg;            /* gridmodelsnow only: initial snow depth (canopy + ground, ground) */
    const int32_t *isnowac, *isnowag;           /* gridmodelsnow only: initial snow age, hours */
    const double *lats, *lons;                  /* array-climate variants only: per-cell latitude / longitude */
} mcf_snow_static;

typedef struct mcf_snow_state { /* snowm of gridmicrosnow1 (:4935-4940), each [rows, cols, tsteps] */
    const double *Tc, *Tg, *totalSWE, *groundsnowdepth, *snowden;
} mcf_snow_state;

/* snowenv: 0 Alpine (default), 1 Maritime, 2 Prairie, 3 Tundra, 4 Taiga (snowdenp :3741-3750).
 * out3d = {Tc, Tg, sdepc, sdepg, sden}, each [rows, cols, tsteps]; out2d = {agec, ageg, meltc, meltg}, each
 * [rows, cols]; NULL entries are skipped. */
int mcf_gridmodelsnow(const mcf_snow_climate* clim, const mcf_snow_point* pt, const mcf_snow_static* st, int32_t snowenv,
                      double* const out3d[5], double* const out2d[4], char* err, size_t errlen);

/* `umu` is the climdata$umu column (:4914).  micro[v] (NULL = out[v] FALSE) are runmicro's [rows, cols, tsteps]
 * outputs, updated IN PLACE where totalSWE > 0. */
int mcf_gridmicrosnow(double reqhgt, const mcf_snow_climate* clim, const double* umu, const mcf_snow_state* sm,
                      const mcf_snow_static* st, double mat, double* const micro[MCF_NOUT], char* err, size_t errlen);

/* Array climate: _microclimf_gridmodelsnow2 (src/microclimfCpp.cpp:4426-4673) and _microclimf_gridmicrosnow2
 * (:5059-5214).  Every series of `clim` (except winddir, length tsteps), of `pt` and `umu` is a
 * [rows, cols, tsteps] array; st->lats / st->lons are required, st->lat / st->lon ignored. */
int mcf_gridmodelsnow2(const mcf_snow_climate* clim, const mcf_snow_point* pt, const mcf_snow_static* st, int32_t snowenv,
                       double* const out3d[5], double* const out2d[4], char* err, size_t errlen);
int mcf_gridmicrosnow2(double reqhgt, const mcf_snow_climate* clim, const double* umu, const mcf_snow_state* sm,
                       const mcf_snow_static* st, double mat, double* const micro[MCF_NOUT], char* err, size_t errlen);

/* Element-wise evaluation of the kernels' own FP64 elementary functions (csrc/mcf_math.cuh) on HOST
 * buffers, for accuracy tests: fn 0 = 1/x, 1 = x/y, 2 = sqrt(x), 3 = exp(x), 4 = 2^x, 5 = log(x),
 * 6 = x^y, 7 = sin(x), 8 = cos(x).  y may be NULL for the one-operand functions. */
int mcf_math_eval(int fn, const double* x, const double* y, int64_t n, double* out, char* err, size_t errlen);

#ifdef __cplusplus
}
#endif
#endif /* MICROCLIMF_B200_H */
