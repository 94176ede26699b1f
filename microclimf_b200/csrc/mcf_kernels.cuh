// microclimf_b200 — kernel argument blocks and launch wrappers (implemented in mcf_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mcf_physics.cuh"

namespace mcf {

#ifndef MCF_TILE
#define MCF_TILE 384
#endif
#ifndef MCF_MINB
#define MCF_MINB 1
#endif
constexpr int kTile = MCF_TILE;  // cells per CTA tile = threads per CTA (one thread per cell)
constexpr int kMinBlocks = MCF_MINB; // resident CTAs per SM the grid kernel is compiled for
constexpr int kNOut = 10;

// Per-hour calendar record for the array-climate modes (solar position is per cell there): everything
// of solpositionCpp (ref :48-57) that depends on the date and time only.
struct HourCal {
    double eot;     // equation of time, minutes (ref soltimeCpp :42-43)
    double sd, cd;  // sin / cos of the solar declination (ref :55)
    double lt;      // local time, decimal hours
    int32_t windex; // wind-shelter sector (ref :2443)
    int32_t pad;
};

// One 24-hour block of work: hours k0 .. k0+23 solved with vegetation layer `lyr`.
struct DayBlock {
    int32_t k0;
    int32_t lyr;
};

enum { RQ_ABOVE = 0, RQ_SURFACE = 1, RQ_BELOW = 2 };

// Where the hourly results of the grid kernel go (compile-time parameter of k_grid):
//   SINK_F64     [rows, cols, T] FP64 arrays or a ring of time slots (the reference's return value)
//   SINK_PACK    the same as writetonc's int16 (SURVEY.md NEXT-4)
//   SINK_BIO     nowhere: the 19 bioclim reductions of runbioclimCpp (ref :3457-3560) are accumulated per cell in
//                shared memory while the days are solved, and only the [rows, cols] summaries are written
//   SINK_SUMMARY nowhere: per-cell sum / minimum / maximum over the window's hours of each requested output — the
//                sink of runmicro_big for rasters whose hourly arrays exist nowhere
enum { SINK_F64 = 0, SINK_PACK = 1, SINK_BIO = 2, SINK_SUMMARY = 3 };

// accumulator slots per thread (shared memory, [slot][thread]) of the reducing sinks
enum {
    BIO_S1 = 0, BIO_DTRSUM, BIO_MON0, BIO_B5 = BIO_MON0 + 12, BIO_B6, BIO_QT0, BIO_TZ0 = BIO_QT0 + 4,
    BIO_M12, BIO_B13, BIO_B14, BIO_K, BIO_SD, BIO_SD2, BIO_QS0, BIO_DMX = BIO_QS0 + 4, BIO_DMN, BIO_DSUM, BIO_NSLOT
};
constexpr int kSummarySlots = 30; // [stat 0 sum, 1 min, 2 max][output]
constexpr int kAccSlots = (BIO_NSLOT > kSummarySlots) ? BIO_NSLOT : kSummarySlots;

struct GridArgs {
    int32_t ncells;     // rows*cols of the problem = time-slot stride of every [rows, cols, n] array
    int32_t cell_begin; // cell range solved by this launch
    int32_t cell_end;
    int32_t tsteps;
    int32_t nlyr;
    double reqhgt2; // max(reqhgt, 1e-5)   (ref :2246-2247)
    double zref, lat;
    double tfact;            // tadd = log(twi)/tfact - mean   (ref soildCppm :975-1019)
    int32_t has_tadd_mean;   // caller-supplied whole-raster mean (band sharding), else dscal[1]/dscal[2]
    double tadd_mean;
    const double* dscal;     // device scalars written by the prep kernels: [0] series max of tc (modes 1/3,
                             // ref :2159-2168), [1] sum and [2] count of log(twi)/tfact over non-NaN cells
    const HourRec* hours;    // modes 1/3: [tsteps]
    // modes 2/4: [tsteps * ncells] arrays
    const double* clim[9]; // tc es ea tdew pk swdown difrad lwdown windspeed
    const double* pnt[6];  // soilm G umu kp muGp dtrp
    const double* lats;
    const double* lons;
    const double* mxtc_cell; // per-cell max of tc over time (ref :2467-2471)
    const HourCal* cal;
    // coarse-grid climate (ARR == 2): clim[] / pnt[] are [clim_rows, clim_cols, tsteps] and are interpolated
    // bilinearly per cell-hour; clim[1..3] (es, ea, tdew) and clim[8] (windspeed) are unused
    int32_t rows;            // fine rows: cell = i + rows * j
    int32_t clim_rows, clim_cols, altcorrect;
    double clim_row0, clim_drow, clim_col0, clim_dcol;
    const double* relhum;
    const double* wu;
    const double* wv;
    const double* elevd;     // [ncells] (altcorrect != 0)
    const double* pfac;      // [ncells] (altcorrect != 0)
    const double* cpack;     // [tsteps][clim_rows * clim_cols][16]: the 14 coarse series as one 128-byte record per
                             // (hour, node), built once per call by k_pack_coarse (k_grid reads these, not clim[] / pnt[])
    // statics
    const double* veg[10];  // hgt pai x gsmax leafr leaft clump leafd paia leafden   [nlyr * ncells]
    const double* soil[13]; // Smin Smax gref soilb Psie Vq Vm Mc rho slope aspect twi svfa
    const double* wsa;      // [8 * ncells]
    const double* hor;      // [24 * ncells]
    // work list
    const DayBlock* blocks;
    int32_t block0, nblocks;
    long long hour0, ring_hours;
    // outputs: hour slot s of cell c is element s * out_stride + (c - out_cell0) of each buffer (the whole raster:
    // out_stride = ncells, out_cell0 = 0; a cell chunk with its own compact buffers: its width and first cell)
    int32_t out_stride, out_cell0;
    double* out[kNOut];     // pack != 0: each is an int16_t* in disguise (the packed integer sink)
    uint32_t outmask;
    int32_t pack;
    // reducing sinks.  SINK_BIO: red[b] = bio(b+1) [ncells] (NULL = not requested), red_mask = requested outputs;
    // SINK_SUMMARY: red[stat * 10 + v] [ncells], stat 0 sum / 1 min / 2 max of output v over the window's hours,
    // outmask = outputs summarised, red_accumulate != 0 merges into what the buffers hold (successive windows)
    double* red[kSummarySlots];
    uint32_t red_mask;
    int32_t red_accumulate;
    int32_t bio_air;            // 1: Tz, 0: tleaf (ref runbioclim1Cpp :3570-3590)
    int32_t bio_soil_gap;       // some hour of [0, tsteps) is never computed: the 336-hour soil sd is NA
    uint32_t bio_q_na;          // bit q: quarter q indexes an hour that is never computed (its sums are NA)
    const uint32_t* bio_qcnt;   // [tsteps] multiplicity of hour k in wetq | dryq << 8 | hotq << 16 | colq << 24
    // scratch
    double* stash;          // [gridDim.x][24][kStashVars][kTile]
    unsigned int* tile_counter;
    double* tg_scratch;     // RQ_BELOW: [tsteps][cell_end - cell_begin] ground temperature series
    double* dd_sum;         // RQ_BELOW: [cell_end - cell_begin] sum of damping depths
};

struct BelowArgs {
    int32_t width;  // cells in this chunk
    int32_t ncells; // slot stride of the Tgp / Tbp / hgt arrays
    int32_t tz_stride, tz_cell0; // Tz: element k * tz_stride + (cell - tz_cell0)
    int32_t cell_begin;
    int32_t tsteps;
    int32_t arr;    // Tgp/Tbp are per-cell arrays (modes 2/4) or per-hour vectors
    int32_t complete;
    int32_t hiy;
    double reqhgt, mat;
    const double* tg;     // [tsteps][width]
    const double* dd_sum; // [width]
    const double* Tgp;
    const double* Tbp;
    const double* hgt;    // first-layer vegetation height: NaN => skipped cell
    double* daily;        // scratch [2][ndays][width]
    double* Tz;           // [tsteps][ncells]
};

struct BioArgs {
    int32_t width, tsteps;
    int32_t stride;      // hour stride of Tz / soilm (>= width)
    const double* Tz;    // [tsteps][stride]
    const double* soilm; // [tsteps][stride]
    const int32_t* q[4]; // wetq dryq hotq colq (device)
    int32_t nq[4];
    double* bio[19];     // each [ncells], written at cell_begin + c
    int32_t cell_begin;
    uint32_t mask;
};

// launch wrappers (all asynchronous on `stream`)
cudaError_t launch_prep_hours(const int32_t* year, const int32_t* month, const int32_t* day, const double* hour,
                              const double* const clim[10], const double* const pnt[6], double lat, double lon,
                              int tsteps, bool arr, HourRec* hours, HourCal* cal, double* mxtc_out,
                              cudaStream_t stream);
cudaError_t launch_mxtc_cell(const double* tc, int ncells, int tsteps, double* mxtc_cell, cudaStream_t stream);
cudaError_t launch_mxtc_cell_coarse(const GridArgs& a, double* mxtc_cell, cudaStream_t stream);
size_t coarse_pack_doubles(const GridArgs& a);
cudaError_t launch_pack_coarse(const GridArgs& a, double* out, cudaStream_t stream);
// coarse [clim_rows, clim_cols, tsteps] -> fine [ncells, tsteps] with the grid kernel's own interpolation
cudaError_t launch_interp_coarse(const GridArgs& a, const double* coarse, double* fine, cudaStream_t stream);
cudaError_t launch_twi_sum(const double* twi, int64_t n, double tfact, double* sum_count /* [2] */,
                           cudaStream_t stream);
cudaError_t launch_grid(const GridArgs& a, int arr /* 0 table, 1 fine arrays, 2 coarse arrays */, int rq, int grid,
                        cudaStream_t stream, int sink = -1 /* default: SINK_PACK if a.pack else SINK_F64 */);
int grid_blocks_per_sm(bool arr, int rq);
// the pair build of the FP64 kernel (mcf_kernels_pair.inl): two threads per cell, per-cell invariants in shared memory
bool pair_eligible(int arr, int rq, int sink);
int pair_tile();
size_t pair_scratch_doubles(); // per CTA: day stash + reduction exchange
cudaError_t launch_grid_pair(const GridArgs& a, int arr, int rq, int grid, cudaStream_t stream, int sink = -1);
// FP32 build (modes 1/3, reqhgt >= 0): narrowed hour table, FP32 stash and outputs
cudaError_t launch_narrow_hours(const HourRec* in, int n, void* out, cudaStream_t stream);
size_t hourrec_f32_bytes();
int f32_blocks_per_sm();
int f32_tile();
cudaError_t launch_grid_f32(const GridArgs& a, const void* hoursf, float* const outf[kNOut], float* stashf, int arr, int rq,
                            int grid, cudaStream_t stream);
cudaError_t launch_narrow32(const double* src, float* dst, int64_t n, cudaStream_t stream);
cudaError_t launch_fill32(float* p, int64_t n, cudaStream_t stream);
cudaError_t launch_below(const BelowArgs& a, cudaStream_t stream);
cudaError_t launch_bioclim(const BioArgs& a, cudaStream_t stream);
cudaError_t launch_fill_na(double* p, int64_t n, cudaStream_t stream);
cudaError_t launch_fill16(int16_t* p, int64_t n, int16_t v, cudaStream_t stream);
cudaError_t launch_pack16(const double* src, int16_t* dst, int64_t n, double rd, cudaStream_t stream);
// terrain preparation (mcf_terrain.cu)
cudaError_t launch_scale_dtm(const double* dtm, int64_t n, double reso, double* out, cudaStream_t st);
cudaError_t launch_horizon(const double* d, int rows, int cols, int ndir, const double* offs, double thr, bool windcoef,
                           double* out, cudaStream_t st);
cudaError_t launch_skyview(const double* hor, int64_t nc, int ndir, double* svf, cudaStream_t st);
cudaError_t launch_blend16to8(const double* a, int64_t nc, double* out, cudaStream_t st);
cudaError_t launch_horn(const double* z, int rows, int cols, double dx, double dy, double* slope, double* aspect,
                        cudaStream_t st);
// block mean over fact x fact cells into `coarse` ([ceil(rows/fact), ceil(cols/fact), nl]), then bilinear back to `out`
cudaError_t launch_smooth(const double* src, int rows, int cols, int nl, int fact, double* coarse, double* out, cudaStream_t st);
cudaError_t launch_math_eval(int fn, const double* x, const double* y, int64_t n, double* out, cudaStream_t stream);
cudaError_t launch_fp64_peak(double* sink, int grid, int iters, cudaStream_t stream);

} // namespace mcf
