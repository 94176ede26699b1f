// microclimf_b200 — terrain preparation kernels (SURVEY.md §8f NEXT-2): the pure-R stencils that turn the
// DTM into the `hor`, `svfa` and wind-shelter inputs of the grid solver.
//
//   .horizon   R/internal.R:909-925   tangent of the horizon angle in one direction: max over 10 steps of
//                                     (dtm[shifted by step^2 cells] - dtm) / step^2, zero padding, NA -> 0
//   svfa       R/internal.R:1146-1148 0.5 cos(2 tan(mean(atan(hor)))) + 0.5 over the 24 directions
//   .windcoef  R/internal.R:949-968   the same stencil with a height threshold, 1 - atan(17 hor) / 1.65
//   blend      R/internal.R:983-989   16 -> 8 wind directions
//
// R forms the shifted window with `from:to` on non-integer bounds and truncating REAL subscripts, so the
// source row of output row i is trunc((101 - cos(azi) step^2) + i) evaluated in double precision PER
// ELEMENT (the rounding of that sum decides the cell near x.5 boundaries); the kernels evaluate the very
// same expression.  One thread per (cell, direction); cells are the fastest axis (R layout), so the ten
// gathers of a warp are ten contiguous 256-byte segments of the scaled DTM (L2-resident stencil reach:
// 100 cells).  HBM-bound: 8 B in (re-read from L2) + 8 B out per (cell, direction).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "mcf_kernels.cuh"

namespace mcf {

__global__ void k_scale_dtm(const double* __restrict__ dtm, int64_t n, double reso, double* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = dtm[i];
        if (isnan(v)) v = 0.0;
        out[i] = v / reso;
    }
}

// offs: [ndir][10][2] = (101 - cos(azi) s^2, 101 + sin(azi) s^2); thr > 0 enables the .windcoef threshold
__global__ void __launch_bounds__(256) k_horizon(const double* __restrict__ d, int rows, int cols, int ndir,
                                                 const double* __restrict__ offs, double thr, int windcoef,
                                                 double* __restrict__ out) {
    const int64_t nc = (int64_t)rows * cols;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nc * ndir) return;
    const int a = (int)(t / nc);
    const int64_t cell = t - (int64_t)a * nc;
    const int i = (int)(cell % rows), j = (int)(cell / rows);
    const double centre = d[cell];
    double hor = 0.0;
#pragma unroll 1
    for (int step = 1; step <= 10; ++step) {
        const double fr = __ldg(&offs[(a * 10 + step - 1) * 2 + 0]);
        const double fc = __ldg(&offs[(a * 10 + step - 1) * 2 + 1]);
        const double s2 = (double)(step * step);
        // R: REAL subscript (1-based, into the padded array) truncated toward zero
        const int ri = (int)trunc(fr + (double)i) - 101; // back to 0-based rows of the unpadded DTM
        const int ci = (int)trunc(fc + (double)j) - 101;
        double src = 0.0; // the 100-cell zero border
        if (ri >= 0 && ri < rows && ci >= 0 && ci < cols) src = d[(int64_t)ci * rows + ri];
        const double v = (src - centre) / s2;
        hor = (hor < v) ? v : hor; // pmax
        if (windcoef && hor < (thr / s2)) hor = 0.0;
    }
    out[t] = windcoef ? 1.0 - atan(0.17 * 100 * hor) / 1.65 : hor;
}

__global__ void k_skyview(const double* __restrict__ hor, int64_t nc, int ndir, double* __restrict__ svf) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    double s = 0.0;
    for (int a = 0; a < ndir; ++a) s += atan(hor[c + nc * a]);
    const double msl = tan(s / ndir);
    svf[c] = 0.5 * cos(2 * msl) + 0.5;
}

__global__ void k_blend16to8(const double* __restrict__ a, int64_t nc, double* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nc * 8) return;
    const int i = (int)(t / nc); // 0-based output direction
    const int64_t c = t - (int64_t)i * nc;
    const int mid = 2 * i, nxt = 2 * i + 1, prv = (i == 0) ? 15 : 2 * i - 1;
    out[t] = 0.5 * a[c + nc * mid] + 0.25 * a[c + nc * nxt] + 0.25 * a[c + nc * prv];
}

// terra::terrain(v = "slope" | "aspect", neighbors = 8) (R/internal.R:1124-1129, R/Cppwrappers.R:483-484): Horn's
// (1981) third-order finite difference on the 3 x 3 neighbourhood.  R layout: cell = i + rows * j, row i counts from
// the NORTHERN edge, column j from the western.  Edge cells and cells with a missing neighbour are NA (NaN here), as
// terra leaves them; degrees.  Aspect is the downslope bearing clockwise from north, 90 for a flat cell.
// HBM-bound: the 9 gathers of a warp are 3 x 3 contiguous segments (L1/L2 reuse), 8 B in + 16 B out per cell.
__global__ void __launch_bounds__(256) k_horn(const double* __restrict__ z, int rows, int cols, double dx, double dy,
                                              double* __restrict__ slope, double* __restrict__ aspect) {
    const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= (int64_t)rows * cols) return;
    const int i = (int)(cell % rows), j = (int)(cell / rows);
    double sl = nan(""), as = nan("");
    if (i > 0 && i < rows - 1 && j > 0 && j < cols - 1) {
        const double* c0 = z + (int64_t)(j - 1) * rows + i; // west column
        const double* c1 = z + (int64_t)j * rows + i;
        const double* c2 = z + (int64_t)(j + 1) * rows + i; // east column
        const double a = c0[-1], d = c0[0], g = c0[1];      // north-west, west, south-west
        const double b = c1[-1], e = c1[0], h = c1[1];
        const double c = c2[-1], f = c2[0], k = c2[1];
        const double dzdx = ((c + 2.0 * f + k) - (a + 2.0 * d + g)) / (8.0 * dx); // towards the east
        const double dzdy = ((a + 2.0 * b + c) - (g + 2.0 * h + k)) / (8.0 * dy); // towards the north
        if (!isnan(e) && !isnan(dzdx) && !isnan(dzdy)) {
            sl = atan(sqrt(dzdx * dzdx + dzdy * dzdy)) * (180.0 / 3.14159265358979323846);
            double r = 0.5 * 3.14159265358979323846 - atan2(-dzdy, -dzdx);
            r = fmod(r, 2.0 * 3.14159265358979323846);
            if (r < 0.0) r += 2.0 * 3.14159265358979323846;
            if (dzdx == 0.0 && dzdy == 0.0) r = 0.5 * 3.14159265358979323846;
            as = r * (180.0 / 3.14159265358979323846);
        }
    }
    if (slope) slope[cell] = sl;
    if (aspect) aspect[cell] = as;
}

// terra::aggregate(fact, fun = "mean") followed by terra::resample(method = "bilinear") back onto the fine raster
// (R/internal.R:979-981: the smoothing of each wind-shelter direction).  Block means over fact x fact cells (ragged edge
// blocks average the cells they have); bilinear interpolation between the block centres, constant beyond their hull.
__global__ void k_block_mean(const double* __restrict__ src, int rows, int cols, int nl, int fact, int orows, int ocols,
                             double* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t onc = (int64_t)orows * ocols;
    if (t >= onc * nl) return;
    const int l = (int)(t / onc);
    const int64_t oc = t - (int64_t)l * onc;
    const int oi = (int)(oc % orows), oj = (int)(oc / orows);
    const int i1 = min((oi + 1) * fact, rows), j1 = min((oj + 1) * fact, cols);
    double s = 0.0;
    int n = 0;
    for (int j = oj * fact; j < j1; ++j)
        for (int i = oi * fact; i < i1; ++i) {
            s += src[(int64_t)l * rows * cols + (int64_t)j * rows + i];
            ++n;
        }
    out[t] = s / n;
}
__global__ void k_bilinear_up(const double* __restrict__ coarse, int orows, int ocols, int nl, int fact, int rows, int cols,
                              double* __restrict__ out) {
    const int64_t nc = (int64_t)rows * cols;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nc * nl) return;
    const int l = (int)(t / nc);
    const int64_t cell = t - (int64_t)l * nc;
    const int i = (int)(cell % rows), j = (int)(cell / rows);
    // fractional coarse row / column of the fine cell centre (coarse cells are `fact` fine cells wide)
    double fy = ((double)i + 0.5) / fact - 0.5, fx = ((double)j + 0.5) / fact - 0.5;
    fy = fmin(fmax(fy, 0.0), (double)(orows - 1));
    fx = fmin(fmax(fx, 0.0), (double)(ocols - 1));
    int y0 = min((int)floor(fy), max(orows - 2, 0)), x0 = min((int)floor(fx), max(ocols - 2, 0));
    const int y1 = min(y0 + 1, orows - 1), x1 = min(x0 + 1, ocols - 1);
    const double wy = fy - y0, wx = fx - x0;
    const double* c = coarse + (int64_t)l * orows * ocols;
    const double top = c[(int64_t)x0 * orows + y0] * (1.0 - wx) + c[(int64_t)x1 * orows + y0] * wx;
    const double bot = c[(int64_t)x0 * orows + y1] * (1.0 - wx) + c[(int64_t)x1 * orows + y1] * wx;
    out[t] = top * (1.0 - wy) + bot * wy;
}

cudaError_t launch_horn(const double* z, int rows, int cols, double dx, double dy, double* slope, double* aspect,
                        cudaStream_t st) {
    const int64_t n = (int64_t)rows * cols;
    k_horn<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(z, rows, cols, dx, dy, slope, aspect);
    return cudaGetLastError();
}
cudaError_t launch_smooth(const double* src, int rows, int cols, int nl, int fact, double* coarse, double* out, cudaStream_t st) {
    const int orows = (rows + fact - 1) / fact, ocols = (cols + fact - 1) / fact;
    const int64_t on = (int64_t)orows * ocols * nl, n = (int64_t)rows * cols * nl;
    k_block_mean<<<(unsigned)((on + 255) / 256), 256, 0, st>>>(src, rows, cols, nl, fact, orows, ocols, coarse);
    k_bilinear_up<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(coarse, orows, ocols, nl, fact, rows, cols, out);
    return cudaGetLastError();
}

cudaError_t launch_scale_dtm(const double* dtm, int64_t n, double reso, double* out, cudaStream_t st) {
    int64_t b = (n + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    k_scale_dtm<<<(int)b, 256, 0, st>>>(dtm, n, reso, out);
    return cudaGetLastError();
}
cudaError_t launch_horizon(const double* d, int rows, int cols, int ndir, const double* offs, double thr, bool windcoef,
                           double* out, cudaStream_t st) {
    const int64_t n = (int64_t)rows * cols * ndir;
    k_horizon<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d, rows, cols, ndir, offs, thr, windcoef ? 1 : 0, out);
    return cudaGetLastError();
}
cudaError_t launch_skyview(const double* hor, int64_t nc, int ndir, double* svf, cudaStream_t st) {
    k_skyview<<<(unsigned)((nc + 255) / 256), 256, 0, st>>>(hor, nc, ndir, svf);
    return cudaGetLastError();
}
cudaError_t launch_blend16to8(const double* a, int64_t nc, double* out, cudaStream_t st) {
    k_blend16to8<<<(unsigned)((nc * 8 + 255) / 256), 256, 0, st>>>(a, nc, out);
    return cudaGetLastError();
}

} // namespace mcf
