// microclimf_b200 — the FP32 build of the grid kernel (included at the end of mcf_kernels.cu, inside namespace mcf).
//
// BASELINE north_star: "an optional FP32 build must stay within 0.05 degC and 0.5 % radiation".  Same execution model as
// k_grid (persistent 384-thread CTAs, one thread per cell, TMA-fed hour-table ring, two 24-hour passes per cell-day), for
// all four drivers and every height.  What differs:
//   * the hour loops run in FP32 with SFU transcendentals (mcf_physics_f32.cuh, generated from the FP64 physics);
//   * per-cell invariants are still computed in FP64 (cell_setup: two-stream solution, logarithms) and narrowed once;
//   * static inputs stay FP64 in HBM (they are R's arrays), the hour table is narrowed once per launch (k_narrow_hours),
//     the day stash and the outputs are FP32: 40 algorithmic bytes per cell-hour instead of 80;
//   * array climate (modes 2/4, ARR = 1 fine arrays / 2 coarse grid): the hour record of a cell-hour is assembled in FP64
//     exactly as k_grid assembles it (hour_from_arrays: interpolation, altitude correction, solar position are
//     cancellation-prone) and narrowed; the physics that consumes it is FP32;
//   * below ground (RQ_BELOW) this kernel is not used: the time-axis pass is DISCONTINUOUS in its inputs (a rolling mean
//     over n = round(-118.35 z / mean damping depth) hours, ref Tbelowgroundv :1481-1482; a ratio of daily ranges in the
//     incomplete branch, :1511) and single-precision ground temperatures / damping depths flip n or move the ratio in a
//     few cells per thousand by 0.1-0.25 degC (measured, tools/diag_f32_below.py).  The FP32 build therefore runs the
//     FP64 hour loops there (a pass that has no canopy physics and is cheap) and narrows Tz and soil moisture to FP32.
// (mcf_physics_f32.cuh is included at the top of mcf_kernels.cu)

#ifndef MCF_F32_TILE
#define MCF_F32_TILE 512
#endif
constexpr int kTileF = MCF_F32_TILE; // cells per CTA tile of the FP32 kernel (16 warps at 128 registers)
#ifndef MCF_F32_MINB
#define MCF_F32_MINB 1
#endif
struct GridArgsF {
    GridArgs g;                 // everything of the FP64 launch (outputs unused)
    const f32::HourRecF* hoursf; // [tsteps]
    float* outf[kNOut];
    float* stashf;              // [gridDim.x][24][kStashVars][kTileF]
};

__global__ void k_narrow_hours(const HourRec* __restrict__ in, int n, f32::HourRecF* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    f32::HourRecF h;
    f32::narrow(in[k], h);
    out[k] = h;
}

__device__ __forceinline__ void st_stash_f(float* p, float v) {
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(stash_policy()) : "memory");
}
__device__ __forceinline__ float ld_stash_f(const float* p) {
    float v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(stash_policy()));
    return v;
}
__device__ __forceinline__ void discard_line_f(const float* p, float loaded) {
    asm volatile("discard.global.L2 [%0], 128; // after %1" ::"l"(p), "f"(loaded) : "memory");
}

template <int ARR, int RQ>
__global__ void __launch_bounds__(kTileF, MCF_F32_MINB) k_grid_f32(const __grid_constant__ GridArgsF af) {
    using f32::HourRecF;
    const GridArgs& a = af.g;
    constexpr int kStages = 4, kAhead = 2;
    __shared__ __align__(128) HourRecF slab_ring[kStages][24];
    __shared__ __align__(8) uint64_t full_bar[kStages];
    __shared__ __align__(8) uint64_t empty_bar[kStages];
    __shared__ int s_tile;
    const int tid = threadIdx.x;
    const int ntiles = (a.cell_end - a.cell_begin + kTileF - 1) / kTileF;
    float* const stash = af.stashf + (size_t)blockIdx.x * (24 * kStashVars * kTileF) + tid;
    unsigned int q0 = 0;
    if (!ARR && tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kTileF / 32);
        }
        mbar_fence_init();
    }
    const uint32_t om = a.outmask;
    const float NA = __int_as_float(0x7FC00000); // quiet NaN: FP32 has no NA payload convention
    auto issue_fill = [&](unsigned int q, int bi) {
        const int s = (int)(q % kStages);
        const unsigned int fill = q / kStages;
        if (fill > 0) mbar_wait(&empty_bar[s], (fill - 1) & 1u);
        const DayBlock nb = a.blocks[a.block0 + bi];
        mbar_expect_tx(&full_bar[s], 24 * sizeof(HourRecF));
        tma_load_1d(&slab_ring[s][0], af.hoursf + nb.k0, 24 * sizeof(HourRecF), &full_bar[s]);
    };
    for (;;) {
        __syncthreads();
        if (tid == 0) s_tile = (int)atomicAdd(a.tile_counter, 1u);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= ntiles) break;
        const int cell = a.cell_begin + tile * kTileF + tid;
        const bool valid = cell < a.cell_end;
        const int cc = valid ? cell : a.cell_end - 1;
        const bool active = valid && !isnan(__ldg(&a.veg[0][cc]));
        const unsigned amask = __ballot_sync(0xffffffffu, active); // lanes taking the solving branch (converged here)
        const double tmean = a.has_tadd_mean ? a.tadd_mean : a.dscal[1] / a.dscal[2];
        const double tadd = log(__ldg(&a.soil[11][cc])) / a.tfact - tmean;
        double lat = a.lat, lon = 0.0, sl = 0.0, cl = 1.0;
        float dTmx = (float)(-0.6273 * a.dscal[0] + 49.79);
        CoarseCell ccell;
        if (ARR == 2) coarse_setup(a, cc, ccell);
        if (ARR) { // modes 2/4: per-cell latitude / longitude and series maximum of tc (ref :2458, :2467-2471)
            lat = __ldg(&a.lats[cc]);
            lon = __ldg(&a.lons[cc]);
            dTmx = (float)(-0.6273 * __ldg(&a.mxtc_cell[cc]) + 49.79);
            sincos(lat * kPi / 180.0, &sl, &cl);
        }
        f32::CellInvF v;
        int cur_lyr = -1;
        double ddsum = 0.0;
        if (!ARR && tid == 0)
            for (int bi = 0; bi < kAhead && bi < a.nblocks; ++bi) issue_fill(q0 + bi, bi);
        for (int bi = 0; bi < a.nblocks; ++bi) {
            const DayBlock blk = a.blocks[a.block0 + bi];
            const unsigned int q = q0 + bi;
            const int buf = (int)(q % kStages);
            const HourRecF* const slab_day = &slab_ring[buf][0];
            if (!ARR) {
                if (tid == 0 && bi + kAhead < a.nblocks) issue_fill(q + kAhead, bi + kAhead);
                mbar_wait(&full_bar[buf], (q / kStages) & 1u);
            }
            const long long slot0 = ((long long)blk.k0 - a.hour0) % a.ring_hours;
            if (blk.lyr != cur_lyr) {
                cur_lyr = blk.lyr;
                CellIn ci;
                load_cell(a, cc, cur_lyr, tadd, ci);
                CellInv v64;
                cell_setup(ci, a.reqhgt2, a.zref, lat, v64); // FP64: once per cell and layer
                f32::narrow(v64, v);
            }
            if (!active) {
                if (valid) {
                    for (int hr = 0; hr < 24; ++hr) {
                        long long slot = slot0 + hr;
                        if (slot >= a.ring_hours) slot -= a.ring_hours;
                        const size_t o = (size_t)slot * a.ncells + cell;
#pragma unroll
                        for (int qq = 0; qq < kNOut; ++qq)
                            if (om & (1u << qq)) __stcs(&af.outf[qq][o], NA);
                    }
                }
            } else {
                const size_t o_first = (size_t)slot0 * a.ncells + cell;
                const long long wrap_at = a.ring_hours - slot0;
                float Rmx = -999.9f, tmx = -999.0f, tmn = 999.0f;
                float ws_n = 0.0f, ha_n = 0.0f;
                if (!ARR) {
                    ws_n = (float)__ldg(&a.wsa[(size_t)slab_day[0].windex * a.ncells + cell]);
                    ha_n = (float)__ldg(&a.hor[(size_t)slab_day[0].sindex * a.ncells + cell]);
                }
                size_t o = o_first;
#pragma unroll 1
                for (int hr = 0; hr < 24; ++hr) {
                    HourRecF hloc;
                    if (ARR) {
                        HourRec h64{};
                        hour_from_arrays<ARR>(a, blk.k0 + hr, cell, sl, cl, lon, true, ccell, h64);
                        f32::narrow(h64, hloc);
                    }
                    const HourRecF& h = ARR ? hloc : slab_day[hr];
                    if (hr == wrap_at) o = cell;
                    float ws, ha;
                    if (ARR) {
                        ws = (float)__ldg(&a.wsa[(size_t)h.windex * a.ncells + cell]);
                        ha = (float)__ldg(&a.hor[(size_t)h.sindex * a.ncells + cell]);
                    } else {
                        ws = ws_n;
                        ha = ha_n;
                        const HourRecF& hn = slab_day[hr < 23 ? hr + 1 : 23];
                        ws_n = (float)__ldg(&a.wsa[(size_t)hn.windex * a.ncells + cell]);
                        ha_n = (float)__ldg(&a.hor[(size_t)hn.sindex * a.ncells + cell]);
                    }
                    float si;
                    if (ARR && h.zend > 90.0f) si = 0.0f; // shadowmask = false in modes 2/4
                    else si = h.cosz * v.cs + h.sinz * (h.cosazi * v.ssca + h.sinazi * v.sssa);
                    if (si < 0.0f) si = 0.0f;
                    if (ha > h.tan_sa) si = 0.0f;
                    const float soild = f32::soil_distribute(v, h.soilmp);
                    if (om & (1u << 3)) __stcs(&af.outf[3][o], soild);
                    f32::Rad r;
                    if (h.Rsw > 0.0f) {
                        r = f32::shortwave(v, h, si);
                    } else {
                        r.radGsw = 0.0f; r.radCsw = 0.0f; r.Rbdown = 0.0f; r.Rddown = 0.0f; r.Rdup = 0.0f; r.Lhalf = 0.0f;
                    }
                    if (om & (1u << 5)) __stcs(&af.outf[5][o], r.Rbdown);
                    if (om & (1u << 6)) __stcs(&af.outf[6][o], r.Rddown);
                    if (om & (1u << 8)) __stcs(&af.outf[8][o], r.Rdup);
                    float radGlw;
                    if (v.pai > 0.0f) radGlw = f32::kEm * (v.trdif * v.svfa * h.Rlw + (1.0f - v.trdif) * h.Rem);
                    else radGlw = f32::kEm * v.svfa * h.Rlw;
                    const f32::Wind w = f32::wind_hour(v, h.u2, h.umu, ws);
                    if (om & (1u << 4)) __stcs(&af.outf[4][o], w.uz);
                    const float radabs = r.radGsw + radGlw;
                    const float matric = -v.psie_abs * f32::fexp(-v.soilb * f32::flog(soild * v.inv_Smax));
                    float surfwet = f32::fexp((0.018f * matric) * h.invRT);
                    if (surfwet > 1.0f) surfwet = 1.0f;
                    float m_unused;
                    const float Tg0 = f32::pm_ts(h, dTmx, radabs, w.gHa, w.gHa, 0.0f, surfwet, m_unused);
                    const float Rnet = radabs - f32::kEm * f32::kSb * f32::radem4(Tg0);
                    const float Rval = fabsf(Rnet);
                    if (Rmx < Rval) Rmx = Rval;
                    if (tmx < Tg0) tmx = Tg0;
                    if (tmn > Tg0) tmn = Tg0;
                    float* st = stash + (size_t)hr * (kStashVars * kTileF);
                    st_stash_f(&st[0 * kTileF], radabs);
                    st_stash_f(&st[1 * kTileF], surfwet);
                    st_stash_f(&st[2 * kTileF], r.radCsw);
                    st_stash_f(&st[3 * kTileF], r.Lhalf);
                    st_stash_f(&st[4 * kTileF], soild);
                    st_stash_f(&st[5 * kTileF], w.uf);
                    o += a.ncells;
                }
                const float dtr = tmx - tmn;
                // pass 2 walks the day backwards (last-in-first-out on the stash) and drops each stash line from L2
                // after its only read, as k_grid does; a 128-byte line is one warp's 32 floats
                o = (23 >= wrap_at) ? (size_t)cell + (size_t)(23 - wrap_at) * a.ncells : o_first + (size_t)23 * a.ncells;
#pragma unroll 1
                for (int hr = 23; hr >= 0; --hr) {
                    HourRecF hloc;
                    if (ARR) {
                        HourRec h64{};
                        hour_from_arrays<ARR>(a, blk.k0 + hr, cell, sl, cl, lon, false, ccell, h64);
                        f32::narrow(h64, hloc);
                    }
                    const HourRecF& h = ARR ? hloc : slab_day[hr];
                    const float* st = stash + (size_t)hr * (kStashVars * kTileF);
                    __syncwarp(amask); // one converged load instruction per line: complete for its first lane = complete for all
                    const float radabs = ld_stash_f(&st[0 * kTileF]), surfwet = ld_stash_f(&st[1 * kTileF]);
                    const float radCsw = ld_stash_f(&st[2 * kTileF]), Lhalf = ld_stash_f(&st[3 * kTileF]);
                    const float soild = ld_stash_f(&st[4 * kTileF]);
                    f32::Wind w; // uz / gHa from the stashed friction velocity (ref windCpp :1199-1217)
                    w.uf = ld_stash_f(&st[5 * kTileF]);
                    if ((tid & 31) == 0) {
                        discard_line_f(&st[0 * kTileF], radabs);
                        discard_line_f(&st[1 * kTileF], surfwet);
                        discard_line_f(&st[2 * kTileF], radCsw);
                        discard_line_f(&st[3 * kTileF], Lhalf);
                        discard_line_f(&st[4 * kTileF], soild);
                        discard_line_f(&st[5 * kTileF], w.uf);
                    }
                    w.uz = w.uf * v.uz_coef;
                    if (w.uz > h.u2) w.uz = h.u2;
                    w.gHa = w.uf * v.gHa_coef;
                    if (w.gHa < 0.0001f) w.gHa = 0.0001f;
                    const float cs = (2400.0f * v.rho / 2.64f + 4180.0f * soild);
                    const float ph = (v.rho * (1.0f - soild) + soild) * 1000.0f;
                    const float c2 = 1.06f * v.rho * soild;
                    const float kcon = v.c1 + c2 * soild - (v.c1 - v.c4) * f32::fexp(-f32::pow4(v.c3 * soild));
                    const float kap = f32::fdiv(kcon, cs * ph);
                    const float kap2 = kap * (float)(2.0 / kOmdy);
                    float iDD;
                    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(iDD) : "f"(kap2));
                    const float dtR = dtr * h.inv_dtrp;
                    const float Gmu = dtR * (kcon * h.muGp_kp) * iDD;
                    float G = h.Gp * Gmu;
                    if (G > 0.6f * Rmx) G = 0.6f * Rmx;
                    if (G < -0.6f * Rmx) G = -0.6f * Rmx;
                    float m_unused;
                    const float Tg = f32::pm_ts(h, dTmx, radabs, w.gHa, w.gHa, G, surfwet, m_unused);
                    if (RQ == RQ_BELOW) {
                        // the time-axis pass runs in FP64 over the whole series: ground temperature and damping depth
                        a.tg_scratch[(size_t)(blk.k0 + hr) * (a.cell_end - a.cell_begin) + (cell - a.cell_begin)] = (double)Tg;
                        ddsum += (double)(kap2 * iDD); // sqrt(x) = x rsqrt(x)
                    } else {
                        const float radClw = f32::kEm * v.svfa * h.Rlw;
                        const f32::Above tv = f32::above_ground(v, h, dTmx, soild, Tg, G, w, radCsw, radClw, Lhalf);
                        if (om & (1u << 0)) __stcs(&af.outf[0][o], (RQ == RQ_ABOVE) ? tv.Tz : Tg);
                        if (om & (1u << 7)) __stcs(&af.outf[7][o], tv.lwdn);
                        if (om & (1u << 9)) __stcs(&af.outf[9][o], tv.lwup);
                        if (RQ == RQ_ABOVE) {
                            if (om & (1u << 1)) __stcs(&af.outf[1][o], tv.tleaf);
                            if (om & (1u << 2)) __stcs(&af.outf[2][o], tv.rh);
                        }
                    }
                    if (hr == wrap_at) o = (size_t)(a.ring_hours - 1) * a.ncells + cell; // back across the ring's seam
                    else o -= a.ncells;
                }
            }
            if (!ARR) {
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(&empty_bar[buf]);
            }
        }
        q0 += (unsigned int)a.nblocks;
        if (RQ == RQ_BELOW && active) a.dd_sum[cell - a.cell_begin] = ddsum;
    }
}

// FP64 -> FP32 (the below-ground Tz, which the time-axis pass produces in FP64); NA / NaN -> quiet NaN
__global__ void k_narrow32(const double* __restrict__ src, float* __restrict__ dst, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = (float)src[i];
}
cudaError_t launch_narrow32(const double* src, float* dst, int64_t n, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_narrow32<<<(int)blocks, 256, 0, stream>>>(src, dst, n);
    return cudaGetLastError();
}
__global__ void k_fill32(float* __restrict__ p, int64_t n) {
    const float NA = __int_as_float(0x7FC00000);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = NA;
}
cudaError_t launch_fill32(float* p, int64_t n, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_fill32<<<(int)blocks, 256, 0, stream>>>(p, n);
    return cudaGetLastError();
}

cudaError_t launch_narrow_hours(const HourRec* in, int n, void* out, cudaStream_t stream) {
    k_narrow_hours<<<(n + 255) / 256, 256, 0, stream>>>(in, n, (f32::HourRecF*)out);
    return cudaGetLastError();
}
size_t hourrec_f32_bytes() { return sizeof(f32::HourRecF); }
int f32_blocks_per_sm() { return MCF_F32_MINB; }
int f32_tile() { return kTileF; }

cudaError_t launch_grid_f32(const GridArgs& a, const void* hoursf, float* const outf[kNOut], float* stashf, int arr, int rq,
                            int grid, cudaStream_t stream) {
    GridArgsF af;
    af.g = a;
    af.hoursf = (const f32::HourRecF*)hoursf;
    for (int i = 0; i < kNOut; ++i) af.outf[i] = outf[i];
    af.stashf = stashf;
#define MCF_F32_RQ(ARR)                                                                          \
    do {                                                                                         \
        if (rq == RQ_ABOVE) k_grid_f32<ARR, RQ_ABOVE><<<grid, kTileF, 0, stream>>>(af);          \
        else if (rq == RQ_SURFACE) k_grid_f32<ARR, RQ_SURFACE><<<grid, kTileF, 0, stream>>>(af); \
        else return cudaErrorInvalidValue; /* below ground the FP32 build runs the FP64 hour loops (mcf_api.cu) */ \
    } while (0)
    if (arr == 0) MCF_F32_RQ(0);
    else if (arr == 1) MCF_F32_RQ(1);
    else MCF_F32_RQ(2);
#undef MCF_F32_RQ
    return cudaGetLastError();
}
