// microclimf_b200 — the pair build of the grid kernel (included by mcf_kernels.cu, inside namespace mcf).
//
// k_grid keeps the ~75 per-cell invariants of a vegetation layer (CellInv) in registers: 168 registers per thread + spills,
// 12 warps per SM, and a dependent FP64 chain per warp that the schedulers cannot cover (stall `wait` 42 %, FP64 pipe
// 46 % busy, profiles/r02_kgrid_headline10.txt).  This build trades registers for shared memory and warps:
//   * the invariants live in SHARED memory, [field][cell] (a warp's 32 cells are 256 contiguous bytes per field:
//     conflict-free), written once per tile and layer by cell_setup and read where they are used through the proxy type
//     CellInvS (mcf_physics.cuh), which the physics templates take in place of the register struct;
//   * TWO threads per cell: the hours of a day are independent of each other within each of the reference's two passes
//     (ref src/microclimfCpp.cpp:2214-2262, :2264-2305), so the thread of half p takes the hours hr = p, p + 2, ... of BOTH
//     passes — its day stash stays private — and the only exchange per cell-day is the daily reduction (Rmx, tmx, tmn;
//     ref :2196-2263) between the two partner WARPS: three doubles through an L2-resident scratch and one 64-thread
//     named barrier (bar.sync 1 + warp pair, 64);
//   * 256 cells x 2 = 512 threads per CTA at 128 registers: 16 warps per SM (measured best: 320 cells / 96 registers and
//     192 cells / 168 registers are 10 % and 7 % slower);
//   * the two math tables of mcf_math.cuh as conflict-free shared-memory replicas (TAB = 1): with 196 KB of the SM carved
//     out for shared memory the __ldg gathers of the tables miss what is left of the L1.
// Same tile counter, TMA-fed hour-table ring (full / empty mbarriers), stash policy and output layout as k_grid.
// Serves the per-hour-table drivers (modes 1/3), reqhgt >= 0, FP64 and packed sinks: the headline path.  Everything else
// runs k_grid: the reducing sinks need the shared memory for their accumulators, array climate needs the registers for
// its per-thread hour record (-DMCF_PAIR_ARR=1 builds it here too: measured slower), below ground has no canopy physics.
// FP64 pipe 54 % busy, 1.54e10 cell-hours/s per B200 (k_grid: 1.32e10); profiles/r02_kpair_headline10.txt.

#ifndef MCF_PAIR_CELLS
#define MCF_PAIR_CELLS 256
#endif
#ifndef MCF_PAIR_STAGES
#define MCF_PAIR_STAGES 4
#endif
constexpr int kPairCells = MCF_PAIR_CELLS;  // cells per tile
constexpr int kPairThreads = 2 * kPairCells;
constexpr int kPairStages = MCF_PAIR_STAGES;
#ifndef MCF_PAIR_UNROLL1
#define MCF_PAIR_UNROLL1 1
#endif
constexpr int kPairUnroll1 = MCF_PAIR_UNROLL1; // hours of pass 1 in flight per thread (2: spills at 128 registers)
static_assert(kPairCells % 32 == 0 && kPairCells / 32 <= 15, "one named barrier per warp pair (ids 1..15)");

// dynamic shared memory of the pair kernel
constexpr size_t kPairRingBytes = sizeof(HourRec) * 24 * kPairStages;
constexpr size_t kPairInvDBytes = sizeof(double) * 2 * ((kInvD + 1) / 2) * kPairCells; // [field][cell] (MCF_INV_PAIRED: [field pair][cell][2])
constexpr size_t kPairInvIBytes = sizeof(int) * kPairCells; // the three small integers of CellInv packed into one word
constexpr int kPairTab = MathTab<CellInvS<kPairCells>>::value;
constexpr size_t kPairTabBytes = kPairTab ? kMathSmemBytes : 0; // math-table replicas, at offset 0 (mcf_math.cuh, TAB = 1)
constexpr size_t kPairSmemBytes =
    kPairTabBytes + kPairRingBytes + kPairInvDBytes + kPairInvIBytes + 2 * kPairStages * sizeof(uint64_t) + 16;
static_assert(kPairSmemBytes <= 232448, "pair kernel: shared memory beyond 227 KB");
// per-CTA global scratch: the day stash [24][kStashVars][kPairCells] and the reduction exchange [2 days][2 halves][3][kPairCells]
constexpr size_t kPairStashDoubles = (size_t)24 * kStashVars * kPairCells;
constexpr size_t kPairXchDoubles = (size_t)2 * 2 * 3 * kPairCells;
constexpr size_t kPairScratchDoubles = kPairStashDoubles + kPairXchDoubles;

__device__ __forceinline__ void pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ double ld_sector(const double* p) { // streamed once: keep it out of the (small) L1
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_xch(double* p, double v) { asm volatile("st.global.cg.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ double ld_xch(const double* p) {
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

template <int ARR, int RQ, int SINK, bool ALLOUT>
__global__ void __launch_bounds__(kPairThreads, 1) k_grid_pair(const __grid_constant__ GridArgs a) {
    static_assert(RQ != RQ_BELOW && (SINK == SINK_F64 || SINK == SINK_PACK), "pair kernel: hourly sinks, reqhgt >= 0");
    unsigned char* const pair_smem = mcf_dyn_smem + kPairTabBytes;
    HourRec(*const slab_ring)[24] = reinterpret_cast<HourRec(*)[24]>(pair_smem);
    double* const inv_d = reinterpret_cast<double*>(pair_smem + kPairRingBytes);
    int* const inv_i = reinterpret_cast<int*>(pair_smem + kPairRingBytes + kPairInvDBytes);
    uint64_t* const full_bar = reinterpret_cast<uint64_t*>(pair_smem + kPairRingBytes + kPairInvDBytes + kPairInvIBytes);
    uint64_t* const empty_bar = full_bar + kPairStages;
    int* const s_tile = reinterpret_cast<int*>(empty_bar + kPairStages);
    constexpr int kStages = kPairStages, kAhead = kPairStages / 2;
    constexpr bool PACK = (SINK == SINK_PACK);

    const int tid = threadIdx.x;
    const int half = (tid >= kPairCells) ? 1 : 0; // uniform per warp
    const int ci = tid - half * kPairCells;       // this thread's cell of the tile
    const int bar_id = 1 + (ci >> 5);
    const int ntiles = (a.cell_end - a.cell_begin + kPairCells - 1) / kPairCells;
    double* const scratch = a.stash + (size_t)blockIdx.x * kPairScratchDoubles;
    double* const stash = scratch + ci;
    double* const xch = scratch + kPairStashDoubles + ci;
    const CellInvS<kPairCells> v(inv_d + kInvCellStep * ci, inv_i + ci);
    unsigned int q0 = 0;

    if (!ARR && tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kPairThreads / 32);
        }
        mbar_fence_init();
    }
    if (kPairTab) math_tables_to_smem(); // ordered before their first use by the tile loop's __syncthreads
    const uint32_t om = a.outmask;
    const double NA = na_real();
    const int zero = a.pack >> 8; // 0, but not to the compiler (settle)

    auto issue_fill = [&](unsigned int q, int bi) {
        const int s = (int)(q % kStages);
        const unsigned int fill = q / kStages;
        if (fill > 0) mbar_wait(&empty_bar[s], (fill - 1) & 1u);
        const DayBlock nb = a.blocks[a.block0 + bi];
        mbar_expect_tx(&full_bar[s], 24 * sizeof(HourRec));
        tma_load_1d(&slab_ring[s][0], a.hours + nb.k0, 24 * sizeof(HourRec), &full_bar[s]);
    };

    for (;;) {
        __syncthreads();
        if (tid == 0) *s_tile = (int)atomicAdd(a.tile_counter, 1u);
        __syncthreads();
        const int tile = *s_tile;
        if (tile >= ntiles) break;
        const int cell = a.cell_begin + tile * kPairCells + ci;
        const bool valid = cell < a.cell_end;
        const int cc = valid ? cell : a.cell_end - 1;
        const bool active = valid && !isnan(__ldg(&a.veg[0][cc])); // ref :2182-2183, :2765-2766
        const unsigned amask = __ballot_sync(0xffffffffu, active);
        double lat = a.lat, lon = 0.0, dTmx = -0.6273 * a.dscal[0] + 49.79;
        double sl = 0.0, cl = 1.0; // sin / cos of the cell's latitude (modes 2/4)
        CoarseCell ccell;
        if (ARR == 2) coarse_setup(a, cc, ccell);
        if (ARR) {
            lat = __ldg(&a.lats[cc]);
            lon = __ldg(&a.lons[cc]);
            dTmx = -0.6273 * __ldg(&a.mxtc_cell[cc]) + 49.79;
            sincos(lat * kPi / 180.0, &sl, &cl);
        }
        int cur_lyr = -1;

        if (!ARR && tid == 0)
            for (int bi = 0; bi < kAhead && bi < a.nblocks; ++bi) issue_fill(q0 + bi, bi);

        for (int bi = 0; bi < a.nblocks; ++bi) {
            const DayBlock blk = a.blocks[a.block0 + bi];
            const unsigned int q = q0 + bi;
            const int buf = (int)(q % kStages);
            const HourRec* const slab_day = &slab_ring[buf][0];
            if (!ARR) {
                if (tid == 0 && bi + kAhead < a.nblocks) issue_fill(q + kAhead, bi + kAhead);
                mbar_wait(&full_bar[buf], (q / kStages) & 1u);
            }

            if (blk.lyr != cur_lyr) { // uniform over the CTA: the work list is the launch's
                if (cur_lyr >= 0) pair_sync(bar_id); // the partner is done with the previous layer's invariants
                cur_lyr = blk.lyr;
                if (half == 0) {
                    const double tmean = a.has_tadd_mean ? a.tadd_mean : a.dscal[1] / a.dscal[2];
                    const double tadd = log(__ldg(&a.soil[11][cc])) / a.tfact - tmean;
                    CellIn cin;
                    CellInv vr;
                    load_cell(a, cc, cur_lyr, tadd, cin);
                    cell_setup(cin, a.reqhgt2, a.zref, lat, vr);
                    CellInvS<kPairCells>::store(vr, inv_d + kInvCellStep * ci, inv_i + ci);
                }
                pair_sync(bar_id);
            }

            const long long slot0 = ((long long)blk.k0 - a.hour0) % a.ring_hours;
            const size_t ocell = (size_t)(cell - a.out_cell0);
            const size_t o_first = (size_t)slot0 * a.out_stride + ocell;
            const size_t o_unwrap = (size_t)a.ring_hours * a.out_stride; // subtracted once the slot has wrapped
            const int wrap_at = (int)(a.ring_hours - slot0 > 24 ? 24 : a.ring_hours - slot0);
            double Rmx = -999.9, tmx = -999.0, tmn = 999.0;

            if (!active) {
                if (valid) {
                    for (int hr = half; hr < 24; hr += 2) {
                        size_t o = o_first + (size_t)hr * a.out_stride;
                        if (hr >= wrap_at) o -= o_unwrap;
#pragma unroll
                        for (int qq = 0; qq < kNOut; ++qq)
                            if (ALLOUT || (om & (1u << qq))) {
                                if (PACK) reinterpret_cast<int16_t*>(a.out[qq])[o] = (int16_t)-9999;
                                else __stcs(&a.out[qq][o], NA);
                            }
                    }
                }
            } else {
                // ------------------------------------------------------------------ pass 1 (this half's hours)
                double ws_n = 0.0, ha_n = 0.0;
                if (!ARR) {
                    ws_n = ld_sector(&a.wsa[(size_t)slab_day[half].windex * a.ncells + cell]);
                    ha_n = ld_sector(&a.hor[(size_t)slab_day[half].sindex * a.ncells + cell]);
                    ws_n = settle(ws_n, zero), ha_n = settle(ha_n, zero); // as for the stash loads of pass 2, below
                }
#pragma unroll(kPairUnroll1)
                for (int hr = half; hr < 24; hr += 2) {
                    HourRec hloc;
                    if (ARR) hour_from_arrays<ARR>(a, blk.k0 + hr, cell, sl, cl, lon, true, ccell, hloc);
                    const HourRec& h = ARR ? hloc : slab_day[hr];
                    size_t o = o_first + (size_t)hr * a.out_stride;
                    if (hr >= wrap_at) o -= o_unwrap;
                    double ws, ha;
                    if (ARR) {
                        ws = ld_sector(&a.wsa[(size_t)h.windex * a.ncells + cell]);
                        ha = ld_sector(&a.hor[(size_t)h.sindex * a.ncells + cell]);
                    } else {
                        ws = ws_n, ha = ha_n;
                        const HourRec& hn = slab_day[hr < 22 ? hr + 2 : hr];
                        ws_n = ld_sector(&a.wsa[(size_t)hn.windex * a.ncells + cell]);
                        ha_n = ld_sector(&a.hor[(size_t)hn.sindex * a.ncells + cell]);
                    }
                    // terrain-adjusted solar index with horizon shading (ref :2218-2223 / :2499-2504)
                    double si;
                    if (ARR && h.zend > 90.0) si = 0.0; // shadowmask = false in modes 2/4
                    else si = h.cosz * v.cs + h.sinz * (h.cosazi * v.ssca + h.sinazi * v.sssa);
                    if (si < 0.0) si = 0.0;
                    if (ha > h.tan_sa) si = 0.0;
                    const double soild = soil_distribute(v, h.soilmp);
                    if (ALLOUT || (om & (1u << 3))) put<3, SINK>(a, o, soild, nullptr);
                    Rad r;
                    if (h.Rsw > 0.0) {
                        r = shortwave(v, h, si);
                    } else {
                        r.radGsw = 0.0; r.radCsw = 0.0; r.Rbdown = 0.0; r.Rddown = 0.0; r.Rdup = 0.0; r.Lhalf = 0.0;
                    }
                    if (ALLOUT || (om & (1u << 5))) put<5, SINK>(a, o, r.Rbdown, nullptr);
                    if (ALLOUT || (om & (1u << 6))) put<6, SINK>(a, o, r.Rddown, nullptr);
                    if (ALLOUT || (om & (1u << 8))) put<8, SINK>(a, o, r.Rdup, nullptr);
                    // longwave absorbed by the ground (ref :1165-1175)
                    double radGlw;
                    if (v.pai > 0.0) radGlw = kL.em * (v.trdif * v.svfa * h.Rlw + (1.0 - v.trdif) * h.Rem);
                    else radGlw = kL.em * v.svfa * h.Rlw;
                    const Wind w = wind_hour(v, h.u2, h.umu, ws);
                    if (ALLOUT || (om & (1u << 4))) put<4, SINK>(a, o, w.uz, nullptr);
                    // ground surface temperature with G = 0 (ref soiltempG0 :1262-1275)
                    const double radabs = r.radGsw + radGlw;
                    const double matric = -v.psie_abs * mexp_nc<kPairTab>(-v.soilb * mlog<kPairTab>(soild * v.inv_Smax));
                    double surfwet = mexp_lo<kPairTab>((kL.wet_a * matric) * h.invRT);
                    if (surfwet > 1.0) surfwet = 1.0;
                    double m_unused;
                    const double Tg0 = pm_ts(h, dTmx, radabs, w.gHa, w.gHa, 0.0, surfwet, m_unused);
                    const double Rnet = radabs - kL.emsb * radem4(Tg0);
                    const double Rval = fabs(Rnet);
                    if (Rmx < Rval) Rmx = Rval;
                    if (tmx < Tg0) tmx = Tg0;
                    if (tmn > Tg0) tmn = Tg0;
                    double* st = stash + (size_t)hr * (kStashVars * kPairCells);
                    st_stash(&st[0 * kPairCells], radabs);
                    st_stash(&st[1 * kPairCells], surfwet);
                    st_stash(&st[2 * kPairCells], r.radCsw);
                    st_stash(&st[3 * kPairCells], r.Lhalf);
                    st_stash(&st[4 * kPairCells], soild);
                    st_stash(&st[5 * kPairCells], w.uf);
                }
            }
            // ---------------------------------------------------------------------- daily reduction across the pair
            {
                double* mine = xch + (size_t)(((q & 1u) * 2 + half) * 3) * kPairCells;
                st_xch(&mine[0], Rmx);
                st_xch(&mine[kPairCells], tmx);
                st_xch(&mine[2 * kPairCells], tmn);
                pair_sync(bar_id);
                const double* theirs = xch + (size_t)(((q & 1u) * 2 + (half ^ 1)) * 3) * kPairCells;
                const double r2 = ld_xch(&theirs[0]), x2 = ld_xch(&theirs[kPairCells]), n2 = ld_xch(&theirs[2 * kPairCells]);
                if (Rmx < r2) Rmx = r2;
                if (tmx < x2) tmx = x2;
                if (tmn > n2) tmn = n2;
            }
            if (active) {
                // ------------------------------------------------------------------ pass 2 (backwards: LIFO stash)
                const double dtr = tmx - tmn;
                const int hlast = 22 + half;
                const double* st0 = stash + (size_t)hlast * (kStashVars * kPairCells);
                __syncwarp(amask);
                double radabs_n = ld_stash(&st0[0 * kPairCells]), surfwet_n = ld_stash(&st0[1 * kPairCells]);
                double radCsw_n = ld_stash(&st0[2 * kPairCells]), Lhalf_n = ld_stash(&st0[3 * kPairCells]);
                double soild_n = ld_stash(&st0[4 * kPairCells]), uf_n = ld_stash(&st0[5 * kPairCells]);
                // ptxas gives these loads and the loop's look-ahead loads the same scoreboard, and the loop's first
                // consumer waits on it for the entry path — i.e., every iteration, on the look-ahead loads it has just
                // issued (5 % of all stall samples, profiles/r02_kpair_v2).  Consuming the values here retires the
                // scoreboard before the loop is entered.
                radabs_n = settle(radabs_n, zero), surfwet_n = settle(surfwet_n, zero), radCsw_n = settle(radCsw_n, zero);
                Lhalf_n = settle(Lhalf_n, zero), soild_n = settle(soild_n, zero), uf_n = settle(uf_n, zero);
#pragma unroll 1
                for (int hr = hlast; hr >= 0; hr -= 2) {
                    HourRec hloc;
                    if (ARR) hour_from_arrays<ARR>(a, blk.k0 + hr, cell, sl, cl, lon, false, ccell, hloc);
                    const HourRec& h = ARR ? hloc : slab_day[hr];
                    size_t o = o_first + (size_t)hr * a.out_stride;
                    if (hr >= wrap_at) o -= o_unwrap;
                    const double radabs = radabs_n, surfwet = surfwet_n, radCsw = radCsw_n, Lhalf = Lhalf_n;
                    const double soild = soild_n;
                    Wind w;
                    w.uf = uf_n;
                    {
                        // the next (earlier) hour of this half, one iteration ahead; the first hour re-reads itself
                        const double* st = stash + (size_t)(hr >= 2 ? hr - 2 : hr) * (kStashVars * kPairCells);
                        __syncwarp(amask); // one converged warp load before a lane discards the line (see k_grid)
                        radabs_n = ld_stash(&st[0 * kPairCells]);
                        surfwet_n = ld_stash(&st[1 * kPairCells]);
                        radCsw_n = ld_stash(&st[2 * kPairCells]);
                        Lhalf_n = ld_stash(&st[3 * kPairCells]);
                        soild_n = ld_stash(&st[4 * kPairCells]);
                        uf_n = ld_stash(&st[5 * kPairCells]);
                        if ((tid & 15) == 0) {
                            const double* sd = stash + (size_t)hr * (kStashVars * kPairCells);
                            discard_line(&sd[0 * kPairCells], radabs);
                            discard_line(&sd[1 * kPairCells], surfwet);
                            discard_line(&sd[2 * kPairCells], radCsw);
                            discard_line(&sd[3 * kPairCells], Lhalf);
                            discard_line(&sd[4 * kPairCells], soild);
                            discard_line(&sd[5 * kPairCells], w.uf);
                        }
                    }
                    w.uz = w.uf * v.uz_coef;
                    if (w.uz > h.u2) w.uz = h.u2;
                    w.gHa = w.uf * v.gHa_coef;
                    if (w.gHa < kL.gha_lo) w.gHa = kL.gha_lo;
                    // soil conductivity and damping depth (ref soilcondCpp :1249-1260)
                    const double rho = v.rho;
                    const double cs = (v.cs0 + 4180.0 * soild);
                    const double ph = (rho * (1.0 - soild) + soild) * 1000.0;
                    const double c2 = kL.c2_a * rho * soild;
                    const double c1 = v.c1;
                    const double kcon = c1 + c2 * soild - v.c14 * mexp_lo<kPairTab>(-pow4(v.c3 * soild));
                    const double kap = mdiv(kcon, cs * ph);
                    const double iDD = mrsqrt(kap * kL.two_omdy);
                    // ground heat flux scaled from the point model (ref soiltemp_hrCpp :1277-1296)
                    const double dtR = dtr * h.inv_dtrp;
                    const double Gmu = dtR * (kcon * h.muGp_kp) * iDD;
                    double G = h.Gp * Gmu;
                    if (G > kL.g_cap * Rmx) G = kL.g_cap * Rmx;
                    if (G < -kL.g_cap * Rmx) G = -kL.g_cap * Rmx;
                    double m_unused;
                    const double Tg = pm_ts(h, dTmx, radabs, w.gHa, w.gHa, G, surfwet, m_unused);
                    const double radClw = kL.em * v.svfa * h.Rlw;
                    const Above tv = above_ground(v, h, dTmx, soild, Tg, G, w, radCsw, radClw, Lhalf);
                    if (ALLOUT || (om & (1u << 0))) put<0, SINK>(a, o, (RQ == RQ_ABOVE) ? tv.Tz : Tg, nullptr);
                    if (ALLOUT || (om & (1u << 7))) put<7, SINK>(a, o, tv.lwdn, nullptr);
                    if (ALLOUT || (om & (1u << 9))) put<9, SINK>(a, o, tv.lwup, nullptr);
                    if (RQ == RQ_ABOVE) {
                        if (ALLOUT || (om & (1u << 1))) put<1, SINK>(a, o, tv.tleaf, nullptr);
                        if (ALLOUT || (om & (1u << 2))) put<2, SINK>(a, o, tv.rh, nullptr);
                    }
                }
            }
            if (!ARR) {
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(&empty_bar[buf]);
            }
        }
        q0 += (unsigned int)a.nblocks;
    }
}

template <int ARR, int RQ, int SINK, bool ALLOUT>
static cudaError_t launch_pair_t(const GridArgs& a, int grid, cudaStream_t stream) {
    // the opt-in is per device and per function: remembered per (instantiation, device) so that a process that moves
    // between devices (mcf_set_device) configures each of them
    static bool configured[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        e = cudaFuncSetAttribute(k_grid_pair<ARR, RQ, SINK, ALLOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPairSmemBytes);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    k_grid_pair<ARR, RQ, SINK, ALLOUT><<<grid, kPairThreads, kPairSmemBytes, stream>>>(a);
    return cudaGetLastError();
}
int pair_tile() { return kPairCells; }
size_t pair_scratch_doubles() { return kPairScratchDoubles; }
#ifndef MCF_PAIR_ARR
#define MCF_PAIR_ARR 0 // 1: array climate (modes 2/4) also runs the pair build
#endif
bool pair_eligible(int arr, int rq, int sink) {
    return (arr == 0 || MCF_PAIR_ARR) && rq != RQ_BELOW && (sink == SINK_F64 || sink == SINK_PACK);
}
template <int ARR>
static cudaError_t launch_pair_arr(const GridArgs& a, int rq, int grid, cudaStream_t stream, int sink) {
    if (rq == RQ_ABOVE) {
        if (ARR == 0 && sink == SINK_F64 && a.outmask == 0x3FFu) return launch_pair_t<0, RQ_ABOVE, SINK_F64, true>(a, grid, stream);
        if (sink == SINK_F64) return launch_pair_t<ARR, RQ_ABOVE, SINK_F64, false>(a, grid, stream);
        return launch_pair_t<ARR, RQ_ABOVE, SINK_PACK, false>(a, grid, stream);
    }
    if (sink == SINK_F64) return launch_pair_t<ARR, RQ_SURFACE, SINK_F64, false>(a, grid, stream);
    return launch_pair_t<ARR, RQ_SURFACE, SINK_PACK, false>(a, grid, stream);
}
cudaError_t launch_grid_pair(const GridArgs& a, int arr, int rq, int grid, cudaStream_t stream, int sink) {
    if (sink < 0) sink = (a.pack == 1) ? SINK_PACK : SINK_F64;
#if MCF_PAIR_ARR
    if (arr == 1) return launch_pair_arr<1>(a, rq, grid, stream, sink);
    if (arr == 2) return launch_pair_arr<2>(a, rq, grid, stream, sink);
#endif
    if (arr != 0) return cudaErrorInvalidValue;
    return launch_pair_arr<0>(a, rq, grid, stream, sink);
}
