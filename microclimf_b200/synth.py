"""Seeded synthetic inputs for the grid solver (SURVEY.md §8d recipe).

There is no network and no R here, so benchmarks and parity tests run on synthetic rasters of the
shapes BASELINE.json names.  Everything is generated on the host with numpy from a fixed seed, so
the CUDA path, the CPU oracle and the compiled reference all see bit-identical inputs.

Ranges follow the reference's own data and defaults:
  * forcing: hourly series, T = 10 + 8 sin(season) + 5 sin(diurnal) deg C, clear-sky x 0.5 shortwave
    with a diffuse floor (as tests/testthat/test-microclimatemodel_wrapper.R:16-23), LW 300-380,
    wind >= 0.5 m/s (R/Cppwrappers.R:118), es/ea/tdew from the R formulas (R/internal.R:501-521);
  * point model columns (umu, kp, muGp, dtrp, G, soilm, Tg, Tbp) are smooth, mutually consistent
    series built from the same soil-conductivity formula the point model uses
    (src/microclimfCpp.cpp:5291-5301);
  * vegetation: ~10 % bare cells (hgt = pai = 0 together, R/internal.R:1101-1104), paia / leafden from
    the .foliageden gamma profile (R/internal.R:937-946), zref >= max(hgt);
  * soil: the 11 rows of the reference's bundled `soilparameters` table; ~2 % NA cells (NaN in hgt).
"""
from __future__ import annotations

import numpy as np

from .problem import GridProblem

# The 11 rows of the reference's bundled `soilparameters` table (data/soilparameters.rda, documented at
# R/data.R:70-94), columns Smax, Smin, b, psi_e, Vq, Vm, Mc, rho — read out of the .rda with a small XDR
# scan (Sand, Loamy sand, Sandy loam, Loam, Silt loam, Sandy clay loam, Clay loam, Silty clay loam,
# Sandy clay, Silty clay, Clay).
_SOIL_TABLE = np.column_stack([
    [0.399, 0.402, 0.403, 0.422, 0.447, 0.388, 0.419, 0.441, 0.381, 0.368, 0.394],      # Smax
    [0.049, 0.054, 0.058, 0.074, 0.067, 0.089, 0.091, 0.089, 0.103, 0.073, 0.073],      # Smin
    [1.7, 2.1, 3.1, 4.5, 4.7, 4.0, 5.2, 6.6, 6.0, 7.9, 7.6],                            # b
    [0.7, 0.9, 1.5, 1.1, 2.1, 2.8, 2.6, 3.3, 2.9, 3.4, 3.7],                            # psi_e
    [0.3, 0.24, 0.18, 0.12, 0.0, 0.14, 0.06, 0.04, 0.15, 0.0, 0.0],                     # Vq
    [0.3, 0.355, 0.41, 0.44, 0.47, 0.464, 0.509, 0.508, 0.4655, 0.624, 0.6],            # Vm
    [0.01, 0.035, 0.06, 0.0844, 0.124, 0.3648, 0.5422, 0.3948, 0.505, 0.55, 1.0],       # Mc
    [1.5978, 1.5871, 1.579, 1.5135, 1.3586, 1.6175, 1.5296, 1.4725, 1.6422, 1.6707, 1.6043],  # rho
])

TORAD = np.pi / 180.0


def _julday(year, month, day):
    """Astronomical Julian day, int arithmetic as src/microclimfCpp.cpp:28-37."""
    dd = day + 0.5
    madj = month + (month < 3) * 12
    yadj = year + (month < 3) * -1
    j = np.trunc(365.25 * (yadj + 4716)) + np.trunc(30.6001 * (madj + 1)) + dd - 1524.5
    c = yadj // 100  # integer division on ints (C++ int / int), yadj > 0
    b = 2 - c + c // 4
    return (j + (j > 2299160) * b).astype(np.int64)


def solar_zenith(lat, lon, year, month, day, hour):
    """Zenith angle in degrees (host helper for synthetic clear-sky radiation only;
    formula of solpositionCpp, src/microclimfCpp.cpp:48-57)."""
    jd = _julday(year, month, day)
    m = 6.24004077 + 0.01720197 * (jd - 2451545.0)
    eot = -7.659 * np.sin(m) + 9.863 * np.sin(2 * m + 3.5932)
    st = hour + (4.0 * lon + eot) / 60.0
    latr = lat * np.pi / 180.0
    tt = 0.261799 * (st - 12)
    dec = (np.pi * 23.5 / 180) * np.cos(2 * np.pi * ((jd - 159.5) / 365.25))
    coh = np.sin(dec) * np.sin(latr) + np.cos(dec) * np.cos(latr) * np.cos(tt)
    return np.degrees(np.arccos(np.clip(coh, -1.0, 1.0)))


def _satvap_r(tc):
    """R/internal.R:501-507 (.satvap)."""
    es = 0.61078 * np.exp(17.27 * tc / (tc + 237.3))
    ei = 0.61078 * np.exp(21.875 * tc / (tc + 265.5))
    return np.where(tc < 0, ei, es)


def _dewpoint_r(ea, tc):
    """R/internal.R:509-521 (.dewpoint)."""
    L = 2.501e6 - 2340 * tc
    tdew = 1 / (1 / 273.15 - (461.5 / L) * np.log(ea / 0.6112)) - 273.15
    tfrost = 1 / (1 / 273.15 - (461.5 / 2.834e6) * np.log(ea / 0.61078)) - 273.15
    return np.where(tdew < 0, tfrost, tdew)


def calendar(tsteps: int, year: int = 2023, start_doy: int = 0):
    """Hourly obstime columns starting at 00:00 on day-of-year `start_doy` (0-based)."""
    mdays = np.array([31, 29 if year % 4 == 0 else 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31])
    doy_month = np.repeat(np.arange(1, 13), mdays)
    doy_day = np.concatenate([np.arange(1, d + 1) for d in mdays])
    k = np.arange(tsteps)
    d = (start_doy + k // 24) % len(doy_month)
    return (np.full(tsteps, year, np.int32), doy_month[d].astype(np.int32), doy_day[d].astype(np.int32),
            (k % 24).astype(np.float64), d)


def forcing(tsteps: int, lat: float = 50.0, lon: float = -5.0, seed: int = 20240321, year: int = 2023,
            start_doy: int = 0, day_list=None):
    """Per-hour forcing + point-model columns (modes 1/3).  Returns dict of 1-D arrays keyed by ABI
    field names.  `day_list` (0-based days of year) selects specific days (bioclim's 14 days)."""
    rng = np.random.default_rng(seed)
    if day_list is not None:
        day_list = np.asarray(day_list)
        tsteps = 24 * len(day_list)
        yr, mo, dy, hr, _ = calendar(24 * 366 if year % 4 == 0 else 24 * 365, year, 0)
        idx = (day_list[:, None] * 24 + np.arange(24)[None, :]).ravel()
        yr, mo, dy, hr, doy = yr[idx], mo[idx], dy[idx], hr[idx], np.repeat(day_list, 24)
    else:
        yr, mo, dy, hr, doy = calendar(tsteps, year, start_doy)
    k = np.arange(tsteps)
    season = np.sin(2 * np.pi * (doy - 110) / 365.0)
    diurnal = np.sin((hr - 8) / 24 * 2 * np.pi)
    synoptic = np.repeat(rng.normal(0, 1.5, tsteps // 24 + 1), 24)[:tsteps]
    tc = 10 + 8 * season + 5 * diurnal + synoptic
    es = _satvap_r(tc)
    rh = np.clip(75 - 15 * diurnal + np.repeat(rng.normal(0, 5, tsteps // 24 + 1), 24)[:tsteps], 25, 100)
    ea = es * rh / 100
    tdew = _dewpoint_r(ea, tc)
    pk = 101.3 + np.repeat(rng.normal(0, 0.6, tsteps // 24 + 1), 24)[:tsteps]
    zen = solar_zenith(lat, lon, yr.astype(np.int64), mo.astype(np.int64), dy.astype(np.int64), hr)
    cz = np.cos(zen * TORAD)
    # clear-sky broadband (shape of clearskyradCpp, src/microclimfCpp.cpp:5220-5244), then clouds
    day = zen <= 90.0
    m = np.where(day, 35 * cz * (1224.0 * cz * cz + 1.0) ** -0.5, 0.0)
    od = (1.021 - 0.084 * np.sqrt(m * 0.00949 * pk + 0.051)) * (1 - 0.077 * (2.0 * m) ** 0.3) * 0.935 * m
    csr = np.where(day, 1352.778 * cz * od, 0.0).clip(min=0)
    cloud = np.repeat(rng.uniform(0.25, 1.0, tsteps // 24 + 1), 24)[:tsteps]
    sw = csr * cloud
    dif = sw * np.clip(1.15 - cloud, 0.3, 1.0)
    lw = 340 + 30 * season + 10 * (1 - cloud) * 4
    u2 = np.maximum(0.5, 3.5 + 2.5 * np.repeat(rng.normal(0, 1, tsteps // 24 + 1), 24)[:tsteps] + diurnal)
    wdir = np.mod(200 + np.cumsum(rng.normal(0, 12, tsteps)), 360.0)
    # point-model columns
    soilm = np.clip(0.30 - 0.10 * season + np.repeat(rng.normal(0, 0.01, tsteps // 24 + 1), 24)[:tsteps], 0.12, 0.40)
    rho, Vm, Vq, Mc = 1.53, 0.509, 0.06, 0.5422
    frs = Vm + Vq
    c1 = (0.57 + 1.73 * Vq + 0.93 * Vm) / (1.0 - 0.74 * Vq - 0.49 * Vm) - 2.8 * frs * (1.0 - frs)
    c3 = 1.0 + 2.6 * Mc ** -0.5
    c4 = 0.03 + 0.7 * frs * frs
    cs = 2400 * rho / 2.64 + 4180 * soilm
    ph = (rho * (1 - soilm) + soilm) * 1000
    kp = c1 + 1.06 * rho * soilm * soilm - (c1 - c4) * np.exp(-(c3 * soilm) ** 4)
    omdy = 2 * np.pi / (24 * 3600)
    mug = np.sqrt(2 * (kp / (cs * ph)) / omdy)
    t0p = tc + 0.012 * sw - 1.5
    Tg = tc + 0.010 * sw - 1.0
    dtrp = np.repeat((t0p.reshape(-1, 24).max(1) - t0p.reshape(-1, 24).min(1)), 24) if tsteps % 24 == 0 else \
        np.full(tsteps, 10.0)
    G = 0.12 * sw - 25 + 15 * diurnal
    umu = np.clip(1.0 + 0.15 * np.sin(k / 7.0) - 0.1 * (sw > 0), 0.6, 1.4)
    Tbp = np.repeat(Tg.reshape(-1, 24).mean(1), 24) if tsteps % 24 == 0 else Tg.copy()
    return dict(year=yr, month=mo, day=dy, hour=hr, temp=tc, es=es, ea=ea, tdew=tdew, pres=pk, swdown=sw,
                difrad=dif, lwdown=lw, windspeed=u2, winddir=wdir, p_soilm=soilm, p_Tg=Tg, p_Tbp=Tbp, p_G=G,
                p_umu=umu, p_kp=kp, p_muGp=mug, p_dtrp=dtrp)


def foliage_density(reqhgt, hgt, pai):
    """.foliageden (R/internal.R:937-946): gamma(shape 1.5, rate 1.5/7) profile over rescaled depth."""
    from scipy.stats import gamma

    shape, rate = 1.5, 1.5 / 7
    with np.errstate(divide="ignore", invalid="ignore"):
        x = ((hgt - reqhgt) / hgt) * 10
        td = gamma.cdf(10, shape, scale=1 / rate)
        rfd = gamma.pdf(x, shape, scale=1 / rate) / td
        leafden = (pai / hgt) * rfd * 10
        paia = gamma.cdf(x, shape, scale=1 / rate) * (pai / td)
    return paia, leafden


def static_layers(rows: int, cols: int, reqhgt: float, seed: int = 20240321, nlyr: int = 1, zref: float = 30.0,
                  bare_frac: float = 0.10, na_frac: float = 0.02, tall_frac: float = 0.35):
    """Vegetation + soil + terrain layers, dict of flat R-order arrays keyed by ABI field names."""
    rng = np.random.default_rng(seed + 1)
    nc = rows * cols
    ii, jj = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    # smooth DTM-like field -> slope/aspect with realistic spatial coherence
    fx, fy = 2 * np.pi / max(rows, 8), 2 * np.pi / max(cols, 8)
    dtm = 150 + 80 * np.sin(1.3 * fx * ii) * np.cos(0.9 * fy * jj) + 40 * np.sin(3.1 * fx * ii + 2.2 * fy * jj)
    gy = np.gradient(dtm, 10.0, axis=0) if rows > 1 else np.zeros_like(dtm)
    gx = np.gradient(dtm, 10.0, axis=1) if cols > 1 else np.zeros_like(dtm)
    slope = np.degrees(np.arctan(np.hypot(gx, gy))).clip(0, 40)
    aspect = np.mod(np.degrees(np.arctan2(-gx, gy)), 360.0)
    flat = rng.random((rows, cols)) < 0.03
    slope[flat] = 0.0
    F = lambda a: np.ascontiguousarray(a.ravel(order="F"), dtype=np.float64)
    out = {}
    # terrain-derived
    out["slope"], out["aspect"] = F(slope), F(aspect)
    out["twi"] = F(np.exp(rng.uniform(np.log(2), np.log(20), (rows, cols))))
    hor = rng.uniform(0, 0.6, (rows, cols, 24)) * (slope[..., None] / 40.0 + 0.15)
    out["hor"] = F(hor)
    out["svfa"] = F((0.5 * np.cos(2 * np.tan(np.arctan(hor).mean(axis=2))) + 0.5).clip(0.3, 1))
    out["wsa"] = F(rng.uniform(0.3, 1.0, (rows, cols, 8)))
    # soil
    st = rng.integers(0, len(_SOIL_TABLE), (rows, cols))
    for c, n in enumerate(("Smax", "Smin", "soilb", "Psie", "Vq", "Vm", "Mc", "rho")):
        out[n] = F(_SOIL_TABLE[st, c])
    out["gref"] = F(rng.uniform(0.1, 0.3, (rows, cols)))
    # vegetation (per layer: seasonal pai scaling)
    bare = rng.random((rows, cols)) < bare_frac
    na = rng.random((rows, cols)) < na_frac
    tall = rng.random((rows, cols)) < tall_frac
    hgt0 = np.where(tall, rng.uniform(1.0, min(25.0, zref * 0.8), (rows, cols)), rng.uniform(0.02, 0.9, (rows, cols)))
    pai0 = rng.uniform(0.1, 6.0, (rows, cols))
    xx = rng.uniform(0.3, 2.0, (rows, cols))
    xx[rng.random((rows, cols)) < 0.05] = 1.0  # the x == 1 branch of cankCpp / twostreamdifCpp
    gs = rng.uniform(0.05, 0.4, (rows, cols))
    lr = rng.uniform(0.2, 0.4, (rows, cols))
    cl0 = rng.uniform(0.0, 0.6, (rows, cols))
    cl0[rng.random((rows, cols)) < 0.1] = 0.0  # exercise the clump == 0 branches
    ld = rng.uniform(0.005, 0.3, (rows, cols))
    veg = {n: np.empty((rows, cols, nlyr)) for n in ("hgt", "pai", "x", "gsmax", "leafr", "leaft", "clump", "leafd",
                                                    "paia", "leafden")}
    for l in range(nlyr):
        sc = 1.0 if nlyr == 1 else 0.55 + 0.45 * np.sin(np.pi * (l + 0.5) / nlyr)
        hgt = np.where(bare, 0.0, hgt0)
        pai = np.where(bare, 0.0, pai0 * sc)
        hgt = np.where(na, np.nan, hgt)
        paia, leafden = foliage_density(max(reqhgt, 0.0), hgt, pai)
        above = ~(max(reqhgt, 0.0) < hgt)  # reqhgt >= hgt (or NaN): nothing above, density unused
        paia = np.where(above, 0.0, paia)
        leafden = np.where(above, 0.0, leafden)
        paia = np.where(bare, 0.0, paia)
        leafden = np.where(bare, 0.0, leafden)
        veg["hgt"][..., l], veg["pai"][..., l], veg["x"][..., l] = hgt, pai, xx
        veg["gsmax"][..., l], veg["leafr"][..., l], veg["leaft"][..., l] = gs, lr, 0.5 * lr
        veg["clump"][..., l], veg["leafd"][..., l] = cl0 * (1.0 - 0.3 * (1 - sc)), ld
        veg["paia"][..., l], veg["leafden"][..., l] = paia, leafden
    for n, a in veg.items():
        out[n] = F(a)
    return out


def make_problem(rows: int, cols: int, tsteps: int, reqhgt: float = 0.05, mode: int = 1, seed: int = 20240321,
                 nlyr: int = 1, zref: float = 30.0, lat: float = 50.0, lon: float = -5.0, complete: bool = True,
                 start_doy: int = 0, day_list=None, coarse: int = 8) -> GridProblem:
    """A full synthetic problem.  Modes 2/4 expand the per-hour series to [rows, cols, T] arrays with a
    smooth spatial modulation (what .cca + resample hands the reference, R/internal.R:523-542)."""
    f = forcing(tsteps, lat, lon, seed, start_doy=start_doy, day_list=day_list)
    tsteps = len(f["hour"])
    layered = mode in (3, 4)
    s = static_layers(rows, cols, reqhgt, seed, nlyr if layered else 1, zref)
    p = GridProblem(mode=mode, rows=rows, cols=cols, tsteps=tsteps, reqhgt=reqhgt, zref=zref, lat=lat, lon=lon,
                    Sminp=0.091, Smaxp=0.419, tfact=1.5, mat=float(np.mean(f["temp"])), complete=complete,
                    nlyr=nlyr if layered else 1)
    if layered:
        ndays = tsteps // 24
        edges = np.linspace(0, ndays, nlyr + 1).astype(np.int64)
        p.lyr_st = (edges[:-1] * 24).astype(np.int32)
        p.lyr_ed = (edges[1:] * 24 - 1).astype(np.int32)
    for n, a in s.items():
        p.arrays[n] = a
    if mode in (1, 3):
        for n, a in f.items():
            p.arrays[n] = np.ascontiguousarray(a)
    else:
        ii, jj = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
        u = (ii / max(rows - 1, 1)).ravel(order="F")
        v = (jj / max(cols - 1, 1)).ravel(order="F")
        for n in ("year", "month", "day", "hour", "winddir"):
            p.arrays[n] = np.ascontiguousarray(f[n])
        # [T, ncells] C-order == R's [rows, cols, T] column-major
        tcg = f["temp"][:, None] + (1.5 * (u - 0.5) - 1.0 * (v - 0.5))[None, :]
        esg = _satvap_r(tcg)
        eag = np.minimum(f["ea"][:, None] * (1 + 0.05 * (u - 0.5))[None, :], esg)
        p.arrays["temp"] = tcg.ravel()
        p.arrays["es"] = esg.ravel()
        p.arrays["ea"] = eag.ravel()
        p.arrays["tdew"] = _dewpoint_r(eag, tcg).ravel()
        p.arrays["pres"] = (f["pres"][:, None] - 0.3 * u[None, :]).ravel()
        swg = f["swdown"][:, None] * (1 - 0.1 * v)[None, :]
        p.arrays["swdown"] = swg.ravel()
        p.arrays["difrad"] = (f["difrad"][:, None] * (1 - 0.1 * v)[None, :]).ravel()
        p.arrays["lwdown"] = (f["lwdown"][:, None] + 5 * (u - 0.5)[None, :]).ravel()
        p.arrays["windspeed"] = np.maximum(0.5, f["windspeed"][:, None] * (1 + 0.2 * (v - 0.5))[None, :]).ravel()
        for n, amp in (("p_soilm", 0.02), ("p_Tg", 1.0), ("p_Tbp", 0.5), ("p_G", 5.0), ("p_umu", 0.05),
                       ("p_kp", 0.05), ("p_muGp", 0.005), ("p_dtrp", 1.0)):
            p.arrays[n] = (f[n][:, None] + amp * (u - 0.5)[None, :]).ravel()
        p.arrays["lats"] = np.ascontiguousarray(lat + 0.05 * (u - 0.5))
        p.arrays["lons"] = np.ascontiguousarray(lon + 0.08 * (v - 0.5))
    p.validate()
    return p


def bioclim_days(year: int = 2023):
    """14 days as .biosel would pick (R/internal.R:1690-1727): one mid-month day per month, then a hot
    and a cold day; and the four quarter index vectors (three consecutive day-blocks of 24 hours,
    0-based, 72 entries, as .getselq R/internal.R:1764-1774)."""
    mid = np.array([14, 45, 73, 104, 134, 165, 195, 226, 257, 287, 318, 348])
    days = np.concatenate([mid, [200, 20]])

    def q(first_month):
        ms = [(first_month + i) % 12 for i in range(3)]
        return np.concatenate([m * 24 + np.arange(24) for m in ms]).astype(np.int32)

    return days, dict(wetq=q(10), dryq=q(4), hotq=q(5), colq=q(11))


def make_coarse_problem(rows: int, cols: int, tsteps: int, reqhgt: float = 0.05, mode: int = 2, crows: int = 5,
                        ccols: int = 4, altcorrect: int = 0, seed: int = 20240321, nlyr: int = 1, zref: float = 30.0,
                        lat: float = 50.0, lon: float = -5.0, complete: bool = True) -> GridProblem:
    """Modes 2/4 with the climate and point-model series on a COARSE [crows, ccols] grid covering the same extent
    as the fine raster (what runpointmodela produces, one series per coarse cell), to be interpolated by the kernels
    (mcf_problem.clim_rows > 0).  oracle/prep_oracle.materialise_coarse expands it the way .runmodel2Cpp does."""
    assert mode in (2, 4)
    base = make_problem(rows, cols, tsteps, reqhgt=reqhgt, mode=mode - 1, seed=seed, nlyr=nlyr, zref=zref, lat=lat,
                        lon=lon, complete=complete)
    f = {n: base.arrays[n] for n in ("temp", "pres", "swdown", "difrad", "lwdown", "windspeed", "winddir", "ea", "es",
                                     "p_soilm", "p_Tg", "p_Tbp", "p_G", "p_umu", "p_kp", "p_muGp", "p_dtrp")}
    T = base.tsteps
    rng = np.random.default_rng(seed + 7)
    p = base.replace(mode=mode, clim_rows=crows, clim_cols=ccols, altcorrect=altcorrect)
    p.arrays = {n: a for n, a in base.arrays.items() if n not in f and n not in ("tdew",)}
    p.clim_drow, p.clim_dcol = crows / rows, ccols / cols
    p.clim_row0, p.clim_col0 = 0.5 * p.clim_drow - 0.5, 0.5 * p.clim_dcol - 0.5
    ncc = crows * ccols

    def field(series, amp, lo=None):
        a = series[:, None] + amp * rng.uniform(-1, 1, ncc)[None, :] * (1 + 0.2 * np.sin(np.arange(T) / 5.0))[:, None]
        if lo is not None:
            a = np.maximum(a, lo)
        return np.ascontiguousarray(a).ravel()  # [T, ccols * crows] C-order == R's [crows, ccols, T]

    p.arrays["temp"] = field(f["temp"], 1.5)
    rh = np.clip(100 * f["ea"] / f["es"], 20, 100)
    p.arrays["relhum"] = np.clip(field(rh, 6.0), 5.0, 100.0)
    p.arrays["pres"] = field(f["pres"], 0.4)
    sw = field(f["swdown"], 0.0) * np.repeat(1 + 0.1 * rng.uniform(-1, 1, ncc)[None, :], T, axis=0).ravel()
    p.arrays["swdown"] = sw
    p.arrays["difrad"] = np.minimum(field(f["difrad"], 0.0), sw)
    p.arrays["lwdown"] = field(f["lwdown"], 6.0)
    u2 = field(f["windspeed"], 0.4, lo=0.5)
    wd = field(f["winddir"], 15.0)
    p.arrays["wu"], p.arrays["wv"] = u2 * np.cos(wd * np.pi / 180), u2 * np.sin(wd * np.pi / 180)
    # R/internal.R:1256-1261: the wind direction handed to the solver is that of the mean coarse wind vector
    wuv = p.arrays["wu"].reshape(T, ncc).mean(axis=1)
    wvv = p.arrays["wv"].reshape(T, ncc).mean(axis=1)
    p.arrays["winddir"] = np.mod(np.arctan2(wvv, wuv) * 180 / np.pi, 360.0)
    for n, amp in (("p_soilm", 0.02), ("p_Tg", 1.0), ("p_Tbp", 0.5), ("p_G", 5.0), ("p_umu", 0.05), ("p_kp", 0.05),
                   ("p_muGp", 0.005)):
        p.arrays[n] = field(f[n], amp)
    p.arrays["p_dtrp"] = field(f["p_dtrp"], 1.0, lo=0.5)
    ii, jj = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    u = (ii / max(rows - 1, 1)).ravel(order="F")
    v = (jj / max(cols - 1, 1)).ravel(order="F")
    p.arrays["lats"] = np.ascontiguousarray(lat + 0.05 * (u - 0.5))
    p.arrays["lons"] = np.ascontiguousarray(lon + 0.08 * (v - 0.5))
    if altcorrect:
        dtm = 150 + 120 * np.sin(3 * u) * np.cos(2 * v)                      # fine elevations
        p.arrays["elevd"] = np.ascontiguousarray(40 * np.sin(5 * u + 2 * v))  # resample(dtmc) - dtm
        p.arrays["pfac"] = np.ascontiguousarray(((293 - 0.0065 * dtm) / 293) ** 5.26)
    p.validate()
    return p


def make_snow_inputs(rows: int, cols: int, tsteps: int, seed: int = 20240321, reqhgt: float = 0.05, lat: float = 61.0,
                     lon: float = 10.0, zref: float = 30.0):
    """Synthetic winter scenario for the snow operators (gridmodelsnow1 / gridmicrosnow1, src/microclimfCpp.cpp:4172,
    4894): sub-zero to just-above-zero air, precipitation events, a mix of snow-free, thinly and deeply covered cells,
    vegetation shorter and taller than the pack.  Returns dict(obstime, climdata, pointm, vegp, other)."""
    rng = np.random.default_rng(seed + 11)
    f = forcing(tsteps, lat, lon, seed, start_doy=20)
    T = len(f["hour"])
    temp = f["temp"] - 9.0 + 3.0 * np.sin(np.arange(T) / 37.0)
    es = _satvap_r(temp)
    relhum = np.clip(100 * f["ea"] / _satvap_r(f["temp"]), 35, 100)
    precip = np.where(rng.random(T) < 0.18, rng.gamma(1.5, 0.8, T), 0.0)
    climdata = dict(temp=temp, relhum=relhum, pres=f["pres"], swdown=f["swdown"], difrad=f["difrad"], lwdown=f["lwdown"] - 40,
                    windspeed=f["windspeed"], winddir=f["winddir"], precip=precip, umu=f["p_umu"])
    obstime = dict(year=f["year"], month=f["month"], day=f["day"], hour=f["hour"])
    Tcp = temp + 2.0 * np.sin(2 * np.pi * (f["hour"] - 9) / 24) - 1.0
    rsw = 0.35 * f["swdown"]
    rlw = 0.97 * (f["lwdown"] - 40)
    pointm = dict(Gp=f["p_G"] * 0.5, Tc=Tcp, RswabsG=rsw, RlwabsG=rlw, umu=f["p_umu"], tr=np.full(T, 0.4))
    s = static_layers(rows, cols, reqhgt, seed, 1, zref)
    R = lambda n: s[n].reshape(cols, rows).T  # noqa: E731  (flat R order -> [rows, cols])
    vegp = {n: R(n) for n in ("pai", "hgt", "leaft", "clump", "paia", "leafd", "leafden")}
    dc = np.where(rng.random((rows, cols)) < 0.25, 0.0, rng.uniform(0.02, 0.9, (rows, cols)))
    dg = dc * rng.uniform(0.4, 1.0, (rows, cols))
    other = dict(slope=R("slope"), aspect=R("aspect"), skyview=R("svfa"), wsa=s["wsa"].reshape(8, cols, rows).transpose(2, 1, 0),
                 hor=s["hor"].reshape(24, cols, rows).transpose(2, 1, 0), lat=lat, lon=lon, zref=zref, Smax=R("Smax"),
                 isnowdc=dc, isnowdg=dg, isnowac=rng.integers(0, 400, (rows, cols)).astype(np.int32),
                 isnowag=rng.integers(0, 400, (rows, cols)).astype(np.int32))
    return dict(obstime=obstime, climdata=climdata, pointm=pointm, vegp=vegp, other=other)
