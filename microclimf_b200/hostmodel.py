"""Host-side mirror of the reference's R drivers for the grid solver: `runmicro()`, `runmicro_big()`,
`subsetpointmodel()`, `checkinputs()` and the internal packing `.runmodel1Cpp` / `.runmodel3Cpp`.

These are the callers of the hot path (SURVEY.md §8f NEXT-1).  In the reference they are R functions
(R/Cppwrappers.R:376-543, R/internal.R:1020-1170, 1345-1460, 3290-3349, R/dataprep.R:31-102, 206-397) that
unpack terra rasters into matrices, derive the static layers and call `.Call(_microclimf_runmicroNCpp)`.
There is no R toolchain in this build's environment, so — per the tier rules — the host side is written in
Python with the same function names, argument names, defaults and error messages, on top of the same
operator mirror (`api.runmicro1Cpp` / `api.runmicro3Cpp`) the Rcpp stub of INTEGRATION.md binds.

What runs where:
  * GPU (this package's CUDA kernels): the solver itself, `.horizon` x 24 + sky view, `.windcoef` x 16;
  * host C++ (`mcf_flowacc`): the sequential flow-accumulation sweep behind `.topidx`;
  * host numpy: everything that is O(cells) bookkeeping in R (NA cleaning, layer selection, soil lookup,
    foliage density, slope/aspect, 16 -> 8 blend).

Third-party steps (terra slope/aspect, aggregate, resample; sf reprojection) are restated from their
published algorithms in `spatial.py` — PARITY UNPINNED there (no R, terra or sf here to compare with).
Everything from the `.Call` boundary down is pinned against the compiled reference on identical inputs
(tests/test_bundled_*.py).
"""
from __future__ import annotations

import math
import os
import warnings
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import api
from .problem import GridProblem
from .spatial import (Raster, aggregate_mean, as_raster, latlong_from_raster, latslons_from_raster, mask,
                      resample_bilinear)
from .tables import SOILPARAMETERS

VEG_NAMES = ("pai", "hgt", "x", "gsmax", "leafr", "clump", "leafd", "leaft")
WEATHER_COLS = ("temp", "relhum", "pres", "swdown", "difrad", "lwdown", "windspeed", "winddir", "precip")


# ---------------------------------------------------------------------------------------------
# micropoint
# ---------------------------------------------------------------------------------------------
@dataclass
class Micropoint:
    """The reference's S3 class `micropoint` (R/Cppwrappers.R:143-147): output of runpointmodel().

    weather : dict of the climdata columns (R/data.R `climdata`) + `obs_time` (numpy datetime64[s], UTC)
    dfo     : dict of point-model series: umu, kp, muGp, dtrp, G, soilm, Tg, Tc (+ DDp, T0p)
    subs    : 1-based indices of the retained hours within `tmeorig` (as in R)
    """
    weather: Dict[str, np.ndarray]
    dfo: Dict[str, np.ndarray]
    Tbz: Optional[np.ndarray]
    lat: float
    long: float
    zref: float
    subs: np.ndarray
    tmeorig: np.ndarray
    matemp: float


def _obstime(tme: np.ndarray) -> Dict[str, np.ndarray]:
    """obstime data.frame of the drivers (R/internal.R:1084-1086) from datetime64 values."""
    t = np.asarray(tme).astype("datetime64[s]")
    Y = t.astype("datetime64[Y]")
    M = t.astype("datetime64[M]")
    D = t.astype("datetime64[D]")
    year = Y.astype(int) + 1970
    month = (M.astype(int) % 12) + 1
    day = (D - M.astype("datetime64[D]")).astype(int) + 1
    hour = (t - D.astype("datetime64[s]")).astype(np.float64) / 3600.0
    return dict(year=year.astype(np.int32), month=month.astype(np.int32), day=day.astype(np.int32), hour=hour)


def subsetpointmodel(pointmodel: Micropoint, tstep: str = "month", what: str = "tmax", days=None, Tc=None) -> Micropoint:
    """ref subsetpointmodel (R/dataprep.R:31-102): keep the hottest / coldest / median day of each month
    (or year), or the listed days."""
    dfo = pointmodel.dfo
    tme = np.asarray(pointmodel.weather["obs_time"]).astype("datetime64[s]")
    ot = _obstime(tme)

    def extractday(Tcv, sel):
        if what == "tmax":
            s2 = int(np.argmax(Tcv[sel]))
        elif what == "tmin":
            s2 = int(np.argmin(Tcv[sel]))
        elif what == "tmedian":
            o = np.argsort(Tcv[sel], kind="stable")
            n = len(o) // 2
            s2 = int(o[n - 1])
        else:
            raise ValueError("what must be one of tmax, tmin or tmedian")
        k = sel[s2]
        return np.nonzero((ot["year"] == ot["year"][k]) & (ot["month"] == ot["month"][k]) & (ot["day"] == ot["day"][k]))[0]

    if days is None:
        Tcv = np.asarray(dfo["Tc"] if Tc is None else Tc)
        yrs = list(dict.fromkeys(ot["year"].tolist()))
        parts: List[np.ndarray] = []
        if tstep == "year":
            for y in yrs:
                parts.append(extractday(Tcv, np.nonzero(ot["year"] == y)[0]))
        elif tstep == "month":
            for y in yrs:
                sely = np.nonzero(ot["year"] == y)[0]
                for m in dict.fromkeys(ot["month"][sely].tolist()):
                    # the reference indexes `sel` within the year's subset but applies it to the whole series
                    # (R/dataprep.R:60-84); identical for the first year, reproduced as written for later ones
                    sel = np.nonzero(ot["month"][sely] == m)[0]
                    parts.append(extractday(Tcv, sel))
        else:
            raise ValueError("tstep must be one of year or month")
        ai = np.concatenate(parts)
    else:
        d = np.asarray(days, dtype=np.int64)
        ai = (np.repeat((d - 1) * 24, 24) + np.tile(np.arange(1, 25), d.size)) - 1
    weather = {k: np.asarray(v)[ai] for k, v in pointmodel.weather.items()}
    dfo2 = {k: np.asarray(v)[ai] for k, v in dfo.items()}
    Tbz = None if pointmodel.Tbz is None else np.asarray(pointmodel.Tbz)[ai]
    return Micropoint(weather=weather, dfo=dfo2, Tbz=Tbz, lat=pointmodel.lat, long=pointmodel.long, zref=pointmodel.zref,
                      subs=np.asarray(pointmodel.subs)[ai], tmeorig=pointmodel.tmeorig, matemp=pointmodel.matemp)


# ---------------------------------------------------------------------------------------------
# small R helpers
# ---------------------------------------------------------------------------------------------
def _getmode(v) -> float:
    """ref .getmode (R/internal.R:106-110)."""
    v = np.asarray(v, dtype=np.float64).ravel()
    v = v[~np.isnan(v)]
    uniq, first, counts = np.unique(v, return_index=True, return_counts=True)
    order = np.argsort(first, kind="stable")
    return float(uniq[order][np.argmax(counts[order])])


def _r_round(x):
    return np.round(x)  # R's round(x, 0) and numpy both round half to even


def _satvap(tc):
    """ref .satvap (R/internal.R:501-507)."""
    tc = np.asarray(tc, dtype=np.float64)
    es = 0.61078 * np.exp(17.27 * tc / (tc + 237.3))
    ei = 0.61078 * np.exp(21.875 * tc / (tc + 265.5))
    return np.where(tc < 0, ei, es)


def _dewpoint(ea, tc):
    """ref .dewpoint (R/internal.R:509-521)."""
    ea, tc = np.asarray(ea, dtype=np.float64), np.asarray(tc, dtype=np.float64)
    e0 = 611.2 / 1000
    L = (2.501 * 10 ** 6) - (2340 * tc)
    it = 1 / 273.15 - (461.5 / L) * np.log(ea / e0)
    Tdew = 1 / it - 273.15
    e0 = 610.78 / 1000
    L = 2.834 * 10 ** 6
    it = 1 / 273.15 - (461.5 / L) * np.log(ea / e0)
    Tfrost = 1 / it - 273.15
    return np.where(Tdew < 0, Tfrost, Tdew)


def _jday(year, month, day):
    """ref .jday (R/internal.R:440-449)."""
    year, month, day = (np.asarray(a, dtype=np.float64) for a in (year, month, day))
    dd = day + 0.5
    madj = month + (month < 3) * 12
    yadj = year + (month < 3) * -1
    j = np.trunc(365.25 * (yadj + 4716)) + np.trunc(30.6001 * (madj + 1)) + dd - 1524.5
    b = 2 - np.trunc(yadj / 100) + np.trunc(np.trunc(yadj / 100) / 4)
    return np.trunc(j + (j > 2299160) * b)


def _solalt(lt, lat, lon, jd):
    """ref .soltime / .solalt (R/internal.R:451-467)."""
    m = 6.24004077 + 0.01720197 * (jd - 2451545)
    eot = -7.659 * np.sin(m) + 9.863 * np.sin(2 * m + 3.5932)
    st = lt + (4 * lon + eot) / 60
    tt = 0.261799 * (st - 12)
    d = (np.pi * 23.5 / 180) * np.cos(2 * np.pi * ((jd - 159.5) / 365.25))
    sh = np.sin(d) * np.sin(lat * np.pi / 180) + np.cos(d) * np.cos(lat * np.pi / 180) * np.cos(tt)
    return (180 * np.arctan(sh / np.sqrt(1 - sh ** 2))) / np.pi


def _clearskyrad(ot, lat, lon, tc, rh, pk):
    """ref .clearskyrad (R/internal.R:469-484)."""
    jd = _jday(ot["year"], ot["month"], ot["day"])
    sa = _solalt(ot["hour"], lat, lon, jd) * np.pi / 180
    with np.errstate(invalid="ignore", divide="ignore"):
        m = 35 * np.sin(sa) * ((1224 * np.sin(sa) ** 2 + 1) ** (-0.5))
        TrTpg = 1.021 - 0.084 * (m * 0.00949 * pk + 0.051) ** 0.5
        xx = np.log(rh / 100) + ((17.27 * tc) / (237.3 + tc))
        Td = (237.3 * xx) / (17.27 - xx)
        u = np.exp(0.1133 - np.log(3.78) + 0.0393 * Td)
        Tw = 1 - 0.077 * (u * m) ** 0.3
        Ta = 0.935 * m
        Ic = 1352.778 * np.sin(sa) * TrTpg * Tw * Ta
    return np.where(np.isnan(Ic), 0.0, Ic)


def _foliageden(z, hgt, pai, paia=None, shape=1.5, rate=None):
    """ref .foliageden (R/internal.R:937-946): gamma(shape, rate) foliage profile over rescaled depth."""
    from scipy.special import gammainc, gammaln

    rate = shape / 7 if rate is None else rate
    with np.errstate(divide="ignore", invalid="ignore"):
        x = ((hgt - z) / hgt) * 10

        def pgamma(q):
            q = np.asarray(q, dtype=np.float64)
            return np.where(np.isnan(q), np.nan, np.where(q > 0, gammainc(shape, np.where(q > 0, q, 0.0) * rate), 0.0))

        def dgamma(q):
            q = np.asarray(q, dtype=np.float64)
            qq = np.where(q > 0, q, 1.0)
            dens = np.exp(shape * math.log(rate) + (shape - 1) * np.log(qq) - rate * qq - gammaln(shape))
            dens = np.where(np.isinf(q), 0.0, dens)
            return np.where(np.isnan(q), np.nan, np.where(q > 0, dens, 0.0))

        td = float(pgamma(10.0))
        rfd = dgamma(x) / td
        tdf = (pai / hgt) * rfd * 10
        if paia is None:
            paia = pgamma(x) * (pai / td)
    return dict(leafden=tdf, pai_a=paia)


def _intr(r: Raster, n: int, subs) -> np.ndarray:
    """ref .intr (R/internal.R:186-194): pick, for each of n steps, the nearest of the raster's layers."""
    nr = r.nlyr
    s = _r_round(np.linspace(0.50001, nr + 0.5, n)).astype(int)
    s = np.clip(s, 1, nr)
    s = s[np.asarray(subs, dtype=int) - 1]
    return r.values[:, :, s - 1]


def _vegpdmx(vegp: Dict[str, Raster]) -> int:
    return max(vegp[k].nlyr for k in VEG_NAMES)


def _unpack(dtm, vegp, soilc):
    """ref .unpack (R/internal.R:141-167): everything to rasters on the DTM's grid."""
    dtm = as_raster(dtm)
    vegp = {k: as_raster(vegp[k], dtm) for k in VEG_NAMES}
    soilc = {k: as_raster(soilc[k], dtm) for k in ("soiltype", "groundr")}
    for r in list(vegp.values()) + list(soilc.values()):
        if (r.nrows, r.ncols) != (dtm.nrows, dtm.ncols):
            raise ValueError("vegp / soilc layers must have the x and y dimensions of dtm")
    return dtm, vegp, soilc


def _cleanvars(vegp, soilc, dtm):
    """ref .cleanvars (R/internal.R:1020-1063): propagate NAs between layers, zero incomplete vegetation."""
    def cleanr(r, s, v=np.nan):
        out = r.values.copy()
        out[s, :] = v
        return r.like(out)

    s = (np.isnan(vegp["pai"].matrix()) | np.isnan(vegp["hgt"].matrix()) | np.isnan(soilc["soiltype"].matrix())
         | np.isnan(soilc["groundr"].matrix()) | np.isnan(dtm.matrix()))
    vegp = dict(vegp)
    soilc = dict(soilc)
    for k in VEG_NAMES:
        vegp[k] = cleanr(vegp[k], s)
    for k in ("soiltype", "groundr"):
        soilc[k] = cleanr(soilc[k], s)
    dtm = cleanr(dtm, s)
    with np.errstate(invalid="ignore"):
        z = ((vegp["pai"].matrix() == 0) | (vegp["hgt"].matrix() == 0) | np.isnan(vegp["gsmax"].matrix())
             | np.isnan(vegp["leafr"].matrix()) | np.isnan(vegp["clump"].matrix()) | np.isnan(vegp["leafd"].matrix())
             | np.isnan(vegp["leaft"].matrix()))
    vegp["pai"] = cleanr(vegp["pai"], z, 0.0)
    # `vegp$hgt <- .cleanr(vegp$hgt, 0)` passes 0 as the INDEX set (R/internal.R:1059): a no-op, kept as one
    vegp["pai"] = mask(vegp["pai"], dtm)
    vegp["hgt"] = mask(vegp["hgt"], dtm)
    return vegp, dtm, soilc


def _cleanvegp(vegp):
    """ref .cleanvegp (R/internal.R:119-139)."""
    vegp = dict(vegp)
    hm = vegp["hgt"].matrix().copy()
    pai = vegp["pai"].values.copy()
    with np.errstate(invalid="ignore"):
        for i in range(pai.shape[2]):
            pm = pai[:, :, i]
            hm[(pm == 0) & (hm > 0)] = 0
            pm[(hm == 0) & (pm > 0)] = 0
    vegp["pai"] = vegp["pai"].like(pai)
    vegp["hgt"] = vegp["hgt"].like(hm)
    return vegp


def checkinputs(weather, vegp, soilc, dtm, windhgt: float = 2):
    """ref checkinputs (R/dataprep.R:206-397): validates units / ranges (same messages) and applies the
    same corrections (humidity cap, diffuse <= total shortwave, clear-sky excess, wind direction modulo)."""
    dtm, vegp, soilc = _unpack(dtm, vegp, soilc)
    weather = {k: np.array(v) for k, v in weather.items()}
    for nm in ("obs_time",) + WEATHER_COLS:
        if nm not in weather:
            raise ValueError(f"Cannot find {nm} in weather")
    for nm in WEATHER_COLS:
        if np.isnan(np.asarray(weather[nm], dtype=np.float64)).any():
            raise ValueError("weather contains NAs")
    if not dtm.crs:
        raise ValueError("dtm must have a coordinate reference system specified")
    lat, lon = latlong_from_raster(dtm)
    ot = _obstime(weather["obs_time"])

    def check_vals(x, mn, mx, char, unit):
        x = np.asarray(x, dtype=np.float64)
        if np.isnan(x).any():
            raise ValueError(f"Missing values in weather${char}")
        if (x < mn).any():
            raise ValueError(f"{x.min()} outside range of typical {char} values. Units should be {unit}")
        if (x > mx).any():
            raise ValueError(f"{x.max()} outside range of typical {char} values. Units should be {unit}")

    mnelev = float(np.nanmin(dtm.matrix()))
    mxp = 108.5 * ((293 - 0.0065 * mnelev) / 293) ** 5.26
    mnp = 87 * ((293 - 0.0065 * mnelev) / 293) ** 5.26
    check_vals(weather["temp"], -50, 65, "temperature", "deg C")
    rh = np.asarray(weather["relhum"], dtype=np.float64)
    if rh.max() > 100:
        warnings.warn("relative humidity values capped at 100")
    weather["relhum"] = np.minimum(rh, 100.0)
    check_vals(weather["relhum"], 0, 100, "relative humidity", "percentage (0-100)")
    me = float(np.mean(weather["relhum"]))
    if me < 5 or me > 100:
        raise ValueError(f"Mean relative humidity of {me} implausible. Units should be percentage (0-100)")
    check_vals(weather["pres"], mnp, mxp, "pressure", "kPa ~101.3")
    check_vals(weather["swdown"], 0, 1350, "shortwave radiation", "W / m^2")
    check_vals(weather["difrad"], 0, 1350, "diffuse radiation", "W / m^2")
    check_vals(weather["lwdown"], 0, 600, "longwave radiation", "W / m^2")
    check_vals(weather["windspeed"], 0, 100, "wind speed", "m/s")
    ws = np.asarray(weather["windspeed"], dtype=np.float64)
    if windhgt != 2:
        ws = (ws * 4.87) / math.log(67.8 * windhgt - 5.42)
    if ws.max() > 30:
        warnings.warn(f"Maximum wind speed seems quite high. Check units are m/s and for {windhgt} m above ground")
    sw = np.asarray(weather["swdown"], dtype=np.float64)
    dif = np.asarray(weather["difrad"], dtype=np.float64).copy()
    dirr = sw - dif
    sel = dirr < 0
    if sel.any():
        dif[sel] = sw[sel]
        warnings.warn("Diffuse radiation values higher than shortwave radiation, and so was set to shortwave radiation values")
    csr = _clearskyrad(ot, lat, lon, np.asarray(weather["temp"], dtype=np.float64), weather["relhum"],
                       np.asarray(weather["pres"], dtype=np.float64))
    sel = dirr > csr  # `dirr` is not refreshed after the diffuse fix, as in the reference
    if sel.any():
        warnings.warn("Direct radiation values higher than expected clear-sky radiation values. Assigning excess as diffuse radiation")
        extra = np.ceil((dirr[sel] - csr[sel]) * 100) / 100
        dif[sel] = np.round(dif[sel] + extra, 3)
    weather["difrad"] = dif
    if (sw > csr + 50).any():
        warnings.warn("Short wave radiation values significantly higher than expected clear-sky radiation values")
    wd = np.asarray(weather["winddir"], dtype=np.float64)
    if wd.min() < 0 or wd.max() > 360:
        weather["winddir"] = np.mod(wd, 360)
        warnings.warn("wind direction adjusted to range 0-360 using modulo operation")
    for k in ("x", "gsmax", "leafr", "leafd"):
        if vegp[k].nlyr > 1:
            raise ValueError(f"time variant vegp${k} not supported")
    if vegp["clump"].nlyr > 1 and vegp["clump"].nlyr != vegp["pai"].nlyr:
        raise ValueError("clump must be a single numeric value or have the same dimensions as vegp$pai")
    with np.errstate(invalid="ignore"):
        if np.nanmax(vegp["leafr"].values + vegp["leaft"].values) > 1:
            raise ValueError("leaf reflectance + transmittance cannot be greater than one")
        st = soilc["soiltype"].values
        if np.nanmax(st) > 11 or np.nanmin(st) < 1:
            raise ValueError("Unrecognised soil type")

    def vals(r):
        v = r.values.ravel()
        return v[~np.isnan(v)]

    check_vals(vals(soilc["groundr"]), 0, 1, "soil reflectivity", "range 0 to 1")
    check_vals(vals(vegp["leafr"]), 0, 1, "leaf reflectivity", "range 0 to 1")
    check_vals(vals(vegp["clump"]), 0, 1, "vegetation clumping factor", "range 0 to 1")
    xx = vals(vegp["leafd"])
    check_vals(xx, 0, 5, "leaf diamater", "metres")
    if xx.mean() > 1:
        warnings.warn(f"Mean leaf diameter of {xx.mean()} seems large. Check units are in metres")
    check_vals(vals(vegp["gsmax"]), 0, 2, "maximum stomatal conductance", "mol / m^2 /s")
    xx = vals(vegp["pai"])
    if xx.min() < 0:
        raise ValueError("Minimum vegp$pai must be greater than or equal to zero")
    if xx.max() > 15:
        warnings.warn(f"Maximum vegp$pai of {xx.max()} seems high")
    vegp = _cleanvegp(vegp)
    return dict(weather=weather, vegp=vegp, soilc=soilc)


def _soilinit(soilc: Dict[str, Raster]) -> Dict[str, np.ndarray]:
    """ref .soilinit (R/internal.R:304-335): per-cell soil parameters from the soil-type lookup table (or
    from a layer of that name supplied in soilc)."""
    st = soilc["soiltype"].matrix()
    u = np.unique(st[~np.isnan(st)])
    out = {}
    for varn, key in (("rho", "rho"), ("Vm", "Vm"), ("Vq", "Vq"), ("Mc", "Mc"), ("psi_e", "psi_e"), ("b", "soilb"),
                      ("Smax", "Smax"), ("Smin", "Smin")):
        if varn in soilc:
            out[key] = as_raster(soilc[varn]).matrix().copy()
            continue
        m = np.full(st.shape, np.nan)
        for ui in u:
            row = SOILPARAMETERS["Number"].index(int(ui))
            m[st == ui] = SOILPARAMETERS[varn][row]
        out[key] = m
    return out


def terrain(dtm: Raster, v: str = "slope", unit: str = "degrees") -> Raster:
    """terra::terrain(dtm, v = "slope" | "aspect") (R/internal.R:1124-1129; R/Cppwrappers.R:483-484): Horn's stencil on the
    GPU (mcf_slope_aspect); NA on the edge and beside missing cells, as terra leaves them."""
    sl, asp = api.slope_aspect(dtm.matrix(), dtm.res[0], dtm.res[1])
    if v not in ("slope", "aspect"):
        raise ValueError("terrain: v must be 'slope' or 'aspect'")
    m = sl if v == "slope" else asp
    return dtm.like(m * (math.pi / 180.0) if unit == "radians" else m)


def _topidx(dtm: Raster) -> Raster:
    """ref .topidx (R/internal.R:861-874): topographic wetness index a / tan(slope) — mcf_topidx: Horn slope on the GPU,
    the sequential flow-accumulation sweep of flowaccCpp in host C++."""
    return dtm.like(api.topidx(dtm.matrix(), dtm.res[0], dtm.res[1]))


def _windsheltera(dtm: Raster, whgt: float, s) -> np.ndarray:
    """ref .windsheltera (R/internal.R:970-991), one call, all on the GPU (mcf_windshelter): .windcoef in 16 directions,
    block-mean + bilinear smoothing of each (terra aggregate / resample, restated), blended to 8 directions."""
    if s is None or (isinstance(s, float) and math.isnan(s)):
        s = min(dtm.nrows, dtm.ncols)
        s = 10 if s > 10 else s
    return api.windshelter(dtm.matrix(), dtm.res[0], whgt, int(s))


# ---------------------------------------------------------------------------------------------
# .runmodel1Cpp / .runmodel3Cpp: build the argument list of runmicro1Cpp / runmicro3Cpp
# ---------------------------------------------------------------------------------------------
@dataclass
class ModelCall:
    """The argument list handed to the operator (what `.Call(_microclimf_runmicroNCpp, ...)` receives)."""
    mode: int
    args: Dict[str, object] = field(default_factory=dict)
    prob: Optional[GridProblem] = None  # modes 2/4: the coarse-grid problem handed to the kernels

    def problem(self):
        if self.prob is not None:
            return self.prob
        a = self.args
        return api._problem(self.mode, a.get("dfsel"), a["obstime"], a["climdata"], a["pointm"], a["vegp"], a["soilc"],
                            a["reqhgt"], a["zref"], a["lat"], a["lon"], None, None, a["Sminp"], a["Smaxp"], a["tfact"],
                            a["complete"], a["mat"])

    def run(self, packed: bool = False) -> Dict[str, np.ndarray]:
        """The `.Call`: FP64 arrays as the reference returns them, or (packed = True) the integers writetonc
        would store for them, produced directly by the kernels (api.run_problem_packed)."""
        a = self.args
        if packed:
            return api.run_problem_packed(self.problem(), out=a["out"])
        if self.prob is not None:
            return api.run_problem(self.prob, out=a["out"])
        if self.mode == 1:
            return api.runmicro1Cpp(a["obstime"], a["climdata"], a["pointm"], a["vegp"], a["soilc"], a["reqhgt"], a["zref"],
                                    a["lat"], a["lon"], a["Sminp"], a["Smaxp"], a["tfact"], a["complete"], a["mat"], a["out"])
        return api.runmicro3Cpp(a["dfsel"], a["obstime"], a["climdata"], a["pointm"], a["vegp"], a["soilc"], a["reqhgt"],
                                a["zref"], a["lat"], a["lon"], a["Sminp"], a["Smaxp"], a["tfact"], a["complete"], a["mat"],
                                a["out"])


def prepare_model(micropoint: Micropoint, vegp, soilc, dtm, reqhgt: float = 0.05, runchecks: bool = True, pai_a=None,
                  tfact: float = 1.5, out: Sequence[bool] = (True,) * 10, slr=None, apr=None, hor=None, twi=None,
                  wsa=None, svf=None) -> ModelCall:
    """ref .runmodel1Cpp (R/internal.R:1065-1170) and .runmodel3Cpp (R/internal.R:1345-1460): everything up
    to, but not including, the `.Call`."""
    dtm, vegp, soilc = _unpack(dtm, vegp, soilc)
    layered = _vegpdmx(vegp) > 1  # dispatch of .runmicronosnow (R/internal.R:3331-3346)
    vegp, dtm, soilc = _cleanvars(vegp, soilc, dtm)
    weather = micropoint.weather
    if runchecks:
        rc = checkinputs(weather, vegp, soilc, dtm)
        weather, vegp, soilc = rc["weather"], rc["vegp"], rc["soilc"]
    obstime = _obstime(weather["obs_time"])
    temp = np.asarray(weather["temp"], dtype=np.float64)
    es = _satvap(temp)
    ea = es * np.asarray(weather["relhum"], dtype=np.float64) / 100
    climdata = dict(temp=temp, es=es, ea=ea, tdew=_dewpoint(ea, temp))
    for k in ("pres", "swdown", "difrad", "lwdown", "windspeed", "winddir"):
        climdata[k] = np.asarray(weather[k], dtype=np.float64)
    pointm = {k: np.asarray(v, dtype=np.float64) for k, v in micropoint.dfo.items()}
    nT = temp.size
    pointm["Tbp"] = np.asarray(micropoint.Tbz, dtype=np.float64) if reqhgt < 0 else np.zeros(nT)
    # ---- vegetation, soil, terrain layers
    subs = np.asarray(micropoint.subs, dtype=int)
    lat, lon = latlong_from_raster(dtm)
    vg, sc, dfsel, layered2 = _static_layers(vegp, soilc, dtm, micropoint, reqhgt, pai_a, slr, apr, hor, twi, wsa, svf,
                                             micropoint.zref, zero_pairs=True)
    assert layered2 == layered
    Sminp, Smaxp = _getmode(sc["Smin"]), _getmode(sc["Smax"])
    complete = len(subs) == len(micropoint.tmeorig)
    out = [bool(o) for o in out]
    if reqhgt == 0:
        out = [o and m for o, m in zip(out, (1, 0, 0, 1, 0, 1, 1, 1, 1, 1))]
    if reqhgt < 0:
        out = [o and m for o, m in zip(out, (1, 0, 0, 1, 0, 0, 0, 0, 0, 0))]
    args = dict(obstime=obstime, climdata=climdata, pointm=pointm, vegp=vg, soilc=sc, reqhgt=float(reqhgt),
                zref=float(micropoint.zref), lat=lat, lon=lon, Sminp=Sminp, Smaxp=Smaxp, tfact=float(tfact),
                complete=complete, mat=float(micropoint.matemp), out=[bool(o) for o in out])
    if layered:
        args["dfsel"] = dfsel
    return ModelCall(mode=3 if layered else 1, args=args)


def _static_layers(vegp, soilc, dtm, micropoint, reqhgt, pai_a, slr, apr, hor, twi, wsa, svf, zref, zero_pairs):
    """The vegetation / soil / terrain part shared by .runmodel1Cpp ... .runmodel4Cpp (R/internal.R:1098-1156,
    1278-1330): returns (vegp dict, soilc dict, dfsel | None, layered)."""
    n = len(micropoint.tmeorig)
    subs = np.asarray(micropoint.subs, dtype=int)
    dmx = _vegpdmx(vegp)
    layered = dmx > 1
    s_all = np.clip(_r_round(np.linspace(0.50001, dmx + 0.5, n)).astype(int), 1, dmx)
    subs2 = np.array(list(dict.fromkeys(s_all[subs - 1].tolist())), dtype=int)
    vg = {k: _intr(vegp[k], dmx, subs2) for k in VEG_NAMES}
    s_raw = _r_round(np.linspace(0.50001, dmx + 0.5, n)).astype(int)[subs - 1]
    if s_raw.size % 24 != 0:
        raise ValueError("weather needs to include data for entire days (24 hours)")
    lsubs = np.repeat(np.array([_getmode(row) for row in s_raw.reshape(-1, 24)]), 24).astype(int)
    if zero_pairs:  # only .runmodel1Cpp / .runmodel3Cpp do this (R/internal.R:1101-1104, 1382-1385)
        with np.errstate(invalid="ignore"):
            vg["hgt"] = np.where(vg["pai"] == 0, 0.0, vg["hgt"])
            vg["pai"] = np.where(vg["hgt"] == 0, 0.0, vg["pai"])
    paia_in = None if pai_a is None else _intr(as_raster(pai_a, dtm), n, subs)
    fd = _foliageden(reqhgt, vg["hgt"], vg["pai"], paia_in)
    vg["paia"], vg["leafden"] = fd["pai_a"], fd["leafden"]
    if not layered:
        vg = {k: v[:, :, 0] for k, v in vg.items()}
    dfsel = None
    if layered:
        lyrs = list(dict.fromkeys(lsubs.tolist()))
        st, ed = [], []
        for ly in lyrs:
            s = np.nonzero(lsubs == ly)[0] + 1
            st.append(int(math.floor(s[0] / 24) * 24))
            ed.append(int(math.floor(s[-1] / 24) * 24 - 1))
        dfsel = dict(lyr=np.arange(1, len(lyrs) + 1, dtype=np.int32), st=np.array(st, dtype=np.int32),
                     ed=np.array(ed, dtype=np.int32))
    soilp = _soilinit(soilc)
    sc: Dict[str, np.ndarray] = dict(gref=soilc["groundr"].matrix().copy(), Smin=soilp["Smin"], Smax=soilp["Smax"],
                                     soilb=soilp["soilb"], Psie=soilp["psi_e"], Vq=soilp["Vq"], Vm=soilp["Vm"],
                                     Mc=soilp["Mc"], rho=soilp["rho"])
    slope = terrain(dtm, "slope") if slr is None else as_raster(slr, dtm)
    aspect = terrain(dtm, "aspect") if apr is None else as_raster(apr, dtm)
    twi_r = _topidx(dtm) if twi is None else as_raster(twi, dtm)

    def fill(r: Raster, v: float) -> np.ndarray:
        m = r.matrix().copy()
        m[np.isnan(m)] = v
        return mask(r.like(m), dtm).matrix()

    sc["slope"], sc["aspect"], sc["twi"] = fill(slope, 0.0), fill(aspect, 0.0), fill(twi_r, 1.0)
    if hor is None:
        hor_a, svf_gpu = api.horizon(dtm.matrix(), dtm.res[0], want_svf=True)
    else:
        hor_a, svf_gpu = np.asarray(hor, dtype=np.float64), None
    sc["hor"] = hor_a
    if svf is None:
        if svf_gpu is None:
            msl = np.tan(np.mean(np.arctan(hor_a), axis=2))
            svf_gpu = 0.5 * np.cos(2 * msl) + 0.5
        sc["svfa"] = svf_gpu
    else:
        sc["svfa"] = as_raster(svf, dtm).matrix()
    s = 10 if dtm.res[0] <= 100 else 1
    sc["wsa"] = _windsheltera(dtm, zref, s) if wsa is None else np.asarray(wsa, dtype=np.float64)
    return vg, sc, dfsel, layered


def prepare_model_a(micropointa, vegp, soilc, dtm, dtmc, reqhgt: float = 0.05, runchecks: bool = True, altcorrect: int = 0,
                    pai_a=None, tfact: float = 1.5, out: Sequence[bool] = (True,) * 10, slr=None, apr=None, hor=None,
                    twi=None, wsa=None, svf=None) -> ModelCall:
    """ref .runmodel2Cpp (R/internal.R:1172-1344) and .runmodel4Cpp (:1461-1642): gridded climate.  `micropointa` is
    the list runpointmodela returns — one Micropoint (or None for a sea / NA cell) per cell of the coarse raster
    `dtmc`, in terra's cell order (row by row).

    The reference expands every series to the fine raster here, on the host (`.cca` -> terra::resample).  This build
    hands the COARSE [rows_c, cols_c, hours] arrays to the kernels, which interpolate per cell-hour and derive
    es / ea / tdew, the altitude correction and the wind speed themselves (include/microclimf_b200.h, clim_rows > 0);
    only O(cells) layers (elevation difference, pressure factor) and O(coarse) arithmetic are prepared here."""
    dtm, vegp, soilc = _unpack(dtm, vegp, soilc)
    vegp, dtm, soilc = _cleanvars(vegp, soilc, dtm)
    dtmc = as_raster(dtmc)
    cr, cc = dtmc.nrows, dtmc.ncols
    if len(micropointa) != cr * cc:
        raise ValueError("micropointa must hold one entry per cell of dtmc")
    last = None
    mats = []
    weathers: List[Optional[Dict[str, np.ndarray]]] = []
    for mp in micropointa:
        if mp is None:
            weathers.append(None)
            continue
        w = mp.weather
        if runchecks:
            rc = checkinputs(w, vegp, soilc, dtm, mp.zref)
            w, vegp, soilc = rc["weather"], rc["vegp"], rc["soilc"]
        weathers.append(w)
        mats.append(mp.matemp)
        last = mp
    if last is None:
        raise ValueError("micropointa holds no point-model output")
    T = len(last.weather["temp"])
    obstime = _obstime(last.weather["obs_time"])

    def cca(getter):
        """ref .cca (R/internal.R:523-542) without the resample: [rows_c, cols_c, T], NA where there is no point model."""
        a = np.full((cr, cc, T), np.nan)
        k = 0
        for i in range(cr):
            for j in range(cc):
                if micropointa[k] is not None:
                    a[i, j, :] = getter(k)
                k += 1
        return a

    wv_ = lambda name: cca(lambda k: np.asarray(weathers[k][name], dtype=np.float64))  # noqa: E731
    dv_ = lambda name: cca(lambda k: np.asarray(micropointa[k].dfo[name], dtype=np.float64))  # noqa: E731
    p = GridProblem(mode=2, rows=dtm.nrows, cols=dtm.ncols, tsteps=T, reqhgt=float(reqhgt), zref=float(last.zref),
                    tfact=float(tfact), mat=float(np.mean(mats)), complete=len(last.subs) == len(last.tmeorig))
    for k in ("year", "month", "day", "hour"):
        p.set(k, obstime[k])
    p.clim_rows, p.clim_cols, p.altcorrect = cr, cc, int(altcorrect)
    rxf, ryf = dtm.res
    rxc, ryc = dtmc.res
    p.clim_drow, p.clim_dcol = ryf / ryc, rxf / rxc
    p.clim_row0 = (dtmc.ymax - (dtm.ymax - 0.5 * ryf)) / ryc - 0.5
    p.clim_col0 = ((dtm.xmin + 0.5 * rxf) - dtmc.xmin) / rxc - 0.5
    p.set("temp", wv_("temp"))
    p.set("relhum", wv_("relhum"))
    pk = wv_("pres")
    if altcorrect:
        zc = dtmc.matrix().copy()
        zc[np.isnan(zc)] = 0.0  # R/internal.R:1229
        pk = pk / (((293 - 0.0065 * zc[:, :, None]) / 293) ** 5.26)
        zf = dtm.matrix()
        p.set("pfac", ((293 - 0.0065 * zf) / 293) ** 5.26)
        p.set("elevd", resample_bilinear(dtmc.like(zc), dtm).matrix() - zf)
    p.set("pres", pk)
    for k in ("swdown", "difrad", "lwdown"):
        p.set(k, wv_(k))
    u2, wd = wv_("windspeed"), wv_("winddir")
    wu, wvv = u2 * np.cos(wd * np.pi / 180), u2 * np.sin(wd * np.pi / 180)
    p.set("wu", wu)
    p.set("wv", wvv)
    p.set("winddir", np.mod(np.arctan2(np.nanmean(wvv, axis=(0, 1)), np.nanmean(wu, axis=(0, 1))) * 180 / np.pi, 360))
    for src, dst in (("umu", "p_umu"), ("kp", "p_kp"), ("muGp", "p_muGp"), ("dtrp", "p_dtrp"), ("G", "p_G"),
                     ("soilm", "p_soilm"), ("Tg", "p_Tg")):
        p.set(dst, dv_(src))
    if reqhgt < 0:
        p.set("p_Tbp", cca(lambda k: np.asarray(micropointa[k].Tbz, dtype=np.float64)))
    else:
        p.set("p_Tbp", np.zeros((cr, cc, T)))
    vg, sc, dfsel, layered = _static_layers(vegp, soilc, dtm, last, reqhgt, pai_a, slr, apr, hor, twi, wsa, svf,
                                            last.zref, zero_pairs=False)
    if layered:
        # .sortvegp can hand over more layers than dfsel has rows (short series, R/internal.R:262-270); the drivers
        # index layers 0 .. nrow(dfsel) - 1 only
        nl = min(vg["pai"].shape[2], len(dfsel["st"]))
        vg = {k: (v[:, :, :nl] if np.ndim(v) == 3 else v) for k, v in vg.items()}
        p.mode, p.nlyr = 4, nl
        p.lyr_st, p.lyr_ed = dfsel["st"], dfsel["ed"]
    for k, v in vg.items():
        p.set(k, v)
    for k, v in sc.items():
        p.set("Psie" if k == "Psie" else k, v)
    lats, lons = latslons_from_raster(dtm)
    p.set("lats", lats)
    p.set("lons", lons)
    p.Sminp, p.Smaxp = _getmode(sc["Smin"]), _getmode(sc["Smax"])
    out = [bool(o) for o in out]
    if reqhgt == 0:
        out = [o and bool(m) for o, m in zip(out, (1, 0, 0, 1, 0, 1, 1, 1, 1, 1))]
    if reqhgt < 0:
        out = [o and bool(m) for o, m in zip(out, (1, 0, 0, 1, 0, 0, 0, 0, 0, 0))]
    p.validate()
    return ModelCall(mode=p.mode, args=dict(out=out), prob=p)


def runmicro(micropoint, reqhgt, vegp, soilc, dtm, dtmc=None, altcorrect=0, snow=False, snowmod=None, runchecks=True,
             pai_a=None, tfact=1.5, out=(True,) * 10, slr=None, apr=None, hor=None, twi=None, wsa=None, svf=None,
             method="Cpp", packed=False):
    """ref runmicro (R/Cppwrappers.R:376-396): grid microclimate model.  Returns the reference's named list
    (dict of [rows, cols, hours] arrays) plus `tme`.  As in the reference, `svf` and `method` are accepted
    but not forwarded (R/internal.R:3336).  A list of micropoints (runpointmodela) with `dtmc` takes the gridded-climate path
    (`.runmodel2Cpp` / `.runmodel4Cpp` -> prepare_model_a); `snow = TRUE` with a data.frame micropoint and `snowmod` (snowmodel1's output)
    takes `.runmicrosnow1` (runmicrosnow1), with a list of micropoints and `dtmc` `.runmicrosnow2` (runmicrosnow2).
    `packed = True` (an addition) returns writetonc's integer packing straight from the kernels."""
    if snow:
        if not isinstance(micropoint, Micropoint):
            if dtmc is None:
                raise ValueError("Require dtmc. Please provide\n")
            mout = runmicrosnow2(micropoint, reqhgt, vegp, soilc, dtm, dtmc, snowmod, altcorrect, runchecks, pai_a, tfact, out,
                                 slr, apr, hor, twi, wsa, svf)
            mout["tme"] = np.asarray(next(m for m in micropoint if m is not None).tmeorig)
            return mout
        mout = runmicrosnow1(micropoint, reqhgt, vegp, soilc, dtm, snowmod, runchecks, pai_a, tfact, out, slr, apr, hor, twi,
                             wsa, svf)
        mout["tme"] = np.asarray(micropoint.tmeorig)
        return mout
    if not isinstance(micropoint, Micropoint):
        if dtmc is None:
            raise ValueError("Require dtmc. Please provide\n")
        call = prepare_model_a(micropoint, vegp, soilc, dtm, dtmc, reqhgt, runchecks, altcorrect, pai_a, tfact, out, slr,
                               apr, hor, twi, wsa)
        tme = next(m for m in micropoint if m is not None).tmeorig
    else:
        call = prepare_model(micropoint, vegp, soilc, dtm, reqhgt, runchecks, pai_a, tfact, out, slr, apr, hor, twi, wsa)
        tme = micropoint.tmeorig
    mout = call.run(packed=packed)
    mout["tme"] = np.asarray(tme)
    return mout


# ---------------------------------------------------------------------------------------------
# runmicro_big
# ---------------------------------------------------------------------------------------------
def _checkbiginputs(dtm, vegp, soilc):
    """ref .checkbiginputs (R/internal.R:1644-1662)."""
    ok = ~np.isnan(dtm.matrix())
    bad = [k for k in VEG_NAMES if (ok & np.isnan(vegp[k].matrix())).any()]
    if bad:
        raise ValueError("The following layers of vegp contain NA that are not NA in dtm: " + " ".join(bad) + "\n")
    bad = [k for k in ("soiltype", "groundr") if (ok & np.isnan(soilc[k].matrix())).any()]
    if bad:
        raise ValueError("The following layers of soilc contain NA that are not NA in dtm: " + " ".join(bad) + "\n")


def runmicro_big(micropoint, reqhgt, pathout, vegp, soilc, dtm, dtmc=None, altcorrect=0, tilesize=None, toverlap=0,
                 writeasnc=False, runchecks=True, pai_a=None, tfact=1.5, out=(True,) * 10, gpus=None, sink=None,
                 window_days=5):
    """ref runmicro_big (R/Cppwrappers.R:444-543): whole-area terrain layers once, then the model tile by
    tile, one file per tile in `<pathout>microut/` (`area_RR_CC.npz`, the analogue of the reference's RDS;
    `writeasnc = TRUE` stores writetonc's x100 integer packing, produced by the kernels' packed sink).
    Returns the list of files written.

    `gpus = N` (an addition; the reference's arguments are unchanged) replaces the tile loop by ONE COLUMN BAND PER GPU
    (bigrun.py: one process per GPU, statics distributed over NCCL, the whole-raster twi mean all-reduced, no collective
    during the solve).  `sink` then says what becomes of the hourly results: "packed" (default with `writeasnc`):
    writetonc's integers, one `area_<band>_<window>.nc` per band and `window_days`-day window; "summary" (default
    otherwise): per-cell mean / min / max of every requested output over the series, reduced inside the kernel and
    returned as {name: {stat: [rows, cols]}} — the sink for rasters whose hourly arrays exist nowhere; "arrays": the
    [rows, cols, hours] arrays themselves (small rasters).  The tile-size heuristic and `toverlap` do not apply (bands
    need no overlap: cells are independent).  Unlike the tile loop, which subtracts each tile's own mean of
    log(twi)/tfact (src/microclimfCpp.cpp:993-1004 sees one tile), the band run subtracts the whole area's mean — the
    untiled result."""
    dtm, vegp, soilc = _unpack(dtm, vegp, soilc)
    _checkbiginputs(dtm, vegp, soilc)
    if gpus is not None:
        if not isinstance(micropoint, Micropoint):
            raise ValueError("runmicro_big(gpus = N) takes a data.frame-climate micropoint (runpointmodel)")
        from . import bigrun

        def fill0_(r):
            m = r.matrix().copy()
            m[np.isnan(m)] = 0.0
            return mask(r.like(m), dtm)

        slr, apr = fill0_(terrain(dtm, "slope")), fill0_(terrain(dtm, "aspect"))
        twi = _topidx(dtm)
        hor, svfa = api.horizon(dtm.matrix(), dtm.res[0], want_svf=True)
        wsa = _windsheltera(dtm.like(dtm.matrix() + vegp["hgt"].matrix()), 8, None)
        call = prepare_model(micropoint, vegp, soilc, dtm, reqhgt, runchecks, pai_a, tfact, out, slr, apr, hor, twi, wsa,
                             svf=svfa)
        root = call.problem()
        root.tme = np.asarray(micropoint.weather["obs_time"])
        sink = sink or ("packed" if writeasnc else "summary")
        res = bigrun.run_local(root, int(gpus), sink=sink, out=call.args["out"], pathout=os.path.join(pathout, "microut"),
                               window_days=int(window_days), dtm=dtm)
        res["tme"] = root.tme
        return res
    if tilesize is None:
        nt = len(micropoint.weather["temp"])
        osize = math.sqrt(20000000 / nt) - 2 * toverlap
        sizeo = np.array([10, 20, 50, 100, 200, 500, 1000, 2000])
        tilesize = int(sizeo[np.argmin(np.abs(osize - sizeo))])
    rws, cls = -(-dtm.nrows // tilesize), -(-dtm.ncols // tilesize)
    path2 = os.path.join(pathout, "microut")
    os.makedirs(path2, exist_ok=True)

    def fill0(r):
        m = r.matrix().copy()
        m[np.isnan(m)] = 0.0
        return mask(r.like(m), dtm)

    slr, apr = fill0(terrain(dtm, "slope")), fill0(terrain(dtm, "aspect"))
    twi = _topidx(dtm)
    hor, svfa = api.horizon(dtm.matrix(), dtm.res[0], want_svf=True)
    dsm = dtm.like(dtm.matrix() + vegp["hgt"].matrix())
    wsa = _windsheltera(dsm, 8, None)
    rx, ry = dtm.res
    written = []
    for rw in range(1, rws + 1):
        for cl in range(1, cls + 1):
            # .croprast (R/internal.R:1664-1677): the overlap is subtracted in map units, as written
            xmn = max(dtm.xmin + (cl - 1) * tilesize * rx - toverlap, dtm.xmin)
            xmx = min(dtm.xmin + cl * tilesize * rx + toverlap, dtm.xmax)
            ymn = max(dtm.ymax - rw * tilesize * ry - toverlap, dtm.ymin)
            ymx = min(dtm.ymax - (rw - 1) * tilesize * ry + toverlap, dtm.ymax)
            c0, c1 = int(round((xmn - dtm.xmin) / rx)), int(round((xmx - dtm.xmin) / rx))
            r0, r1 = int(round((dtm.ymax - ymx) / ry)), int(round((dtm.ymax - ymn) / ry))
            dtmi = dtm.crop(r0, r1, c0, c1)
            if np.count_nonzero(~np.isnan(dtmi.matrix())) <= 1:
                continue
            vegpi = {k: v.crop(r0, r1, c0, c1) for k, v in vegp.items()}
            soilci = {k: v.crop(r0, r1, c0, c1) for k, v in soilc.items()}
            want_packed = bool(writeasnc and all(out))
            if writeasnc and not want_packed:
                warnings.warn("Can only write as nc with all variables in out set to TRUE. Writing as RDS\n")
            mout = runmicro(micropoint, reqhgt, vegpi, soilci, dtmi, dtmc, altcorrect, False, None, runchecks, pai_a,
                            tfact, out, slr.crop(r0, r1, c0, c1), apr.crop(r0, r1, c0, c1), hor[r0:r1, c0:c1, :],
                            twi.crop(r0, r1, c0, c1), wsa[r0:r1, c0:c1, :], svf=svfa[r0:r1, c0:c1], packed=want_packed)
            mp1 = micropoint if isinstance(micropoint, Micropoint) else next(m for m in micropoint if m is not None)
            mout["tme"] = np.asarray(mp1.weather["obs_time"])  # `mout$tme <- tme`: the hours actually modelled (wrap:517)
            fo = os.path.join(path2, f"area_{rw:02d}_{cl:02d}")
            ext = [dtmi.xmin, dtmi.xmax, dtmi.ymin, dtmi.ymax]
            if want_packed:
                from .ncwriter import writetonc
                writetonc(mout, fo + ".nc", dtmi, reqhgt)  # the integers come from the kernels' packed sink
                written.append(fo + ".nc")
            else:
                np.savez(fo + ".npz", extent=ext, dtm=dtmi.matrix(), **mout)
                written.append(fo + ".npz")
    return written


# ---------------------------------------------------------------------------------------------
# snow: .snowmodel1 (the 5-day chunk driver around gridmodelsnow1)
# ---------------------------------------------------------------------------------------------
def _sortl(vegp: Dict[str, Raster], sdep: np.ndarray) -> Dict[str, np.ndarray]:
    """ref .sortl (R/internal.R:2389-2420): vegetation averaged over the layers of the hours with snow."""
    out = {}
    for k in ("pai", "hgt", "leaft", "clump"):
        a = vegp[k].values
        dmx = a.shape[2]
        if dmx > 1:
            n = len(sdep)
            s = np.clip(_r_round(np.linspace(0.50001, dmx + 0.5, n)).astype(int), 1, dmx)
            sel = np.nonzero(np.asarray(sdep) > 0)[0]
            s = s[sel] if sel.size else s[:1]
            num, fre = np.unique(s, return_counts=True)
            m = a[:, :, 0] * 0
            for j in range(len(num)):
                m = m + a[:, :, j] * fre[j]  # the reference indexes layer j, not num[j] (R/internal.R:2413): as written
            out[k] = m / fre.sum()
        else:
            out[k] = a[:, :, 0].copy()
    return out


def _tpicalc(af: int, me: int, dtm: Raster, tfact: float) -> np.ndarray:
    """ref .tpicalc (R/internal.R:2479-2493): topographic positioning index for snow redistribution."""
    z = dtm.matrix()
    if af < me / 2 and af >= 1:
        dtmc = resample_bilinear(aggregate_mean(dtm, af, na_rm=True), dtm).matrix()
    else:
        dtmc = z * 0 + np.nanmean(z)
    tpic = np.exp((dtmc - z) * tfact)
    with np.errstate(invalid="ignore"):
        tpic[tpic < 0.05] = 0.1
        tpic[tpic > 10] = 10
    return tpic / np.nanmean(tpic)


def _snow_chunks(op, ot, clim_of, point_of, wss_of, umu, dtm, vg, other, snowenv, snowinitd, tfact, zref, chunk_days,
                 smooth_shelter, af_floor=None):
    """The 5-day chunk loop shared by .snowmodel1 (R/internal.R:2553-2613) and .snowmodel2 (:2948-3006): terrain of
    DTM + ground snow depth before each chunk (slope/aspect in numpy, horizons / sky view / wind shelter as GPU
    stencils), the grid snow operator, then the topographic redistribution of fresh snow.  `clim_of(s)`, `point_of(s)`
    slice the operator's climate and point-model inputs to the chunk's hours, `wss_of(s)` gives its wind speeds."""
    z = dtm.matrix()
    h = ot["hour"].size
    span = 24 * chunk_days
    nchunks = h // span
    shape = z.shape + (h,)
    Tc, Tg, snowdepg, swe, sden = (np.full(shape, np.nan) for _ in range(5))
    dtms = dtm.like(z + (z * 0 + snowinitd) * 0.5)
    for ch in range(nchunks):
        sl = terrain(dtms, "slope").matrix()
        sl[np.isnan(sl)] = 0
        other["slope"] = mask(dtm.like(sl), dtm).matrix()
        ap = terrain(dtms, "aspect").matrix()
        ap[np.isnan(ap)] = 180
        other["aspect"] = mask(dtm.like(ap), dtm).matrix()
        other["hor"], other["skyview"] = api.horizon(dtms.matrix(), dtm.res[0], want_svf=True)
        other["wsa"] = _windsheltera(dtms, zref, 10 if smooth_shelter else 1)
        s = slice(ch * span, min((ch + 1) * span, h))
        ns = s.stop - s.start
        smod = op({k: v[s] for k, v in ot.items()}, clim_of(s), point_of(s), vg, other, snowenv)
        # topographic snow redistribution (R/internal.R:2584-2599 / :2977-2992)
        tpr = 10 * np.mean(wss_of(s)) ** 0.5
        af = int(round(tpr / dtm.res[0]))
        if af_floor is not None and af < af_floor:  # only .snowmodel2 floors the radius (R/internal.R:2980)
            af = af_floor
        tpi = _tpicalc(af, min(dtm.nrows, dtm.ncols), dtms, tfact)
        asd = np.repeat(other["isnowdg"][:, :, None], ns, axis=2)
        dsnow = smod["sdepg"] - asd
        dsnow2 = dsnow * tpi[:, :, None]
        with np.errstate(invalid="ignore"):
            dsnow2 = np.where(dsnow < 0, dsnow, dsnow2)
        asc = np.repeat(other["isnowdc"][:, :, None], ns, axis=2)
        cdsnow = smod["sdepc"] - asc - dsnow
        Tc[:, :, s], Tg[:, :, s], sden[:, :, s] = smod["Tc"], smod["Tg"], smod["sden"]
        swe[:, :, s] = (asc + cdsnow + dsnow2) * smod["sden"]
        snowdepg[:, :, s] = asd + dsnow2
        other["isnowdc"] = (asc + cdsnow + dsnow2)[:, :, -1]
        # `other$isnowac <- (asd + dsnow2)[,,n]` is immediately overwritten by the ages and isnowdg is never advanced
        # (R/internal.R:2606-2609, :2999-3002): reproduced
        other["isnowac"] = np.nan_to_num(smod["agec"]).astype(np.int32)
        other["isnowag"] = np.nan_to_num(smod["ageg"]).astype(np.int32)
        dtms = dtm.like(z + snowdepg[:, :, s.stop - 1])
    return dict(Tc=Tc, Tg=Tg, groundsnowdepth=snowdepg, totalSWE=swe, snowden=sden, umu=umu)


def snowmodel1(weather, pointm, dtm, vegp, soilc, snowenv: str = "Taiga", snowinitd: float = 0, snowinita: float = 0,
               zref: float = 2, tfact: float = 0.02, chunk_days: int = 5, operator=None):
    """ref .snowmodel1 (R/internal.R:2498-2616) from the point where the point snow model has run: `pointm` is the
    data.frame built from pointmodelsnow's output (Gp, Tc, RswabsG, RlwabsG, umu, tr, and sdepc for .sortl) — the
    point model is upstream of this build.  The grid model runs in 5-day chunks; between chunks slope,
    aspect, the 24 horizons, sky view and wind shelter are recomputed from DTM + ground snow depth (GPU stencils) and
    fresh snow is redistributed by the topographic positioning index.  `operator` defaults to the CUDA
    `snow.gridmodelsnow1`; the tests pass the compiled reference's to check the driver end to end."""
    from . import snow as snowops

    op = operator or snowops.gridmodelsnow1
    dtm, vegp, soilc = _unpack(dtm, vegp, soilc)
    tme = np.asarray(weather["obs_time"]).astype("datetime64[s]")
    ot = _obstime(tme)
    ot["hour"] = np.floor(ot["hour"])  # obstime$hour = tme$hour here (R/internal.R:2531), no minutes
    z = dtm.matrix()
    sdep = z * 0 + snowinitd
    sage = z * 0 + snowinita
    lat, lon = latlong_from_raster(dtm)
    vg = _sortl(vegp, np.asarray(pointm["sdepc"])[:tme.size])
    other = dict(zref=float(zref), lat=lat, lon=lon, isnowdc=sdep, isnowac=np.nan_to_num(sage).astype(np.int32),
                 isnowdg=sdep * 0.5, isnowag=np.nan_to_num(sage).astype(np.int32))
    climcols = ("temp", "relhum", "pres", "swdown", "difrad", "lwdown", "windspeed", "winddir", "precip")
    wind = np.asarray(weather["windspeed"], dtype=np.float64)
    return _snow_chunks(op, ot,
                        lambda s: {k: np.asarray(weather[k], dtype=np.float64)[s] for k in climcols},
                        lambda s: {k: np.asarray(pointm[k], dtype=np.float64)[s]
                                   for k in ("Gp", "Tc", "RswabsG", "RlwabsG", "umu", "tr")},
                        lambda s: wind[s], np.asarray(pointm["umu"]), dtm, vg, other, snowenv, snowinitd, tfact, zref,
                        chunk_days, dtm.res[0] <= 100)


def snowmodel2(climdata, pointm, tme, dtm, vegp, soilc, sdept, wuv, wvv, coarse_dims=(10, 10), snowenv: str = "Taiga",
               snowinitd: float = 0, snowinita: float = 0, zref: float = 2, tfact: float = 0.02, chunk_days: int = 5,
               operator=None):
    """ref .snowmodel2 (R/internal.R:2780-3015) from the point where the per-coarse-cell point snow models have run
    and their series have been resampled to the DTM (:2866-2933; `.cca` + terra::resample — spatial.resample_bilinear
    here — and the altitude correction are the caller's, as is pointmodelsnow itself).

    `climdata`: temp, relhum, pres, swdown, difrad, lwdown, windspeed, precip as [rows, cols, hours] arrays and
    winddir [hours]; `pointm`: Gp, Tc, RswabsG, RlwabsG, umu, tr as [rows, cols, hours]; `sdept`: the hourly maximum of
    the point models' canopy snow depth (selects the vegetation layers, :2936-2938); `wuv`, `wvv`: the coarse grid's
    mean wind components per hour (:2918-2919, they set the redistribution radius); `coarse_dims`: rows, cols of the
    climate grid (wind shelter is smoothed only when both are >= 10, :2963-2964).  Returns .snowmodel2's list, masked by
    the DTM (.cleansmod, :3811)."""
    from . import snow as snowops

    op = operator or snowops.gridmodelsnow2
    dtm, vegp, soilc = _unpack(dtm, vegp, soilc)
    tme = np.asarray(tme).astype("datetime64[s]")
    ot = _obstime(tme)
    ot["hour"] = np.floor(ot["hour"])
    z = dtm.matrix()
    sdep = z * 0 + snowinitd
    sage = z * 0 + snowinita
    vg = _sortl(vegp, np.asarray(sdept)[:tme.size])
    vg["leaft"] = np.where(np.isnan(vg["leaft"]), 0.001, vg["leaft"])  # :2971
    lats, lons = latslons_from_raster(dtm)
    other = dict(zref=float(zref), lats=lats, lons=lons, isnowdc=sdep, isnowac=np.nan_to_num(sage).astype(np.int32),
                 isnowdg=sdep * 0.5, isnowag=np.nan_to_num(sage).astype(np.int32))
    arr3 = ("temp", "relhum", "pres", "swdown", "difrad", "lwdown", "windspeed", "precip")
    clim = {k: np.asarray(climdata[k], dtype=np.float64) for k in arr3}
    wdir = np.asarray(climdata["winddir"], dtype=np.float64)
    pnt = {k: np.asarray(pointm[k], dtype=np.float64) for k in ("Gp", "Tc", "RswabsG", "RlwabsG", "umu", "tr")}
    wss = np.sqrt(np.asarray(wuv, dtype=np.float64) ** 2 + np.asarray(wvv, dtype=np.float64) ** 2)
    out = _snow_chunks(op, ot,
                       lambda s: dict({k: v[:, :, s] for k, v in clim.items()}, winddir=wdir[s]),
                       lambda s: {k: v[:, :, s] for k, v in pnt.items()},
                       lambda s: wss[s], pnt["umu"], dtm, vg, other, snowenv, snowinitd, tfact, zref, chunk_days,
                       dtm.res[0] <= 100 and min(coarse_dims) >= 10, af_floor=2)
    land = ~np.isnan(z)
    return {k: np.where(land[:, :, None], v, np.nan) for k, v in out.items()}


def canintfrac(hgt, pai, uf, prec, tc, Li):
    """ref canintfrac (src/microclimfCpp.cpp:5417-5451) around canopysnowintCpp (:3713-3739): fraction of a snowfall
    `prec` (mm SWE) the canopy intercepts, per cell; 0.5 everywhere when prec <= 0, NA where hgt is NA."""
    hgt = np.asarray(hgt, dtype=np.float64)
    pai = np.asarray(pai, dtype=np.float64)
    na = np.isnan(hgt)
    if not prec > 0.0:
        return np.where(na, np.nan, 0.5)
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        h = np.where(hgt < 0.001, 0.001, hgt)
        p = np.where(pai < 0.001, 0.001, pai)
        Be = np.sqrt(0.003 + (0.2 * p) / 2.0)
        uh = uf / Be
        Lc = (0.25 * (p / h)) ** -1.0
        Lm = 2.0 * Be ** 3.0 * Lc
        k1 = Be / Lm
        uzm = (uh / (h * k1)) * (1 - np.exp(-k1 * h))
        uzm = np.where(uzm < uf, uf, uzm)
        rhos = 67.92 + 51.25 * np.exp(tc / 2.59)
        Lstr = 6.2 * (0.26 + 46 / rhos) * p
        kc = 1.0 / (2.0 * np.cos(np.arctan(uzm / 0.8)))
        Cp = 1.0 - np.exp(-kc * p)
        I1 = (Lstr - Li) * (1.0 - np.exp(-(Cp / Lstr) * prec))
        cis = I1 * 0.678
        cis = np.where(cis > prec, prec, cis)
        return np.where(na, np.nan, cis / prec)


def meltmu(skyview, stemp, tc):
    """ref meltmu (src/microclimfCpp.cpp:5454-5492): per-cell multiplier of the point model's temperature melt — the
    positive degree-hours of the snow surface temperature blended towards air temperature by the sky view, over those of
    the point model; 1 everywhere when the point model has none."""
    sv = np.asarray(skyview, dtype=np.float64)
    st = np.asarray(stemp, dtype=np.float64)
    t = np.asarray(tc, dtype=np.float64)
    dhp = st[st > 0.0].sum()
    if not dhp > 0.0:
        return np.ones(sv.shape)
    with np.errstate(invalid="ignore"):
        dhm = np.zeros(sv.shape)
        for k in range(st.size):  # sequential sum, as the reference accumulates it
            s2 = (st[k] - t[k]) * sv + t[k]
            dhm = dhm + np.where(s2 > 0.0, s2, 0.0)
        return np.where(np.isnan(sv), np.nan, dhm / dhp)


def snowmodelq1(weather, pmod, subs, dtm, vegp, soilc, snowenv: str = "Taiga", snowinitd: float = 0, snowinita: float = 0,
                zref: float = 2, tfact: float = 0.02, operator=None):
    """ref .snowmodelq1 (R/internal.R:2627-2778), the quick snow model, from the point where the point snow model has run:
    the grid model runs only on the days of `subs` (1-based hours into the full series, whole days); between two modelled
    days the snow balance comes from the point model's melt terms (`pmod`: G, Tc, RswabsG, RlwabsG, umu, tr, sdepc, sdepg,
    sublmelt, tempmelt, rainmelt, sstemp, sdenc, sdeng over the FULL series; sdepc / sdepg one element longer, as
    pointmodelsnow returns them), scaled per cell by meltmu and the canopy interception fraction.  `weather` is the full
    hourly data.frame (`precip` in mm).  The terrain is that of the bare DTM throughout (:2686-2699)."""
    from . import snow as snowops

    op = operator or snowops.gridmodelsnow1
    dtm, vegp, soilc = _unpack(dtm, vegp, soilc)
    subs = np.asarray(subs, dtype=int)
    tme = np.asarray(weather["obs_time"]).astype("datetime64[s]")
    n_full = tme.size
    ot = _obstime(tme)
    ot["hour"] = np.floor(ot["hour"])
    z = dtm.matrix()
    sage = z * 0 + snowinita
    lat, lon = latlong_from_raster(dtm)
    g = lambda k: np.asarray(pmod[k], dtype=np.float64)  # noqa: E731
    w_full = {k: np.asarray(weather[k], dtype=np.float64) for k in WEATHER_COLS}
    # `weather$prec` (R/internal.R:2675) partially matches the `precip` column
    snow = np.where(w_full["temp"] > 2, 0.0, w_full["precip"])
    vg = _sortl(vegp, g("sdepc")[:n_full])
    vg["leaft"] = np.where(np.isnan(vg["leaft"]), 0.01, vg["leaft"])
    ix = subs - 1
    pointm = {"Gp": g("G")[ix], "Tc": g("Tc")[ix], "RswabsG": g("RswabsG")[ix], "RlwabsG": g("RlwabsG")[ix],
              "umu": g("umu")[ix], "tr": g("tr")[ix]}
    wsub = {k: v[ix] for k, v in w_full.items()}
    osub = {k: v[ix] for k, v in ot.items()}
    other: Dict[str, object] = dict(zref=float(zref), lat=lat, lon=lon, isnowdc=snowinitd * (z * 0 + 1),
                                    isnowac=np.nan_to_num(sage).astype(np.int32),
                                    isnowag=np.nan_to_num(sage).astype(np.int32))
    sl = terrain(dtm, "slope").matrix()
    sl[np.isnan(sl)] = 0
    other["slope"] = mask(dtm.like(sl), dtm).matrix()
    ap = terrain(dtm, "aspect").matrix()
    ap[np.isnan(ap)] = 180
    other["aspect"] = mask(dtm.like(ap), dtm).matrix()
    other["hor"], other["skyview"] = api.horizon(z, dtm.res[0], want_svf=True)
    other["wsa"] = _windsheltera(dtm, zref, 10 if dtm.res[0] <= 100 else 1)
    n = ix.size
    shape = z.shape + (n,)
    Tc, Tg, sdepc, sden = (np.full(shape, np.nan) for _ in range(4))
    sdepg = np.zeros(shape)
    pos = snow[snow > 0]
    msnow = pos.mean() if pos.size else np.nan
    intfrac = canintfrac(vg["hgt"], vg["pai"], 2, msnow, wsub["temp"].mean(), 0)
    other["isnowdg"] = (1 - intfrac) * other["isnowdc"]
    climcols = ("temp", "relhum", "pres", "swdown", "difrad", "lwdown", "windspeed", "winddir", "precip")
    ped = 0
    sbtn = None
    for day in range(n // 24):
        s = slice(day * 24, day * 24 + 24)
        first = int(subs[s.start])
        if first - 1 > 1:
            sbtn = np.arange(ped + 1, first)  # 1-based hours between the previous modelled day and this one
            b = sbtn - 1
            mu = meltmu(other["skyview"], g("sstemp")[b], w_full["temp"][b])
            melt = g("sublmelt")[b].sum() + g("rainmelt")[b].sum() + mu * g("tempmelt")[b].sum()
            balancec = (snow[b] / 1000).sum() - melt
            balanceg = (1 - intfrac) * (snow[b] / 1000).sum() - np.exp(-vg["pai"]) * melt
        else:
            balancec = 0.0
            balanceg = 0.0
        if sbtn is None:
            # R evaluates mean(pmod$sdenc[sbtn]) with `sbtn` undefined on a first day that starts the series: an error
            raise ValueError("snowmodelq1: the first modelled day must not start the series (object 'sbtn' not found in R)")
        bb = sbtn - 1
        with np.errstate(invalid="ignore"):
            dc = other["isnowdc"] + balancec * (1000 / g("sdenc")[bb].mean())
            dg = other["isnowdg"] + balanceg * (1000 / g("sdeng")[bb].mean())
            other["isnowdc"] = np.where(dc < 0, 0.0, dc)
            other["isnowdg"] = np.where(dg < 0, 0.0, dg)
        smod = op({k: v[s] for k, v in osub.items()}, {k: wsub[k][s] for k in climcols}, {k: v[s] for k, v in pointm.items()},
                  vg, other, snowenv)
        dsnow = smod["sdepc"] - other["isnowdc"][:, :, None]
        dsnowg = smod["sdepg"] - other["isnowdg"][:, :, None]
        dsnowc = dsnow - dsnowg
        dtms = dtm.like(z + sdepg[:, :, s.stop - 1])  # (still zero for this day: the DTM itself, as written :2745)
        tpr = 10 * np.mean(wsub["windspeed"][s]) ** 0.5
        af = int(round(tpr / dtm.res[0]))
        tpi = _tpicalc(af, min(dtm.nrows, dtm.ncols), dtms, tfact)
        dsnowg2 = dsnowg * tpi[:, :, None]
        dsnowc2 = dsnowc + dsnowg2
        Tc[:, :, s], Tg[:, :, s], sden[:, :, s] = smod["Tc"], smod["Tg"], smod["sden"]
        with np.errstate(invalid="ignore"):
            sdc = dsnowc2 + other["isnowdc"][:, :, None]
            sdg = dsnowg2 + other["isnowdg"][:, :, None]
            sdc = np.where(sdc < 0, 0.0, sdc)
            sdg = np.where(sdg < 0, 0.0, sdg)
        sdepc[:, :, s], sdepg[:, :, s] = sdc, sdg
        ped = int(subs[s.stop - 1])
        other["isnowdc"], other["isnowdg"] = sdc[:, :, 23], sdg[:, :, 23]
    return dict(Tc=Tc, Tg=Tg, groundsnowdepth=sdepg, totalSWE=sdepc * sden, snowden=sden, umu=pointm["umu"])


def meltmu2(mu, stemp, tc):
    """ref meltmu2 (src/microclimfCpp.cpp:5495-5528): meltmu with the snow surface and air temperatures as
    [rows, cols, hours] arrays; 0.5 where a cell's point series has no positive degree-hours."""
    mu = np.asarray(mu, dtype=np.float64)
    st = np.asarray(stemp, dtype=np.float64)
    t = np.asarray(tc, dtype=np.float64)
    dhp = np.zeros(mu.shape)
    dhm = np.zeros(mu.shape)
    with np.errstate(invalid="ignore"):
        for k in range(st.shape[2]):  # sequential sums, as the reference accumulates them
            dhp = dhp + np.where(st[:, :, k] > 0.0, st[:, :, k], 0.0)
            s2 = (st[:, :, k] - t[:, :, k]) * mu + t[:, :, k]
            dhm = dhm + np.where(s2 > 0.0, s2, 0.0)
        out = np.where(dhp > 0.0, dhm / np.where(dhp > 0.0, dhp, 1.0), 0.5)
    return np.where(np.isnan(mu), np.nan, out)


def snowmodelq2(climdata, pointm, pointm2, tme, subs, dtm, dtmc, vegp, soilc, sdept, wuv, wvv, snowenv: str = "Taiga",
                snowinitd: float = 0, snowinita: float = 0, zref: float = 2, tfact: float = 0.02, operator=None):
    """ref .snowmodelq2 (R/internal.R:3017-3290), the quick snow model with gridded climate, from the point where the
    per-coarse-cell point snow models have run and their series have been turned into arrays (:3117-3180).

    `climdata` (temp, relhum, pres, swdown, difrad, lwdown, windspeed, precip as [rows, cols, n] on the DTM, winddir [n])
    and `pointm` (Gp, Tc, RswabsG, RlwabsG, umu, tr as [rows, cols, n]) hold the n = len(subs) modelled hours, `tme`
    their times; `pointm2` holds the FULL series: sstemp and tc on the DTM, sublmelt, tempmelt, rainmelt, snow, sdenc and
    sdeng on the coarse grid `dtmc` (they are summed over each gap and then resampled, `.resamplemelt` :2620); `sdept`
    the hourly maximum of the point models' canopy snow depth over the full series; `wuv`, `wvv` the coarse grid's mean
    wind components of the modelled hours.  Result masked by the DTM (.cleansmod)."""
    from . import snow as snowops

    op = operator or snowops.gridmodelsnow2
    dtm, vegp, soilc = _unpack(dtm, vegp, soilc)
    dtmc = as_raster(dtmc)
    subs = np.asarray(subs, dtype=int)
    ot = _obstime(np.asarray(tme).astype("datetime64[s]"))
    ot["hour"] = np.floor(ot["hour"])
    z = dtm.matrix()
    vg = _sortl(vegp, np.asarray(sdept))
    vg["leaft"] = np.where(np.isnan(vg["leaft"]), 0.01, vg["leaft"])
    lats, lons = latslons_from_raster(dtm)
    other: Dict[str, object] = dict(zref=float(zref), lats=lats, lons=lons, isnowdc=z * 0 + snowinitd,
                                    isnowac=np.nan_to_num(z * 0 + snowinita).astype(np.int32),
                                    isnowag=np.nan_to_num(z * 0 + snowinita).astype(np.int32))
    sl = terrain(dtm, "slope").matrix()
    sl[np.isnan(sl)] = 0
    other["slope"] = mask(dtm.like(sl), dtm).matrix()
    ap = terrain(dtm, "aspect").matrix()
    ap[np.isnan(ap)] = 180
    other["aspect"] = mask(dtm.like(ap), dtm).matrix()
    other["hor"], other["skyview"] = api.horizon(z, dtm.res[0], want_svf=True)
    other["wsa"] = _windsheltera(dtm, zref, 10 if dtm.res[0] <= 100 else 1)
    arr3 = ("temp", "relhum", "pres", "swdown", "difrad", "lwdown", "windspeed", "precip")
    clim = {k: np.asarray(climdata[k], dtype=np.float64) for k in arr3}
    wdir = np.asarray(climdata["winddir"], dtype=np.float64)
    pnt = {k: np.asarray(pointm[k], dtype=np.float64) for k in ("Gp", "Tc", "RswabsG", "RlwabsG", "umu", "tr")}
    p2 = {k: np.asarray(v, dtype=np.float64) for k, v in pointm2.items()}
    wss = np.sqrt(np.asarray(wuv, dtype=np.float64) ** 2 + np.asarray(wvv, dtype=np.float64) ** 2)
    n = subs.size
    shape = z.shape + (n,)
    Tc, Tg, sdepc, sden = (np.full(shape, np.nan) for _ in range(4))
    sdepg = np.zeros(shape)
    with np.errstate(invalid="ignore"):
        pos = p2["snow"][p2["snow"] > 0]
    msnow = np.nanmean(pos) if pos.size else np.nan
    intfrac = canintfrac(vg["hgt"], vg["pai"], 2, msnow, np.nanmean(p2["tc"]), 0)
    other["isnowdg"] = (1 - intfrac) * other["isnowdc"]

    def resamplemelt(a, b):  # ref .resamplemelt (R/internal.R:2620-2624)
        return resample_bilinear(dtmc.like(a[:, :, b].sum(axis=2)), dtm).matrix()

    ped = 0
    for day in range(n // 24):
        s = slice(day * 24, day * 24 + 24)
        first = int(subs[s.start])
        if first - 1 > 1:
            b = np.arange(ped + 1, first) - 1  # 0-based hours between the previous modelled day and this one
            mu = meltmu2(other["skyview"], p2["sstemp"][:, :, b], p2["tc"][:, :, b])
            melt = resamplemelt(p2["sublmelt"], b) + resamplemelt(p2["rainmelt"], b) + mu * resamplemelt(p2["tempmelt"], b)
            snowsum = resamplemelt(p2["snow"], b)
            balancec = snowsum / 1000 - melt
            balanceg = (1 - intfrac) * snowsum / 1000 - np.exp(-vg["pai"]) * melt
            sdec = resamplemelt(p2["sdenc"], b) / b.size
            sdeg = resamplemelt(p2["sdeng"], b) / b.size
            other["isnowdc"] = other["isnowdc"] + balancec * (1000 / sdec)
            other["isnowdg"] = other["isnowdg"] + balanceg * (1000 / sdeg)
        with np.errstate(invalid="ignore"):
            other["isnowdc"] = np.where(other["isnowdc"] < 0, 0.0, other["isnowdc"])
            other["isnowdg"] = np.where(other["isnowdg"] < 0, 0.0, other["isnowdg"])
        smod = op({k: v[s] for k, v in ot.items()}, dict({k: v[:, :, s] for k, v in clim.items()}, winddir=wdir[s]),
                  {k: v[:, :, s] for k, v in pnt.items()}, vg, other, snowenv)
        dsnow = smod["sdepc"] - other["isnowdc"][:, :, None]
        dsnowg = smod["sdepg"] - other["isnowdg"][:, :, None]
        dsnowc = dsnow - dsnowg
        dtms = dtm.like(z + sdepg[:, :, s.stop - 1])
        tpr = 10 * np.mean(wss[s]) ** 0.5
        af = int(round(tpr / dtm.res[0]))
        tpi = _tpicalc(af, min(dtm.nrows, dtm.ncols), dtms, tfact)
        dsnowg2 = dsnowg * tpi[:, :, None]
        dsnowc2 = dsnowc + dsnowg2
        Tc[:, :, s], Tg[:, :, s], sden[:, :, s] = smod["Tc"], smod["Tg"], smod["sden"]
        with np.errstate(invalid="ignore"):
            sdc = dsnowc2 + other["isnowdc"][:, :, None]
            sdg = dsnowg2 + other["isnowdg"][:, :, None]
            sdc = np.where(sdc < 0, 0.0, sdc)
            sdg = np.where(sdg < 0, 0.0, sdg)
        sdepc[:, :, s], sdepg[:, :, s] = sdc, sdg
        ped = int(subs[s.stop - 1])
        other["isnowdc"], other["isnowdg"] = sdc[:, :, 23], sdg[:, :, 23]
    out = dict(Tc=Tc, Tg=Tg, groundsnowdepth=sdepg, totalSWE=sdepc * sden, snowden=sden, umu=pnt["umu"])
    land = ~np.isnan(z)
    return {k: np.where(land[:, :, None], v, np.nan) for k, v in out.items()}


def runsnowmodel(weather, micropoint, pmod, vegp, soilc, dtm, snowenv: str = "Taiga", method: str = "fast", snowinitd: float = 0,
                 snowinita: float = 0, zref: float = 2, stfact: float = 0.01, arrays=None, operator=None):
    """ref runsnowmodel (R/Cppwrappers.R:718-760): the snow model behind `runmicro(snow = TRUE)`.  The point snow model is
    upstream of this build, so its output travels as an argument: `pmod` = pointmodelsnow's list over the full series
    (G, Tc, RswabsG, RlwabsG, umu, tr, sdepc, sdepg, and for the quick model sublmelt, tempmelt, rainmelt, sstemp, sdenc,
    sdeng).

    data.frame climate (`micropoint` a Micropoint): the full model (`snowmodel1`) when the point model covers every hour;
    with a subset point model `method = "fast"` runs the quick model on the subset days (`snowmodelq1`), `"slow"` the
    full model followed by `subsetsnowmodel` — exactly the reference's dispatch, including which zref each branch uses.
    Gridded climate (`micropoint` a list): `arrays` holds the keyword arguments of `snowmodel2` / `snowmodelq2` that
    `.snowmodel2` / `.snowmodelq2` build from the rasters (climdata, pointm, tme, sdept, wuv, wvv, ... — see those
    functions); the dispatch is the same."""
    vegp = _cleanvegp({k: as_raster(v) for k, v in vegp.items()})
    if isinstance(micropoint, Micropoint):
        g = lambda k: np.asarray(pmod[k], dtype=np.float64)  # noqa: E731
        n = len(weather["temp"])
        pointm = dict(Gp=g("G"), Tc=g("Tc"), RswabsG=g("RswabsG"), RlwabsG=g("RlwabsG"), umu=g("umu"), tr=g("tr"),
                      sdepc=g("sdepc")[:n])
        full = len(micropoint.subs) == len(micropoint.tmeorig)
        if full:
            return snowmodel1(weather, pointm, dtm, vegp, soilc, snowenv, snowinitd, snowinita, micropoint.zref, stfact,
                              operator=operator)
        if method == "fast":
            return snowmodelq1(weather, pmod, micropoint.subs, dtm, vegp, soilc, snowenv, snowinitd, snowinita, zref, stfact,
                               operator=operator)
        smod = snowmodel1(weather, pointm, dtm, vegp, soilc, snowenv, snowinitd, snowinita, zref, stfact, operator=operator)
        return subsetsnowmodel(smod, micropoint.subs)
    if arrays is None:
        raise ValueError("gridded climate: pass the prepared arrays of snowmodel2 / snowmodelq2 as `arrays`")
    one = next(m for m in micropoint if m is not None)
    kw = dict(arrays)
    full = len(one.subs) == len(one.tmeorig)
    if full:
        return snowmodel2(dtm=dtm, vegp=vegp, soilc=soilc, snowenv=snowenv, snowinitd=snowinitd, snowinita=snowinita,
                          zref=micropoint[0].zref if micropoint[0] is not None else one.zref, tfact=stfact, operator=operator, **kw)
    if method == "fast":
        return snowmodelq2(subs=one.subs, dtm=dtm, vegp=vegp, soilc=soilc, snowenv=snowenv, snowinitd=snowinitd,
                           snowinita=snowinita, zref=zref, tfact=stfact, operator=operator, **kw)
    smod = snowmodel2(dtm=dtm, vegp=vegp, soilc=soilc, snowenv=snowenv, snowinitd=snowinitd, snowinita=snowinita, zref=zref,
                      tfact=stfact, operator=operator, **kw)
    return subsetsnowmodel(smod, one.subs)


# ---------------------------------------------------------------------------------------------
# runmicro(snow = TRUE), data.frame climate: .runmicrosnow1
# ---------------------------------------------------------------------------------------------
def subsetsnowmodel(snowmod, subs):
    """ref subsetsnowmodel (R/dataprep.R:148-159); `subs` 1-based hours."""
    ix = np.asarray(subs, dtype=int) - 1
    out = {k: np.asarray(snowmod[k])[:, :, ix] for k in ("Tc", "Tg", "groundsnowdepth", "totalSWE", "snowden")}
    u = np.asarray(snowmod["umu"])
    out["umu"] = u[:, :, ix] if u.ndim == 3 else u[ix]
    return out


def _sortl2(vegp, sdep, reqhgt, pai_a):
    """ref .sortl2 (R/internal.R:2422-2477): snow-hour vegetation averages + foliage density."""
    out = {}
    for k in ("pai", "hgt", "leaft", "clump", "leafd"):
        a = vegp[k].values
        dmx = a.shape[2]
        if dmx > 1:
            n = len(sdep)
            s = np.clip(_r_round(np.linspace(0.50001, dmx + 0.5, n)).astype(int), 1, dmx)
            sel = np.nonzero(np.asarray(sdep) > 0)[0]
            s = s[sel] if sel.size else s[:1]
            num, fre = np.unique(s, return_counts=True)
            m = a[:, :, 0] * 0
            for j in range(len(num)):
                m = m + a[:, :, j] * fre[j]  # layer j, not num[j], as written (R/internal.R:2446)
            out[k] = m / fre.sum()
        else:
            out[k] = a[:, :, 0].copy()
    paia = None if pai_a is None else as_raster(pai_a, vegp["hgt"]).values.mean(axis=2)
    fd = _foliageden(reqhgt, out["hgt"], out["pai"], paia)
    out["paia"], out["leafden"] = fd["pai_a"], fd["leafden"]
    return out


def _hours_of_days(days):
    d = np.asarray(days, dtype=int)
    return np.repeat((d - 1) * 24, 24) + np.tile(np.arange(1, 25), d.size)


def runmicrosnow1(micropoint: Micropoint, reqhgt, vegp, soilc, dtm, smod, runchecks=True, pai_a=None, tfact=1.5,
                  out=(True,) * 10, slr=None, apr=None, hor=None, twi=None, wsa=None, svf=None, snow_operator=None):
    """ref .runmicrosnow1 (R/internal.R:3580-3660): days without snow anywhere go through the ordinary grid model, days
    with snow somewhere through gridmicrosnow1 (seeded with the ordinary model's values where a day is in both sets),
    and the two are merged by day.  `smod` is the snow model's output (snowmodel1)."""
    from . import snow as snowops

    op = snow_operator or snowops.gridmicrosnow1
    dtm_r = as_raster(dtm)
    smod, snowdays, nosnowdays = _snow_day_sets(smod, dtm_r)
    micropoints = subsetpointmodel(micropoint, days=snowdays) if snowdays.size else None
    if nosnowdays.size:
        micropointn = subsetpointmodel(micropoint, days=nosnowdays)
        moutn = prepare_model(micropointn, vegp, soilc, dtm, reqhgt, runchecks, pai_a, tfact, out, slr, apr, hor, twi,
                              wsa).run()
    else:  # .createblanktemplate1
        micropointn = subsetpointmodel(micropoint, days=[1])
        moutn = prepare_model(micropointn, vegp, soilc, dtm, reqhgt, False, None, 1.5, out).run()
        moutn = {k: a * np.nan for k, a in moutn.items()}
    if not snowdays.size:
        return moutn
    # ---- .prepsnowinputs1
    dtm_u, vegp_u, soilc_u = _unpack(dtm, vegp, soilc)
    vegp_u, dtm_u, soilc_u = _cleanvars(vegp_u, soilc_u, dtm_u)
    weather = micropoints.weather
    if runchecks:
        rc = checkinputs(weather, vegp_u, soilc_u, dtm_u)
        weather, vegp_u, soilc_u = rc["weather"], rc["vegp"], rc["soilc"]
    ai = _hours_of_days(snowdays)
    obstime = _obstime(weather["obs_time"])
    climdata = {k: np.asarray(weather[k], dtype=np.float64) for k in WEATHER_COLS}
    climdata["umu"] = np.asarray(smod["umu"], dtype=np.float64)[ai - 1]
    sdept = np.zeros(len(micropoint.tmeorig))
    sdept[np.asarray(micropoints.subs, dtype=int) - 1] = 1
    vg = _sortl2(vegp_u, sdept, reqhgt, pai_a)
    other: Dict[str, object] = {}
    other["slope"] = (terrain(dtm_u, "slope") if slr is None else as_raster(slr, dtm_u)).matrix()
    other["aspect"] = (terrain(dtm_u, "aspect") if apr is None else as_raster(apr, dtm_u)).matrix()
    lat, lon = latlong_from_raster(dtm_u)
    if hor is None:
        other["hor"], sv = api.horizon(dtm_u.matrix(), dtm_u.res[0], want_svf=True)
    else:
        other["hor"] = np.asarray(hor, dtype=np.float64)
        sv = 0.5 * np.cos(2 * np.tan(np.mean(np.arctan(other["hor"]), axis=2))) + 0.5
    other["skyview"] = sv if svf is None else as_raster(svf, dtm_u).matrix()
    other["wsa"] = (_windsheltera(dtm_u, micropoint.zref, 10 if dtm_u.res[0] <= 100 else 1) if wsa is None
                    else np.asarray(wsa, dtype=np.float64))
    other["lat"], other["lon"], other["zref"] = lat, lon, micropoints.zref  # `ll$lon` is NULL in R (the column is `long`)
    other["Smax"] = _soilinit(soilc_u)["Smax"]
    micros = _seed_snow_micro(moutn, snowdays, nosnowdays)
    smods = subsetsnowmodel(smod, ai)
    outm = _snow_out_mask(out, reqhgt, micros)
    # `micros` was seeded just above and is ours: the product operator may update it in place (no host copies)
    kw = {} if snow_operator is not None else {"copy": False}
    mouts = op(reqhgt, obstime, climdata, smods, micros, vg, other, micropoint.matemp, outm, **kw)
    if not nosnowdays.size:
        return mouts
    return _merge_snow_days(moutn, mouts, snowdays, nosnowdays)


# ---------------------------------------------------------------------------------------------
# runmicro(snow = TRUE), gridded climate: .runmicrosnow2
# ---------------------------------------------------------------------------------------------
def subsetpointmodela(pointmodela, tstep: str = "month", what: str = "tmax", days=None):
    """ref subsetpointmodela (R/dataprep.R:114-133): every point model of the coarse grid subset to the same days,
    chosen on the grid-mean canopy temperature."""
    live = [mp for mp in pointmodela if mp is not None]
    Tc = sum(np.asarray(mp.dfo["Tc"], dtype=np.float64) for mp in live) / len(live)
    return [None if mp is None else subsetpointmodel(mp, tstep, what, days, Tc=Tc) for mp in pointmodela]


def _snow_day_sets(smod, dtm_r):
    """Days with snow somewhere / without snow somewhere (applycpp3 min / max over space per hour,
    src/microclimfCpp.cpp:5553, and snowdaysfun :5531) from the masked total SWE; returns the cleaned snow model too."""
    smod = dict(smod)
    swe = np.array(smod["totalSWE"], dtype=np.float64)
    swe[np.isnan(swe)] = 0
    swe[np.isnan(dtm_r.matrix()), :] = np.nan  # mask(totalSWE, dtm)
    smod["totalSWE"] = swe
    with np.errstate(all="ignore"):
        minsnow = np.nanmin(swe, axis=(0, 1))
        maxsnow = np.nanmax(swe, axis=(0, 1))
    ndays = swe.shape[2] // 24
    snowflag = (maxsnow[:ndays * 24].reshape(ndays, 24) > 0.0).any(axis=1)
    nosnowflag = (minsnow[:ndays * 24].reshape(ndays, 24) == 0.0).any(axis=1)
    v = np.arange(1, ndays + 1)
    return smod, v[snowflag], v[nosnowflag]


def _seed_snow_micro(moutn, snowdays, nosnowdays):
    """Blank [rows, cols, snow hours] arrays seeded with the snow-free model on days present in both sets
    (R/internal.R:3430-3443 / :3561-3577)."""
    t1 = snowdays.size * 24
    s1 = np.arange(t1)[np.repeat(np.isin(snowdays, nosnowdays), 24)]
    s2 = np.arange(nosnowdays.size * 24)[np.repeat(np.isin(nosnowdays, snowdays), 24)]
    micros = {}
    for k, a in moutn.items():
        vv = np.full(a.shape[:2] + (t1,), np.nan)
        vv[:, :, s1] = a[:, :, s2]
        micros[k] = vv
    return micros


def _merge_snow_days(moutn, mouts, snowdays, nosnowdays):
    """ref R/internal.R:3634-3656 / :3717-3742: snow days from the snow operator, the other days from the ordinary model."""
    nosnow = np.setdiff1d(np.union1d(snowdays, nosnowdays), snowdays)
    s1 = np.arange(next(iter(moutn.values())).shape[2])[np.repeat(np.isin(nosnowdays, nosnow), 24)]
    nosnowh, snowh = _hours_of_days(nosnow) - 1, _hours_of_days(snowdays) - 1
    n = nosnowh.size + snowh.size
    mout = {}
    for k, a in moutn.items():
        m = np.full(a.shape[:2] + (n,), np.nan)
        m[:, :, nosnowh] = a[:, :, s1]
        if k in mouts:
            m[:, :, snowh] = mouts[k]
        mout[k] = m
    return mout


def _snow_out_mask(out, reqhgt, micros):
    outm = [bool(o) for o in out]
    if reqhgt == 0:
        outm = [i in (0, 3, 5, 6, 7, 8, 9) for i in range(10)]
    elif reqhgt < 0:
        outm = [i in (0, 3) for i in range(10)]
    names = ("Tz", "tleaf", "relhum", "soilm", "windspeed", "Rdirdown", "Rdifdown", "Rlwdown", "Rswup", "Rlwup")
    return [o and (n in micros) for o, n in zip(outm, names)]


def prepsnowinputs2(reqhgt, dtm, dtmc, vegp, soilc, micropoints, altcorrect, runchecks, snowdays, nosnowdays, moutn,
                    slr=None, apr=None, hor=None, svf=None, wsa=None, pai_a=None):
    """ref .prepsnowinputs2 (R/internal.R:3445-3579): the fine-raster [rows, cols, hours] climate arrays gridmicrosnow2
    takes (`.cca` + terra::resample of every series of the coarse grid, with the altitude correction of pressure and
    temperature), the snow-hour vegetation averages, terrain layers and the seeded blank microclimate arrays."""
    dtm, vegp, soilc = _unpack(dtm, vegp, soilc)
    dtmc = as_raster(dtmc)
    cr, cc = dtmc.nrows, dtmc.ncols
    weathers, last = [], None
    for mp in micropoints:
        if mp is None:
            weathers.append(None)
            continue
        w = mp.weather
        if runchecks:
            rc = checkinputs(w, vegp, soilc, dtm, mp.zref)
            w, vegp, soilc = rc["weather"], rc["vegp"], rc["soilc"]
        weathers.append(w)
        last = mp
    if last is None:
        raise ValueError("micropoints holds no point-model output")
    obstime = _obstime(last.weather["obs_time"])
    h = len(last.weather["temp"])

    def cca(series):
        a = np.full((cr, cc, h), np.nan)
        k = 0
        for i in range(cr):
            for j in range(cc):
                if micropoints[k] is not None:
                    a[i, j, :] = np.asarray(series(k), dtype=np.float64)
                k += 1
        return a

    fine = lambda a: resample_bilinear(dtmc.like(a), dtm).values  # noqa: E731  .cca(..., dtmc, dtm)
    wv_ = lambda name: cca(lambda k: weathers[k][name])  # noqa: E731
    tc, pk, rh = wv_("temp"), wv_("pres"), wv_("relhum")
    ea = fine(_satvap(tc) * (rh / 100))
    clim: Dict[str, np.ndarray] = {"temp": fine(tc)}
    if altcorrect == 0:
        clim["pres"] = fine(pk)
    else:
        zc = dtmc.matrix().copy()
        zc[np.isnan(zc)] = 0.0
        zf = dtm.matrix()
        psl = fine(pk / (((293 - 0.0065 * zc[:, :, None]) / 293) ** 5.26))
        clim["pres"] = psl * (((293 - 0.0065 * zf[:, :, None]) / 293) ** 5.26)
        elevd = (resample_bilinear(dtmc.like(zc), dtm).matrix() - zf)[:, :, None]
        if altcorrect == 1:
            tcdif = elevd * (5 / 1000)
        else:  # .lapserate (R/internal.R:546-553)
            tk = clim["temp"] + 273.15
            rv = 0.622 * ea / (clim["pres"] - ea)
            lr = 9.8076 * (1 + (2501000 * rv) / (287 * tk)) / (1003.5 + (0.622 * 2501000 ** 2 * rv) / (287 * tk ** 2))
            tcdif = lr * elevd
        clim["temp"] = tcdif + clim["temp"]
    with np.errstate(invalid="ignore"):
        clim["relhum"] = np.clip((ea / _satvap(clim["temp"])) * 100, 20, 100)
    for k in ("swdown", "difrad", "lwdown"):
        clim[k] = fine(wv_(k))
    u2, wd = wv_("windspeed"), wv_("winddir")
    wu, wvc = u2 * np.cos(wd * np.pi / 180), u2 * np.sin(wd * np.pi / 180)
    clim["windspeed"] = np.sqrt(fine(wu) ** 2 + fine(wvc) ** 2)
    clim["winddir"] = np.mod(np.arctan2(np.nanmean(wvc, axis=(0, 1)), np.nanmean(wu, axis=(0, 1))) * 180 / np.pi, 360)
    clim["precip"] = fine(wv_("precip"))  # `climdata$prec` in the reference; the operator does not read it
    clim["umu"] = fine(cca(lambda k: micropoints[k].dfo["umu"]))
    sdept = np.zeros(len(last.tmeorig))
    sdept[np.asarray(last.subs, dtype=int) - 1] = 1
    vg = _sortl2(vegp, sdept, reqhgt, pai_a)
    other: Dict[str, object] = {}
    other["slope"] = (terrain(dtm, "slope") if slr is None else as_raster(slr, dtm)).matrix()
    other["aspect"] = (terrain(dtm, "aspect") if apr is None else as_raster(apr, dtm)).matrix()
    if hor is None:
        other["hor"], sv = api.horizon(dtm.matrix(), dtm.res[0], want_svf=True)
    else:
        other["hor"] = np.asarray(hor, dtype=np.float64)
        sv = 0.5 * np.cos(2 * np.tan(np.mean(np.arctan(other["hor"]), axis=2))) + 0.5
    other["skyview"] = sv if svf is None else as_raster(svf, dtm).matrix()
    other["wsa"] = (_windsheltera(dtm, last.zref, 10 if dtm.res[0] <= 100 else 1) if wsa is None
                    else np.asarray(wsa, dtype=np.float64))
    other["lats"], other["lons"] = latslons_from_raster(dtm)
    other["zref"] = last.zref
    other["Smax"] = _soilinit(soilc)["Smax"]
    return dict(obstime=obstime, weather=clim, micro=_seed_snow_micro(moutn, snowdays, nosnowdays), vegp=vg, other=other)


def runmicrosnow2(micropoint, reqhgt, vegp, soilc, dtm, dtmc, smod, altcorrect=0, runchecks=True, pai_a=None, tfact=1.5,
                  out=(True,) * 10, slr=None, apr=None, hor=None, twi=None, wsa=None, svf=None, snow_operator=None):
    """ref .runmicrosnow2 (R/internal.R:3661-3745): runmicro(snow = TRUE) with gridded climate.  `micropoint` is
    runpointmodela's list (one Micropoint or None per cell of `dtmc`), `smod` the snow model's output (snowmodel2).
    Days without snow anywhere go through the gridded-climate grid model (prepare_model_a: coarse arrays interpolated
    in the kernels), days with snow somewhere through gridmicrosnow2 on the fine arrays prepsnowinputs2 builds."""
    from . import snow as snowops

    op = snow_operator or snowops.gridmicrosnow2
    dtm_r = as_raster(dtm)
    smod, snowdays, nosnowdays = _snow_day_sets(smod, dtm_r)
    micropoints = subsetpointmodela(micropoint, days=snowdays) if snowdays.size else None
    matemp = float(np.mean([mp.matemp for mp in micropoint if mp is not None]))
    if nosnowdays.size:
        micropointn = subsetpointmodela(micropoint, days=nosnowdays)
        moutn = prepare_model_a(micropointn, vegp, soilc, dtm, dtmc, reqhgt, runchecks, altcorrect, pai_a, tfact, out, slr,
                                apr, hor, twi, wsa).run()
    else:  # .createblanktemplate2
        micropointn = subsetpointmodela(micropoint, days=[1])
        moutn = prepare_model_a(micropointn, vegp, soilc, dtm, dtmc, reqhgt, False, 0, None, 1.5, out).run()
        moutn = {k: a * np.nan for k, a in moutn.items()}
    if not snowdays.size:
        return moutn
    sin = prepsnowinputs2(reqhgt, dtm, dtmc, vegp, soilc, micropoints, altcorrect, runchecks, snowdays, nosnowdays, moutn,
                          slr, apr, hor, svf, wsa, pai_a)
    smods = subsetsnowmodel(smod, _hours_of_days(snowdays))
    outm = _snow_out_mask(out, reqhgt, sin["micro"])
    kw = {} if snow_operator is not None else {"copy": False}  # sin["micro"] is ours (prepsnowinputs2 built it): update in place
    mouts = op(reqhgt, sin["obstime"], sin["weather"], smods, sin["micro"], sin["vegp"], sin["other"], matemp, outm, **kw)
    if not nosnowdays.size:
        return mouts
    return _merge_snow_days(moutn, mouts, snowdays, nosnowdays)


# ---------------------------------------------------------------------------------------------
# runbioclim (data.frame climate): .runbioclim1 / .runbioclim3
# ---------------------------------------------------------------------------------------------
def _biosel(tme, tc):
    """ref .biosel (R/internal.R:1690-1727): the 14 days of the bioclim run — each month's median-temperature day, then
    the (median over years of the) hottest and coldest day.  Returns (selh, seld), 1-based."""
    tc = np.asarray(tc, dtype=np.float64)
    nd = tc.size // 24
    tcd = tc[:nd * 24].reshape(nd, 24).mean(axis=1)
    t = np.asarray(tme).astype("datetime64[s]").astype(np.int64)[:nd * 24].reshape(nd, 24).mean(axis=1)
    tmd = _obstime(t.astype("int64").astype("datetime64[s]"))
    sel_med = []
    for mth in range(1, 13):
        s = np.nonzero(tmd["month"] == mth)[0] + 1
        o = np.argsort(tcd[s - 1], kind="stable") + 1
        n = len(o) // 2
        sel_med.append(int(s[0] - 1 + o[n - 1]))
    yrs = list(dict.fromkeys(tmd["year"].tolist()))
    sel_max, sel_min = [], []
    for y in yrs:
        s = np.nonzero(tmd["year"] == y)[0] + 1
        sel_max.append(int(np.argmax(tcd[s - 1]) + s[0]))
        sel_min.append(int(np.argmin(tcd[s - 1]) + s[0]))
    sel_max = np.array(sel_max)[np.argsort(tcd[np.array(sel_max) - 1], kind="stable")]
    sel_min = np.array(sel_min)[np.argsort(tcd[np.array(sel_min) - 1], kind="stable")]
    n = len(sel_max) // 2
    seld = np.array(sel_med + [int(sel_max[n]), int(sel_min[n])])
    return _hours_of_days(seld), seld


def _quarter(agg, fun):
    """which.max / which.min of stats::filter(agg, rep(1/3, 3), sides = 2, circular = TRUE): 1-based centre month."""
    a = np.asarray(agg, dtype=np.float64)
    f = (np.roll(a, 1) + a + np.roll(a, -1)) / 3.0
    return int(fun(f)) + 1


def _getselq(iq, month):
    """ref .getselq (R/internal.R:1764-1774): 1-based hours of the three months around month iq."""
    imn, imx = (12 if iq - 1 == 0 else iq - 1), (1 if iq + 1 == 13 else iq + 1)
    return np.sort(np.nonzero((month == imn) | (month == iq) | (month == imx))[0] + 1)


def _sortvegp2(vegp, seld, vegpisannual, n):
    """ref .sortvegp2 (R/internal.R:276-302): the vegetation layer in force on each of the 14 selected days."""
    if vegpisannual:
        sd, nd = np.asarray(seld) % 365, 365
    else:
        sd, nd = np.asarray(seld), int(round(n / 24))
    out = {}
    for k in VEG_NAMES:
        a = vegp[k].values
        dmx = a.shape[2]
        if dmx == 1:
            out[k] = np.repeat(a, 14, axis=2)
        else:
            s = np.clip(_r_round(np.linspace(0.50001, dmx + 0.5, nd)).astype(int), 1, dmx)
            s = np.append(s, s[-1])
            # R subscripts: element 0 of `seld %% 365` selects nothing in R; the bundled data never produce it
            out[k] = a[:, :, s[np.maximum(sd, 1) - 1] - 1]
    return out


def runbioclim(climdata, reqhgt, vegp, soilc, dtm, pointmodel, temp="air", zref=2, windhgt=None, soilm=None, runchecks=True,
               pai_a=None, tfact=1.5, out=(True,) * 19, vegpisannual=True, operator=None):
    """ref runbioclim (R/Cppwrappers.R:628-651) for data.frame climate -> .runbioclim1 (static vegetation,
    R/internal.R:1776-1894) / .runbioclim3 (layered vegetation, :2082-2181): quarters from the monthly weather, the 14
    bioclim days, the point model on those 336 hours, the usual static layers, then the fused CUDA operator
    (`api.runbioclim1Cpp` / `runbioclim3Cpp`: grid solve + 19 reductions on the device).

    The point model is upstream of this build: `pointmodel(weather336, reqhgt, dtm, vegp, soilc, zref,
    windhgt, soilm)` must return a `Micropoint` for the 336 selected hours, as `runpointmodel(..., yearG = FALSE)` does
    (R/internal.R:1747).  Returns {"bio1": [rows, cols], ...} masked by the DTM."""
    dtm_u, vegp_u, soilc_u = _unpack(dtm, vegp, soilc)
    layered = _vegpdmx(vegp_u) > 1
    vegp_u, dtm_u, soilc_u = _cleanvars(vegp_u, soilc_u, dtm_u)
    weather = climdata
    if runchecks:
        rc = checkinputs(weather, vegp_u, soilc_u, dtm_u)
        weather, vegp_u, soilc_u = rc["weather"], rc["vegp"], rc["soilc"]
    tme = np.asarray(weather["obs_time"]).astype("datetime64[s]")
    ot = _obstime(tme)
    n = tme.size
    months = sorted(set(ot["month"].tolist()))
    pmean = [np.nanmean(np.asarray(weather["precip"], dtype=np.float64)[ot["month"] == m]) for m in months]
    tsum = [np.nansum(np.asarray(weather["temp"], dtype=np.float64)[ot["month"] == m]) for m in months]
    wq, dq = _quarter(pmean, np.argmax), _quarter(pmean, np.argmin)
    hq, cq = _quarter(tsum, np.argmax), _quarter(tsum, np.argmin)
    selh, seld = _biosel(tme, weather["temp"])
    w336 = {k: np.asarray(v)[selh - 1] for k, v in weather.items()}
    micropoint = pointmodel(w336, reqhgt, dtm_u, vegp_u, soilc_u, zref, zref if windhgt is None else windhgt, soilm)
    weather = micropoint.weather
    obstime = _obstime(weather["obs_time"])
    tempv = np.asarray(weather["temp"], dtype=np.float64)
    es = _satvap(tempv)
    ea = es * np.asarray(weather["relhum"], dtype=np.float64) / 100
    clim = dict(temp=tempv, es=es, ea=ea, tdew=_dewpoint(ea, tempv))
    for k in ("pres", "swdown", "difrad", "lwdown", "windspeed", "winddir"):
        clim[k] = np.asarray(weather[k], dtype=np.float64)
    pointm = {k: np.asarray(v, dtype=np.float64) for k, v in micropoint.dfo.items()}
    pointm["Tbp"] = np.asarray(micropoint.Tbz, dtype=np.float64) if reqhgt < 0 else np.zeros(tempv.size)
    if layered:
        vg = _sortvegp2(vegp_u, seld, vegpisannual, n)
    else:
        vg = {k: vegp_u[k].values[:, :, 0].copy() for k in VEG_NAMES}
        with np.errstate(invalid="ignore"):
            vg["hgt"] = np.where(vg["pai"] == 0, 0.0, vg["hgt"])
            vg["pai"] = np.where(vg["hgt"] == 0, 0.0, vg["pai"])
    fd = _foliageden(reqhgt, vg["hgt"], vg["pai"], None if pai_a is None else as_raster(pai_a, dtm_u).values.squeeze())
    vg["paia"], vg["leafden"] = fd["pai_a"], fd["leafden"]
    soilp = _soilinit(soilc_u)
    sc = dict(gref=soilc_u["groundr"].matrix().copy(), Smin=soilp["Smin"], Smax=soilp["Smax"], soilb=soilp["soilb"],
              Psie=soilp["psi_e"], Vq=soilp["Vq"], Vm=soilp["Vm"], Mc=soilp["Mc"], rho=soilp["rho"])

    def fill(r, v):
        m = r.matrix().copy()
        m[np.isnan(m)] = v
        return mask(r.like(m), dtm_u).matrix()

    sc["slope"], sc["aspect"] = fill(terrain(dtm_u, "slope"), 0.0), fill(terrain(dtm_u, "aspect"), 0.0)
    sc["twi"] = fill(_topidx(dtm_u), 1.0)
    lat, lon = latlong_from_raster(dtm_u)
    sc["hor"], sc["svfa"] = api.horizon(dtm_u.matrix(), dtm_u.res[0], want_svf=True)
    sc["wsa"] = _windsheltera(dtm_u, micropoint.zref, 10 if dtm_u.res[0] <= 100 else 1)
    month = obstime["month"]
    q = [(_getselq(x, month) - 1).astype(np.int32) for x in (wq, dq, hq, cq)]
    fn = operator or (api.runbioclim3Cpp if layered else api.runbioclim1Cpp)  # tests inject the reference's
    bio = fn(obstime, clim, pointm, vg, sc, float(reqhgt), float(micropoint.zref), lat, lon, _getmode(sc["Smin"]),
             _getmode(sc["Smax"]), float(tfact), float(micropoint.matemp), [bool(o) for o in out], q[0], q[1], q[2], q[3],
             temp == "air")
    na = np.isnan(dtm_u.matrix())
    return {k: np.where(na, np.nan, v) for k, v in bio.items()}


def runbioclim_a(micropointa, prech, tcmean, tme, reqhgt, vegp, soilc, dtm, dtmc, temp="air", runchecks=True, altcorrect=0,
                 pai_a=None, tfact=1.5, out=(True,) * 19, vegpisannual=True, operator=None):
    """ref runbioclim with gridded climate -> .runbioclim2 (static vegetation, R/internal.R:1896-2081) / .runbioclim4
    (time-variant vegetation, :2200-2388).

    The point models are upstream of this build: `micropointa` is what `.biomicropoint` returns (:1921) — one Micropoint
    for the 14 bioclim days (336 h) per cell of the coarse raster `dtmc`, None for sea cells.  `prech` and `tcmean` are
    the space means of the climate array's precipitation and temperature for every hour of the full series and `tme`
    its times (they choose the quarters, :1909-1917).  The coarse series go to the fused CUDA operator as they are
    (interpolated in the kernels) instead of being expanded to [rows, cols, 336] arrays on the host."""
    dtm_u, vegp_u, soilc_u = _unpack(dtm, vegp, soilc)
    layered = _vegpdmx(vegp_u) > 1
    ot_full = _obstime(np.asarray(tme).astype("datetime64[s]"))
    months = sorted(set(ot_full["month"].tolist()))
    pr, tc = np.asarray(prech, dtype=np.float64), np.asarray(tcmean, dtype=np.float64)
    pmean = [np.nanmean(pr[ot_full["month"] == m]) for m in months]
    tsum = [np.nansum(tc[ot_full["month"] == m]) for m in months]
    wq, dq = _quarter(pmean, np.argmax), _quarter(pmean, np.argmin)
    hq, cq = _quarter(tsum, np.argmax), _quarter(tsum, np.argmin)
    call = prepare_model_a(micropointa, vegp, soilc, dtm, dtmc, reqhgt, runchecks, altcorrect, pai_a, tfact)
    prob = call.problem()
    if prob.tsteps != 336:
        raise ValueError("micropointa must hold the 14 bioclim days (336 hours)")
    if layered:
        # .runbioclim4 (R/internal.R:2200-2388): the vegetation layer in force on each of the 14 days (.sortvegp2), one
        # day-block per layer, foliage density per layer
        vegp_c, dtm_c, soilc_c = _cleanvars(vegp_u, soilc_u, dtm_u)
        if runchecks:
            one = next(m for m in micropointa if m is not None)
            vegp_c = checkinputs(one.weather, vegp_c, soilc_c, dtm_c, one.zref)["vegp"]
        _, seld = _biosel(np.asarray(tme).astype("datetime64[s]"), tc)
        vg = _sortvegp2(vegp_c, seld, vegpisannual, pr.size)
        fd = _foliageden(reqhgt, vg["hgt"], vg["pai"], None if pai_a is None else as_raster(pai_a, dtm_u).values.squeeze())
        vg["paia"], vg["leafden"] = fd["pai_a"], fd["leafden"]
        prob.mode, prob.nlyr = 4, 14
        prob.lyr_st = (np.arange(14) * 24).astype(np.int32)
        prob.lyr_ed = (np.arange(14) * 24 + 23).astype(np.int32)
        for k, v in vg.items():
            prob.set(k, v)
        prob.validate()
    month = np.asarray(_obstime(next(m for m in micropointa if m is not None).weather["obs_time"])["month"])
    q = [(_getselq(x, month) - 1).astype(np.int32) for x in (wq, dq, hq, cq)]
    fn = operator or api.run_bioclim_problem  # tests inject the reference's on the expanded arrays
    bio = fn(prob, q[0], q[1], q[2], q[3], temp == "air", [bool(o) for o in out])
    na = np.isnan(dtm_u.matrix())
    return {k: np.where(na, np.nan, v) for k, v in bio.items()}

