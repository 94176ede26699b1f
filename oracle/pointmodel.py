"""TEST INFRASTRUCTURE ONLY — the reference's point model, driven the way `runpointmodel` drives it.

The grid solver takes a `micropoint` object as INPUT (R/Cppwrappers.R:376): the output of
`runpointmodel` (R/Cppwrappers.R:59-148), an iterative big-leaf model that SURVEY.md §2 puts out of
scope for the CUDA build.  To exercise the host layer on the reference's bundled example data
(BASELINE configs[0]/[1]) the tests still need a realistic `micropoint`, so this module reproduces
`runpointmodel`'s R glue in Python around the UNMODIFIED compiled reference (`oracle/_ref`:
weatherhgtCpp, soilmCpp, BigLeafCpp, pointmprocess, manCpp).  It is used by
tools/make_bundled_fixtures.py in the build container (where /root/reference exists) to write
tests/golden/bundled_micropoint.npz; nothing in the product imports it.

PARITY UNPINNED for the R-only steps (`stats::spline` of the daily soil moisture, `.getmode`): there is
no R here to check them against.  They only shape the *inputs* of the grid solver; the solver's parity is
pinned separately on identical inputs.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref", "libmicroclimf_ref.so")
_PD = C.POINTER(C.c_double)
_PI = C.POINTER(C.c_int32)

WEATHER_COLS = ("temp", "relhum", "pres", "swdown", "difrad", "lwdown", "windspeed", "winddir", "precip")


def _lib():
    return C.CDLL(_REF)


def _pd(a):
    return a.ctypes.data_as(_PD)


def _pi(a):
    return a.ctypes.data_as(_PI)


def getmode(v):
    """ref .getmode (R/internal.R:106-110): first most frequent non-NA value (in order of appearance)."""
    v = np.asarray(v, dtype=np.float64).ravel()
    v = v[~np.isnan(v)]
    uniq, first, counts = np.unique(v, return_index=True, return_counts=True)
    order = np.argsort(first)  # R's unique() keeps order of appearance; which.max takes the first maximum
    uniq, counts = uniq[order], counts[order]
    return float(uniq[np.argmax(counts)])


def fmm_spline(y, nout):
    """stats::spline(y, n = nout)$y with the default method "fmm" on x = 1..length(y) (R's splines.c,
    Forsythe, Malcolm & Moler: cubic through the points with end third-derivatives from divided differences)."""
    y = np.asarray(y, dtype=np.float64)
    n = y.size
    x = np.arange(1, n + 1, dtype=np.float64)
    b, c, d = np.zeros(n), np.zeros(n), np.zeros(n)
    if n < 3:
        t = (y[1] - y[0]) / (x[1] - x[0])
        b[:] = t
    else:
        nm1 = n - 1
        d[0] = x[1] - x[0]
        c[1] = (y[1] - y[0]) / d[0]
        for i in range(1, nm1):
            d[i] = x[i + 1] - x[i]
            b[i] = 2 * (d[i - 1] + d[i])
            c[i + 1] = (y[i + 1] - y[i]) / d[i]
            c[i] = c[i + 1] - c[i]
        b[0], b[nm1] = -d[0], -d[nm1 - 1]
        c[0] = c[nm1] = 0.0
        if n > 3:
            c[0] = c[2] / (x[3] - x[1]) - c[1] / (x[2] - x[0])
            c[nm1] = c[nm1 - 1] / (x[nm1] - x[nm1 - 2]) - c[nm1 - 2] / (x[nm1 - 1] - x[nm1 - 3])
            c[0] = c[0] * d[0] * d[0] / (x[3] - x[0])
            c[nm1] = -c[nm1] * d[nm1 - 1] * d[nm1 - 1] / (x[nm1] - x[nm1 - 3])
        for i in range(1, n):
            t = d[i - 1] / b[i - 1]
            b[i] = b[i] - t * d[i - 1]
            c[i] = c[i] - t * c[i - 1]
        c[nm1] = c[nm1] / b[nm1]
        for i in range(nm1 - 1, -1, -1):
            c[i] = (c[i] - d[i] * c[i + 1]) / b[i]
        b[nm1] = (y[nm1] - y[nm1 - 1]) / d[nm1 - 1] + d[nm1 - 1] * (c[nm1 - 1] + 2 * c[nm1])
        for i in range(nm1):
            b[i] = (y[i + 1] - y[i]) / d[i] - d[i] * (c[i + 1] + 2 * c[i])
            d[i] = (c[i + 1] - c[i]) / d[i]
            c[i] = 3 * c[i]
        c[nm1] = 3 * c[nm1]
        d[nm1] = d[nm1 - 1]
    xout = np.linspace(x[0], x[-1], nout)
    i = np.clip(np.searchsorted(x, xout, side="right") - 1, 0, n - 1)
    dx = xout - x[i]
    return y[i] + dx * (b[i] + dx * (c[i] + dx * d[i]))


def _weather_matrix(weather):
    return np.ascontiguousarray(np.stack([np.asarray(weather[k], dtype=np.float64) for k in WEATHER_COLS]))


def runpointmodel(weather, obstime, reqhgt, vegp_p, groundp_p, soilparams_row, lat, lon, mxhgt, zref=2.0, windhgt=None,
                  matemp=None, dTmx=25.0, maxiter=20, yearG=True):
    """ref runpointmodel (R/Cppwrappers.R:59-148).  `weather`: dict of the 9 numeric climdata columns;
    `obstime`: dict(year, month, day, hour); `vegp_p` / `groundp_p`: the `.sortvegp(method="P")` /
    `.sortsoilc(method="P")` vectors (R/internal.R:229-240, 341-357); `soilparams_row`: the modal soil
    type's row of `soilparamsp` (rmu, mult, pwr, Smax, Smin, Ksat, a)."""
    L = _lib()
    w = {k: np.array(weather[k], dtype=np.float64) for k in WEATHER_COLS}
    n = w["temp"].size
    if matemp is None:
        matemp = float(np.mean(w["temp"]))
    if windhgt is not None and windhgt != zref:
        w["windspeed"] = w["windspeed"] * np.log(67.8 * zref - 5.42) / np.log(67.8 * windhgt - 5.42)
    if n < 8760:
        yearG = False
    yr = np.ascontiguousarray(obstime["year"], dtype=np.int32)
    mo = np.ascontiguousarray(obstime["month"], dtype=np.int32)
    dy = np.ascontiguousarray(obstime["day"], dtype=np.int32)
    hr = np.ascontiguousarray(obstime["hour"], dtype=np.float64)
    zout = mxhgt if mxhgt > 2 else 2.0
    if zout > zref:
        wm = _weather_matrix(w)
        out3 = np.empty((3, n))
        rc = L.ref_weatherhgt(n, _pi(yr), _pi(mo), _pi(dy), _pd(hr), _pd(wm), C.c_double(zref), C.c_double(zout),
                              C.c_double(zout), C.c_double(lat), C.c_double(lon), _pd(out3))
        if rc == 0 and not np.isnan(out3[0].mean()):
            w["temp"], w["relhum"], w["windspeed"] = out3[0].copy(), out3[1].copy(), out3[2].copy()
        zref = zout
    w["windspeed"] = np.where(w["windspeed"] < 0.5, 0.5, w["windspeed"])
    # soil moisture
    wm = _weather_matrix(w)
    sp = soilparams_row
    daily = np.empty(n // 24 + 2)
    nout = C.c_int32(0)
    rc = L.ref_soilm(n, _pd(wm), C.c_double(sp["rmu"]), C.c_double(sp["mult"]), C.c_double(sp["pwr"]),
                     C.c_double(sp["Smax"]), C.c_double(sp["Smin"]), C.c_double(sp["Ksat"]), C.c_double(sp["a"]),
                     _pd(daily), C.byref(nout))
    assert rc == 0
    soilm = fmm_spline(daily[:nout.value], n)
    # big-leaf model
    vp = np.ascontiguousarray(vegp_p, dtype=np.float64)
    gp = np.ascontiguousarray(groundp_p, dtype=np.float64)
    sm = np.ascontiguousarray(soilm)
    out6 = np.empty((6, n))
    rc = L.ref_bigleaf(n, _pi(yr), _pi(mo), _pi(dy), _pd(hr), _pd(wm), _pd(vp), vp.size, _pd(gp), gp.size, _pd(sm),
                       C.c_double(lat), C.c_double(lon), C.c_double(dTmx), C.c_double(zref), int(maxiter),
                       C.c_double(0.5), C.c_double(0.5), C.c_double(0.1), 1 if yearG else 0, _pd(out6))
    assert rc == 0
    Tc, Tg, G, uf, RabsG = (out6[k].copy() for k in range(5))
    in7 = np.ascontiguousarray(np.stack([w["windspeed"], w["temp"], w["relhum"], w["pres"], uf, soilm, RabsG]))
    p6 = np.empty((6, n))
    rc = L.ref_pointmprocess(n, _pd(in7), C.c_double(zref), C.c_double(vp[0]), C.c_double(vp[1]), C.c_double(gp[4]),
                             C.c_double(gp[5]), C.c_double(gp[6]), C.c_double(gp[7]), _pd(p6))
    assert rc == 0
    dfo = dict(umu=p6[0].copy(), kp=p6[1].copy(), muGp=p6[2].copy(), DDp=p6[3].copy(), T0p=p6[4].copy(),
               dtrp=p6[5].copy(), G=G, soilm=soilm, Tg=Tg, Tc=Tc)
    Tbz = soilbelowT(dfo, reqhgt) if reqhgt < 0 else None
    return dict(weather=w, dfo=dfo, Tbz=Tbz, lat=lat, long=lon, zref=zref, matemp=matemp)


def man(x, n):
    L = _lib()
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    L.ref_man(_pd(x), int(x.size), int(n), _pd(out))
    return out


def soilbelowT(dfo, reqhgt):
    """ref .soilbelowT (R/internal.R:169-184)."""
    n = -118.35 * reqhgt / dfo["DDp"]
    nmn, nmx = int(np.floor(n.min())), int(np.ceil(n.max()))
    Tnmn, Tnmx = man(dfo["Tg"], nmn), man(dfo["Tg"], nmx)
    wgt = (n - nmn) / (nmx - nmn)
    Tb = wgt * Tnmx + (1 - wgt) * Tnmn
    wgt = 0.041596 * (reqhgt / np.mean(dfo["DDp"])) + 0.87142
    return wgt * Tb + (1 - wgt) * np.mean(dfo["Tg"])
