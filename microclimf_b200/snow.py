"""Python mirror of the reference's snow operators (data.frame climate), R/RcppExports.R `gridmodelsnow1` and
`gridmicrosnow1` (src/microclimfCpp.cpp:4172, 4894): same argument lists (data.frames / lists as dicts of numpy
arrays keyed by the reference's names), results as the reference's named lists.  Runs the CUDA kernels of
csrc/mcf_snow.cu through the C ABI (mcf_gridmodelsnow / mcf_gridmicrosnow); no CPU path."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Sequence

import numpy as np

from . import _abi, _lib

_PD = C.POINTER(C.c_double)
_PI = C.POINTER(C.c_int32)
SNOWENV = {"Alpine": 0, "Maritime": 1, "Prairie": 2, "Tundra": 3, "Taiga": 4}
CLIM_COLS = ("temp", "relhum", "pres", "swdown", "difrad", "lwdown", "windspeed", "winddir", "precip")


class SnowClimate(C.Structure):
    _fields_ = [("tsteps", C.c_int32), ("year", _PI), ("month", _PI), ("day", _PI), ("hour", _PD)] + [(n, _PD) for n in CLIM_COLS]


class SnowPoint(C.Structure):
    _fields_ = [(n, _PD) for n in ("Gp", "Tc", "RswabsG", "RlwabsG", "umu")]


class SnowStatic(C.Structure):
    _fields_ = ([("rows", C.c_int32), ("cols", C.c_int32)]
                + [(n, _PD) for n in ("pai", "hgt", "leaft", "clump", "paia", "leafd", "leafden", "Smax", "slope", "aspect",
                                      "skyview", "wsa", "hor")]
                + [("lat", C.c_double), ("lon", C.c_double), ("zref", C.c_double), ("isnowdc", _PD), ("isnowdg", _PD),
                   ("isnowac", _PI), ("isnowag", _PI), ("lats", _PD), ("lons", _PD)])


class SnowState(C.Structure):
    _fields_ = [(n, _PD) for n in ("Tc", "Tg", "totalSWE", "groundsnowdepth", "snowden")]


Out3 = _PD * 5
Out2 = _PD * 4
MODEL_3D = ("Tc", "Tg", "sdepc", "sdepg", "sden")
MODEL_2D = ("agec", "ageg", "meltc", "meltg")


def _f(a):
    a = np.asarray(a, dtype=np.float64)
    return np.ascontiguousarray(a.ravel(order="F") if a.ndim > 1 else a)


def _i(a):
    a = np.asarray(a, dtype=np.int32)
    return np.ascontiguousarray(a.ravel(order="F") if a.ndim > 1 else a)


def pack_climate(obstime, climdata, keep):
    s = SnowClimate()
    s.tsteps = int(np.asarray(obstime["hour"]).size)
    for n in ("year", "month", "day"):
        a = _i(obstime[n]); keep.append(a); setattr(s, n, a.ctypes.data_as(_PI))
    a = _f(obstime["hour"]); keep.append(a); s.hour = a.ctypes.data_as(_PD)
    for n in CLIM_COLS:
        a = _f(climdata[n]); keep.append(a); setattr(s, n, a.ctypes.data_as(_PD))
    return s


def pack_static(vegp, other, keep):
    s = SnowStatic()
    hgt = np.asarray(vegp["hgt"])
    s.rows, s.cols = hgt.shape
    for n in ("pai", "hgt", "leaft", "clump", "paia", "leafd", "leafden"):
        if n in vegp:
            a = _f(vegp[n]); keep.append(a); setattr(s, n, a.ctypes.data_as(_PD))
    for n in ("slope", "aspect", "skyview", "wsa", "hor", "Smax", "isnowdc", "isnowdg", "lats", "lons"):
        if n in other:
            a = _f(other[n]); keep.append(a); setattr(s, n, a.ctypes.data_as(_PD))
    for n in ("isnowac", "isnowag"):
        if n in other:
            a = _i(other[n]); keep.append(a); setattr(s, n, a.ctypes.data_as(_PI))
    if np.ndim(other.get("lat", 0.0)) == 0:
        s.lat, s.lon = float(other.get("lat", 0.0)), float(other.get("lon", 0.0))
    s.zref = float(other["zref"])
    return s


def pack_point(pointm, keep):
    s = SnowPoint()
    for n in ("Gp", "Tc", "RswabsG", "RlwabsG", "umu"):
        a = _f(pointm[n]); keep.append(a); setattr(s, n, a.ctypes.data_as(_PD))
    return s


def pack_state(snowm, keep):
    s = SnowState()
    for n in ("Tc", "Tg", "totalSWE", "groundsnowdepth", "snowden"):
        a = _f(snowm[n]); keep.append(a); setattr(s, n, a.ctypes.data_as(_PD))
    return s


def _bind(L):
    if getattr(L, "_snow_bound", False):
        return
    L.mcf_gridmodelsnow.argtypes = [C.POINTER(SnowClimate), C.POINTER(SnowPoint), C.POINTER(SnowStatic), C.c_int32, Out3, Out2,
                                    C.c_char_p, C.c_size_t]
    L.mcf_gridmodelsnow.restype = C.c_int
    L.mcf_gridmicrosnow.argtypes = [C.c_double, C.POINTER(SnowClimate), _PD, C.POINTER(SnowState), C.POINTER(SnowStatic),
                                    C.c_double, _abi.OutPtrs, C.c_char_p, C.c_size_t]
    L.mcf_gridmicrosnow.restype = C.c_int
    L.mcf_gridmodelsnow2.argtypes = L.mcf_gridmodelsnow.argtypes
    L.mcf_gridmodelsnow2.restype = C.c_int
    L.mcf_gridmicrosnow2.argtypes = L.mcf_gridmicrosnow.argtypes
    L.mcf_gridmicrosnow2.restype = C.c_int
    L._snow_bound = True


def call_gridmodelsnow(fn, obstime, climdata, pointm, vegp, other, snowenv):
    keep = []
    c, p, s = pack_climate(obstime, climdata, keep), pack_point(pointm, keep), pack_static(vegp, other, keep)
    n3 = s.rows * s.cols * c.tsteps
    b3 = [np.empty(n3) for _ in MODEL_3D]
    b2 = [np.empty(s.rows * s.cols) for _ in MODEL_2D]
    err = C.create_string_buffer(512)
    rc = fn(C.byref(c), C.byref(p), C.byref(s), C.c_int32(SNOWENV[snowenv]), Out3(*[b.ctypes.data_as(_PD) for b in b3]),
            Out2(*[b.ctypes.data_as(_PD) for b in b2]), err, C.c_size_t(512))
    if rc != 0:
        raise _lib.McfError(rc, err.value.decode(errors="replace"))
    out = {n: b.reshape((s.rows, s.cols, c.tsteps), order="F") for n, b in zip(MODEL_3D, b3)}
    out.update({n: b.reshape((s.rows, s.cols), order="F") for n, b in zip(MODEL_2D, b2)})
    return out


def call_gridmicrosnow(fn, reqhgt, obstime, climdata, snowm, micro, vegp, other, mat, out, copy=True):
    keep = []
    c, st, s = pack_climate(obstime, climdata, keep), pack_state(snowm, keep), pack_static(vegp, other, keep)
    umu = _f(climdata["umu"])
    bufs = []
    for nm, o in zip(_abi.OUT_NAMES, out):
        # the C entry point works IN PLACE, as the reference does on its arguments; `copy` keeps the caller's arrays intact
        # (R semantics) at the price of a host copy of every requested array
        bufs.append((_f(micro[nm]).copy() if copy else _f(micro[nm])) if o else None)
    err = C.create_string_buffer(512)
    rc = fn(C.c_double(reqhgt), C.byref(c), umu.ctypes.data_as(_PD), C.byref(st), C.byref(s), C.c_double(mat),
            _abi.OutPtrs(*[b.ctypes.data_as(_PD) if b is not None else None for b in bufs]), err, C.c_size_t(512))
    if rc != 0:
        raise _lib.McfError(rc, err.value.decode(errors="replace"))
    return {nm: b.reshape((s.rows, s.cols, c.tsteps), order="F") for nm, b in zip(_abi.OUT_NAMES, bufs) if b is not None}


def gridmodelsnow1(obstime, climdata, pointm, vegp, other, snowenv: str = "Alpine") -> Dict[str, np.ndarray]:
    """src/microclimfCpp.cpp:4172 — snow-pack model over the grid, data.frame climate."""
    L = _lib.lib()
    _bind(L)
    return call_gridmodelsnow(L.mcf_gridmodelsnow, obstime, climdata, pointm, vegp, other, snowenv)


def gridmicrosnow1(reqhgt, obstime, climdata, snowm, micro, vegp, other, mat, out: Sequence[bool],
                   copy: bool = True) -> Dict[str, np.ndarray]:
    """src/microclimfCpp.cpp:4894 — microclimate on snow-covered cell-hours.  `copy=True` (default) leaves `micro`'s arrays
    untouched and returns updated copies; `copy=False` updates F-contiguous float64 arrays in place, as the C entry point
    does (no host copies: the call is then bound by the 200 bytes per cell-hour that cross PCIe)."""
    L = _lib.lib()
    _bind(L)
    return call_gridmicrosnow(L.mcf_gridmicrosnow, reqhgt, obstime, climdata, snowm, micro, vegp, other, mat, out, copy)


def gridmodelsnow2(obstime, climdata, pointm, vegp, other, snowenv: str = "Alpine") -> Dict[str, np.ndarray]:
    """src/microclimfCpp.cpp:4426 — snow-pack model, array climate ([rows, cols, hours] series; other$lats / lons)."""
    L = _lib.lib()
    _bind(L)
    return call_gridmodelsnow(L.mcf_gridmodelsnow2, obstime, climdata, pointm, vegp, other, snowenv)


def gridmicrosnow2(reqhgt, obstime, climdata, snowm, micro, vegp, other, mat, out: Sequence[bool],
                   copy: bool = True) -> Dict[str, np.ndarray]:
    """src/microclimfCpp.cpp:5059 — snow microclimate, array climate (climdata$prec, climdata$umu arrays; other$lat / lon
    matrices, passed here as other["lats"] / other["lons"])."""
    L = _lib.lib()
    _bind(L)
    return call_gridmicrosnow(L.mcf_gridmicrosnow2, reqhgt, obstime, climdata, snowm, micro, vegp, other, mat, out, copy)
