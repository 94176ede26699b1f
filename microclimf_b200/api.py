"""Python mirror of the reference's Rcpp-exported operators for the hot path.

Same names, argument order and meaning as R/RcppExports.R:72-102 (runmicro1Cpp ... runbioclim4Cpp):
data.frames / lists become dicts of numpy arrays keyed by the reference's column names, matrices are
[rows, cols] arrays and 3-D arrays are [rows, cols, n] (any memory order; they are packed to R's
column-major layout here).  The return value is the reference's named list as a dict: each present
element a float64 array of shape (rows, cols, tsteps) (Fortran order, i.e. R's memory layout).

Everything runs through the C ABI (include/microclimf_b200.h) on the GPU; errors surface as McfError,
the analogue of the R condition raised through END_RCPP (src/RcppExports.cpp:251-271).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np

from . import _abi, _lib
from .problem import GridProblem

_PD = C.POINTER(C.c_double)
_PI = C.POINTER(C.c_int32)

# reference column names -> ABI field names
_CLIM_DF = dict(temp="temp", es="es", ea="ea", tdew="tdew", pres="pres", swdown="swdown", difrad="difrad",
                lwdown="lwdown", windspeed="windspeed", winddir="winddir")          # src/microclimfCpp.cpp:2062-2071
_CLIM_ARR = dict(tc="temp", es="es", ea="ea", tdew="tdew", pk="pres", swdown="swdown", difrad="difrad",
                 lwdown="lwdown", windspeed="windspeed", winddir="winddir")         # :2350-2359
_POINT_DF = dict(soilm="p_soilm", Tg="p_Tg", Tbp="p_Tbp", G="p_G", umu="p_umu", kp="p_kp", muGp="p_muGp",
                 dtrp="p_dtrp")                                                      # :2073-2082
_POINT_ARR = dict(soilm="p_soilm", Tg="p_Tg", Tbp="p_Tbp", Gp="p_G", umu="p_umu", kp="p_kp", muGp="p_muGp",
                  dtrp="p_dtrp")                                                     # :2361-2370
_VEG = dict(hgt="hgt", pai="pai", x="x", gsmax="gsmax", leafr="leafr", leaft="leaft", clump="clump", leafd="leafd",
            paia="paia", leafden="leafden")
_SOIL = {n: n for n in _abi.SOIL_FIELDS}


def _problem(mode, dfsel, obstime, climdata, pointm, vegp, soilc, reqhgt, zref, lat, lon, lats, lons, Sminp, Smaxp,
             tfact, complete, mat) -> GridProblem:
    hgt = np.asarray(vegp["hgt"], dtype=np.float64)
    if hgt.ndim < 2:
        raise ValueError("vegp$hgt must be a [rows, cols] matrix or [rows, cols, nlyr] array")
    rows, cols = hgt.shape[0], hgt.shape[1]
    tsteps = int(np.asarray(obstime["hour"]).size)
    layered = mode in (3, 4)
    nlyr = hgt.shape[2] if (layered and hgt.ndim == 3) else 1
    if layered and dfsel is not None and nlyr > len(np.asarray(dfsel["st"])):
        # .sortvegp can hand over more layers than dfsel has rows (short series: several layers share a day and the
        # day-mode keeps one of them, R/internal.R:262-270); the drivers index layers 0 .. nrow(dfsel) - 1 only
        # (src/microclimfCpp.cpp:2770-2778), so the surplus layers are never read
        nlyr = len(np.asarray(dfsel["st"]))
        vegp = {k: (np.asarray(v)[:, :, :nlyr] if np.asarray(v).ndim == 3 else v) for k, v in vegp.items()}
    p = GridProblem(mode=mode, rows=rows, cols=cols, tsteps=tsteps, reqhgt=float(reqhgt), zref=float(zref),
                    lat=float(lat), lon=float(lon), Sminp=float(Sminp), Smaxp=float(Smaxp), tfact=float(tfact),
                    mat=float(mat), complete=bool(complete), nlyr=nlyr)
    if layered:
        p.lyr_st = np.asarray(dfsel["st"], dtype=np.int32)
        p.lyr_ed = np.asarray(dfsel["ed"], dtype=np.int32)
    for n in ("year", "month", "day", "hour"):
        p.set(n, obstime[n])
    arr = mode in (2, 4)
    for table, names in ((climdata, _CLIM_ARR if arr else _CLIM_DF), (pointm, _POINT_ARR if arr else _POINT_DF),
                         (vegp, _VEG), (soilc, _SOIL)):
        for ref_name, abi_name in names.items():
            if ref_name in table and table[ref_name] is not None:
                p.set(abi_name, table[ref_name])
    if arr:
        p.set("lats", lats)
        p.set("lons", lons)
    p.validate()
    return p


def run_problem(p: GridProblem, out: Optional[Sequence[bool]] = None, out_buffers=None) -> Dict[str, np.ndarray]:
    """mcf_runmicro on a host GridProblem; returns the reference's named list.  `out_buffers` optionally
    supplies 10 caller-owned flat float64 arrays (or None) to write into, e.g. views of pinned memory."""
    L = _lib.lib()
    n = p.ncells * p.tsteps
    if out_buffers is not None:
        bufs = list(out_buffers)
        if len(bufs) != _abi.MCF_NOUT:
            raise ValueError("out_buffers must have 10 entries")
        for b in bufs:
            if b is not None and (b.dtype != np.float64 or b.size != n or not b.flags.c_contiguous):
                raise ValueError("each output buffer must be a contiguous float64 array of rows*cols*tsteps")
    else:
        out = [True] * _abi.MCF_NOUT if out is None else [bool(o) for o in out]
        if len(out) != _abi.MCF_NOUT:
            raise ValueError("out must have 10 logicals")
        bufs = [np.empty(n, dtype=np.float64) if o else None for o in out]
    ptrs = _abi.OutPtrs(*[b.ctypes.data_as(_PD) if b is not None else None for b in bufs])
    s, keep = p.as_struct()
    err = C.create_string_buffer(512)
    _lib.check(L.mcf_runmicro(C.byref(s), ptrs, err, 512), err)
    del keep
    return {nm: b.reshape((p.rows, p.cols, p.tsteps), order="F") for nm, b in zip(_abi.OUT_NAMES, bufs)
            if b is not None}


def run_problem_packed(p: GridProblem, out: Optional[Sequence[bool]] = None, out_buffers=None) -> Dict[str, np.ndarray]:
    """mcf_runmicro_packed: the solve with writetonc's integer packing (R/dataprep.R:1064-1069, 1164-1173) applied in
    the kernel's store.  Returns int16 arrays [rows, cols, tsteps] (Fortran order), -9999 = NA."""
    L = _lib.lib()
    n = p.ncells * p.tsteps
    if out_buffers is not None:
        bufs = list(out_buffers)
        for b in bufs:
            if b is not None and (b.dtype != np.int16 or b.size != n or not b.flags.c_contiguous):
                raise ValueError("each output buffer must be a contiguous int16 array of rows*cols*tsteps")
    else:
        out = [True] * _abi.MCF_NOUT if out is None else [bool(o) for o in out]
        bufs = [np.empty(n, dtype=np.int16) if o else None for o in out]
    if len(bufs) != _abi.MCF_NOUT:
        raise ValueError("10 outputs expected")
    P16 = C.POINTER(C.c_int16)
    ptrs = _abi.OutPtrs16(*[b.ctypes.data_as(P16) if b is not None else None for b in bufs])
    s, keep = p.as_struct()
    err = C.create_string_buffer(512)
    _lib.check(L.mcf_runmicro_packed(C.byref(s), ptrs, err, 512), err)
    del keep
    return {nm: b.reshape((p.rows, p.cols, p.tsteps), order="F") for nm, b in zip(_abi.OUT_NAMES, bufs)
            if b is not None}


def run_problem_packed_dev(p: GridProblem, out_tensors, window=None, stream=None) -> None:
    """mcf_runmicro_packed_dev: device-resident problem, 10 CUDA int16 tensors (or None) as outputs; `window` and
    `stream` as for run_problem_dev."""
    import torch

    L = _lib.lib()
    st = torch.cuda.current_stream() if stream is None else stream
    P16 = C.POINTER(C.c_int16)
    ptrs = _abi.OutPtrs16(*[C.cast(C.c_void_p(t.data_ptr()), P16) if t is not None else None for t in out_tensors])
    s, keep = p.as_struct()
    w = None
    if window is not None:
        w = _abi.McfWindow(int(window[0]), int(window[1]), int(window[2]), int(window[3]))
    err = C.create_string_buffer(512)
    _lib.check(L.mcf_runmicro_packed_dev(C.byref(s), ptrs, C.byref(w) if w is not None else None,
                                         C.c_void_p(st.cuda_stream), err, 512), err)
    del keep


def run_problem_f32_dev(p: GridProblem, out_tensors, window=None, stream=None) -> None:
    """mcf_runmicro_f32_dev: the FP32 build (modes 1/3, reqhgt >= 0).  `p` holds CUDA float64 tensors
    (GridProblem.to_device); `out_tensors`: 10 CUDA float32 tensors or None; window / stream as run_problem_dev."""
    import torch

    L = _lib.lib()
    st = torch.cuda.current_stream() if stream is None else stream
    PF = C.POINTER(C.c_float)
    ptrs = _abi.OutPtrsF(*[C.cast(C.c_void_p(t.data_ptr()), PF) if t is not None else None for t in out_tensors])
    s, keep = p.as_struct()
    w = None
    if window is not None:
        w = _abi.McfWindow(int(window[0]), int(window[1]), int(window[2]), int(window[3]))
    err = C.create_string_buffer(512)
    _lib.check(L.mcf_runmicro_f32_dev(C.byref(s), ptrs, C.byref(w) if w is not None else None,
                                      C.c_void_p(st.cuda_stream), err, 512), err)
    del keep


def run_problem_f32(p: GridProblem, out: Optional[Sequence[bool]] = None, out_buffers=None) -> Dict[str, np.ndarray]:
    """mcf_runmicro_f32: the FP32 build through the host-buffer path; float32 arrays [rows, cols, tsteps] (NaN = NA)."""
    L = _lib.lib()
    n = p.ncells * p.tsteps
    if out_buffers is not None:
        bufs = list(out_buffers)
        for b in bufs:
            if b is not None and (b.dtype != np.float32 or b.size != n or not b.flags.c_contiguous):
                raise ValueError("each output buffer must be a contiguous float32 array of rows*cols*tsteps")
    else:
        out = [True] * _abi.MCF_NOUT if out is None else [bool(o) for o in out]
        bufs = [np.empty(n, dtype=np.float32) if o else None for o in out]
    if len(bufs) != _abi.MCF_NOUT:
        raise ValueError("10 outputs expected")
    PF = C.POINTER(C.c_float)
    ptrs = _abi.OutPtrsF(*[b.ctypes.data_as(PF) if b is not None else None for b in bufs])
    s, keep = p.as_struct()
    err = C.create_string_buffer(512)
    _lib.check(L.mcf_runmicro_f32(C.byref(s), ptrs, err, 512), err)
    del keep
    return {nm: b.reshape((p.rows, p.cols, p.tsteps), order="F") for nm, b in zip(_abi.OUT_NAMES, bufs) if b is not None}


def run_bioclim_problem(p: GridProblem, wetq, dryq, hotq, colq, air: bool = True,
                        out: Optional[Sequence[bool]] = None) -> Dict[str, np.ndarray]:
    L = _lib.lib()
    out = [True] * _abi.MCF_NBIO if out is None else [bool(o) for o in out]
    if len(out) != _abi.MCF_NBIO:
        raise ValueError("out must have 19 logicals")
    bufs = [np.empty(p.ncells, dtype=np.float64) if o else None for o in out]
    ptrs = _abi.BioPtrs(*[b.ctypes.data_as(_PD) if b is not None else None for b in bufs])
    qs = [np.ascontiguousarray(np.asarray(q, dtype=np.int32)) for q in (wetq, dryq, hotq, colq)]
    qargs = []
    for q in qs:
        qargs += [q.ctypes.data_as(_PI), C.c_int32(q.size)]
    s, keep = p.as_struct()
    err = C.create_string_buffer(512)
    _lib.check(L.mcf_runbioclim(C.byref(s), *qargs, C.c_int32(1 if air else 0), ptrs, err, 512), err)
    del keep
    return {nm: b.reshape((p.rows, p.cols), order="F") for nm, b in zip(_abi.BIO_NAMES, bufs) if b is not None}


# ---------------------------------------------------------------------------------------------------
# the reference's operators (R/RcppExports.R:72-102)
# ---------------------------------------------------------------------------------------------------
def runmicro1Cpp(obstime, climdata, pointm, vegp, soilc, reqhgt, zref, lat, lon, Sminp, Smaxp, tfact, complete, mat,
                 out):
    """src/microclimfCpp.cpp:2052 — static vegetation, data.frame climate."""
    return run_problem(_problem(1, None, obstime, climdata, pointm, vegp, soilc, reqhgt, zref, lat, lon, None, None,
                                Sminp, Smaxp, tfact, complete, mat), out)


def runmicro2Cpp(obstime, climdata, pointm, vegp, soilc, reqhgt, zref, lats, lons, Sminp, Smaxp, tfact, complete, mat,
                 out):
    """src/microclimfCpp.cpp:2340 — static vegetation, array climate."""
    return run_problem(_problem(2, None, obstime, climdata, pointm, vegp, soilc, reqhgt, zref, 0.0, 0.0, lats, lons,
                                Sminp, Smaxp, tfact, complete, mat), out)


def runmicro3Cpp(dfsel, obstime, climdata, pointm, vegp, soilc, reqhgt, zref, lat, lon, Sminp, Smaxp, tfact, complete,
                 mat, out):
    """src/microclimfCpp.cpp:2624 — layered vegetation, data.frame climate."""
    return run_problem(_problem(3, dfsel, obstime, climdata, pointm, vegp, soilc, reqhgt, zref, lat, lon, None, None,
                                Sminp, Smaxp, tfact, complete, mat), out)


def runmicro4Cpp(dfsel, obstime, climdata, pointm, vegp, soilc, reqhgt, zref, lats, lons, Sminp, Smaxp, tfact,
                 complete, mat, out):
    """src/microclimfCpp.cpp:2926 — layered vegetation, array climate."""
    return run_problem(_problem(4, dfsel, obstime, climdata, pointm, vegp, soilc, reqhgt, zref, 0.0, 0.0, lats, lons,
                                Sminp, Smaxp, tfact, complete, mat), out)


def _bioclim(mode, obstime, climdata, pointm, vegp, soilc, reqhgt, zref, lat, lon, lats, lons, Sminp, Smaxp, tfact,
             mat, out, wetq, dryq, hotq, colq, air):
    dfsel = None
    if mode in (3, 4):  # the reference hard-codes 14 one-day layers (src/microclimfCpp.cpp:3635-3646)
        dfsel = dict(st=np.arange(14) * 24, ed=np.arange(14) * 24 + 23)
    p = _problem(mode, dfsel, obstime, climdata, pointm, vegp, soilc, reqhgt, zref, lat, lon, lats, lons, Sminp,
                 Smaxp, tfact, True, mat)
    return run_bioclim_problem(p, wetq, dryq, hotq, colq, air, out)


def runbioclim1Cpp(obstime, climdata, pointm, vegp, soilc, reqhgt, zref, lat, lon, Sminp, Smaxp, tfact, mat, out,
                   wetq, dryq, hotq, colq, air):
    """src/microclimfCpp.cpp:3563."""
    return _bioclim(1, obstime, climdata, pointm, vegp, soilc, reqhgt, zref, lat, lon, None, None, Sminp, Smaxp, tfact,
                    mat, out, wetq, dryq, hotq, colq, air)


def runbioclim2Cpp(obstime, climdata, pointm, vegp, soilc, reqhgt, zref, lats, lons, Sminp, Smaxp, tfact, mat, out,
                   wetq, dryq, hotq, colq, air):
    return _bioclim(2, obstime, climdata, pointm, vegp, soilc, reqhgt, zref, 0.0, 0.0, lats, lons, Sminp, Smaxp, tfact,
                    mat, out, wetq, dryq, hotq, colq, air)


def runbioclim3Cpp(obstime, climdata, pointm, vegp, soilc, reqhgt, zref, lat, lon, Sminp, Smaxp, tfact, mat, out,
                   wetq, dryq, hotq, colq, air):
    return _bioclim(3, obstime, climdata, pointm, vegp, soilc, reqhgt, zref, lat, lon, None, None, Sminp, Smaxp, tfact,
                    mat, out, wetq, dryq, hotq, colq, air)


def runbioclim4Cpp(obstime, climdata, pointm, vegp, soilc, reqhgt, zref, lats, lons, Sminp, Smaxp, tfact, mat, out,
                   wetq, dryq, hotq, colq, air):
    return _bioclim(4, obstime, climdata, pointm, vegp, soilc, reqhgt, zref, 0.0, 0.0, lats, lons, Sminp, Smaxp, tfact,
                    mat, out, wetq, dryq, hotq, colq, air)


# ---------------------------------------------------------------------------------------------------
# device-resident path (inputs already in HBM): mcf_runmicro_dev / mcf_runbioclim_dev
# ---------------------------------------------------------------------------------------------------
def run_problem_dev(p: GridProblem, out_tensors, window=None, stream=None) -> None:
    """mcf_runmicro_dev.  `p` holds CUDA tensors (GridProblem.to_device); `out_tensors` is a sequence of
    10 CUDA float64 tensors or None.  `window` = (block0, nblocks, hour0, ring_hours) or None for the
    whole series.  Asynchronous on `stream` (a torch.cuda.Stream; default: torch's current stream)."""
    import torch

    L = _lib.lib()
    st = torch.cuda.current_stream() if stream is None else stream
    ptrs = _abi.OutPtrs(*[C.cast(C.c_void_p(t.data_ptr()), _PD) if t is not None else None for t in out_tensors])
    s, keep = p.as_struct()
    w = None
    if window is not None:
        w = _abi.McfWindow(int(window[0]), int(window[1]), int(window[2]), int(window[3]))
    err = C.create_string_buffer(512)
    _lib.check(L.mcf_runmicro_dev(C.byref(s), ptrs, C.byref(w) if w is not None else None,
                                  C.c_void_p(st.cuda_stream), err, 512), err)
    del keep


SUMMARY_STATS = ("mean", "min", "max")


def run_summary(p: GridProblem, out: Optional[Sequence[bool]] = None):
    """mcf_runmicro_summary on a host GridProblem: per-cell mean / minimum / maximum over the computed hours of each
    requested output, reduced inside the grid kernel (nothing hourly is stored or copied).  Returns
    ({name: {"mean" | "min" | "max": [rows, cols] array}}, hours)."""
    L = _lib.lib()
    out = [True] * _abi.MCF_NOUT if out is None else [bool(o) for o in out]
    bufs = [[np.empty(p.ncells, dtype=np.float64) if o else None for o in out] for _ in range(3)]
    ptrs = [_abi.OutPtrs(*[b.ctypes.data_as(_PD) if b is not None else None for b in bb]) for bb in bufs]
    s, keep = p.as_struct()
    hours = C.c_int64(0)
    err = C.create_string_buffer(512)
    _lib.check(L.mcf_runmicro_summary(C.byref(s), ptrs[0], ptrs[1], ptrs[2], C.byref(hours), err, 512), err)
    del keep
    res = {}
    for v, nm in enumerate(_abi.OUT_NAMES):
        if out[v]:
            res[nm] = {st: bufs[k][v].reshape((p.rows, p.cols), order="F") for k, st in enumerate(SUMMARY_STATS)}
    return res, int(hours.value)


def run_summary_dev(p: GridProblem, sums, mins, maxs, window=None, accumulate: bool = False, stream=None) -> int:
    """mcf_runmicro_summary_dev: device-resident problem; `sums`, `mins`, `maxs` are sequences of 10 CUDA float64
    tensors [rows * cols] or None (all three or none per output).  Returns the hours the call added."""
    import torch

    L = _lib.lib()
    st = torch.cuda.current_stream() if stream is None else stream
    ptrs = [_abi.OutPtrs(*[C.cast(C.c_void_p(t.data_ptr()), _PD) if t is not None else None for t in tt])
            for tt in (sums, mins, maxs)]
    s, keep = p.as_struct()
    w = None
    if window is not None:
        w = _abi.McfWindow(int(window[0]), int(window[1]), int(window[2]), int(window[3]))
    hours = C.c_int64(0)
    err = C.create_string_buffer(512)
    _lib.check(L.mcf_runmicro_summary_dev(C.byref(s), ptrs[0], ptrs[1], ptrs[2], C.byref(w) if w is not None else None,
                                          C.c_int32(1 if accumulate else 0), C.byref(hours), C.c_void_p(st.cuda_stream),
                                          err, 512), err)
    del keep
    return int(hours.value)


def run_bioclim_problem_dev(p: GridProblem, wetq, dryq, hotq, colq, air, bio_tensors, stream=None) -> None:
    import torch

    L = _lib.lib()
    st = torch.cuda.current_stream() if stream is None else stream
    ptrs = _abi.BioPtrs(*[C.cast(C.c_void_p(t.data_ptr()), _PD) if t is not None else None for t in bio_tensors])
    qs = [np.ascontiguousarray(np.asarray(q, dtype=np.int32)) for q in (wetq, dryq, hotq, colq)]
    qargs = []
    for q in qs:
        qargs += [q.ctypes.data_as(_PI), C.c_int32(q.size)]
    s, keep = p.as_struct()
    err = C.create_string_buffer(512)
    _lib.check(L.mcf_runbioclim_dev(C.byref(s), *qargs, C.c_int32(1 if air else 0), ptrs,
                                    C.c_void_p(st.cuda_stream), err, 512), err)
    del keep


def fp64_peak_tflops() -> float:
    L = _lib.lib()
    v = C.c_double(0.0)
    err = C.create_string_buffer(512)
    _lib.check(L.mcf_fp64_peak(C.byref(v), err, 512), err)
    return v.value


def kernel_time(reset: bool = False):
    """(total ms, launches) of the grid kernel since the last reset (needs timing enabled)."""
    L = _lib.lib()
    ms, n = C.c_double(0.0), C.c_int64(0)
    rc = L.mcf_kernel_time(C.byref(ms), C.byref(n))
    if rc != 0:
        raise _lib.McfError(rc, "mcf_kernel_time failed")
    if reset:
        L.mcf_kernel_time_reset()
    return ms.value, n.value


def math_eval(fn: int, x, y=None):
    """mcf_math_eval: the kernels' own FP64 elementary functions, element-wise (accuracy tests)."""
    L = _lib.lib()
    x = np.ascontiguousarray(x, dtype=np.float64)
    yy = None if y is None else np.ascontiguousarray(np.broadcast_to(np.asarray(y, dtype=np.float64), x.shape))
    out = np.empty_like(x)
    err = C.create_string_buffer(512)
    _lib.check(L.mcf_math_eval(fn, x.ctypes.data_as(_PD), yy.ctypes.data_as(_PD) if yy is not None else None,
                               x.size, out.ctypes.data_as(_PD), err, 512), err)
    return out


# ---------------------------------------------------------------------------------------------------
# terrain preparation (SURVEY.md NEXT-2): .horizon / sky view / .windcoef of R/internal.R on the GPU
# ---------------------------------------------------------------------------------------------------
def horizon(dtm, reso: float, azimuths=None, want_svf: bool = True):
    """soilc$hor and soilc$svfa as .runmodelNCpp builds them (R/internal.R:1142-1148).  `dtm` is the
    [rows, cols] elevation matrix (row 0 = north).  Returns (hor[rows, cols, nazi], svfa[rows, cols] | None)."""
    L = _lib.lib()
    d = np.asarray(dtm, dtype=np.float64)
    rows, cols = d.shape
    az = np.arange(24) * 15.0 if azimuths is None else np.ascontiguousarray(azimuths, dtype=np.float64)
    flat = np.ascontiguousarray(d.ravel(order="F"))
    hor = np.empty(rows * cols * az.size)
    svf = np.empty(rows * cols) if want_svf else None
    err = C.create_string_buffer(512)
    _lib.check(L.mcf_horizon(flat.ctypes.data_as(_PD), rows, cols, float(reso), int(az.size), az.ctypes.data_as(_PD),
                             hor.ctypes.data_as(_PD), svf.ctypes.data_as(_PD) if svf is not None else None, err, 512), err)
    return hor.reshape((rows, cols, az.size), order="F"), (svf.reshape((rows, cols), order="F") if want_svf else None)


def windcoef(dsm, reso: float, hgt: float, directions=None, blend8: bool = False):
    """.windcoef (R/internal.R:949-968) per direction (default: the 16 of .windsheltera) and optionally the
    16 -> 8 blend (R/internal.R:983-989, without terra's aggregate/resample smoothing)."""
    L = _lib.lib()
    d = np.asarray(dsm, dtype=np.float64)
    rows, cols = d.shape
    dr = np.arange(16) * 22.5 if directions is None else np.ascontiguousarray(directions, dtype=np.float64)
    flat = np.ascontiguousarray(d.ravel(order="F"))
    idx = np.empty(rows * cols * dr.size)
    b8 = np.empty(rows * cols * 8) if blend8 else None
    err = C.create_string_buffer(512)
    _lib.check(L.mcf_windcoef(flat.ctypes.data_as(_PD), rows, cols, float(reso), float(hgt), int(dr.size),
                              dr.ctypes.data_as(_PD), idx.ctypes.data_as(_PD),
                              b8.ctypes.data_as(_PD) if b8 is not None else None, err, 512), err)
    out = idx.reshape((rows, cols, dr.size), order="F")
    return (out, b8.reshape((rows, cols, 8), order="F")) if blend8 else out


def flowacc(dtm):
    """flowaccCpp (src/microclimfCpp.cpp:5368-5414) of a [rows, cols] elevation matrix (NaN = NA)."""
    L = _lib.lib()
    d = np.asarray(dtm, dtype=np.float64)
    rows, cols = d.shape
    flat = np.ascontiguousarray(d.ravel(order="F"))
    fa = np.empty(rows * cols)
    err = C.create_string_buffer(512)
    _lib.check(L.mcf_flowacc(flat.ctypes.data_as(_PD), rows, cols, fa.ctypes.data_as(_PD), err, 512), err)
    return fa.reshape((rows, cols), order="F")


def slope_aspect(dtm, xres: float, yres: float):
    """terra::terrain(v = "slope") / (v = "aspect") of a [rows, cols] elevation matrix (row 0 = north), degrees, NaN where
    terra leaves NA (R/internal.R:1124-1129): Horn's stencil on the GPU."""
    L = _lib.lib()
    d = np.asarray(dtm, dtype=np.float64)
    rows, cols = d.shape
    flat = np.ascontiguousarray(d.ravel(order="F"))
    sl, asp = np.empty(rows * cols), np.empty(rows * cols)
    err = C.create_string_buffer(512)
    _lib.check(L.mcf_slope_aspect(flat.ctypes.data_as(_PD), rows, cols, float(xres), float(yres), sl.ctypes.data_as(_PD),
                                  asp.ctypes.data_as(_PD), err, 512), err)
    return sl.reshape((rows, cols), order="F"), asp.reshape((rows, cols), order="F")


def windshelter(dsm, reso: float, hgt: float, s: int = 10):
    """.windsheltera (R/internal.R:970-991) in one call on the GPU: 16 x .windcoef, block-mean + bilinear smoothing of each
    direction (fact = s), 16 -> 8 blend.  Returns wsa[rows, cols, 8]."""
    L = _lib.lib()
    d = np.asarray(dsm, dtype=np.float64)
    rows, cols = d.shape
    flat = np.ascontiguousarray(d.ravel(order="F"))
    out = np.empty(rows * cols * 8)
    err = C.create_string_buffer(512)
    _lib.check(L.mcf_windshelter(flat.ctypes.data_as(_PD), rows, cols, float(reso), float(hgt), int(s),
                                 out.ctypes.data_as(_PD), err, 512), err)
    return out.reshape((rows, cols, 8), order="F")


def topidx(dtm, xres: float, yres: float):
    """.topidx (R/internal.R:861-874) of a [rows, cols] elevation matrix (NaN = NA)."""
    L = _lib.lib()
    d = np.asarray(dtm, dtype=np.float64)
    rows, cols = d.shape
    flat = np.ascontiguousarray(d.ravel(order="F"))
    out = np.empty(rows * cols)
    err = C.create_string_buffer(512)
    _lib.check(L.mcf_topidx(flat.ctypes.data_as(_PD), rows, cols, float(xres), float(yres), out.ctypes.data_as(_PD), err, 512), err)
    return out.reshape((rows, cols), order="F")
