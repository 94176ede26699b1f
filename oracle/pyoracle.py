"""TEST INFRASTRUCTURE ONLY — ctypes access to the two CPU checkers.

  ref    : oracle/_ref/libmicroclimf_ref.so — the unmodified reference C++ behind oracle/ref_driver.cpp
  oracle : oracle/liboracle.so              — the plain-C restatement (oracle/mcf_oracle.c)

Both take the product's `mcf_problem` struct, so a GridProblem feeds either checker or the CUDA path
unchanged.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from microclimf_b200 import _abi
from microclimf_b200.problem import GridProblem

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_PATH = os.path.join(_HERE, "_ref", "libmicroclimf_ref.so")
ORACLE_PATH = os.path.join(_HERE, "liboracle.so")
# kind "glue" is not a checker: it is the product entered through the Rcpp-typed binding (rglue/microclimf_glue.cpp
# compiled against the Rcpp stand-in, driven by this directory's ref_driver.cpp with the reference's DataFrame / List
# arguments).  It shares the calling code below so that the glue tests read like the reference-parity tests.
GLUE_PATH = os.path.join(_HERE, "..", "rglue", "_build", "libmcf_glue_test.so")
# kind "patched": the reference's own translation unit after rglue/apply_glue.sh (its twelve grid functions removed, the
# glue in their place) behind the full driver — what the R package's shared object holds after the change
PATCHED_PATH = os.path.join(_HERE, "..", "rglue", "_build", "libmicroclimf_patched.so")
_PATHS = {"ref": REF_PATH, "oracle": ORACLE_PATH, "glue": GLUE_PATH, "patched": PATCHED_PATH}
_PREFIX = {"ref": "ref_", "oracle": "oracle_", "glue": "glue_", "patched": "ref_"}
_libs = {}


def have_ref() -> bool:
    return os.path.exists(REF_PATH)


def have_oracle() -> bool:
    return os.path.exists(ORACLE_PATH)


def have_glue() -> bool:
    return os.path.exists(GLUE_PATH)


def have_patched() -> bool:
    return os.path.exists(PATCHED_PATH)


def _lib(kind: str):
    if kind not in _libs:
        path = _PATHS[kind]
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} not built: run `make -C oracle` / `make -C rglue` (or __graft_entry__.build())")
        _libs[kind] = C.CDLL(path)
    return _libs[kind]


def _na_filled(n):
    a = np.empty(n, dtype=np.float64)
    a.view(np.uint64)[:] = 0xDEADBEEFDEADBEEF  # poison: the callee must overwrite everything
    return a


def runmicro(prob: GridProblem, out_mask=None, kind: str = "oracle"):
    """Run the CPU checker; returns {name: array[rows, cols, tsteps] (Fortran order)}."""
    lib = _lib(kind)
    fn = getattr(lib, _PREFIX[kind] + "runmicro")
    fn.restype = C.c_int
    s, keep = prob.as_struct()
    out_mask = [True] * _abi.MCF_NOUT if out_mask is None else list(out_mask)
    n = prob.ncells * prob.tsteps
    bufs = [_na_filled(n) if m else None for m in out_mask]
    ptrs = _abi.OutPtrs(*[b.ctypes.data_as(C.POINTER(C.c_double)) if b is not None else None for b in bufs])
    err = C.create_string_buffer(512)
    rc = fn(C.byref(s), ptrs, err, C.c_size_t(512))
    if rc != 0:
        raise RuntimeError(f"{kind} runmicro failed ({rc}): {err.value.decode()}")
    del keep
    return {nm: b.reshape((prob.rows, prob.cols, prob.tsteps), order="F")
            for nm, b in zip(_abi.OUT_NAMES, bufs) if b is not None}


def runbioclim(prob: GridProblem, quarters: dict, air: bool = True, out_mask=None, kind: str = "oracle"):
    lib = _lib(kind)
    fn = getattr(lib, _PREFIX[kind] + "runbioclim")
    fn.restype = C.c_int
    s, keep = prob.as_struct()
    out_mask = [True] * _abi.MCF_NBIO if out_mask is None else list(out_mask)
    bufs = [_na_filled(prob.ncells) if m else None for m in out_mask]
    ptrs = _abi.BioPtrs(*[b.ctypes.data_as(C.POINTER(C.c_double)) if b is not None else None for b in bufs])
    q = {k: np.ascontiguousarray(np.asarray(v, dtype=np.int32)) for k, v in quarters.items()}
    pi = C.POINTER(C.c_int32)
    err = C.create_string_buffer(512)
    rc = fn(C.byref(s), q["wetq"].ctypes.data_as(pi), C.c_int32(len(q["wetq"])), q["dryq"].ctypes.data_as(pi),
            C.c_int32(len(q["dryq"])), q["hotq"].ctypes.data_as(pi), C.c_int32(len(q["hotq"])),
            q["colq"].ctypes.data_as(pi), C.c_int32(len(q["colq"])), C.c_int32(1 if air else 0), ptrs, err,
            C.c_size_t(512))
    if rc != 0:
        raise RuntimeError(f"{kind} runbioclim failed ({rc}): {err.value.decode()}")
    del keep
    return {nm: b.reshape((prob.rows, prob.cols), order="F") for nm, b in zip(_abi.BIO_NAMES, bufs) if b is not None}


def gridmodelsnow1(obstime, climdata, pointm, vegp, other, snowenv="Alpine", kind="ref"):
    """The compiled reference's gridmodelsnow1 (src/microclimfCpp.cpp:4172) behind the product's snow structs."""
    from microclimf_b200 import snow
    fn = getattr(_lib(kind), _PREFIX[kind] + "gridmodelsnow")
    fn.restype = C.c_int
    return snow.call_gridmodelsnow(fn, obstime, climdata, pointm, vegp, other, snowenv)


def gridmicrosnow1(reqhgt, obstime, climdata, snowm, micro, vegp, other, mat, out, kind="ref"):
    from microclimf_b200 import snow
    fn = getattr(_lib(kind), _PREFIX[kind] + "gridmicrosnow")
    fn.restype = C.c_int
    return snow.call_gridmicrosnow(fn, reqhgt, obstime, climdata, snowm, micro, vegp, other, mat, out)


def gridmodelsnow2(obstime, climdata, pointm, vegp, other, snowenv="Alpine", kind="ref"):
    from microclimf_b200 import snow
    fn = getattr(_lib(kind), _PREFIX[kind] + "gridmodelsnow2")
    fn.restype = C.c_int
    return snow.call_gridmodelsnow(fn, obstime, climdata, pointm, vegp, other, snowenv)


def gridmicrosnow2(reqhgt, obstime, climdata, snowm, micro, vegp, other, mat, out, kind="ref"):
    from microclimf_b200 import snow
    fn = getattr(_lib(kind), _PREFIX[kind] + "gridmicrosnow2")
    fn.restype = C.c_int
    return snow.call_gridmicrosnow(fn, reqhgt, obstime, climdata, snowm, micro, vegp, other, mat, out)


def canintfrac(hgt, pai, uf, prec, tc, Li):
    """The compiled reference's canintfrac (src/microclimfCpp.cpp:5417) on [rows, cols] matrices."""
    lib = _lib("ref")
    h = np.asfortranarray(hgt, dtype=np.float64)
    p = np.asfortranarray(pai, dtype=np.float64)
    out = np.empty(h.shape, dtype=np.float64, order="F")
    PD = C.POINTER(C.c_double)
    lib.ref_canintfrac.argtypes = [PD, PD, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double, PD]
    rc = lib.ref_canintfrac(h.ctypes.data_as(PD), p.ctypes.data_as(PD), h.shape[0], h.shape[1], float(uf), float(prec), float(tc),
                            float(Li), out.ctypes.data_as(PD))
    if rc != 0:
        raise RuntimeError("ref_canintfrac failed")
    return out


def meltmu(skyview, stemp, tc):
    """The compiled reference's meltmu (src/microclimfCpp.cpp:5454)."""
    lib = _lib("ref")
    sv = np.asfortranarray(skyview, dtype=np.float64)
    st = np.ascontiguousarray(stemp, dtype=np.float64)
    t = np.ascontiguousarray(tc, dtype=np.float64)
    out = np.empty(sv.shape, dtype=np.float64, order="F")
    PD = C.POINTER(C.c_double)
    lib.ref_meltmu.argtypes = [PD, C.c_int32, C.c_int32, PD, PD, C.c_int32, PD]
    rc = lib.ref_meltmu(sv.ctypes.data_as(PD), sv.shape[0], sv.shape[1], st.ctypes.data_as(PD), t.ctypes.data_as(PD), st.size,
                        out.ctypes.data_as(PD))
    if rc != 0:
        raise RuntimeError("ref_meltmu failed")
    return out


def meltmu2(mu, stemp, tc):
    """The compiled reference's meltmu2 (src/microclimfCpp.cpp:5495); stemp / tc are [rows, cols, n]."""
    lib = _lib("ref")
    m = np.asfortranarray(mu, dtype=np.float64)
    st = np.asfortranarray(stemp, dtype=np.float64)
    t = np.asfortranarray(tc, dtype=np.float64)
    out = np.empty(m.shape, dtype=np.float64, order="F")
    PD = C.POINTER(C.c_double)
    lib.ref_meltmu2.argtypes = [PD, C.c_int32, C.c_int32, PD, PD, C.c_int32, PD]
    rc = lib.ref_meltmu2(m.ctypes.data_as(PD), m.shape[0], m.shape[1], st.ctypes.data_as(PD), t.ctypes.data_as(PD), st.shape[2],
                         out.ctypes.data_as(PD))
    if rc != 0:
        raise RuntimeError("ref_meltmu2 failed")
    return out
