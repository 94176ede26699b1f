"""Loader for the built C-ABI shared library (microclimf_b200/csrc/libmicroclimf_b200.so).

There is no CPU fallback anywhere in this package: if the library is missing, or no sm_100 device is
usable, every compute entry point raises.  The library is built IN-TREE (`make -C microclimf_b200/csrc`
or `__graft_entry__.build()`), never JIT-compiled into a cache.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MCF_LIB_PATH") or os.path.join(_HERE, "csrc", "libmicroclimf_b200.so")  # env: dev experiments
_lib = None


class McfError(RuntimeError):
    """Raised for any non-zero status of the C ABI (the Rcpp stub's Rcpp::stop equivalent)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"microclimf_b200 error {code}: {msg}")
        self.code = code
        self.msg = msg


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                f"{LIB_PATH} is not built. Run `make -C microclimf_b200/csrc` (needs nvcc); there is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        pd, pi = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        pp, pw = C.POINTER(_abi.McfProblem), C.POINTER(_abi.McfWindow)
        L.mcf_abi_version.restype = C.c_int
        L.mcf_device_count.restype = C.c_int
        L.mcf_set_device.argtypes = [C.c_int]
        L.mcf_set_device.restype = C.c_int
        L.mcf_launch_count.restype = C.c_int64
        L.mcf_launch_count_reset.restype = None
        L.mcf_kernel_time.argtypes = [pd, C.POINTER(C.c_int64)]
        L.mcf_kernel_time.restype = C.c_int
        L.mcf_kernel_time_reset.restype = None
        L.mcf_kernel_timing_enable.argtypes = [C.c_int]
        L.mcf_kernel_timing_enable.restype = None
        L.mcf_fp64_peak.argtypes = [pd, C.c_char_p, C.c_size_t]
        L.mcf_fp64_peak.restype = C.c_int
        L.mcf_horizon.argtypes = [pd, C.c_int32, C.c_int32, C.c_double, C.c_int32, pd, pd, pd, C.c_char_p, C.c_size_t]
        L.mcf_horizon.restype = C.c_int
        L.mcf_windcoef.argtypes = [pd, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int32, pd, pd, pd, C.c_char_p, C.c_size_t]
        L.mcf_windcoef.restype = C.c_int
        L.mcf_windshelter.argtypes = [pd, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int32, pd, C.c_char_p, C.c_size_t]
        L.mcf_windshelter.restype = C.c_int
        L.mcf_slope_aspect.argtypes = [pd, C.c_int32, C.c_int32, C.c_double, C.c_double, pd, pd, C.c_char_p, C.c_size_t]
        L.mcf_slope_aspect.restype = C.c_int
        L.mcf_topidx.argtypes = [pd, C.c_int32, C.c_int32, C.c_double, C.c_double, pd, C.c_char_p, C.c_size_t]
        L.mcf_topidx.restype = C.c_int
        L.mcf_flowacc.argtypes = [pd, C.c_int32, C.c_int32, pd, C.c_char_p, C.c_size_t]
        L.mcf_flowacc.restype = C.c_int
        L.mcf_math_eval.argtypes = [C.c_int, pd, pd, C.c_int64, pd, C.c_char_p, C.c_size_t]
        L.mcf_math_eval.restype = C.c_int
        L.mcf_runmicro.argtypes = [pp, _abi.OutPtrs, C.c_char_p, C.c_size_t]
        L.mcf_runmicro.restype = C.c_int
        L.mcf_runmicro_dev.argtypes = [pp, _abi.OutPtrs, pw, C.c_void_p, C.c_char_p, C.c_size_t]
        L.mcf_runmicro_dev.restype = C.c_int
        L.mcf_runmicro_packed.argtypes = [pp, _abi.OutPtrs16, C.c_char_p, C.c_size_t]
        L.mcf_runmicro_packed.restype = C.c_int
        L.mcf_runmicro_packed_dev.argtypes = [pp, _abi.OutPtrs16, pw, C.c_void_p, C.c_char_p, C.c_size_t]
        L.mcf_runmicro_packed_dev.restype = C.c_int
        L.mcf_runmicro_f32_dev.argtypes = [pp, _abi.OutPtrsF, pw, C.c_void_p, C.c_char_p, C.c_size_t]
        L.mcf_runmicro_f32_dev.restype = C.c_int
        L.mcf_runmicro_f32.argtypes = [pp, _abi.OutPtrsF, C.c_char_p, C.c_size_t]
        L.mcf_runmicro_f32.restype = C.c_int
        L.mcf_runmicro_summary.argtypes = [pp, _abi.OutPtrs, _abi.OutPtrs, _abi.OutPtrs, C.POINTER(C.c_int64), C.c_char_p, C.c_size_t]
        L.mcf_runmicro_summary.restype = C.c_int
        L.mcf_runmicro_summary_dev.argtypes = [pp, _abi.OutPtrs, _abi.OutPtrs, _abi.OutPtrs, pw, C.c_int32, C.POINTER(C.c_int64),
                                               C.c_void_p, C.c_char_p, C.c_size_t]
        L.mcf_runmicro_summary_dev.restype = C.c_int
        q = [pi, C.c_int32] * 4
        L.mcf_runbioclim.argtypes = [pp] + q + [C.c_int32, _abi.BioPtrs, C.c_char_p, C.c_size_t]
        L.mcf_runbioclim.restype = C.c_int
        L.mcf_runbioclim_dev.argtypes = [pp] + q + [C.c_int32, _abi.BioPtrs, C.c_void_p, C.c_char_p, C.c_size_t]
        L.mcf_runbioclim_dev.restype = C.c_int
        L.mcf_twi_partial.argtypes = [pd, C.c_int64, C.c_double, pd, C.POINTER(C.c_int64), C.c_char_p, C.c_size_t]
        L.mcf_twi_partial.restype = C.c_int
        if L.mcf_abi_version() != _abi.MCF_ABI_VERSION:
            raise RuntimeError("libmicroclimf_b200.so ABI version mismatch: rebuild the library")
        _lib = L
    return _lib


def check(rc: int, err) -> None:
    if rc != _abi.MCF_OK:
        raise McfError(rc, err.value.decode(errors="replace"))
