"""ctypes mirror of include/microclimf_b200.h (struct layouts and constants).

Kept free of any library loading so that the CPU checkers under oracle/ can share the same
`mcf_problem` layout (oracle/ref_driver.cpp and oracle/mcf_oracle.c take the product's struct).
"""
from __future__ import annotations

import ctypes as C

MCF_ABI_VERSION = 2
MCF_OK, MCF_ERR_ARG, MCF_ERR_CUDA, MCF_ERR_NOMEM = 0, 1, 2, 3
MCF_NOUT = 10
MCF_NBIO = 19
NA_REAL_BITS = 0x7FF00000000007A2

# order of the reference's `out` logical vector (src/microclimfCpp.cpp:2131-2140, 2325-2336)
OUT_NAMES = ("Tz", "tleaf", "relhum", "soilm", "windspeed", "Rdirdown", "Rdifdown", "Rlwdown", "Rswup", "Rlwup")
BIO_NAMES = tuple(f"bio{i}" for i in range(1, 20))

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int32)

# (field, ctype) in header order
_PROBLEM_FIELDS = [
    ("mode", C.c_int32), ("rows", C.c_int32), ("cols", C.c_int32), ("tsteps", C.c_int32),
    ("nlyr", C.c_int32), ("complete", C.c_int32),
    ("reqhgt", C.c_double), ("zref", C.c_double), ("lat", C.c_double), ("lon", C.c_double),
    ("Sminp", C.c_double), ("Smaxp", C.c_double), ("tfact", C.c_double), ("mat", C.c_double),
    ("lyr_st", _pi), ("lyr_ed", _pi),
    ("year", _pi), ("month", _pi), ("day", _pi), ("hour", _pd),
]
CLIM_FIELDS = ("temp", "es", "ea", "tdew", "pres", "swdown", "difrad", "lwdown", "windspeed", "winddir")
POINTM_FIELDS = ("p_soilm", "p_Tg", "p_Tbp", "p_G", "p_umu", "p_kp", "p_muGp", "p_dtrp")
VEG_FIELDS = ("hgt", "pai", "x", "gsmax", "leafr", "leaft", "clump", "leafd", "paia", "leafden")
SOIL_FIELDS = ("Smin", "Smax", "gref", "soilb", "Psie", "Vq", "Vm", "Mc", "rho", "slope", "aspect", "twi",
               "svfa", "wsa", "hor")
for _n in CLIM_FIELDS + POINTM_FIELDS + VEG_FIELDS + SOIL_FIELDS + ("lats", "lons"):
    _PROBLEM_FIELDS.append((_n, _pd))
_PROBLEM_FIELDS += [("has_twi_mean", C.c_int32), ("twi_mean", C.c_double)]
# coarse-grid climate (ABI 2)
COARSE_FIELDS = ("relhum", "wu", "wv", "elevd", "pfac")
_PROBLEM_FIELDS += [("clim_rows", C.c_int32), ("clim_cols", C.c_int32), ("clim_row0", C.c_double),
                    ("clim_drow", C.c_double), ("clim_col0", C.c_double), ("clim_dcol", C.c_double),
                    ("altcorrect", C.c_int32)]
for _n in COARSE_FIELDS:
    _PROBLEM_FIELDS.append((_n, _pd))


class McfProblem(C.Structure):
    _fields_ = _PROBLEM_FIELDS


class McfWindow(C.Structure):
    _fields_ = [("block0", C.c_int32), ("nblocks", C.c_int32), ("hour0", C.c_int64), ("ring_hours", C.c_int64)]


OutPtrs = _pd * MCF_NOUT
OutPtrs16 = C.POINTER(C.c_int16) * MCF_NOUT
OutPtrsF = C.POINTER(C.c_float) * MCF_NOUT
PACKED_NA = -9999
PACK_SCALE = (100.0, 100.0, 1.0, 100.0, 100.0, 1.0, 1.0, 1.0, 1.0, 1.0)  # writetonc, R/dataprep.R:1164-1173
BioPtrs = _pd * MCF_NBIO

# every symbol include/microclimf_b200.h declares (tests check the built library exports them all)
EXPORTED_SYMBOLS = (
    "mcf_runmicro", "mcf_runbioclim", "mcf_runmicro_dev", "mcf_runbioclim_dev", "mcf_twi_partial",
    "mcf_abi_version", "mcf_device_count", "mcf_set_device", "mcf_launch_count", "mcf_launch_count_reset",
    "mcf_kernel_time", "mcf_kernel_time_reset", "mcf_kernel_timing_enable", "mcf_fp64_peak", "mcf_math_eval", "mcf_release_workspace", "mcf_horizon", "mcf_windcoef", "mcf_flowacc", "mcf_runmicro_packed", "mcf_runmicro_packed_dev", "mcf_gridmodelsnow", "mcf_gridmicrosnow", "mcf_gridmodelsnow2", "mcf_gridmicrosnow2", "mcf_runmicro_f32_dev",
    "mcf_runmicro_summary", "mcf_runmicro_summary_dev", "mcf_windshelter", "mcf_slope_aspect", "mcf_topidx", "mcf_runmicro_f32",
)
