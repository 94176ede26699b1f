"""Host-side description of one grid-model problem (the argument list of the reference's
runmicroNCpp drivers, src/microclimfCpp.cpp:2052/2340/2624/2926) and its packing into the C ABI's
`mcf_problem` (include/microclimf_b200.h).

Arrays are kept in R layout: matrices [rows, cols] and arrays [rows, cols, n] are column-major, so a
numpy array created with order="F" (or any array whose .ravel(order="F") is the R vector) maps 1:1.
Internally every field is stored flat (1-D, float64 / int32, C-contiguous) in that R order.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np

from . import _abi

_F64 = C.POINTER(C.c_double)
_I32 = C.POINTER(C.c_int32)

SERIES_FIELDS = _abi.CLIM_FIELDS[:-1] + _abi.POINTM_FIELDS  # per-hour (modes 1/3) or per-cell-hour (2/4)
OBSTIME_FIELDS = ("year", "month", "day", "hour")


def _flat_f64(a) -> np.ndarray:
    a = np.asarray(a, dtype=np.float64)
    if a.ndim > 1:
        a = a.ravel(order="F")
    return np.ascontiguousarray(a)


def _flat_i32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32).ravel())


@dataclass
class GridProblem:
    mode: int
    rows: int
    cols: int
    tsteps: int
    reqhgt: float
    zref: float
    lat: float = 0.0
    lon: float = 0.0
    Sminp: float = 0.0
    Smaxp: float = 0.0
    tfact: float = 1.5
    mat: float = 10.0
    complete: bool = True
    nlyr: int = 1
    lyr_st: Optional[np.ndarray] = None
    lyr_ed: Optional[np.ndarray] = None
    twi_mean: Optional[float] = None
    # coarse-grid climate (modes 2/4): series are [clim_rows, clim_cols, tsteps]; see include/microclimf_b200.h
    clim_rows: int = 0
    clim_cols: int = 0
    clim_row0: float = 0.0
    clim_drow: float = 0.0
    clim_col0: float = 0.0
    clim_dcol: float = 0.0
    altcorrect: int = 0
    arrays: Dict[str, object] = field(default_factory=dict)  # flat numpy arrays or CUDA torch tensors

    # ------------------------------------------------------------------ construction helpers
    @property
    def ncells(self) -> int:
        return self.rows * self.cols

    @property
    def array_climate(self) -> bool:
        return self.mode in (2, 4)

    @property
    def layered(self) -> bool:
        return self.mode in (3, 4)

    @property
    def coarse(self) -> bool:
        return self.array_climate and self.clim_rows > 0

    def set(self, name: str, value) -> None:
        if name in ("year", "month", "day"):
            self.arrays[name] = _flat_i32(value)
        else:
            self.arrays[name] = _flat_f64(value)

    def expected_len(self, name: str) -> int:
        nc, T = self.ncells, self.tsteps
        if name in OBSTIME_FIELDS or name == "winddir":
            return T
        if name in SERIES_FIELDS or name in ("relhum", "wu", "wv"):
            if self.coarse:
                return self.clim_rows * self.clim_cols * T
            return nc * T if self.array_climate else T
        if name in ("elevd", "pfac"):
            return nc
        if name in _abi.VEG_FIELDS:
            return nc * (self.nlyr if self.layered else 1)
        if name == "wsa":
            return nc * 8
        if name == "hor":
            return nc * 24
        if name in _abi.SOIL_FIELDS or name in ("lats", "lons"):
            return nc
        raise KeyError(name)

    def validate(self) -> None:
        if self.mode not in (1, 2, 3, 4):
            raise ValueError("mode must be 1..4")
        clim = list(_abi.CLIM_FIELDS)
        if self.coarse:
            clim = ["temp", "relhum", "pres", "swdown", "difrad", "lwdown", "wu", "wv", "winddir"]
            if self.altcorrect not in (0, 1, 2):
                raise ValueError("altcorrect must be 0, 1 or 2")
            if self.altcorrect:
                clim += ["elevd", "pfac"]
        required = list(OBSTIME_FIELDS) + clim + list(_abi.VEG_FIELDS) + list(_abi.SOIL_FIELDS)
        required += ["p_soilm", "p_G", "p_umu", "p_kp", "p_muGp", "p_dtrp"]
        if self.reqhgt < 0:
            required += ["p_Tg", "p_Tbp"]
        if self.array_climate:
            required += ["lats", "lons"]
        for n in required:
            if n not in self.arrays:
                raise ValueError(f"missing input '{n}'")
        for n, a in self.arrays.items():
            ln = int(a.numel()) if hasattr(a, "numel") else int(a.size)
            if ln != self.expected_len(n):
                raise ValueError(f"input '{n}' has {ln} elements, expected {self.expected_len(n)}")
        if self.layered:
            if self.lyr_st is None or self.lyr_ed is None or len(self.lyr_st) != self.nlyr:
                raise ValueError("layered modes need lyr_st/lyr_ed of length nlyr")

    # ------------------------------------------------------------------ packing
    def as_struct(self):
        """Returns (McfProblem, keepalive list). Pointers are host or device according to storage."""
        keep = []
        s = _abi.McfProblem()
        s.mode, s.rows, s.cols, s.tsteps = self.mode, self.rows, self.cols, self.tsteps
        s.nlyr = self.nlyr if self.layered else 1
        s.complete = 1 if self.complete else 0
        s.reqhgt, s.zref, s.lat, s.lon = self.reqhgt, self.zref, self.lat, self.lon
        s.Sminp, s.Smaxp, s.tfact, s.mat = self.Sminp, self.Smaxp, self.tfact, self.mat
        if self.layered:
            st, ed = _flat_i32(self.lyr_st), _flat_i32(self.lyr_ed)
            keep += [st, ed]
            s.lyr_st = st.ctypes.data_as(_I32)
            s.lyr_ed = ed.ctypes.data_as(_I32)
        for name, a in self.arrays.items():
            is_int = name in ("year", "month", "day")
            if hasattr(a, "data_ptr"):  # torch tensor (device-resident)
                if is_int:
                    # calendar fields stay on the host (see mcf_runmicro_dev contract)
                    a = a.cpu().numpy()
                else:
                    keep.append(a)
                    setattr(s, name, C.cast(C.c_void_p(a.data_ptr()), _F64))
                    continue
            keep.append(a)
            setattr(s, name, a.ctypes.data_as(_I32 if is_int else _F64))
        if self.twi_mean is not None:
            s.has_twi_mean, s.twi_mean = 1, float(self.twi_mean)
        if self.coarse:
            s.clim_rows, s.clim_cols = int(self.clim_rows), int(self.clim_cols)
            s.clim_row0, s.clim_drow = float(self.clim_row0), float(self.clim_drow)
            s.clim_col0, s.clim_dcol = float(self.clim_col0), float(self.clim_dcol)
            s.altcorrect = int(self.altcorrect)
        return s, keep

    # ------------------------------------------------------------------ transforms
    def to_device(self, device="cuda"):
        """Copy of this problem whose FP64 arrays are CUDA tensors (calendar ints stay numpy)."""
        import torch

        out = self._clone_meta()
        for n, a in self.arrays.items():
            if n in ("year", "month", "day"):
                out.arrays[n] = a
            elif hasattr(a, "data_ptr"):
                out.arrays[n] = a.to(device)
            else:
                out.arrays[n] = torch.from_numpy(a).to(device)
        return out

    def _clone_meta(self) -> "GridProblem":
        return GridProblem(mode=self.mode, rows=self.rows, cols=self.cols, tsteps=self.tsteps, reqhgt=self.reqhgt,
                           zref=self.zref, lat=self.lat, lon=self.lon, Sminp=self.Sminp, Smaxp=self.Smaxp,
                           tfact=self.tfact, mat=self.mat, complete=self.complete, nlyr=self.nlyr,
                           lyr_st=None if self.lyr_st is None else np.array(self.lyr_st, dtype=np.int32),
                           lyr_ed=None if self.lyr_ed is None else np.array(self.lyr_ed, dtype=np.int32),
                           twi_mean=self.twi_mean, clim_rows=self.clim_rows, clim_cols=self.clim_cols,
                           clim_row0=self.clim_row0, clim_drow=self.clim_drow, clim_col0=self.clim_col0,
                           clim_dcol=self.clim_dcol, altcorrect=self.altcorrect)

    def band(self, c0: int, c1: int) -> "GridProblem":
        """Column band [c0, c1) of a host problem: the unit of multi-GPU sharding.  In R layout a
        column band is a contiguous slab of every [rows, cols, ...] slice."""
        out = self._clone_meta()
        out.cols = c1 - c0
        out.clim_col0 = self.clim_col0 + self.clim_dcol * c0  # the coarse grid is replicated, the mapping shifts
        R, Cc, T = self.rows, self.cols, self.tsteps
        for n, a in self.arrays.items():
            ln = self.expected_len(n)
            if ln in (T,) and (n in OBSTIME_FIELDS or n == "winddir" or (n in SERIES_FIELDS and not self.array_climate)):
                out.arrays[n] = a
                continue
            if self.coarse and (n in SERIES_FIELDS or n in ("relhum", "wu", "wv")):
                out.arrays[n] = a
                continue
            nsl = ln // (R * Cc)
            v = np.asarray(a).reshape(nsl, Cc, R)[:, c0:c1, :]
            out.arrays[n] = np.ascontiguousarray(v).ravel()
        return out

    def replace(self, **kw) -> "GridProblem":
        out = self._clone_meta()
        out.arrays = dict(self.arrays)
        for k, v in kw.items():
            setattr(out, k, v)
        return out
