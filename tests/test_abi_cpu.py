"""CPU-side checks: the C-ABI library loads and exports every symbol include/microclimf_b200.h
declares, the ctypes mirror matches the header, host-side packing / banding logic."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from microclimf_b200 import _abi, _lib, bands, synth
from microclimf_b200.problem import GridProblem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "microclimf_b200.h")


def _build_if_needed():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()


def test_library_exports_every_declared_symbol():
    _build_if_needed()
    src = open(HEADER).read()
    declared = set(re.findall(r"^\s*(?:int|int64_t|void)\s+(mcf_\w+)\s*\(", src, flags=re.M))
    assert declared == set(_abi.EXPORTED_SYMBOLS)
    L = C.CDLL(_lib.LIB_PATH)
    for s in declared:
        assert hasattr(L, s), s
    L.mcf_abi_version.restype = C.c_int
    assert L.mcf_abi_version() == _abi.MCF_ABI_VERSION


def test_struct_mirror_matches_header():
    src = open(HEADER).read()
    body = src[src.index("typedef struct mcf_problem {") + len("typedef struct mcf_problem {"):src.index("} mcf_problem;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        parts = decl.split(",")
        names.append(parts[0].split()[-1].lstrip("*"))
        for extra in parts[1:]:
            names.append(extra.strip().lstrip("*"))
    assert names == [f[0] for f in _abi.McfProblem._fields_]
    nptr = 2 + 4 + len(_abi.CLIM_FIELDS) + len(_abi.POINTM_FIELDS) + len(_abi.VEG_FIELDS) + len(_abi.SOIL_FIELDS) + 2
    coarse = 8 + 32 + 8 + 8 * len(_abi.COARSE_FIELDS)  # clim_rows/cols, 4 mapping doubles, altcorrect (+ pad), 5 pointers
    assert C.sizeof(_abi.McfProblem) == 24 + 64 + 8 * nptr + 16 + coarse


def test_no_device_fails_loudly():
    """Without a GPU every compute entry point must fail (no CPU fallback)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    _build_if_needed()
    from microclimf_b200 import api

    p = synth.make_problem(4, 4, 24, reqhgt=0.05, mode=1)
    with pytest.raises(_lib.McfError) as ei:
        api.run_problem(p)
    assert ei.value.code == _abi.MCF_ERR_CUDA


def test_problem_validation_and_layout():
    p = synth.make_problem(5, 7, 48, reqhgt=0.05, mode=3, nlyr=2)
    assert p.arrays["hgt"].size == 5 * 7 * 2 and p.arrays["hor"].size == 5 * 7 * 24
    s, keep = p.as_struct()
    assert s.rows == 5 and s.cols == 7 and s.nlyr == 2 and s.mode == 3
    got = np.ctypeslib.as_array(s.hor, shape=(5 * 7 * 24,))
    np.testing.assert_array_equal(got, p.arrays["hor"])
    bad = p.replace()
    bad.arrays["pai"] = bad.arrays["pai"][:-1]
    with pytest.raises(ValueError):
        bad.validate()
    q = GridProblem(mode=2, rows=2, cols=2, tsteps=24, reqhgt=0.0, zref=2.0)
    with pytest.raises(ValueError):
        q.validate()


def test_band_slicing_is_r_layout():
    p = synth.make_problem(6, 9, 24, reqhgt=0.05, mode=2)
    R, Cc, T = p.rows, p.cols, p.tsteps
    b = p.band(3, 7)
    assert b.cols == 4
    full = p.arrays["temp"].reshape(R, Cc, T, order="F")
    np.testing.assert_array_equal(b.arrays["temp"].reshape(R, 4, T, order="F"), full[:, 3:7, :])
    hor = p.arrays["hor"].reshape(R, Cc, 24, order="F")
    np.testing.assert_array_equal(b.arrays["hor"].reshape(R, 4, 24, order="F"), hor[:, 3:7, :])
    np.testing.assert_array_equal(b.arrays["winddir"], p.arrays["winddir"])
    assert bands.band_ranges(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert bands.band_ranges(8192, 8)[-1] == (7168, 8192)


def test_api_mirror_builds_same_problem():
    """The runmicroNCpp mirrors pack named lists exactly like GridProblem does."""
    from microclimf_b200 import api

    p = synth.make_problem(4, 5, 48, reqhgt=0.05, mode=4, nlyr=2)
    A = p.arrays

    def shp3(a, n):
        return a.reshape(p.rows, p.cols, n, order="F")

    obstime = {k: A[k] for k in ("year", "month", "day", "hour")}
    clim = dict(tc=shp3(A["temp"], 48), es=shp3(A["es"], 48), ea=shp3(A["ea"], 48), tdew=shp3(A["tdew"], 48),
                pk=shp3(A["pres"], 48), swdown=shp3(A["swdown"], 48), difrad=shp3(A["difrad"], 48),
                lwdown=shp3(A["lwdown"], 48), windspeed=shp3(A["windspeed"], 48), winddir=A["winddir"])
    pointm = dict(soilm=shp3(A["p_soilm"], 48), Tg=shp3(A["p_Tg"], 48), Tbp=shp3(A["p_Tbp"], 48),
                  Gp=shp3(A["p_G"], 48), umu=shp3(A["p_umu"], 48), kp=shp3(A["p_kp"], 48),
                  muGp=shp3(A["p_muGp"], 48), dtrp=shp3(A["p_dtrp"], 48))
    vegp = {k: shp3(A[k], 2) for k in _abi.VEG_FIELDS}
    soilc = {k: A[k].reshape(p.rows, p.cols, order="F") for k in _abi.SOIL_FIELDS if k not in ("wsa", "hor")}
    soilc["wsa"] = shp3(A["wsa"], 8)
    soilc["hor"] = shp3(A["hor"], 24)
    q = api._problem(4, dict(st=p.lyr_st, ed=p.lyr_ed), obstime, clim, pointm, vegp, soilc, p.reqhgt, p.zref, 0, 0,
                     A["lats"].reshape(p.rows, p.cols, order="F"), A["lons"].reshape(p.rows, p.cols, order="F"),
                     p.Sminp, p.Smaxp, p.tfact, p.complete, p.mat)
    assert set(q.arrays) == set(p.arrays)
    for k in p.arrays:
        np.testing.assert_array_equal(q.arrays[k], p.arrays[k], err_msg=k)
