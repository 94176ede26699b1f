"""The re-derived snow physics (csrc/mcf_snow_physics.cuh + mcf_snow_drivers.cuh) compiled for the HOST by
tests/hostcheck/snow_host.cpp and compared with the compiled reference (gridmodelsnow1/2, gridmicrosnow1/2,
src/microclimfCpp.cpp:4172-5214) — no GPU needed.  This pins the ALGEBRA of the re-derivation (hoisted hour terms, the
specialised two-stream solution, the cancelled latent heat, the trig-free interception and canopy integrals); the device
build swaps libm for the MUFU-seeded functions of mcf_math.cuh and is pinned by tests/test_snow_gpu.py.  The host build
is test infrastructure: the product library never contains a CPU path."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import parity
from microclimf_b200 import _abi, snow, synth
from oracle import pyoracle

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
pytestmark = pytest.mark.skipif(not pyoracle.have_ref(), reason="compiled reference absent (snow has no C restatement)")


@pytest.fixture(scope="module")
def host():
    out = os.path.join(HERE, "hostcheck", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libsnow_host.so")
    src = os.path.join(HERE, "hostcheck", "snow_host.cpp")
    hdrs = [os.path.join(ROOT, "microclimf_b200", "csrc", h) for h in ("mcf_snow_physics.cuh", "mcf_snow_drivers.cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in [src] + hdrs):
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-shared", "-o", so, src],
                       check=True)
    L = C.CDLL(so)
    for f in ("host_gridmodelsnow", "host_gridmodelsnow2", "host_gridmicrosnow", "host_gridmicrosnow2"):
        getattr(L, f).restype = C.c_int
    return L


def _snowm(r):
    with np.errstate(invalid="ignore"):
        return dict(Tc=r["Tc"], Tg=r["Tg"], totalSWE=r["sdepc"] * r["sden"], groundsnowdepth=r["sdepg"], snowden=r["sden"])


@pytest.mark.parametrize("snowenv,seed", [("Alpine", 5), ("Tundra", 6), ("Taiga", 7), ("Maritime", 8)])
def test_snow_pack_recurrence(host, snowenv, seed):
    s = synth.make_snow_inputs(17, 13, 24 * 9, seed=seed)
    want = pyoracle.gridmodelsnow1(s["obstime"], s["climdata"], s["pointm"], s["vegp"], s["other"], snowenv)
    got = snow.call_gridmodelsnow(host.host_gridmodelsnow, s["obstime"], s["climdata"], s["pointm"], s["vegp"], s["other"], snowenv)
    ok, rows = parity.compare(got, want)
    assert ok, "\n" + parity.fmt(rows)


@pytest.mark.parametrize("reqhgt", [0.05, 0.0, 0.3, 1.0, 6.0])
def test_snow_microclimate(host, reqhgt):
    s = synth.make_snow_inputs(15, 11, 24 * 6, seed=9, reqhgt=max(reqhgt, 0.0))
    model = pyoracle.gridmodelsnow1(s["obstime"], s["climdata"], s["pointm"], s["vegp"], s["other"])
    snowm = _snowm(model)
    rng = np.random.default_rng(3)
    micro = {n: rng.uniform(-5, 5, model["Tc"].shape) for n in _abi.OUT_NAMES}
    out = [True] * 10 if reqhgt > 0 else [True, False, False, True, False, True, True, True, True, True]
    want = pyoracle.gridmicrosnow1(reqhgt, s["obstime"], s["climdata"], snowm, micro, s["vegp"], s["other"], 4.0, out)
    got = snow.call_gridmicrosnow(host.host_gridmicrosnow, reqhgt, s["obstime"], s["climdata"], snowm, micro, s["vegp"], s["other"],
                                  4.0, out)
    assert set(got) == set(want)
    ok, rows = parity.compare(got, want)
    assert ok, "\n" + parity.fmt(rows)


def test_array_climate_variants(host):
    """gridmodelsnow2 / gridmicrosnow2 (ref :4426, :5059): hour records formed per cell-hour, per-cell albedo scan, daily
    extremes gathered per cell; ragged tail without extremes."""
    from test_snow_gpu import _array_inputs

    rows, cols = 13, 9
    s = synth.make_snow_inputs(rows, cols, 24 * 5 + 7, seed=31)
    clim, pointm, other = _array_inputs(s, rows, cols)
    want = pyoracle.gridmodelsnow2(s["obstime"], clim, pointm, s["vegp"], other, "Maritime")
    got = snow.call_gridmodelsnow(host.host_gridmodelsnow2, s["obstime"], clim, pointm, s["vegp"], other, "Maritime")
    ok, rws = parity.compare(got, want)
    assert ok, "\n" + parity.fmt(rws)
    snowm = _snowm(want)
    rng = np.random.default_rng(3)
    micro = {n: rng.uniform(-5, 5, want["Tc"].shape) for n in _abi.OUT_NAMES}
    for reqhgt in (0.05, 0.6):
        w = pyoracle.gridmicrosnow2(reqhgt, s["obstime"], clim, snowm, micro, s["vegp"], other, 3.0, [True] * 10)
        g = snow.call_gridmicrosnow(host.host_gridmicrosnow2, reqhgt, s["obstime"], clim, snowm, micro, s["vegp"], other, 3.0,
                                    [True] * 10)
        ok, rws = parity.compare(g, w)
        assert ok, "\n" + parity.fmt(rws)
