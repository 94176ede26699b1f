"""The optional FP32 build (BASELINE north_star: "must stay within 0.05 degC and 0.5 % radiation"): k_grid_f32 against the
FP64 CPU checker on the same seeded inputs.  Tolerances written out below; NaN masks must be identical."""
import numpy as np
import pytest

from microclimf_b200 import _abi, api, synth
from oracle import pyoracle

pytestmark = pytest.mark.gpu
KIND = "ref" if pyoracle.have_ref() else "oracle"
TEMP_TOL = 0.05            # degC: Tz, tleaf
RAD_REL, RAD_ABS = 0.005, 0.05   # 0.5 % of the value (+ 0.05 W/m^2 floor for near-zero fluxes)
RH_TOL = 0.5               # percentage points: 0.05 degC moves the saturation pressure by ~0.3 %
OTHER_REL = 0.005          # soil moisture, wind speed


def run_f32(p, mask=None):
    import torch
    d = p.to_device()
    mask = [True] * 10 if mask is None else mask
    outs = [torch.empty(p.ncells * p.tsteps, dtype=torch.float32, device="cuda") if m else None for m in mask]
    for t in outs:
        if t is not None:
            t.fill_(float("nan"))
    api.run_problem_f32_dev(d, outs)
    torch.cuda.synchronize()
    return {n: t.cpu().numpy().astype(np.float64).reshape((p.rows, p.cols, p.tsteps), order="F")
            for n, t in zip(_abi.OUT_NAMES, outs) if t is not None}


def check(got, want):
    worst = {}
    for name, w in want.items():
        g = got[name]
        assert np.array_equal(np.isnan(g), np.isnan(w)), f"{name}: NaN mask differs"
        ok = ~np.isnan(w)
        d = np.abs(g[ok] - w[ok])
        if name in ("Tz", "tleaf"):
            lim = np.full(d.shape, TEMP_TOL)
        elif name == "relhum":
            lim = np.full(d.shape, RH_TOL)
        elif name.startswith("R"):
            lim = RAD_REL * np.abs(w[ok]) + RAD_ABS
        else:
            lim = OTHER_REL * np.abs(w[ok]) + 1e-4
        worst[name] = (float(d.max()), float((d / lim).max()))
    bad = {k: v for k, v in worst.items() if v[1] > 1.0}
    assert not bad, f"beyond the FP32 budget: {bad}\nall: {worst}"
    return worst


@pytest.mark.parametrize("mode", [1, 3])
@pytest.mark.parametrize("reqhgt", [0.05, 0.0, 5.0])
def test_fp32_build_within_budget(mode, reqhgt):
    p = synth.make_problem(41, 33, 24 * 4, reqhgt=reqhgt, mode=mode, nlyr=2)
    mask = [True] * 10 if reqhgt > 0 else [True, False, False, True, False, True, True, True, True, True]
    want = pyoracle.runmicro(p, out_mask=mask, kind=KIND)
    got = run_f32(p, mask)
    # hours beyond the whole days are never written by the device path: compare the computed hours
    worst = check(got, want)
    print(mode, reqhgt, {k: f"{v[0]:.2e}" for k, v in worst.items()})


def test_fp32_seasons_and_latitudes():
    for lat, doy in ((10.0, 80), (65.0, 172), (-35.0, 355)):
        p = synth.make_problem(24, 20, 48, reqhgt=0.5, mode=1, lat=lat, lon=20.0, start_doy=doy)
        check(run_f32(p), pyoracle.runmicro(p, kind=KIND))


@pytest.mark.parametrize("mode", [2, 4])
@pytest.mark.parametrize("reqhgt", [0.05, 0.0, 5.0])
def test_fp32_array_climate(mode, reqhgt):
    """modes 2/4 (fine [rows, cols, hours] arrays): the hour record of each cell-hour is assembled in FP64 and narrowed."""
    p = synth.make_problem(29, 23, 24 * 3, reqhgt=reqhgt, mode=mode, nlyr=2)
    mask = [True] * 10 if reqhgt > 0 else [True, False, False, True, False, True, True, True, True, True]
    check(run_f32(p, mask), pyoracle.runmicro(p, out_mask=mask, kind=KIND))


@pytest.mark.parametrize("altcorrect", [0, 2])
def test_fp32_coarse_grid_climate(altcorrect):
    """coarse-grid climate interpolated in the kernel (ABI 2) against the FP64 reference on the expanded arrays."""
    from oracle import prep_oracle
    p = synth.make_coarse_problem(27, 21, 24 * 3, reqhgt=0.05, mode=2, crows=4, ccols=3, altcorrect=altcorrect)
    want = pyoracle.runmicro(prep_oracle.materialise_coarse(p), kind=KIND)
    check(run_f32(p), want)


@pytest.mark.parametrize("mode,reqhgt,complete", [(1, -0.05, True), (1, -0.6, True), (3, -0.1, False), (2, -0.2, True)])
def test_fp32_below_ground(mode, reqhgt, complete):
    """reqhgt < 0: the time-axis pass is discontinuous in its inputs (rolling-mean length n = round(-118.35 z / mean
    damping depth); a ratio of daily ranges when the series is incomplete), so the FP32 build runs the FP64 hour loops
    there and narrows: Tz and soil moisture agree with the FP64 reference to single-precision rounding."""
    p = synth.make_problem(21, 17, 24 * 6, reqhgt=reqhgt, mode=mode, nlyr=2, complete=complete)
    mask = [True, False, False, True, False, False, False, False, False, False]
    worst = check(run_f32(p, mask), pyoracle.runmicro(p, out_mask=mask, kind=KIND))
    assert worst["Tz"][0] < 1e-4 and worst["soilm"][0] < 1e-6, worst


def test_fp32_host_entry_point_matches_device_entry_point():
    """mcf_runmicro_f32 (host buffers, windowed copy-back) == mcf_runmicro_f32_dev, bit for bit; hours beyond the whole
    days are NaN."""
    p = synth.make_problem(37, 19, 24 * 5 + 5, reqhgt=0.05, mode=1)
    dev = run_f32(p)
    host = api.run_problem_f32(p)
    for nm in _abi.OUT_NAMES:
        assert host[nm].dtype == np.float32
        np.testing.assert_array_equal(host[nm].astype(np.float64), dev[nm], err_msg=nm)
        assert np.isnan(host[nm][:, :, 120:]).all()
    pb = synth.make_problem(9, 8, 24 * 4, reqhgt=-0.1, mode=1)
    hb = api.run_problem_f32(pb, out=[True, False, False, True] + [False] * 6)
    check({k: v.astype(np.float64) for k, v in hb.items()}, pyoracle.runmicro(pb, out_mask=[True, False, False, True] + [False] * 6, kind=KIND))
