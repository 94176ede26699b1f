"""The summary sink (mcf_runmicro_summary[_dev]): per-cell mean / min / max over the computed hours, reduced inside
the grid kernel.  Checked against the same statistics of the CPU checker's hourly arrays (ref runmicro1..4Cpp,
src/microclimfCpp.cpp:2052-3223) — the reference has no such operator, its arrays are the oracle."""
import numpy as np
import pytest

import sampling
from microclimf_b200 import _abi, _lib, api, synth
from oracle import pyoracle

pytestmark = pytest.mark.gpu
KIND = "ref" if pyoracle.have_ref() else "oracle"


def _stats(arr, hours):
    """mean / min / max over the first `hours` slices, with the sink's NaN rules (NaN poisons the mean, the extremes
    skip it and start from +inf / -inf)."""
    a = arr[:, :, :hours]
    with np.errstate(invalid="ignore"):
        return {"mean": a.sum(axis=2) / hours, "min": np.fmin.reduce(a, axis=2, initial=np.inf),
                "max": np.fmax.reduce(a, axis=2, initial=-np.inf)}


def _close(got, want, what):
    gn, wn = np.isnan(got), np.isnan(want)
    assert np.array_equal(gn, wn), what
    m = ~wn & np.isfinite(want)
    assert np.array_equal(got[~wn & ~m], want[~wn & ~m]), what  # infinities of all-NaN series
    err = np.abs(got[m] - want[m])
    assert np.all(err <= 1e-6 + 1e-6 * np.abs(want[m])), f"{what}: max err {err.max():.3g}"


@pytest.mark.parametrize("mode", [1, 2, 3, 4])
@pytest.mark.parametrize("reqhgt", [0.05, 0.0, 5.0])
def test_summary_matches_statistics_of_reference_arrays(mode, reqhgt):
    p = synth.make_problem(23, 19, 24 * 5 + 7, reqhgt=reqhgt, mode=mode, nlyr=3)  # 7 trailing hours are never computed
    want = pyoracle.runmicro(p, kind=KIND)
    got, hours = api.run_summary(p)
    nblocks = 5 if mode <= 2 else sum((e - s + 1) // 24 for s, e in zip(p.lyr_st, p.lyr_ed))
    assert hours == 24 * nblocks
    na_cell = np.isnan(p.arrays["hgt"][:p.ncells].reshape(p.rows, p.cols, order="F"))
    for nm in _abi.OUT_NAMES:
        w = want[nm]
        written = ~np.all(np.isnan(w), axis=(0, 1))  # hours the reference computed (whole days / layer spans)
        if not written.any():  # tleaf / relhum at the surface: NA throughout
            for st in api.SUMMARY_STATS:
                assert (got[nm][st].view(np.uint64) == _abi.NA_REAL_BITS).all(), (nm, st)
            continue
        assert written.sum() == hours, nm
        ws = _stats(w[:, :, written], hours)
        for st in api.SUMMARY_STATS:
            g = got[nm][st]
            assert (g.view(np.uint64)[na_cell] == _abi.NA_REAL_BITS).all(), (nm, st)
            _close(g[~na_cell], ws[st][~na_cell], f"{nm} {st} mode {mode} reqhgt {reqhgt}")


def test_summary_windows_accumulate_and_mask():
    """Successive windows merged with accumulate = 1 equal one launch over the whole series; outputs that are not
    requested are not touched; reqhgt < 0 is refused."""
    import torch

    p = synth.make_problem(31, 9, 24 * 6, reqhgt=0.05, mode=1)
    dp = p.to_device()
    nc = p.ncells

    def bufs(mask):
        return [[torch.full((nc,), -7.0, dtype=torch.float64, device="cuda") if m else None for m in mask] for _ in range(3)]

    full = [True] * 10
    a = bufs(full)
    assert api.run_summary_dev(dp, *a) == 144
    b = bufs(full)
    assert api.run_summary_dev(dp, *b, window=(0, 2, 0, 144)) == 48
    assert api.run_summary_dev(dp, *b, window=(2, 3, 0, 144), accumulate=True) == 72
    assert api.run_summary_dev(dp, *b, window=(5, 1, 0, 144), accumulate=True) == 24
    torch.cuda.synchronize()
    for k in range(3):
        for v in range(10):
            x, y = a[k][v].cpu().numpy(), b[k][v].cpu().numpy()
            if k == 0:
                np.testing.assert_allclose(y, x, rtol=1e-13, equal_nan=True)
            else:
                np.testing.assert_array_equal(y, x)
    mask = [v in (0, 3) for v in range(10)]
    c = bufs(mask)
    api.run_summary_dev(dp, *c)
    torch.cuda.synchronize()
    for k in range(3):
        np.testing.assert_array_equal(c[k][0].cpu().numpy(), a[k][0].cpu().numpy())
        np.testing.assert_array_equal(c[k][3].cpu().numpy(), a[k][3].cpu().numpy())
    below = synth.make_problem(5, 5, 48, reqhgt=-0.1, mode=1)
    with pytest.raises(_lib.McfError) as ei:
        api.run_summary(below)
    assert ei.value.code == _abi.MCF_ERR_ARG


def test_year_long_band_summary_sampled_against_reference():
    """What runmicro_big's summary sink holds after a whole year on a band-sized raster (2048 x 512 cells x 8760 h, one
    launch): 300 sampled cells re-solved hour by hour on the CPU checker (26 M cell-hours would be the whole band)."""
    import torch

    rows, cols, T = 2048, 512, 8760
    p = synth.make_problem(rows, cols, T, reqhgt=0.05, mode=1, seed=31)
    pick = sampling.pick_cells(p, 300, seed=8)
    p.twi_mean = sampling.sample_twi_mean(p, pick)
    dp = p.to_device()
    nc = p.ncells
    s = [[torch.empty(nc, dtype=torch.float64, device="cuda") for _ in range(10)] for _ in range(3)]
    hours = api.run_summary_dev(dp, *s)
    torch.cuda.synchronize()
    assert hours == 8760
    sub = sampling.subproblem(p, pick)
    want = pyoracle.runmicro(sub, kind=KIND)
    idx = torch.from_numpy(pick).cuda()
    na_cell = np.isnan(sub.arrays["hgt"]).reshape(-1, 1)
    for v, nm in enumerate(_abi.OUT_NAMES):
        ws = _stats(want[nm], hours)
        for k, st in enumerate(api.SUMMARY_STATS):
            g = s[k][v][idx].cpu().numpy().reshape(-1, 1)
            if st == "mean":
                g = np.where(np.isnan(g), g, g / hours)
            assert np.isnan(g[na_cell]).all()
            _close(g[~na_cell], ws[st][~na_cell], f"{nm} {st}")
