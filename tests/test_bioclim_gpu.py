"""runbioclim1..4Cpp parity (ref src/microclimfCpp.cpp:3457-3700) through mcf_runbioclim."""
import numpy as np
import pytest

import parity
from microclimf_b200 import _abi, _lib, api, synth
from oracle import pyoracle

pytestmark = pytest.mark.gpu
KIND = "ref" if pyoracle.have_ref() else "oracle"


def _problem(mode, reqhgt, rows=21, cols=18):
    days, q = synth.bioclim_days()
    p = synth.make_problem(rows, cols, 336, reqhgt=reqhgt, mode=mode, nlyr=14, day_list=days)
    return p, q


@pytest.mark.parametrize("mode", [1, 2, 3, 4])
@pytest.mark.parametrize("reqhgt,air", [(0.05, True), (0.05, False), (0.0, True), (-0.1, True), (5.0, True)])
def test_bioclim_parity(mode, reqhgt, air):
    p, q = _problem(mode, reqhgt)
    want = pyoracle.runbioclim(p, q, air=air, kind=KIND)
    got = api.run_bioclim_problem(p, q["wetq"], q["dryq"], q["hotq"], q["colq"], air=air)
    ok, rows = parity.compare(got, want)
    assert ok, "\n" + parity.fmt(rows)


def test_bioclim_output_mask_and_quirks():
    """bio3 = bio2 / bio7 without the x100 (ref :3528-3536); NA cells carry R's NA payload.  The reference
    reads bio2/bio5/bio6 from unallocated 0x0 matrices when bio3/bio7 are requested without them
    (undefined behaviour, :3533-3534); the checker is therefore only run with the dependencies on, and
    the CUDA path is checked to give the same bio3/bio7 with the dependencies off."""
    p, q = _problem(1, 0.05)
    mask = [False] * 19
    for b in (2, 3, 5, 6, 7, 15):
        mask[b - 1] = True
    want = pyoracle.runbioclim(p, q, air=True, out_mask=mask, kind=KIND)
    got = api.run_bioclim_problem(p, q["wetq"], q["dryq"], q["hotq"], q["colq"], air=True, out=mask)
    assert set(got) == {"bio2", "bio3", "bio5", "bio6", "bio7", "bio15"}
    ok, rows = parity.compare(got, want)
    assert ok, "\n" + parity.fmt(rows)
    lean = [False] * 19
    lean[2] = lean[6] = True
    g2 = api.run_bioclim_problem(p, q["wetq"], q["dryq"], q["hotq"], q["colq"], air=True, out=lean)
    assert set(g2) == {"bio3", "bio7"}
    np.testing.assert_array_equal(g2["bio3"], got["bio3"])
    np.testing.assert_array_equal(g2["bio7"], got["bio7"])
    full = api.run_bioclim_problem(p, q["wetq"], q["dryq"], q["hotq"], q["colq"], air=True)
    np.testing.assert_allclose(full["bio3"], full["bio2"] / full["bio7"], rtol=1e-12, equal_nan=True)
    na = np.isnan(p.arrays["hgt"][:p.ncells].reshape(p.rows, p.cols, order="F"))
    assert (full["bio1"].view(np.uint64)[na] == _abi.NA_REAL_BITS).all()


def test_bioclim_short_quarters_and_errors():
    p, q = _problem(1, 0.05, 9, 9)
    q2 = dict(q, wetq=q["wetq"][:48])  # the reference still divides by 72 (:3325)
    want = pyoracle.runbioclim(p, q2, air=True, kind=KIND)
    got = api.run_bioclim_problem(p, q2["wetq"], q2["dryq"], q2["hotq"], q2["colq"], air=True)
    ok, rows = parity.compare(got, want)
    assert ok, "\n" + parity.fmt(rows)
    bad = synth.make_problem(4, 4, 48, reqhgt=0.05, mode=1)
    with pytest.raises(_lib.McfError):
        api.run_bioclim_problem(bad, q["wetq"], q["dryq"], q["hotq"], q["colq"])


@pytest.mark.parametrize("mode,tsteps", [(1, 336 + 48), (2, 336 + 24), (3, 336 + 24), (1, 336 + 13)])
def test_bioclim_longer_series_and_uncovered_hours(mode, tsteps):
    """runbioclimCpp takes whatever length it is given (ref :3461-3464): the soil statistics run over ALL hours, the
    temperature statistics over the first 336.  Hours no whole day (or none of the 14 hard-coded layers, ref :3635-3646)
    covers stay NA in the reference's arrays: the 336-hour soil sd, and with it bio15, is then NA."""
    days, q = synth.bioclim_days()
    extra = (tsteps - 336 + 23) // 24
    p = synth.make_problem(11, 7, 24 * (14 + extra), reqhgt=0.05, mode=mode, nlyr=14,
                           day_list=np.concatenate([days, days[:extra]]))
    if tsteps != p.tsteps:  # ragged tail: drop the last hours of every per-hour series
        T0, nc = p.tsteps, p.ncells
        for n, a in list(p.arrays.items()):
            ln = p.expected_len(n)
            if ln == T0:
                p.arrays[n] = np.ascontiguousarray(a[:tsteps])
            elif ln == T0 * nc and n not in _abi.VEG_FIELDS and n not in _abi.SOIL_FIELDS:
                p.arrays[n] = np.ascontiguousarray(a.reshape(T0, nc)[:tsteps].ravel())
        p.tsteps = tsteps
        p.validate()
    want = pyoracle.runbioclim(p, q, air=True, kind=KIND)
    got = api.run_bioclim_problem(p, q["wetq"], q["dryq"], q["hotq"], q["colq"], air=True)
    ok, rows = parity.compare(got, want)
    assert ok, "\n" + parity.fmt(rows)


def test_bioclim_unordered_and_repeated_quarter_indices():
    """The quarter vectors are plain index lists (ref bioclim8 :3317-3327): any order, repeats allowed."""
    p, q = _problem(1, 0.05, 9, 8)
    rng = np.random.default_rng(5)
    q2 = {k: rng.integers(0, 336, 72).astype(np.int32) for k in q}
    q2["wetq"][:10] = q2["wetq"][10]  # one hour eleven times
    want = pyoracle.runbioclim(p, q2, air=True, kind=KIND)
    got = api.run_bioclim_problem(p, q2["wetq"], q2["dryq"], q2["hotq"], q2["colq"], air=True)
    ok, rows = parity.compare(got, want)
    assert ok, "\n" + parity.fmt(rows)


@pytest.mark.parametrize("reqhgt,air", [(0.05, True), (-0.08, True), (0.05, False)])
def test_bioclim_config3_raster_sampled_against_reference(reqhgt, air):
    """BASELINE configs[2]: runbioclim on a synthetic 2048 x 2048 raster (the reference would materialise two
    [2048, 2048, 336] arrays = 22.6 GB; here nothing hourly exists above ground, and below ground the series is reduced
    chunk by chunk).  300 sampled cells are re-solved by the CPU checker as a 300 x 1 raster (tests/sampling.py);
    every other cell must be NA exactly where the raster is NA and finite elsewhere."""
    import sampling

    days, q = synth.bioclim_days()
    p = synth.make_problem(2048, 2048, 336, reqhgt=reqhgt, mode=1, day_list=days, seed=77)
    pick = sampling.pick_cells(p, 300, seed=3)
    p.twi_mean = sampling.sample_twi_mean(p, pick)
    got = api.run_bioclim_problem(p, q["wetq"], q["dryq"], q["hotq"], q["colq"], air=air)
    sub = sampling.subproblem(p, pick)
    want = pyoracle.runbioclim(sub, q, air=air, kind=KIND)
    na = np.isnan(p.arrays["hgt"]).reshape(p.rows, p.cols, order="F")
    sampled = {}
    for nm in _abi.BIO_NAMES:
        g = got[nm]
        assert (g.view(np.uint64)[na] == _abi.NA_REAL_BITS).all(), nm
        assert np.isfinite(g[~na]).all(), nm
        sampled[nm] = g.ravel(order="F")[pick].reshape(-1, 1)
    ok, rows = parity.compare(sampled, want)
    assert ok, "\n" + parity.fmt(rows)
