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
