"""Coarse-grid climate (SURVEY.md H3 / build-plan step 6): modes 2/4 with the climate and point-model series on
runpointmodela's coarse grid, interpolated bilinearly inside the kernels (mcf_problem.clim_rows > 0) instead of being
expanded to [rows, cols, tsteps] arrays on the host as `.runmodel2Cpp` does (R/internal.R:1219-1277).

Checker: the UNMODIFIED compiled reference (runmicro2Cpp / runmicro4Cpp) on the fine arrays that
oracle/prep_oracle.materialise_coarse builds the way the R code does.  Tolerance 1e-6 abs / 1e-6 rel."""
import numpy as np
import pytest

import parity
from microclimf_b200 import api, synth
from oracle import prep_oracle, pyoracle

KIND = "ref" if pyoracle.have_ref() else "oracle"


def test_coarse_problem_layout_cpu():
    p = synth.make_coarse_problem(21, 17, 48, mode=2, crows=5, ccols=4, altcorrect=2)
    assert p.coarse and p.expected_len("temp") == 5 * 4 * 48 and p.expected_len("elevd") == 21 * 17
    s, keep = p.as_struct()
    assert (s.clim_rows, s.clim_cols, s.altcorrect) == (5, 4, 2) and bool(s.relhum) and not bool(s.es)
    # fine cell centres map to coarse fractional indices as terra::resample places them
    assert np.isclose(p.clim_row0 + p.clim_drow * 0, 0.5 * 5 / 21 - 0.5)
    b = p.band(4, 9)
    assert b.cols == 5 and np.isclose(b.clim_col0, p.clim_col0 + p.clim_dcol * 4) and b.arrays["temp"] is p.arrays["temp"]
    q = prep_oracle.materialise_coarse(p)
    assert not q.coarse and q.expected_len("temp") == 21 * 17 * 48
    # the materialised band equals the band of the materialised problem
    qb = prep_oracle.materialise_coarse(b)
    for n in ("temp", "pres", "windspeed", "p_G"):
        assert np.array_equal(qb.arrays[n], q.band(4, 9).arrays[n]), n
    with pytest.raises(ValueError):
        bad = synth.make_coarse_problem(8, 8, 24, mode=2)
        del bad.arrays["relhum"]
        bad.validate()


def _check(p, out_mask=None):
    want = pyoracle.runmicro(prep_oracle.materialise_coarse(p), out_mask=out_mask, kind=KIND)
    got = api.run_problem(p, out=out_mask)
    ok, rows = parity.compare(got, want)
    assert ok, "\n" + parity.fmt(rows)
    return got


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [2, 4])
@pytest.mark.parametrize("altcorrect", [0, 1, 2])
@pytest.mark.parametrize("reqhgt", [0.05, 0.0, 5.0])
def test_coarse_climate_matches_reference_on_expanded_arrays(mode, altcorrect, reqhgt):
    p = synth.make_coarse_problem(37, 29, 24 * 3, reqhgt=reqhgt, mode=mode, crows=5, ccols=4, altcorrect=altcorrect,
                                  nlyr=2)
    _check(p)


@pytest.mark.gpu
@pytest.mark.parametrize("complete", [True, False])
def test_coarse_climate_below_ground(complete):
    p = synth.make_coarse_problem(23, 19, 24 * 4, reqhgt=-0.1, mode=2, crows=4, ccols=3, altcorrect=1, complete=complete)
    _check(p)


@pytest.mark.gpu
def test_coarse_climate_degenerate_grids_and_bands():
    # a single coarse cell (constant field) and a coarse grid finer than the raster along one axis
    _check(synth.make_coarse_problem(16, 12, 48, mode=2, crows=1, ccols=1))
    _check(synth.make_coarse_problem(6, 40, 48, mode=2, crows=9, ccols=3, altcorrect=2))
    # column bands reproduce the whole raster (the coarse grid is replicated, the mapping shifts)
    p = synth.make_coarse_problem(24, 30, 48, mode=2, crows=4, ccols=5, altcorrect=2)
    whole = api.run_problem(p)
    from microclimf_b200 import bands
    lib_sum = bands.twi_partial_host(p.arrays["twi"], p.tfact)
    for c0, c1 in ((0, 11), (11, 30)):
        b = p.band(c0, c1)
        b.twi_mean = lib_sum[0] / lib_sum[1]
        got = api.run_problem(b)
        # not bit-identical: the band gets the whole-raster twi mean from the host reduction and its coarse column
        # origin is clim_col0 + dcol * c0 (one more rounding) — equal to rounding level
        ok, rows = parity.compare(got, {k: v[:, c0:c1, :] for k, v in whole.items()}, atol=1e-9, rtol=1e-9)
        assert ok, "\n" + parity.fmt(rows)


@pytest.mark.gpu
def test_coarse_climate_packed_sink():
    from oracle import packing_oracle
    p = synth.make_coarse_problem(20, 16, 48, mode=4, crows=3, ccols=3, altcorrect=2, nlyr=2)
    fp = api.run_problem(p)
    pk = api.run_problem_packed(p)
    for k in fp:
        assert np.array_equal(pk[k], packing_oracle.pack(k, fp[k])), k
