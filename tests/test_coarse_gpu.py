"""Coarse-grid climate (SURVEY.md H3 / build-plan step 6): modes 2/4 with the climate and point-model series on
runpointmodela's coarse grid, interpolated bilinearly inside the kernels (mcf_problem.clim_rows > 0) instead of being
expanded to [rows, cols, tsteps] arrays on the host as `.runmodel2Cpp` does (R/internal.R:1219-1277).

Checker: the UNMODIFIED compiled reference (runmicro2Cpp / runmicro4Cpp) on the fine arrays that
oracle/prep_oracle.materialise_coarse builds the way the R code does.  Tolerance 1e-6 abs / 1e-6 rel."""
import numpy as np
import pytest

import parity
from microclimf_b200 import _abi, api, synth
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


def test_sampled_expansion_equals_sample_of_expansion_cpu():
    """materialise_coarse(p, pick) — what the raster-size test feeds the checker — is the sampled full expansion."""
    p = synth.make_coarse_problem(19, 13, 48, mode=2, crows=4, ccols=3, altcorrect=2)
    pick = np.array([0, 5, 18, 19, 100, 246])
    full = prep_oracle.materialise_coarse(p)
    sub = prep_oracle.materialise_coarse(p, pick)
    assert (sub.rows, sub.cols) == (len(pick), 1) and not sub.coarse
    nc = p.ncells
    for n, a in sub.arrays.items():
        ln = full.expected_len(n)
        f = np.asarray(full.arrays[n])
        if ln % nc == 0 and ln >= nc and ln != p.tsteps:
            np.testing.assert_array_equal(a, f.reshape(ln // nc, nc)[:, pick].ravel(), err_msg=n)
        else:
            np.testing.assert_array_equal(a, f, err_msg=n)


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


@pytest.mark.gpu
def test_config5_raster_size_sampled_against_reference():
    """BASELINE configs[4]: gridded climate on a 4096 x 4096 raster (41 x 41 climate grid, altcorrect 2), device-resident,
    one day.  The reference's host-side expansion would be 15 arrays x 16.8 M cells x 24 h; here 300 sampled cells are
    expanded by the numpy restatement of `.runmodel2Cpp` (oracle/prep_oracle.materialise_coarse with `pick`) and solved
    by the CPU checker as a 300 x 1 raster with the sample's own twi mean (tests/sampling.py); every other cell must be
    NA exactly where the raster is NA and finite elsewhere."""
    import torch

    import sampling
    from oracle import prep_oracle

    rows = cols = 4096
    T = 24
    p = synth.make_coarse_problem(rows, cols, T, reqhgt=0.05, mode=2, crows=41, ccols=41, altcorrect=2, seed=11)
    pick = sampling.pick_cells(p, 300, seed=8)
    p.twi_mean = sampling.sample_twi_mean(p, pick)
    dp = p.to_device()
    nc = p.ncells
    outs = [torch.empty(T * nc, dtype=torch.float64, device="cuda") for _ in range(10)]
    api.run_problem_dev(dp, outs)
    torch.cuda.synchronize()
    sub = prep_oracle.materialise_coarse(p, pick)
    want = pyoracle.runmicro(sub, kind=KIND)
    na = torch.from_numpy(np.isnan(np.asarray(p.arrays["hgt"])[:nc])).cuda()
    idx = torch.from_numpy(pick).cuda()
    got = {}
    for nm, o in zip(_abi.OUT_NAMES, outs):
        full = o.view(T, nc)
        assert bool(torch.isnan(full[:, na]).all()) and bool(torch.isfinite(full[:, ~na]).all()), nm
        got[nm] = np.ascontiguousarray(full[:, idx].cpu().numpy().T).reshape(len(pick), 1, T)
    ok, rws = parity.compare(got, want)
    assert ok, "\n" + parity.fmt(rws)
