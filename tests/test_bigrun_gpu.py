"""runmicro_big on column bands (microclimf_b200/bigrun.py; ref R/Cppwrappers.R:444-543): the three sinks on one GPU
against the plain whole-raster run, the driver through hostmodel.runmicro_big on the bundled example, and — where the
box has two GPUs — the same through two spawned ranks over NCCL."""
import glob
import os

import numpy as np
import pytest

from microclimf_b200 import _abi, api, bands, bigrun, synth

pytestmark = pytest.mark.gpu


def _two_gpus():
    import torch

    return torch.cuda.device_count() >= 2


def _check_sinks(p, gpus, tmp_path):
    from oracle import packing_oracle

    # the bands' all-reduced mean of log(twi)/tfact may differ from the whole-raster kernel reduction in the last bit:
    # give both runs the same number, so that the comparison below can be bit-exact
    s_, n_ = bands.twi_partial_host(p.arrays["twi"], p.tfact)
    p.twi_mean = s_ / n_
    whole = api.run_problem(p)
    T = p.tsteps
    hours = (T // 24) * 24
    r = bigrun.run_local(p.replace(), gpus, sink="arrays")
    for nm in _abi.OUT_NAMES:
        np.testing.assert_array_equal(r["arrays"][nm], whole[nm], err_msg=nm)
    r = bigrun.run_local(p.replace(), gpus, sink="summary")
    assert r["hours"] == hours
    na = np.isnan(p.arrays["hgt"][:p.ncells]).reshape(p.rows, p.cols, order="F")
    for nm in _abi.OUT_NAMES:
        w = whole[nm][:, :, :hours]
        s = r["summary"][nm]
        assert np.isnan(s["mean"][na]).all(), nm
        # NaN hours of a solved cell poison its mean and are skipped by its extremes (the sink's rule)
        with np.errstate(invalid="ignore"):
            np.testing.assert_allclose(s["mean"][~na], w[~na].sum(axis=1) / hours, rtol=1e-12, atol=1e-12, err_msg=nm)
        # the reducing sinks run in the register build of the grid kernel (k_grid), the hourly arrays come from the pair
        # build (k_grid_pair): same physics source, different FMA contraction, so the extremes agree to rounding
        np.testing.assert_allclose(s["min"][~na], np.fmin.reduce(w[~na], axis=1, initial=np.inf), rtol=1e-12, atol=1e-12, err_msg=nm)
        np.testing.assert_allclose(s["max"][~na], np.fmax.reduce(w[~na], axis=1, initial=-np.inf), rtol=1e-12, atol=1e-12, err_msg=nm)
    out_dir = str(tmp_path / f"packed{gpus}")
    r = bigrun.run_local(p.replace(), gpus, sink="packed", pathout=out_dir, window_days=2)
    assert r["hours"] == hours
    files = sorted(glob.glob(os.path.join(out_dir, "area_*.npz")))
    rngs = r["bands"]
    nwin = -(-(T // 24) // 2)
    assert len(files) == len(rngs) * nwin
    for b, (c0, c1) in enumerate(rngs):
        for w in range(nwin):
            z = np.load(os.path.join(out_dir, f"area_{b + 1:02d}_{w + 1:03d}.npz"))
            k0 = int(z["first_hour"])
            assert k0 == w * 48
            for nm in _abi.OUT_NAMES:
                a = z[nm]
                want = packing_oracle.pack(nm, whole[nm][:, c0:c1, k0:k0 + a.shape[2]])
                assert np.array_equal(a, want), (nm, b, w)


def test_band_sinks_one_gpu(tmp_path):
    p = synth.make_problem(29, 23, 24 * 5 + 3, reqhgt=0.05, mode=1)
    _check_sinks(p, 1, tmp_path)


def test_band_sinks_layered_with_gap(tmp_path):
    p = synth.make_problem(13, 9, 24 * 8, reqhgt=0.05, mode=3, nlyr=2)
    p.lyr_st = np.array([0, 120], dtype=np.int32)  # day 3 and 4 are never computed
    p.lyr_ed = np.array([71, 191], dtype=np.int32)
    r = bigrun.run_local(p.replace(), 1, sink="summary")
    assert r["hours"] == 144
    whole = api.run_problem(p)
    covered = np.r_[0:72, 120:192]
    na = np.isnan(p.arrays["hgt"][:p.ncells]).reshape(p.rows, p.cols, order="F")
    np.testing.assert_allclose(r["summary"]["Tz"]["mean"][~na], whole["Tz"][:, :, covered][~na].mean(axis=1), rtol=1e-12)
    out_dir = str(tmp_path / "gap")
    r = bigrun.run_local(p.replace(), 1, sink="packed", pathout=out_dir, window_days=5)
    firsts = sorted(int(np.load(f)["first_hour"]) for f in r["files"])
    assert firsts == [0, 120]  # windows never span the gap


@pytest.mark.skipif(not _two_gpus(), reason="needs two GPUs (gpurun --gpus 2)")
def test_band_sinks_two_ranks_over_nccl(tmp_path):
    p = synth.make_problem(31, 21, 24 * 4, reqhgt=0.05, mode=1)
    _check_sinks(p, 2, tmp_path)


def test_hostmodel_runmicro_big_bands_on_bundled_example(tmp_path):
    """hostmodel.runmicro_big(gpus = 1): the R driver's arguments on the bundled dtmcaerth example (BASELINE configs[0]);
    the band run equals runmicro() on the whole area with the whole-area terrain layers."""
    import test_bundled_example as tb

    from microclimf_b200 import hostmodel

    dtm, vegp, soilc, mp, _ = tb.load_example()
    vegp = tb._fill_reflectance(vegp, dtm)  # .checkbiginputs refuses the bundled NA reflectances of bare ground
    sub = hostmodel.subsetpointmodel(mp, days=[100, 101, 250])
    res = hostmodel.runmicro_big(sub, 0.05, str(tmp_path) + "/", vegp, soilc, dtm, gpus=1)
    mout = hostmodel.runmicro(sub, 0.05, vegp, soilc, dtm)
    assert res["hours"] == 72
    # runmicro_big derives the wind shelter from dtm + vegetation height at 8 m (R/Cppwrappers.R:493-494), runmicro from
    # the dtm at zref: compare what does not depend on wind
    for nm in ("soilm", "Rdirdown", "Rdifdown", "Rswup"):
        m = ~np.isnan(mout[nm][:, :, 0])
        np.testing.assert_allclose(res["summary"][nm]["mean"][m], mout[nm][m].mean(axis=1), rtol=1e-9, atol=1e-9, err_msg=nm)
        np.testing.assert_allclose(res["summary"][nm]["max"][m], mout[nm][m].max(axis=1), rtol=1e-9, atol=1e-9, err_msg=nm)
    na = np.isnan(dtm.matrix())
    assert np.isnan(res["summary"]["Tz"]["mean"][na]).all() and np.isfinite(res["summary"]["Tz"]["mean"][~na]).all()
