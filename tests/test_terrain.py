"""Terrain preparation (SURVEY.md NEXT-2): CUDA stencils vs the numpy restatement of the R code
(oracle/terrain_oracle.py, PARITY UNPINNED — no R interpreter here; see its header)."""
import numpy as np
import pytest

from oracle import terrain_oracle as T


def _dtm(rows, cols, seed=5):
    rng = np.random.default_rng(seed)
    ii, jj = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    d = 120 + 60 * np.sin(ii / 9.0) * np.cos(jj / 13.0) + 25 * np.sin((ii + 2 * jj) / 5.0) + rng.normal(0, 1.5, (rows, cols))
    d[rng.random((rows, cols)) < 0.01] = np.nan
    return d


def test_oracle_properties():
    flat = np.full((12, 9), 37.0)
    h = T.horizon24(flat, 10.0)
    # a flat plateau only sees the zero padding outside the raster: horizon 0 everywhere, sky view 1
    assert h.max() == 0.0 and np.allclose(T.skyview(h), 1.0)
    # a wall to the north (row 0 = north) raises the horizon of the azimuth-0 layer only for cells south of it
    wall = np.zeros((40, 8))
    wall[5, :] = 100.0
    h0 = T.horizon(wall, 0.0, 1.0)
    assert h0[6, 3] == 100.0 and h0[9, 3] == 100.0 / 4 and h0[4, 3] == 0.0
    h180 = T.horizon(wall, 180.0, 1.0)
    assert h180[4, 3] == 100.0 and h180[6, 3] == 0.0
    w = T.windcoef(wall, 0.0, 2.0, 1.0)
    assert w[6, 3] < 0.1 and w[4, 3] == 1.0
    b = T.blend16to8(T.windcoef16(wall, 2.0, 1.0))
    assert b.shape == (40, 8, 8) and (b <= 1.0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("rows,cols,reso", [(64, 48, 10.0), (150, 211, 1.0), (7, 5, 30.0)])
def test_horizon_skyview_parity(rows, cols, reso):
    from microclimf_b200 import api

    d = _dtm(rows, cols)
    hor, svf = api.horizon(d, reso)
    want = T.horizon24(d, reso)
    np.testing.assert_array_equal(hor, want)  # pure IEEE arithmetic on identical gathers: bit-exact
    np.testing.assert_allclose(svf, T.skyview(want), rtol=0, atol=1e-9)  # cos(2 tan(mean)) is ill-conditioned for steep relief: last-bit differences of the mean are amplified
    # arbitrary azimuths, including ones whose shifts fall on x.5 boundaries
    az = np.array([7.5, 30.0, 60.0, 123.4, 270.0, 359.9])
    hor2, _ = api.horizon(d, reso, azimuths=az, want_svf=False)
    np.testing.assert_array_equal(hor2, np.stack([T.horizon(d, a, reso) for a in az], axis=2))


@pytest.mark.gpu
def test_windcoef_parity():
    from microclimf_b200 import api

    d = _dtm(90, 70, seed=9)
    idx, b8 = api.windcoef(d, 5.0, 2.0, blend8=True)
    want = T.windcoef16(d, 2.0, 5.0)
    np.testing.assert_allclose(idx, want, rtol=0, atol=1e-14)
    np.testing.assert_allclose(b8, T.blend16to8(want), rtol=0, atol=1e-14)


@pytest.mark.gpu
def test_terrain_feeds_the_grid_model():
    """hor / svfa computed on the GPU from a DTM drop into the grid model's inputs (same layout)."""
    import parity
    from microclimf_b200 import api, synth
    from oracle import pyoracle

    p = synth.make_problem(24, 20, 48, reqhgt=0.05, mode=1)
    d = _dtm(24, 20, seed=2)
    hor, svf = api.horizon(d, 10.0)
    p.arrays["hor"] = np.ascontiguousarray(hor.ravel(order="F"))
    p.arrays["svfa"] = np.ascontiguousarray(svf.ravel(order="F"))
    got = api.run_problem(p)
    want = pyoracle.runmicro(p, kind="ref" if pyoracle.have_ref() else "oracle")
    ok, rows = parity.compare(got, want)
    assert ok, "\n" + parity.fmt(rows)
