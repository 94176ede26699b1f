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


# ---------------------------------------------------------------------------------------------------------------------
# slope / aspect, aggregate + resample, wind shelter, TWI on the device.  terra's routines are third-party arithmetic with
# no vectors to pin against (PARITY UNPINNED): what CAN be pinned is pinned on analytic surfaces, where
# Horn's stencil, the horizon search and bilinear interpolation have closed-form answers; on rough surfaces the kernels
# are compared with the numpy restatements of the same published definitions (microclimf_b200/spatial.py).
# ---------------------------------------------------------------------------------------------------------------------
def _plane(rows, cols, gx, gy, dx, dy):
    """z = gx * x + gy * y with x east, y north; row 0 is the northern edge."""
    ii, jj = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    return gx * (jj * dx) + gy * ((rows - 1 - ii) * dy)


@pytest.mark.gpu
@pytest.mark.parametrize("gx,gy", [(0.3, 0.0), (0.0, -0.2), (0.15, 0.25), (-0.4, 0.1), (0.0, 0.0)])
def test_slope_aspect_known_answers_on_planes(gx, gy):
    """Horn's stencil is exact on a plane: slope = atan |grad z|, aspect = bearing of the DOWNSLOPE direction clockwise
    from north (90 on flat ground), NA on the one-cell edge."""
    from microclimf_b200 import api

    dx, dy = 10.0, 12.0
    sl, asp = api.slope_aspect(_plane(33, 41, gx, gy, dx, dy), dx, dy)
    inner = (slice(1, -1), slice(1, -1))
    assert np.isnan(sl[0]).all() and np.isnan(sl[-1]).all() and np.isnan(asp[:, 0]).all() and np.isnan(asp[:, -1]).all()
    np.testing.assert_allclose(sl[inner], np.degrees(np.arctan(np.hypot(gx, gy))), rtol=0, atol=1e-10)
    want = 90.0 if gx == 0 and gy == 0 else np.degrees(np.arctan2(-gx, -gy)) % 360.0
    np.testing.assert_allclose(asp[inner], want, rtol=0, atol=1e-9)


@pytest.mark.gpu
def test_slope_aspect_on_a_cone_and_against_the_restatement():
    from microclimf_b200 import api, spatial

    n, k = 101, 0.35
    ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    r = np.hypot(ii - 50, jj - 50) * 5.0
    sl, asp = api.slope_aspect(1000.0 - k * r, 5.0, 5.0)
    far = r > 150  # away from the apex Horn's 3 x 3 average of a cone's gradient is within 1e-3 of the true one
    far[[0, -1], :] = False
    far[:, [0, -1]] = False
    np.testing.assert_allclose(sl[far], np.degrees(np.arctan(k)), atol=0.05)
    bearing_out = np.degrees(np.arctan2(jj - 50, -(ii - 50))) % 360.0  # direction away from the apex = downslope
    dif = np.abs((asp[far] - bearing_out[far] + 180) % 360 - 180)
    assert dif.max() < 0.5
    d = _dtm(77, 58, seed=12)
    ras = spatial.Raster(d, 0, 58 * 7.0, 0, 77 * 9.0, "")
    sl, asp = api.slope_aspect(d, 7.0, 9.0)
    np.testing.assert_allclose(sl, spatial.terrain(ras, "slope").matrix(), rtol=0, atol=1e-10, equal_nan=True)
    np.testing.assert_allclose(asp, spatial.terrain(ras, "aspect").matrix(), rtol=0, atol=1e-9, equal_nan=True)


@pytest.mark.gpu
def test_horizon_known_answer_on_planes():
    """On a plane the horizon tangent towards a cardinal azimuth is the (positive part of the) directional derivative,
    whichever of the 10 search steps finds it — as long as the search stays inside the raster (beyond it R pads zeros)."""
    from microclimf_b200 import api

    reso = 10.0
    z = _plane(260, 250, 0.2, -0.1, reso, reso) + 500.0
    hor, _ = api.horizon(z, reso, azimuths=np.array([0.0, 90.0, 180.0, 270.0]), want_svf=False)
    core = (slice(101, -101), slice(101, -101))  # every search step (up to 100 cells) stays inside
    for k, want in enumerate((0.0, 0.2, 0.1, 0.0)):  # north: -0.1 -> 0; east: +0.2; south: +0.1; west: -0.2 -> 0
        np.testing.assert_allclose(hor[core][..., k], want, rtol=0, atol=1e-12)


@pytest.mark.gpu
def test_windshelter_and_topidx_against_restatements():
    from microclimf_b200 import api, spatial

    d = _dtm(123, 97, seed=21)
    d[np.isnan(d)] = 100.0
    reso = 4.0
    ras = spatial.Raster(d, 0, 97 * reso, 0, 123 * reso, "")
    # .windsheltera = 16 x .windcoef -> aggregate(10, mean) -> resample(bilinear) -> 16 to 8 blend
    idx = api.windcoef(d, reso, 2.0)
    sm = spatial.resample_bilinear(spatial.aggregate_mean(ras.like(idx), 10), ras).values
    np.testing.assert_allclose(api.windshelter(d, reso, 2.0, 10), T.blend16to8(sm), rtol=0, atol=1e-12)
    np.testing.assert_allclose(api.windshelter(d, reso, 2.0, 1), T.blend16to8(idx), rtol=0, atol=1e-14)
    # a constant field survives block mean + bilinear exactly; a plane survives it away from the ragged edge blocks
    # .topidx
    B = spatial.terrain(ras, "slope", unit="radians").matrix().copy()
    minslope = np.arctan(0.02 / reso)
    with np.errstate(invalid="ignore"):
        B[B < minslope] = minslope
    B[np.isnan(B)] = np.nanmedian(B)
    a = np.maximum((api.flowacc(d) + 1) * reso * reso, 1.0)
    np.testing.assert_allclose(api.topidx(d, reso, reso), a / np.tan(B), rtol=1e-12)
    sea = d.copy()
    sea[:20, :30] = np.nan
    assert np.array_equal(np.isnan(api.topidx(sea, reso, reso)), np.isnan(sea))
