"""Host logic of the snow drivers that needs no GPU: day sets, seeding and merging of snow / snow-free days
(R/internal.R:3581-3660), `.prepsnowinputs2`'s climate arrays (:3445-3579), `subsetpointmodela` (R/dataprep.R:114)."""
import numpy as np

from microclimf_b200 import hostmodel
from microclimf_b200.spatial import Raster, resample_bilinear
from test_bundled_example import _micropointa, cpu_slope_aspect, cpu_terrain, load_example


def test_snow_day_sets_seed_and_merge():
    z = np.ones((3, 2)); z[0, 0] = np.nan
    dtm = Raster(z, 0, 20, 0, 30, "")
    swe = np.zeros((3, 2, 96))
    swe[:, :, 24:48] = 5.0                 # day 2: snow everywhere
    swe[1, 1, 48:72] = 2.0                 # day 3: snow somewhere
    swe[0, 0, :] = np.nan                  # sea cell: ignored by the masked min / max
    smod, snowdays, nosnowdays = hostmodel._snow_day_sets(dict(totalSWE=swe), dtm)
    assert snowdays.tolist() == [2, 3] and nosnowdays.tolist() == [1, 3, 4]
    assert np.isnan(smod["totalSWE"][0, 0]).all() and smod["totalSWE"][1, 0, 0] == 0
    moutn = {"Tz": np.arange(3 * 2 * 72, dtype=float).reshape(3, 2, 72)}          # the three snow-free days 1, 3, 4
    seeded = hostmodel._seed_snow_micro(moutn, snowdays, nosnowdays)
    assert seeded["Tz"].shape == (3, 2, 48) and np.isnan(seeded["Tz"][:, :, :24]).all()
    assert np.array_equal(seeded["Tz"][:, :, 24:], moutn["Tz"][:, :, 24:48])       # day 3 is in both sets
    mouts = {"Tz": -np.ones((3, 2, 48))}
    merged = hostmodel._merge_snow_days(moutn, mouts, snowdays, nosnowdays)
    assert merged["Tz"].shape == (3, 2, 96)
    assert np.array_equal(merged["Tz"][:, :, :24], moutn["Tz"][:, :, :24])          # day 1 from the ordinary model
    assert np.all(merged["Tz"][:, :, 24:72] == -1)                                  # days 2-3 from the snow operator
    assert np.array_equal(merged["Tz"][:, :, 72:], moutn["Tz"][:, :, 48:72])        # day 4
    assert hostmodel._snow_out_mask([True] * 10, 0.0, {"Tz": 0, "soilm": 0, "Rswup": 0}) == \
        [True, False, False, True, False, False, False, False, True, False]
    assert hostmodel._hours_of_days([2, 4]).tolist() == list(range(25, 49)) + list(range(73, 97))


def test_prepsnowinputs2_climate_arrays():
    dtm, vegp, soilc, mp, clim = load_example()
    sub = hostmodel.subsetpointmodel(mp, days=[10, 11, 12, 13])
    sub.tmeorig = sub.weather["obs_time"]
    sub.subs = np.arange(1, 97)
    mpa, dtmc = _micropointa(sub, dtm)
    mps = hostmodel.subsetpointmodela(mpa, days=np.array([2, 3]))
    assert all(len(m.weather["temp"]) == 48 for m in mps)
    assert all(np.array_equal(m.subs, mps[0].subs) for m in mps)                    # the same days for every coarse cell
    hor, wsa = cpu_terrain(dtm, mp.zref)
    moutn = {"Tz": np.zeros((dtm.nrows, dtm.ncols, 48))}
    r0 = hostmodel.prepsnowinputs2(0.05, dtm, dtmc, vegp, soilc, mps, 0, True, np.array([2, 3]), np.array([1, 4]), moutn,
                                   hor=hor, wsa=wsa, **cpu_slope_aspect(dtm))
    w = r0["weather"]
    tc = np.stack([np.asarray(m.weather["temp"], dtype=float) for m in mps]).reshape(dtmc.nrows, dtmc.ncols, 48)
    np.testing.assert_allclose(w["temp"], resample_bilinear(dtmc.like(tc), dtm).values, rtol=0, atol=1e-12)
    land = ~np.isnan(dtm.matrix())
    assert np.nanmin(w["relhum"]) >= 20 and np.nanmax(w["relhum"]) <= 100 and w["winddir"].shape == (48,)
    assert set(r0["vegp"]) == {"pai", "hgt", "leaft", "clump", "leafd", "paia", "leafden"}
    assert r0["micro"]["Tz"].shape == (dtm.nrows, dtm.ncols, 48) and r0["other"]["lats"].shape == land.shape
    # altitude correction: fixed lapse rate moves temperature by 5 K per km of (coarse - fine) elevation
    r1 = hostmodel.prepsnowinputs2(0.05, dtm, dtmc, vegp, soilc, mps, 1, True, np.array([2, 3]), np.array([1, 4]), moutn,
                                   hor=hor, wsa=wsa, **cpu_slope_aspect(dtm))
    zc = np.nan_to_num(dtmc.matrix())
    elevd = resample_bilinear(dtmc.like(zc), dtm).matrix() - dtm.matrix()
    np.testing.assert_allclose((r1["weather"]["temp"] - w["temp"])[land], np.repeat((elevd * 0.005)[:, :, None], 48, 2)[land],
                               rtol=0, atol=1e-12)
    assert np.all(r1["weather"]["pres"][land] != w["pres"][land]) or np.allclose(elevd[land], 0)
