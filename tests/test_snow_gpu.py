"""Snow operators (SURVEY.md NEXT-3): gridmodelsnow1 (hourly snow-pack recurrence) and gridmicrosnow1 (microclimate on
snow-covered cell-hours), CUDA kernels of csrc/mcf_snow.cu against the UNMODIFIED compiled reference
(src/microclimfCpp.cpp:4172-4424, 4894-5057 in oracle/_ref).  Tolerance 1e-6 abs / 1e-6 rel, identical NA masks."""
import numpy as np
import pytest

import parity
from microclimf_b200 import _abi, api, snow, synth
from oracle import pyoracle

needs_ref = pytest.mark.skipif(not pyoracle.have_ref(), reason="compiled reference absent (snow has no C restatement)")


@needs_ref
def test_reference_snow_scenario_cpu():
    """The scenario exercises accumulation, melt-out, snow-free hours and vegetation above / below the pack."""
    s = synth.make_snow_inputs(9, 7, 24 * 6)
    r = pyoracle.gridmodelsnow1(s["obstime"], s["climdata"], s["pointm"], s["vegp"], s["other"])
    ok = ~np.isnan(s["vegp"]["hgt"])
    d = r["sdepc"][ok]
    assert np.isfinite(d).all() and d.max() > 0.3 and (d == 0).any() and (np.diff(d, axis=1) > 0).any()
    assert np.isnan(r["meltg"]).all()  # never initialised by the reference (bioclimfill + accumulate)
    assert (r["Tg"][ok] <= 1e-12).mean() > 0.5


def _snowm(r):
    return dict(Tc=r["Tc"], Tg=r["Tg"], totalSWE=r["sdepc"] * r["sden"], groundsnowdepth=r["sdepg"], snowden=r["sden"])


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("snowenv", ["Alpine", "Tundra", "Taiga"])
def test_gridmodelsnow1_parity(snowenv):
    s = synth.make_snow_inputs(23, 17, 24 * 8, seed=5)
    want = pyoracle.gridmodelsnow1(s["obstime"], s["climdata"], s["pointm"], s["vegp"], s["other"], snowenv)
    got = snow.gridmodelsnow1(s["obstime"], s["climdata"], s["pointm"], s["vegp"], s["other"], snowenv)
    ok, rows = parity.compare(got, want)
    assert ok, "\n" + parity.fmt(rows)


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("reqhgt", [0.05, 0.0, 1.0, 0.3])
def test_gridmicrosnow1_parity(reqhgt):
    """reqhgt 0.05 / 0.0 fall below the pack for deep snow (belowpointsnow) and above it elsewhere; 1.0 and 0.3 cut through
    the canopy-above-snow and above-canopy branches of snowabovepoint."""
    s = synth.make_snow_inputs(19, 13, 24 * 5, seed=9, reqhgt=max(reqhgt, 0.0))
    model = pyoracle.gridmodelsnow1(s["obstime"], s["climdata"], s["pointm"], s["vegp"], s["other"])
    snowm = _snowm(model)
    rng = np.random.default_rng(3)
    shape = model["Tc"].shape
    micro = {n: rng.uniform(-5, 5, shape) for n in _abi.OUT_NAMES}
    out = [True] * 10 if reqhgt > 0 else [True, False, False, True, False, True, True, True, True, True]
    want = pyoracle.gridmicrosnow1(reqhgt, s["obstime"], s["climdata"], snowm, micro, s["vegp"], s["other"], 4.0, out)
    got = snow.gridmicrosnow1(reqhgt, s["obstime"], s["climdata"], snowm, micro, s["vegp"], s["other"], 4.0, out)
    assert set(got) == set(want)
    ok, rows = parity.compare(got, want)
    assert ok, "\n" + parity.fmt(rows)
    # cell-hours without snow keep runmicro's values
    nosnow = ~(snowm["totalSWE"] > 0)
    assert nosnow.any() and np.array_equal(got["Tz"][nosnow], micro["Tz"][nosnow])


@pytest.mark.gpu
@needs_ref
def test_snowmodel1_chunk_driver_matches_reference_operator():
    """hostmodel.snowmodel1 (the 5-day chunk loop of .snowmodel1, R/internal.R:2498-2616) driven by the CUDA operator and
    by the compiled reference's gridmodelsnow1: same terrain updates, same redistribution, same result."""
    from microclimf_b200 import hostmodel
    from microclimf_b200.spatial import Raster
    rows, cols, days = 30, 26, 10
    s = synth.make_snow_inputs(rows, cols, 24 * days, seed=21)
    rng = np.random.default_rng(2)
    ii, jj = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    z = 300 + 40 * np.sin(ii / 5.0) * np.cos(jj / 4.0) + rng.normal(0, 0.5, (rows, cols))
    mk = lambda v: Raster(v, 0, cols * 10.0, 0, rows * 10.0, "")  # noqa: E731
    hgt = np.nan_to_num(s["vegp"]["hgt"], nan=0.5)
    vegp = {k: mk(np.nan_to_num(s["vegp"].get(k, hgt), nan=0.3)) for k in hostmodel.VEG_NAMES if k in s["vegp"] or k == "hgt"}
    for k in hostmodel.VEG_NAMES:
        vegp.setdefault(k, mk(np.full((rows, cols), 0.3)))
    soilc = dict(soiltype=mk(np.full((rows, cols), 4.0)), groundr=mk(np.full((rows, cols), 0.15)))
    T = 24 * days
    tme = (np.datetime64("2023-01-20T00:00:00") + np.arange(T) * np.timedelta64(3600, "s")).astype("datetime64[s]")
    weather = dict(s["climdata"], obs_time=tme)
    pointm = dict(s["pointm"], sdepc=np.full(T, 0.2))
    a = hostmodel.snowmodel1(weather, pointm, mk(z), vegp, soilc, snowenv="Alpine", snowinitd=0.1, zref=30.0)
    b = hostmodel.snowmodel1(weather, pointm, mk(z), vegp, soilc, snowenv="Alpine", snowinitd=0.1, zref=30.0,
                             operator=pyoracle.gridmodelsnow1)
    ok, rows_ = parity.compare({k: a[k] for k in ("Tc", "Tg", "groundsnowdepth", "totalSWE", "snowden")},
                               {k: b[k] for k in ("Tc", "Tg", "groundsnowdepth", "totalSWE", "snowden")})
    assert ok, "\n" + parity.fmt(rows_)
    assert np.isfinite(a["totalSWE"]).all() and a["totalSWE"].max() > 0


@pytest.mark.gpu
@needs_ref
def test_runmicro_snow_true_merges_snow_and_snowfree_days():
    """runmicro(snow = TRUE) on the bundled raster (R/Cppwrappers.R:380-383 -> .runmicrosnow1): a synthetic snow-model
    output with snow on days 2-3 of 4; snow days come from gridmicrosnow1, snow-free days from the ordinary model, and
    the whole thing agrees with the same driver run on the compiled reference's snow operator."""
    from microclimf_b200 import hostmodel
    from test_bundled_example import load_example
    dtm, vegp, soilc, mp, clim = load_example()
    sub = hostmodel.subsetpointmodel(mp, days=[10, 11, 12, 13])
    sub.tmeorig = sub.weather["obs_time"]          # a 4-day model in its own right
    sub.subs = np.arange(1, 97)
    sub.weather["temp"] = sub.weather["temp"] - 8.0
    rng = np.random.default_rng(4)
    shp = (50, 50, 96)
    swe = np.zeros(shp)
    swe[:, :, 24:72] = rng.uniform(5, 60, shp[:2])[:, :, None] * (rng.random(shp[:2]) < 0.8)[:, :, None]
    den = np.full(shp, 250.0)
    smod = dict(Tc=np.minimum(sub.weather["temp"][None, None, :] + rng.normal(0, 0.5, shp), 0.0),
                Tg=np.minimum(sub.weather["temp"][None, None, :] * 0.5 + rng.normal(0, 0.3, shp), 0.0),
                totalSWE=swe, groundsnowdepth=swe / den * 0.8, snowden=den, umu=sub.dfo["umu"])
    a = hostmodel.runmicro(sub, 0.05, vegp, soilc, dtm, snow=True, snowmod=smod)
    b = hostmodel.runmicrosnow1(sub, 0.05, vegp, soilc, dtm, smod, snow_operator=pyoracle.gridmicrosnow1)
    assert a["Tz"].shape == shp
    ok, rows = parity.compare({k: v for k, v in a.items() if k != "tme"}, b)
    assert ok, "\n" + parity.fmt(rows)
    plain = hostmodel.runmicro(sub, 0.05, vegp, soilc, dtm)
    land = ~np.isnan(dtm.matrix())
    assert np.array_equal(a["Tz"][:, :, :24][land], plain["Tz"][:, :, :24][land])        # snow-free day: ordinary model
    snowy = (swe[:, :, 30] > 0) & land
    assert snowy.any() and not np.allclose(a["Tz"][:, :, 30][snowy], plain["Tz"][:, :, 30][snowy])
    assert np.all(a["soilm"][:, :, 30][snowy] == 0.419)                                     # Smax under snow (:5029)


def _array_inputs(s, rows, cols):
    """Expand the data.frame scenario to [rows, cols, hours] arrays with a smooth spatial modulation."""
    rng = np.random.default_rng(8)
    ii, jj = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    u, v = ii / max(rows - 1, 1) - 0.5, jj / max(cols - 1, 1) - 0.5
    ex = lambda a, amp: np.asarray(a)[None, None, :] + amp * (u + 0.5 * v)[:, :, None]  # noqa: E731
    c = s["climdata"]
    clim = dict(temp=ex(c["temp"], 2.0), relhum=np.clip(ex(c["relhum"], 6.0), 20, 100), pres=ex(c["pres"], 0.4),
                swdown=ex(c["swdown"], 0.0) * (1 + 0.1 * v)[:, :, None], lwdown=ex(c["lwdown"], 5.0),
                windspeed=np.maximum(ex(c["windspeed"], 0.5), 0.3), winddir=c["winddir"],
                precip=np.asarray(c["precip"])[None, None, :] * (rng.random((rows, cols)) < 0.85)[:, :, None],
                umu=ex(c["umu"], 0.05))
    clim["difrad"] = np.minimum(ex(c["difrad"], 0.0), clim["swdown"])
    p = s["pointm"]
    pointm = dict(Gp=ex(p["Gp"], 3.0), Tc=ex(p["Tc"], 1.0), RswabsG=np.maximum(ex(p["RswabsG"], 0.0), 0), RlwabsG=ex(p["RlwabsG"], 4.0),
                  umu=clim["umu"], tr=clim["umu"])
    other = dict(s["other"])
    other["lats"] = s["other"]["lat"] + 0.4 * u
    other["lons"] = s["other"]["lon"] + 0.6 * v
    return clim, pointm, other


@pytest.mark.gpu
@needs_ref
def test_array_climate_snow_parity():
    """gridmodelsnow2 / gridmicrosnow2 (src/microclimfCpp.cpp:4426, 5059): per-cell albedo scan, daily extremes and solar
    position."""
    rows, cols = 17, 13
    s = synth.make_snow_inputs(rows, cols, 24 * 5 + 7, seed=31)   # ragged tail: the last 7 hours have no daily extremes
    clim, pointm, other = _array_inputs(s, rows, cols)
    want = pyoracle.gridmodelsnow2(s["obstime"], clim, pointm, s["vegp"], other, "Maritime")
    got = snow.gridmodelsnow2(s["obstime"], clim, pointm, s["vegp"], other, "Maritime")
    ok, rws = parity.compare(got, want)
    assert ok, "\n" + parity.fmt(rws)
    assert np.isnan(want["meltc"]).all()  # not initialised in the array-climate driver
    snowm = _snowm(want)
    rng = np.random.default_rng(3)
    micro = {n: rng.uniform(-5, 5, want["Tc"].shape) for n in _abi.OUT_NAMES}
    for reqhgt in (0.05, 0.6):
        w = pyoracle.gridmicrosnow2(reqhgt, s["obstime"], clim, snowm, micro, s["vegp"], other, 3.0, [True] * 10)
        g = snow.gridmicrosnow2(reqhgt, s["obstime"], clim, snowm, micro, s["vegp"], other, 3.0, [True] * 10)
        ok, rws = parity.compare(g, w)
        assert ok, "\n" + parity.fmt(rws)


@pytest.mark.gpu
@needs_ref
def test_snowmodel2_chunk_driver_matches_reference_operator():
    """hostmodel.snowmodel2 (the 5-day chunk loop of .snowmodel2, R/internal.R:2948-3010, on fine-raster climate arrays)
    driven by the CUDA gridmodelsnow2 and by the compiled reference's: same terrain updates, redistribution radius floor
    and DTM mask."""
    from microclimf_b200 import hostmodel
    from microclimf_b200.spatial import Raster
    rows, cols, days = 24, 20, 10
    T = 24 * days
    s = synth.make_snow_inputs(rows, cols, T, seed=23)
    clim, pointm, _ = _array_inputs(s, rows, cols)
    rng = np.random.default_rng(5)
    ii, jj = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    z = 250 + 30 * np.sin(ii / 4.0) * np.cos(jj / 5.0) + rng.normal(0, 0.5, (rows, cols))
    z[0, :3] = np.nan  # sea cells: masked in the result (.cleansmod)
    crs = ('PROJCRS["OSGB36 / British National Grid",BASEGEOGCRS["OSGB36",DATUM["Ordnance Survey of Great Britain 1936",'
           'ELLIPSOID["Airy 1830",6377563.396,299.3249646]]],CONVERSION["British National Grid",METHOD["Transverse Mercator"],'
           'PARAMETER["Latitude of natural origin",49],PARAMETER["Longitude of natural origin",-2],'
           'PARAMETER["Scale factor at natural origin",0.9996012717],PARAMETER["False easting",400000],'
           'PARAMETER["False northing",-100000]]]')
    mk = lambda v: Raster(v, 170000.0, 170000.0 + cols * 10.0, 12000.0, 12000.0 + rows * 10.0, crs)  # noqa: E731
    hgt = np.nan_to_num(s["vegp"]["hgt"], nan=0.5)
    vegp = {k: mk(np.nan_to_num(s["vegp"].get(k, hgt), nan=0.3)) for k in hostmodel.VEG_NAMES if k in s["vegp"] or k == "hgt"}
    for k in hostmodel.VEG_NAMES:
        vegp.setdefault(k, mk(np.full((rows, cols), 0.3)))
    soilc = dict(soiltype=mk(np.full((rows, cols), 4.0)), groundr=mk(np.full((rows, cols), 0.15)))
    tme = (np.datetime64("2023-01-20T00:00:00") + np.arange(T) * np.timedelta64(3600, "s")).astype("datetime64[s]")
    wuv = np.asarray(s["climdata"]["windspeed"]) * 0.6
    wvv = np.asarray(s["climdata"]["windspeed"]) * 0.5
    kw = dict(sdept=np.full(T, 0.2), wuv=wuv, wvv=wvv, coarse_dims=(3, 3), snowenv="Prairie", snowinitd=0.1, zref=30.0)
    a = hostmodel.snowmodel2(clim, pointm, tme, mk(z), vegp, soilc, **kw)
    b = hostmodel.snowmodel2(clim, pointm, tme, mk(z), vegp, soilc, operator=pyoracle.gridmodelsnow2, **kw)
    keys = ("Tc", "Tg", "groundsnowdepth", "totalSWE", "snowden")
    ok, rows_ = parity.compare({k: a[k] for k in keys}, {k: b[k] for k in keys})
    assert ok, "\n" + parity.fmt(rows_)
    land = ~np.isnan(z)
    assert np.isfinite(a["totalSWE"][land]).all() and np.nanmax(a["totalSWE"]) > 0
    assert np.isnan(a["Tc"][0, 0]).all() and a["umu"].shape == (rows, cols, T)


@pytest.mark.gpu
@needs_ref
def test_runmicro_snow_true_gridded_climate():
    """runmicro(snow = TRUE) with a list of micropoints and dtmc (R/Cppwrappers.R:380-383 -> .runmicrosnow2): snow-free
    days through the coarse-grid kernels, snow days through gridmicrosnow2 on the arrays .prepsnowinputs2 builds; the
    driver agrees with itself run on the compiled reference's snow operator, and its snow-free days with the plain
    gridded-climate model."""
    from microclimf_b200 import hostmodel
    from test_bundled_example import _micropointa, load_example
    dtm, vegp, soilc, mp, clim = load_example()
    sub = hostmodel.subsetpointmodel(mp, days=[10, 11, 12, 13])
    sub.tmeorig = sub.weather["obs_time"]
    sub.subs = np.arange(1, 97)
    sub.weather["temp"] = sub.weather["temp"] - 8.0
    mpa, dtmc = _micropointa(sub, dtm)
    rng = np.random.default_rng(6)
    shp = (50, 50, 96)
    swe = np.zeros(shp)
    swe[:, :, 24:72] = rng.uniform(5, 60, shp[:2])[:, :, None] * (rng.random(shp[:2]) < 0.8)[:, :, None]
    den = np.full(shp, 250.0)
    tair = sub.weather["temp"][None, None, :]
    smod = dict(Tc=np.minimum(tair + rng.normal(0, 0.5, shp), 0.0), Tg=np.minimum(tair * 0.5 + rng.normal(0, 0.3, shp), 0.0),
                totalSWE=swe, groundsnowdepth=swe / den * 0.8, snowden=den,
                umu=np.repeat(np.asarray(sub.dfo["umu"])[None, None, :], 50, 0).repeat(50, 1))
    for altcorrect in (0, 2):
        a = hostmodel.runmicro(mpa, 0.05, vegp, soilc, dtm, dtmc=dtmc, altcorrect=altcorrect, snow=True, snowmod=smod)
        b = hostmodel.runmicrosnow2(mpa, 0.05, vegp, soilc, dtm, dtmc, smod, altcorrect=altcorrect,
                                    snow_operator=pyoracle.gridmicrosnow2)
        assert a["Tz"].shape == shp
        ok, rws = parity.compare({k: v for k, v in a.items() if k != "tme"}, b)
        assert ok, "\n" + parity.fmt(rws)
    plain = hostmodel.runmicro(mpa, 0.05, vegp, soilc, dtm, dtmc=dtmc, altcorrect=2)
    land = ~np.isnan(dtm.matrix())
    assert np.array_equal(a["Tz"][:, :, :24][land], plain["Tz"][:, :, :24][land])
    snowy = (swe[:, :, 30] > 0) & land
    assert snowy.any() and not np.allclose(a["Tz"][:, :, 30][snowy], plain["Tz"][:, :, 30][snowy])
    with pytest.raises(ValueError, match="Require dtmc"):
        hostmodel.runmicro(mpa, 0.05, vegp, soilc, dtm, snow=True, snowmod=smod)


@needs_ref
def test_quick_snow_model_helpers_match_reference_cpu():
    """hostmodel.canintfrac / meltmu against the compiled reference's canintfrac (src/microclimfCpp.cpp:5417) and meltmu
    (:5454), special cases included (no snowfall, NA cells, bare cells, no positive degree-hours)."""
    from microclimf_b200 import hostmodel
    rng = np.random.default_rng(12)
    hgt = rng.uniform(0, 25, (9, 7)); hgt[0, 0] = np.nan; hgt[1, 1] = 0.0
    pai = rng.uniform(0, 6, (9, 7)); pai[2, 2] = 0.0
    for prec, tc, li in ((0.0, -3.0, 0.0), (0.4, -8.0, 0.0), (15.0, 1.5, 0.3)):
        a, b = hostmodel.canintfrac(hgt, pai, 2, prec, tc, li), pyoracle.canintfrac(hgt, pai, 2, prec, tc, li)
        np.testing.assert_allclose(a, b, rtol=1e-13, atol=1e-15, equal_nan=True)
    sv = rng.uniform(0.3, 1, (9, 7)); sv[0, 0] = np.nan
    st, tc = rng.normal(-1, 3, 60), rng.normal(0, 4, 60)
    np.testing.assert_allclose(hostmodel.meltmu(sv, st, tc), pyoracle.meltmu(sv, st, tc), rtol=1e-13, equal_nan=True)
    assert np.array_equal(hostmodel.meltmu(sv, -np.abs(st), tc), pyoracle.meltmu(sv, -np.abs(st), tc))  # all ones
    st3 = rng.normal(-1, 3, (9, 7, 40)); st3[1, 1, :] = -np.abs(st3[1, 1, :])   # a cell without positive degree-hours: 0.5
    tc3 = rng.normal(0, 4, (9, 7, 40))
    np.testing.assert_allclose(hostmodel.meltmu2(sv, st3, tc3), pyoracle.meltmu2(sv, st3, tc3), rtol=1e-13, equal_nan=True)


@pytest.mark.gpu
@needs_ref
def test_snowmodelq1_quick_driver_matches_reference_operator():
    """hostmodel.snowmodelq1 (.snowmodelq1, R/internal.R:2627-2778): the grid snow model on 3 of 12 days, the point
    model's melt terms bridging the gaps; CUDA operator against the compiled reference's in the same driver."""
    from microclimf_b200 import hostmodel
    from microclimf_b200.spatial import Raster
    rows, cols, days = 22, 18, 12
    T = 24 * days
    s = synth.make_snow_inputs(rows, cols, T, seed=29)
    rng = np.random.default_rng(9)
    ii, jj = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    z = 320 + 35 * np.sin(ii / 4.0) * np.cos(jj / 3.0) + rng.normal(0, 0.5, (rows, cols))
    mk = lambda v: Raster(v, 0, cols * 10.0, 0, rows * 10.0, "")  # noqa: E731
    hgt = np.nan_to_num(s["vegp"]["hgt"], nan=0.5)
    vegp = {k: mk(np.nan_to_num(s["vegp"].get(k, hgt), nan=0.3)) for k in hostmodel.VEG_NAMES if k in s["vegp"] or k == "hgt"}
    for k in hostmodel.VEG_NAMES:
        vegp.setdefault(k, mk(np.full((rows, cols), 0.3)))
    soilc = dict(soiltype=mk(np.full((rows, cols), 4.0)), groundr=mk(np.full((rows, cols), 0.15)))
    tme = (np.datetime64("2023-01-10T00:00:00") + np.arange(T) * np.timedelta64(3600, "s")).astype("datetime64[s]")
    weather = dict(s["climdata"], obs_time=tme)
    p = s["pointm"]
    pmod = dict(G=p["Gp"], Tc=p["Tc"], RswabsG=p["RswabsG"], RlwabsG=p["RlwabsG"], umu=p["umu"], tr=p["tr"],
                sdepc=np.full(T + 1, 0.25), sdepg=np.full(T + 1, 0.2), sublmelt=np.full(T, 2e-6),
                tempmelt=np.maximum(np.asarray(s["climdata"]["temp"]), 0) * 4e-5, rainmelt=np.full(T, 1e-6),
                sstemp=np.minimum(np.asarray(s["climdata"]["temp"]) + 0.5, 1.0), sdenc=np.full(T, 210.0), sdeng=np.full(T, 260.0))
    sel_days = np.array([3, 7, 11])
    subs = (np.repeat((sel_days - 1) * 24, 24) + np.tile(np.arange(1, 25), sel_days.size))
    kw = dict(snowenv="Tundra", snowinitd=0.15, zref=30.0)
    a = hostmodel.snowmodelq1(weather, pmod, subs, mk(z), vegp, soilc, **kw)
    b = hostmodel.snowmodelq1(weather, pmod, subs, mk(z), vegp, soilc, operator=pyoracle.gridmodelsnow1, **kw)
    keys = ("Tc", "Tg", "groundsnowdepth", "totalSWE", "snowden")
    ok, rows_ = parity.compare({k: a[k] for k in keys}, {k: b[k] for k in keys})
    assert ok, "\n" + parity.fmt(rows_)
    assert a["Tc"].shape == (rows, cols, 72) and np.nanmax(a["totalSWE"]) > 0
    with pytest.raises(ValueError, match="sbtn"):
        hostmodel.snowmodelq1(weather, pmod, np.arange(1, 25), mk(z), vegp, soilc, **kw)


@pytest.mark.gpu
@needs_ref
def test_snowmodelq2_quick_driver_gridded_climate():
    """hostmodel.snowmodelq2 (.snowmodelq2, R/internal.R:3017-3290): the quick model on fine-raster climate arrays with the
    gap balances from coarse-grid point-model series (meltmu2, .resamplemelt); CUDA gridmodelsnow2 against the compiled
    reference's in the same driver."""
    from microclimf_b200 import hostmodel
    from microclimf_b200.spatial import Raster, aggregate_mean
    rows, cols, days = 20, 16, 9
    T = 24 * days
    s = synth.make_snow_inputs(rows, cols, T, seed=37)
    clim, pointm, _ = _array_inputs(s, rows, cols)
    rng = np.random.default_rng(10)
    ii, jj = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    z = 280 + 30 * np.sin(ii / 4.0) * np.cos(jj / 3.0) + rng.normal(0, 0.5, (rows, cols))
    z[-1, -2:] = np.nan
    crs = ('PROJCRS["OSGB36 / British National Grid",BASEGEOGCRS["OSGB36",DATUM["Ordnance Survey of Great Britain 1936",'
           'ELLIPSOID["Airy 1830",6377563.396,299.3249646]]],CONVERSION["British National Grid",METHOD["Transverse Mercator"],'
           'PARAMETER["Latitude of natural origin",49],PARAMETER["Longitude of natural origin",-2],'
           'PARAMETER["Scale factor at natural origin",0.9996012717],PARAMETER["False easting",400000],'
           'PARAMETER["False northing",-100000]]]')
    mk = lambda v: Raster(v, 170000.0, 170000.0 + cols * 10.0, 12000.0, 12000.0 + rows * 10.0, crs)  # noqa: E731
    dtm = mk(z)
    dtmc = aggregate_mean(mk(np.nan_to_num(z, nan=280.0)), 4)           # 5 x 4 coarse grid
    cr, cc = dtmc.nrows, dtmc.ncols
    hgt = np.nan_to_num(s["vegp"]["hgt"], nan=0.5)
    vegp = {k: mk(np.nan_to_num(s["vegp"].get(k, hgt), nan=0.3)) for k in hostmodel.VEG_NAMES if k in s["vegp"] or k == "hgt"}
    for k in hostmodel.VEG_NAMES:
        vegp.setdefault(k, mk(np.full((rows, cols), 0.3)))
    soilc = dict(soiltype=mk(np.full((rows, cols), 4.0)), groundr=mk(np.full((rows, cols), 0.15)))
    tme = (np.datetime64("2023-01-10T00:00:00") + np.arange(T) * np.timedelta64(3600, "s")).astype("datetime64[s]")
    sel_days = np.array([3, 6, 9])
    subs = np.repeat((sel_days - 1) * 24, 24) + np.tile(np.arange(1, 25), sel_days.size)
    ix = subs - 1
    csub = {k: (v[:, :, ix] if np.ndim(v) == 3 else np.asarray(v)[ix]) for k, v in clim.items()}
    psub = {k: v[:, :, ix] for k, v in pointm.items()}
    tair = np.asarray(s["climdata"]["temp"])
    coarse = lambda a, amp: a[None, None, :] + amp * rng.normal(0, 1, (cr, cc))[:, :, None]  # noqa: E731
    pointm2 = dict(sstemp=np.minimum(clim["temp"] + 0.5, 1.0), tc=clim["temp"],
                   sublmelt=np.abs(coarse(np.full(T, 2e-6), 2e-7)), tempmelt=np.abs(coarse(np.maximum(tair, 0) * 4e-5, 1e-6)),
                   rainmelt=np.abs(coarse(np.full(T, 1e-6), 1e-7)),
                   snow=np.where(coarse(tair, 0.3) > 2, 0.0, np.abs(coarse(np.asarray(s["climdata"]["precip"]), 0.05))),
                   sdenc=coarse(np.full(T, 210.0), 5.0), sdeng=coarse(np.full(T, 260.0), 5.0))
    wuv = np.asarray(s["climdata"]["windspeed"])[ix] * 0.6
    wvv = np.asarray(s["climdata"]["windspeed"])[ix] * 0.5
    args = (csub, psub, pointm2, tme[ix], subs, dtm, dtmc, vegp, soilc, np.full(T, 0.2), wuv, wvv)
    kw = dict(snowenv="Maritime", snowinitd=0.12, zref=30.0)
    a = hostmodel.snowmodelq2(*args, **kw)
    b = hostmodel.snowmodelq2(*args, operator=pyoracle.gridmodelsnow2, **kw)
    keys = ("Tc", "Tg", "groundsnowdepth", "totalSWE", "snowden")
    ok, rows_ = parity.compare({k: a[k] for k in keys}, {k: b[k] for k in keys})
    assert ok, "\n" + parity.fmt(rows_)
    assert a["Tc"].shape == (rows, cols, 72) and np.nanmax(a["totalSWE"]) > 0 and np.isnan(a["Tc"][-1, -1]).all()


@pytest.mark.gpu
def test_runsnowmodel_dispatch():
    """runsnowmodel (R/Cppwrappers.R:718-760), data.frame climate: full point model -> snowmodel1; subset point model ->
    the quick model (method "fast") or the full model followed by subsetsnowmodel ("slow")."""
    from microclimf_b200 import hostmodel
    from microclimf_b200.hostmodel import Micropoint
    from microclimf_b200.spatial import Raster
    rows, cols, days = 14, 12, 10
    T = 24 * days
    s = synth.make_snow_inputs(rows, cols, T, seed=41)
    rng = np.random.default_rng(11)
    ii, jj = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    z = 300 + 25 * np.sin(ii / 3.0) * np.cos(jj / 3.0) + rng.normal(0, 0.5, (rows, cols))
    mk = lambda v: Raster(v, 0, cols * 10.0, 0, rows * 10.0, "")  # noqa: E731
    hgt = np.nan_to_num(s["vegp"]["hgt"], nan=0.5)
    vegp = {k: mk(np.nan_to_num(s["vegp"].get(k, hgt), nan=0.3)) for k in hostmodel.VEG_NAMES if k in s["vegp"] or k == "hgt"}
    for k in hostmodel.VEG_NAMES:
        vegp.setdefault(k, mk(np.full((rows, cols), 0.3)))
    soilc = dict(soiltype=mk(np.full((rows, cols), 4.0)), groundr=mk(np.full((rows, cols), 0.15)))
    tme = (np.datetime64("2023-01-10T00:00:00") + np.arange(T) * np.timedelta64(3600, "s")).astype("datetime64[s]")
    weather = dict(s["climdata"], obs_time=tme)
    p = s["pointm"]
    tair = np.asarray(s["climdata"]["temp"])
    pmod = dict(G=p["Gp"], Tc=p["Tc"], RswabsG=p["RswabsG"], RlwabsG=p["RlwabsG"], umu=p["umu"], tr=p["tr"],
                sdepc=np.full(T + 1, 0.25), sdepg=np.full(T + 1, 0.2), sublmelt=np.full(T, 2e-6),
                tempmelt=np.maximum(tair, 0) * 4e-5, rainmelt=np.full(T, 1e-6), sstemp=np.minimum(tair + 0.5, 1.0),
                sdenc=np.full(T, 210.0), sdeng=np.full(T, 260.0))
    mp_full = Micropoint(weather=weather, dfo={}, Tbz=None, lat=50.0, long=-5.0, zref=30.0, subs=np.arange(1, T + 1),
                         tmeorig=tme, matemp=5.0)
    sel_days = np.array([4, 8])
    subs = np.repeat((sel_days - 1) * 24, 24) + np.tile(np.arange(1, 25), sel_days.size)
    mp_sub = Micropoint(weather=weather, dfo={}, Tbz=None, lat=50.0, long=-5.0, zref=30.0, subs=subs, tmeorig=tme, matemp=5.0)
    kw = dict(snowenv="Alpine", snowinitd=0.1, zref=30.0)
    keys = ("Tc", "Tg", "groundsnowdepth", "totalSWE", "snowden")
    same = lambda a, b: all(np.array_equal(a[k], b[k], equal_nan=True) for k in keys)  # noqa: E731
    cv = hostmodel._cleanvegp(vegp)
    pointm = dict(Gp=pmod["G"], Tc=pmod["Tc"], RswabsG=pmod["RswabsG"], RlwabsG=pmod["RlwabsG"], umu=pmod["umu"], tr=pmod["tr"],
                  sdepc=pmod["sdepc"][:T])
    full = hostmodel.runsnowmodel(weather, mp_full, pmod, vegp, soilc, mk(z), **kw)
    want_full = hostmodel.snowmodel1(weather, pointm, mk(z), cv, soilc, "Alpine", 0.1, 0, 30.0, 0.01)
    assert same(full, want_full) and full["Tc"].shape == (rows, cols, T)
    fast = hostmodel.runsnowmodel(weather, mp_sub, pmod, vegp, soilc, mk(z), method="fast", **kw)
    assert same(fast, hostmodel.snowmodelq1(weather, pmod, subs, mk(z), cv, soilc, "Alpine", 0.1, 0, 30.0, 0.01))
    slow = hostmodel.runsnowmodel(weather, mp_sub, pmod, vegp, soilc, mk(z), method="slow", **kw)
    assert same(slow, hostmodel.subsetsnowmodel(want_full, subs)) and slow["Tc"].shape == (rows, cols, 48)
    with pytest.raises(ValueError, match="prepared arrays"):
        hostmodel.runsnowmodel(weather, [mp_sub], pmod, vegp, soilc, mk(z))
