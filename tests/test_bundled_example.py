"""BASELINE configs[0] / [1]: `runmicro` on the reference's bundled example data (dtmcaerth + climdata + vegp +
soilc, 50 x 50 cells, 12 vegetation layers => the runmicro3Cpp path) through the host layer
(microclimf_b200/hostmodel.py, the mirror of R/Cppwrappers.R:376 and R/internal.R:1345-1460).

The inputs come from tests/golden/bundled/bundled_example.npz (made by tools/make_bundled_fixtures.py from
/root/reference/data/*.rda and the compiled reference's point model).  Parity is checked at the `.Call`
boundary: the very argument list `prepare_model` builds is run through the CUDA path and through the
UNMODIFIED compiled reference (oracle/_ref), tolerance 1e-6 abs / 1e-6 rel (tests/parity.py).
"""
import os

import numpy as np
import pytest

import parity
from microclimf_b200 import api, hostmodel
from microclimf_b200.hostmodel import Micropoint
from microclimf_b200.spatial import Raster
from oracle import pyoracle, terrain_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = os.path.join(HERE, "golden", "bundled", "bundled_example.npz")


def load_example(reqhgt=0.05):
    z = np.load(FIX)
    xmin, xmax, ymin, ymax = z["extent"]
    crs = str(z["crs"])
    mk = lambda v: Raster(v, xmin, xmax, ymin, ymax, crs)  # noqa: E731
    dtm = mk(z["dtm"])
    vegp = {k: mk(z["vegp_" + k]) for k in hostmodel.VEG_NAMES}
    soilc = {k: mk(z["soilc_" + k]) for k in ("soiltype", "groundr")}
    tme = z["obs_time"].astype("datetime64[s]")
    weather = {k: z["mpw_" + k] for k in hostmodel.WEATHER_COLS}
    weather["obs_time"] = tme
    dfo = {k[4:]: z[k] for k in z.files if k.startswith("dfo_")}
    mp = Micropoint(weather=weather, dfo=dfo, Tbz=z["mp_Tbz_m005"] if reqhgt < 0 else None, lat=float(z["mp_lat"]),
                    long=float(z["mp_long"]), zref=float(z["mp_zref"]), subs=np.arange(1, tme.size + 1), tmeorig=tme,
                    matemp=float(z["mp_matemp"]))
    clim = {k: z["clim_" + k] for k in hostmodel.WEATHER_COLS}
    clim["obs_time"] = tme
    return dtm, vegp, soilc, mp, clim


def cpu_slope_aspect(dtm):
    """slr / apr through the numpy restatement of terra::terrain (spatial.terrain), so host-logic tests run without a GPU."""
    from microclimf_b200 import spatial
    return dict(slr=spatial.terrain(dtm, "slope"), apr=spatial.terrain(dtm, "aspect"))


def cpu_terrain(dtm, zref):
    """hor / wsa through the numpy restatement, so host-logic tests run without a GPU."""
    d = dtm.matrix()
    hor = terrain_oracle.horizon24(d, dtm.res[0])
    w16 = terrain_oracle.windcoef16(d, zref, dtm.res[0])
    from microclimf_b200.spatial import aggregate_mean, resample_bilinear
    sm = resample_bilinear(aggregate_mean(dtm.like(w16), 10), dtm).values
    return hor, terrain_oracle.blend16to8(sm)


def to_problem(call):
    a = call.args
    return api._problem(call.mode, a.get("dfsel"), a["obstime"], a["climdata"], a["pointm"], a["vegp"], a["soilc"],
                        a["reqhgt"], a["zref"], a["lat"], a["lon"], None, None, a["Sminp"], a["Smaxp"], a["tfact"],
                        a["complete"], a["mat"])


# --------------------------------------------------------------------------------------------- CPU
def test_fixture_shapes_and_geolocation():
    dtm, vegp, soilc, mp, clim = load_example()
    assert dtm.dim == (50, 50, 1) and vegp["pai"].nlyr == 12 and vegp["clump"].nlyr == 12
    assert clim["temp"].size == 8760
    lat, lon = hostmodel.latlong_from_raster(dtm)
    # Caerthillian Cove, Lizard peninsula (R/data.R: dtmcaerth): 49.97 N, 5.21 W
    assert abs(lat - 49.9674) < 1e-3 and abs(lon + 5.2147) < 1e-3


def test_subsetpointmodel_month_tmax():
    dtm, vegp, soilc, mp, clim = load_example()
    sub = hostmodel.subsetpointmodel(mp, tstep="month", what="tmax")
    assert sub.weather["temp"].size == 288 and sub.subs.size == 288
    # each retained block is one whole calendar day and holds that month's maximum of Tc
    ot = hostmodel._obstime(mp.weather["obs_time"])
    for m in range(12):
        blk = sub.subs[m * 24:(m + 1) * 24] - 1
        assert np.all(np.diff(blk) == 1) and ot["hour"][blk[0]] == 0
        msel = ot["month"] == m + 1
        assert mp.dfo["Tc"][blk].max() == mp.dfo["Tc"][msel].max()
    days = hostmodel.subsetpointmodel(mp, days=[1, 200])
    assert np.array_equal(days.subs, np.r_[np.arange(1, 25), np.arange(199 * 24 + 1, 200 * 24 + 1)])


def test_prepare_model_layers_and_masks():
    dtm, vegp, soilc, mp, clim = load_example()
    sub = hostmodel.subsetpointmodel(mp, tstep="month", what="tmax")
    hor, wsa = cpu_terrain(dtm, mp.zref)
    twi = dtm.like(np.where(np.isnan(dtm.matrix()), np.nan, 5.0))  # flow accumulation needs the built library
    call = hostmodel.prepare_model(sub, vegp, soilc, dtm, reqhgt=0.05, hor=hor, wsa=wsa, twi=twi, **cpu_slope_aspect(dtm))
    a = call.args
    assert call.mode == 3
    # 12 monthly layers, each spanning one 24-hour block (R/internal.R:1388-1399)
    assert a["vegp"]["pai"].shape == (50, 50, 12)
    assert np.array_equal(a["dfsel"]["st"], np.arange(12) * 24) and np.array_equal(a["dfsel"]["ed"], np.arange(12) * 24 + 23)
    assert a["complete"] is False and a["out"] == [True] * 10
    # NA cells of the DTM are NA in hgt (the solver's skip rule) and zero-height cells have zero pai
    na = np.isnan(dtm.matrix())
    assert na.sum() == 128 and np.array_equal(np.isnan(a["vegp"]["hgt"][:, :, 0]), na)
    assert np.all(a["vegp"]["pai"][a["vegp"]["hgt"] == 0] == 0)
    assert np.nanmax(a["vegp"]["paia"] - a["vegp"]["pai"]) <= 1e-12
    # soil parameters come from the lookup table
    st = soilc["soiltype"].matrix()
    from microclimf_b200.tables import SOILPARAMETERS
    k = int(st[10, 10])
    assert a["soilc"]["Smax"][10, 10] == SOILPARAMETERS["Smax"][k - 1] and a["soilc"]["soilb"][10, 10] == SOILPARAMETERS["b"][k - 1]
    # reqhgt = 0 / < 0 mask the outputs as the reference does (R/internal.R:1159-1166)
    call0 = hostmodel.prepare_model(sub, vegp, soilc, dtm, reqhgt=0.0, hor=hor, wsa=wsa, twi=twi, **cpu_slope_aspect(dtm))
    assert call0.args["out"] == [True, False, False, True, False, True, True, True, True, True]
    p = to_problem(call)
    p.validate()


def test_checkinputs_messages():
    dtm, vegp, soilc, mp, clim = load_example()
    bad = dict(clim)
    bad["temp"] = clim["temp"] + 100
    with pytest.raises(ValueError, match="outside range of typical temperature values"):
        hostmodel.checkinputs(bad, vegp, soilc, dtm)
    bad = dict(clim)
    bad["pres"] = clim["pres"] * 10
    with pytest.raises(ValueError, match="pressure"):
        hostmodel.checkinputs(bad, vegp, soilc, dtm)
    bad = {k: v for k, v in clim.items() if k != "lwdown"}
    with pytest.raises(ValueError, match="Cannot find lwdown in weather"):
        hostmodel.checkinputs(bad, vegp, soilc, dtm)
    with pytest.warns(UserWarning):
        ok = hostmodel.checkinputs(clim, vegp, soilc, dtm)
    assert np.all(ok["weather"]["difrad"] <= clim["swdown"] + 1e-9 + (ok["weather"]["difrad"] - clim["difrad"]))


@pytest.mark.skipif(not pyoracle.have_ref(), reason="compiled reference absent")
def test_reference_on_bundled_example_ranges():
    """The compiled reference on the prepared bundled inputs gives physically sensible fields (the bounds of
    the reference's own wrapper test, tests/testthat/test-microclimatemodel_wrapper.R:82-90, loosened to a
    heterogeneous landscape)."""
    dtm, vegp, soilc, mp, clim = load_example()
    sub = hostmodel.subsetpointmodel(mp, days=[172])  # midsummer
    hor, wsa = cpu_terrain(dtm, mp.zref)
    twi = dtm.like(np.where(np.isnan(dtm.matrix()), np.nan, 5.0))
    call = hostmodel.prepare_model(sub, vegp, soilc, dtm, reqhgt=0.05, hor=hor, wsa=wsa, twi=twi, **cpu_slope_aspect(dtm))
    ref = pyoracle.runmicro(to_problem(call), kind="ref")
    ok = ~np.isnan(dtm.matrix())
    tair = sub.weather["temp"]
    Tz = ref["Tz"][ok]
    assert np.isfinite(Tz).all() and np.abs(Tz - tair[None, :]).max() < 25
    rh = ref["relhum"][ok]
    assert rh.min() > 5 and rh.max() <= 100
    assert (ref["windspeed"][ok] <= sub.weather["windspeed"][None, :] + 1e-9).all()


# --------------------------------------------------------------------------------------------- GPU
def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def _parity(call):
    got = call.run()
    want = pyoracle.runmicro(to_problem(call), out_mask=call.args["out"], kind="ref" if pyoracle.have_ref() else "oracle")
    ok, rows = parity.compare(got, want)
    assert ok, "\n" + parity.fmt(rows)
    return got


def _fill_reflectance(vegp, dtm):
    vegp = dict(vegp)
    for k, v in (("leafr", 0.3), ("leaft", 0.15)):
        m = vegp[k].values.copy()
        m[np.isnan(m) & ~np.isnan(dtm.values)] = v
        vegp[k] = vegp[k].like(m)
    return vegp


@pytest.mark.gpu
def test_flowacc_and_terrain_on_bundled_dtm():
    dtm, vegp, soilc, mp, clim = load_example()
    hor, svf = api.horizon(dtm.matrix(), dtm.res[0])
    assert np.array_equal(hor, terrain_oracle.horizon24(dtm.matrix(), dtm.res[0]))
    np.testing.assert_allclose(svf, terrain_oracle.skyview(hor), rtol=0, atol=1e-14)
    twi = hostmodel._topidx(dtm).matrix()
    assert np.array_equal(np.isnan(twi), np.isnan(dtm.matrix())) and np.nanmin(twi) > 0


@pytest.mark.gpu
@pytest.mark.parametrize("reqhgt", [0.05, 0.0, 2.0])
def test_config0_month_tmax_days(reqhgt):
    """configs[0]: the roxygen example of runmicro (R/Cppwrappers.R:353-362): 12 hottest days, 288 h."""
    dtm, vegp, soilc, mp, clim = load_example()
    sub = hostmodel.subsetpointmodel(mp, tstep="month", what="tmax")
    call = hostmodel.prepare_model(sub, vegp, soilc, dtm, reqhgt=reqhgt)
    got = _parity(call)
    assert got["Tz"].shape == (50, 50, 288)
    mout = hostmodel.runmicro(sub, reqhgt, vegp, soilc, dtm)
    assert "tme" in mout and np.array_equal(np.isnan(mout["Tz"][:, :, 0]), np.isnan(dtm.matrix()))
    for k, v in got.items():
        assert np.array_equal(mout[k], v, equal_nan=True)


@pytest.mark.gpu
@pytest.mark.parametrize("reqhgt", [-0.05, 0.05])
def test_config1_full_year(reqhgt):
    """configs[1]: full year hourly (8760 h, complete = TRUE), below ground and within the canopy (above-canopy heights: config0 test)."""
    dtm, vegp, soilc, mp, clim = load_example(reqhgt)
    call = hostmodel.prepare_model(mp, vegp, soilc, dtm, reqhgt=reqhgt)
    assert call.args["complete"] is True and call.args["vegp"]["pai"].shape[2] == 12
    got = _parity(call)
    assert got["Tz"].shape == (50, 50, 8760)


@pytest.mark.gpu
def test_runmicro_big_tiles_match_untiled_statics(tmp_path):
    """runmicro_big (R/Cppwrappers.R:444-543): 4 tiles of 25 x 25; each tile file equals runmicro on the cropped
    inputs with the whole-area terrain layers."""
    dtm, vegp, soilc, mp, clim = load_example()
    sub = hostmodel.subsetpointmodel(mp, days=[100, 250])
    # the bundled leafr / leaft are NA on bare ground: .checkbiginputs refuses that (R/internal.R:1644-1662) ...
    with pytest.raises(ValueError, match="contain NA that are not NA in dtm: leafr leaft"):
        hostmodel.runmicro_big(sub, 0.05, str(tmp_path) + "/", vegp, soilc, dtm, tilesize=25)
    vegp = _fill_reflectance(vegp, dtm)  # ... so a user fills them first
    files = hostmodel.runmicro_big(sub, 0.05, str(tmp_path) + "/", vegp, soilc, dtm, tilesize=25)
    assert [os.path.basename(f) for f in files] == ["area_01_01.npz", "area_01_02.npz", "area_02_01.npz", "area_02_02.npz"]
    t = np.load(files[3])
    assert t["Tz"].shape == (25, 25, 48) and np.array_equal(t["dtm"], dtm.matrix()[25:, 25:], equal_nan=True)
    ok = ~np.isnan(t["dtm"])
    assert np.isfinite(t["Tz"][ok]).all()


@pytest.mark.gpu
def test_runmicro_big_writeasnc_packs_like_writetonc(tmp_path):
    """writeasnc = TRUE: each tile is a netCDF file holding the integers writetonc stores (R/dataprep.R:1064-1069,
    1164-1173), produced by the kernels' packed sink, with writetonc's dimensions, names, units and missing value."""
    from oracle import packing_oracle

    dtm, vegp, soilc, mp, clim = load_example()
    sub = hostmodel.subsetpointmodel(mp, days=[180])
    vegp = _fill_reflectance(vegp, dtm)
    files = hostmodel.runmicro_big(sub, 0.05, str(tmp_path) + "/", vegp, soilc, dtm, tilesize=50, writeasnc=True)
    assert [os.path.basename(f) for f in files] == ["area_01_01.nc"]
    from scipy.io import netcdf_file
    nc = netcdf_file(files[0], "r", mmap=False)
    t = {k: nc.variables[k][:] for k in ("Tz", "Rdirdown", "Rdifdown", "Rswup", "windspeed")}   # [time, north, east]
    assert nc.variables["Tz"].units == b"deg C x 100" and nc.variables["Tz"].missing_value == -9999
    assert np.allclose(nc.variables["east"][:], dtm.xmin + 0.5 + np.arange(50)) and nc.variables["time"].units.startswith(b"hours since 1970")
    assert "soilm" not in nc.variables  # not among writetonc's default variables above ground (R/dataprep.R:1112)
    fp = hostmodel.runmicro(sub, 0.05, vegp, soilc, dtm)  # tile == whole raster here, so the statics coincide...
    assert t["Tz"].shape == (24, 50, 50) and t["Tz"].dtype.kind == "i"
    # ...except the wind shelter, which runmicro_big derives from dtm + vegetation height at 8 m (R/Cppwrappers.R:493-494):
    # compare the wind-independent radiation streams exactly
    for name in ("Rdirdown", "Rdifdown", "Rswup"):
        want = np.transpose(packing_oracle.pack(name, fp[name]), (2, 0, 1))
        assert np.array_equal(t[name], want), name
    na = np.isnan(dtm.matrix())
    assert np.all(t["Tz"][:, na] == packing_oracle.NA) and np.all(t["Tz"][:, ~na] != packing_oracle.NA)


# --------------------------------------------------------------------------------------------- runbioclim
def _pointmodel336(w336, reqhgt, dtm, vegp, soilc, zref, windhgt, soilm):
    """runpointmodel(weather2, ..., yearG = FALSE) on the 14 bioclim days (R/internal.R:1747) through the compiled
    reference's point model (oracle/pointmodel.py, test infrastructure)."""
    from microclimf_b200.tables import SOILPARAMSP
    from oracle import pointmodel
    vm = lambda k: float(np.nanmean(vegp[k].values))  # noqa: E731
    vegp_p = [vm("hgt"), vm("pai"), vm("x"), vm("clump"), vm("leafr"), vm("leaft"), vm("leafd"), 0.97, vm("gsmax"), 100.0]
    sl = hostmodel._soilinit(soilc)
    sn = int(pointmodel.getmode(soilc["soiltype"].values))
    gm = pointmodel.getmode
    groundp_p = [gm(soilc["groundr"].values), 0.0, 180.0, 0.97, gm(sl["rho"]), gm(sl["Vm"]), gm(sl["Vq"]), gm(sl["Mc"]), gm(sl["soilb"]),
                 gm(sl["psi_e"]), gm(sl["Smax"]), gm(sl["Smin"]), SOILPARAMSP["alpha"][sn - 1], SOILPARAMSP["n"][sn - 1],
                 SOILPARAMSP["Ksat"][sn - 1]]
    sprow = {k: SOILPARAMSP[k][sn - 1] for k in ("rmu", "mult", "pwr", "Smax", "Smin", "Ksat", "a")}
    lat, lon = hostmodel.latlong_from_raster(dtm)
    tme = w336["obs_time"]
    mp = pointmodel.runpointmodel(w336, hostmodel._obstime(tme), reqhgt, vegp_p, groundp_p, sprow, lat, lon,
                                  float(np.nanmax(vegp["hgt"].values)), zref=zref, yearG=False)
    w = dict(mp["weather"], obs_time=tme)
    return Micropoint(weather=w, dfo=mp["dfo"], Tbz=mp["Tbz"], lat=lat, long=lon, zref=mp["zref"], subs=np.arange(1, tme.size + 1),
                      tmeorig=tme, matemp=mp["matemp"])


def test_bioclim_day_selection_cpu():
    """.biosel / quarters (R/internal.R:1690-1727, 1796-1801) on the bundled weather: 14 whole days, one per month in
    calendar order, then the year's hottest and coldest day."""
    dtm, vegp, soilc, mp, clim = load_example()
    selh, seld = hostmodel._biosel(clim["obs_time"], clim["temp"])
    assert selh.size == 336 and seld.size == 14
    ot = hostmodel._obstime(clim["obs_time"])
    assert [int(ot["month"][(d - 1) * 24]) for d in seld[:12]] == list(range(1, 13))
    tcd = clim["temp"].reshape(365, 24).mean(axis=1)
    assert seld[12] == np.argmax(tcd) + 1 and seld[13] == np.argmin(tcd) + 1
    # the quarter around January: the 14 selected days whose month is 12, 1 or 2 (the coldest day, in February, included)
    m14 = np.array([int(ot["month"][(d - 1) * 24]) for d in seld])
    want = np.concatenate([np.arange(i * 24 + 1, i * 24 + 25) for i in np.nonzero(np.isin(m14, (12, 1, 2)))[0]])
    assert np.array_equal(hostmodel._getselq(1, ot["month"][selh - 1]), want) and want.size == 96


@pytest.mark.gpu
@pytest.mark.skipif(not pyoracle.have_ref(), reason="compiled reference (its point model) absent")
def test_runbioclim_bundled():
    """BASELINE configs[2] in miniature: runbioclim on the bundled raster (12 vegetation layers => .runbioclim3 /
    runbioclim3Cpp), CUDA operator against the compiled reference's on the same prepared arguments."""
    dtm, vegp, soilc, mp, clim = load_example()

    def ref_op(obstime, climdata, pointm, vg, sc, reqhgt, zref, lat, lon, Sminp, Smaxp, tfact, mat, out, wetq, dryq, hotq, colq, air):
        p = api._problem(3, dict(st=np.arange(14) * 24, ed=np.arange(14) * 24 + 23), obstime, climdata, pointm, vg, sc, reqhgt, zref,
                         lat, lon, None, None, Sminp, Smaxp, tfact, True, mat)
        return pyoracle.runbioclim(p, dict(wetq=wetq, dryq=dryq, hotq=hotq, colq=colq), air=air, out_mask=out, kind="ref")

    got = hostmodel.runbioclim(clim, 0.05, vegp, soilc, dtm, _pointmodel336)
    want = hostmodel.runbioclim(clim, 0.05, vegp, soilc, dtm, _pointmodel336, operator=ref_op)
    assert set(got) == {f"bio{i}" for i in range(1, 20)}
    ok, rows = parity.compare(got, want)
    assert ok, "\n" + parity.fmt(rows)
    land = ~np.isnan(dtm.matrix())
    assert np.isfinite(got["bio1"][land]).all() and 5 < np.nanmean(got["bio1"]) < 20   # annual mean temperature, Cornwall
    assert np.all(got["bio5"][land] >= got["bio6"][land])


# --------------------------------------------------------------------------------------------- gridded climate
def _micropointa(mp, dtm, cr=2, cc=2):
    """A runpointmodela-like list: the bundled point model perturbed per coarse cell (row by row, terra order)."""
    from microclimf_b200.spatial import aggregate_mean
    dtmc = aggregate_mean(dtm.like(np.where(np.isnan(dtm.values), 0.0, dtm.values)), dtm.nrows // cr)
    out = []
    for k in range(cr * cc):
        w = {n: np.array(v) for n, v in mp.weather.items()}
        w["temp"] = w["temp"] + 0.4 * k
        w["relhum"] = np.clip(w["relhum"] - 2.0 * k, 10, 100)
        w["windspeed"] = w["windspeed"] * (1 + 0.05 * k)
        w["winddir"] = np.mod(w["winddir"] + 10.0 * k, 360)
        dfo = {n: np.array(v) for n, v in mp.dfo.items()}
        dfo["G"] = dfo["G"] * (1 + 0.03 * k)
        dfo["Tg"] = dfo["Tg"] + 0.3 * k
        out.append(Micropoint(weather=w, dfo=dfo, Tbz=mp.Tbz, lat=mp.lat, long=mp.long, zref=mp.zref, subs=mp.subs,
                              tmeorig=mp.tmeorig, matemp=mp.matemp + 0.4 * k))
    return out, dtmc


def test_gridded_climate_mapping_cpu():
    """prepare_model_a places fine cell centres on the coarse grid as terra::resample does."""
    from microclimf_b200.spatial import resample_bilinear
    from oracle import prep_oracle
    dtm, vegp, soilc, mp, clim = load_example()
    sub = hostmodel.subsetpointmodel(mp, days=[172])
    mpa, dtmc = _micropointa(sub, dtm)
    hor, wsa = cpu_terrain(dtm, mp.zref)
    twi = dtm.like(np.where(np.isnan(dtm.matrix()), np.nan, 5.0))
    call = hostmodel.prepare_model_a(mpa, vegp, soilc, dtm, dtmc, reqhgt=0.05, altcorrect=2, hor=hor, wsa=wsa, twi=twi,
                                     **cpu_slope_aspect(dtm))
    p = call.prob
    assert (p.mode, p.clim_rows, p.clim_cols, p.nlyr) == (4, 2, 2, 1) or p.mode == 4
    tc = p.arrays["temp"].reshape(p.tsteps, 2, 2)  # [k, cj, ci]
    want = resample_bilinear(dtmc.like(tc[5].T), dtm).matrix()
    got = prep_oracle.resample(p, p.arrays["temp"])[5].reshape(dtm.ncols, dtm.nrows).T
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)
    assert np.isclose(p.mat, mp.matemp + 0.4 * 1.5)


@pytest.mark.gpu
@pytest.mark.parametrize("altcorrect", [0, 2])
def test_gridded_climate_runmicro(altcorrect):
    """runmicro with a list of micropoints and dtmc (R/Cppwrappers.R:389-391 -> .runmodel4Cpp): the kernels' fused
    expansion against the compiled reference on the arrays .runmodel4Cpp would have built."""
    from oracle import prep_oracle
    dtm, vegp, soilc, mp, clim = load_example()
    sub = hostmodel.subsetpointmodel(mp, days=[30, 172])
    mpa, dtmc = _micropointa(sub, dtm)
    call = hostmodel.prepare_model_a(mpa, vegp, soilc, dtm, dtmc, reqhgt=0.05, altcorrect=altcorrect)
    got = call.run()
    want = pyoracle.runmicro(prep_oracle.materialise_coarse(call.prob), out_mask=call.args["out"],
                             kind="ref" if pyoracle.have_ref() else "oracle")
    ok, rows = parity.compare(got, want)
    assert ok, "\n" + parity.fmt(rows)
    mout = hostmodel.runmicro(mpa, 0.05, vegp, soilc, dtm, dtmc=dtmc, altcorrect=altcorrect)
    assert np.array_equal(mout["Tz"], got["Tz"], equal_nan=True)
    with pytest.raises(ValueError, match="Require dtmc"):
        hostmodel.runmicro(mpa, 0.05, vegp, soilc, dtm)


@pytest.mark.gpu
@pytest.mark.parametrize("altcorrect", [0, 2])
def test_runbioclim_gridded_climate(altcorrect):
    """runbioclim with gridded climate (.runbioclim2 / .runbioclim4, R/internal.R:1896-2081, 2200-2388): the coarse series
    of the 14 bioclim days interpolated in the kernels + the 19 reductions, against the compiled reference's
    runbioclim2Cpp / runbioclim4Cpp on the [rows, cols, 336] arrays the R code would have expanded."""
    from oracle import prep_oracle
    dtm, vegp, soilc, mp, clim = load_example()
    vegp1 = {k: (v.like(v.values[:, :, :1]) if v.values.shape[2] > 1 else v) for k, v in vegp.items()}   # static vegetation
    vegp1 = _fill_reflectance(vegp1, dtm)
    tme = np.asarray(clim["obs_time"]).astype("datetime64[s]")
    selh, seld = hostmodel._biosel(tme, clim["temp"])
    w336 = {k: np.asarray(v)[selh - 1] for k, v in clim.items()}
    d_u, v_u, s_u = hostmodel._unpack(dtm, vegp1, soilc)
    mp336 = _pointmodel336(w336, 0.05, d_u, v_u, s_u, 2, 2, None)
    mpa, dtmc = _micropointa(mp336, dtm)

    def ref_op(prob, wetq, dryq, hotq, colq, air, out):
        return pyoracle.runbioclim(prep_oracle.materialise_coarse(prob), dict(wetq=wetq, dryq=dryq, hotq=hotq, colq=colq),
                                   air=air, out_mask=out, kind="ref" if pyoracle.have_ref() else "oracle")

    args = (mpa, clim["precip"], clim["temp"], tme, 0.05, vegp1, soilc, dtm, dtmc)
    got = hostmodel.runbioclim_a(*args, altcorrect=altcorrect)
    want = hostmodel.runbioclim_a(*args, altcorrect=altcorrect, operator=ref_op)
    assert set(got) == {f"bio{i}" for i in range(1, 20)}
    ok, rows = parity.compare(got, want)
    assert ok, "\n" + parity.fmt(rows)
    land = ~np.isnan(dtm.matrix())
    assert np.isfinite(got["bio1"][land]).all() and np.all(got["bio5"][land] >= got["bio6"][land])
    if altcorrect == 0:
        # time-variant vegetation (.runbioclim4, R/internal.R:2200-2388): the bundled 12 layers -> 14 one-day layers
        vegp12 = _fill_reflectance(vegp, dtm)
        got4 = hostmodel.runbioclim_a(mpa, clim["precip"], clim["temp"], tme, 0.05, vegp12, soilc, dtm, dtmc)
        want4 = hostmodel.runbioclim_a(mpa, clim["precip"], clim["temp"], tme, 0.05, vegp12, soilc, dtm, dtmc, operator=ref_op)
        ok, rows = parity.compare(got4, want4)
        assert ok, "\n" + parity.fmt(rows)
        assert not np.allclose(got4["bio1"][land], got["bio1"][land])      # the seasonal layers matter
