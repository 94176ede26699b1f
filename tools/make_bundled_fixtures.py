"""Builds tests/golden/bundled/bundled_example.npz from the reference's bundled example data
(/root/reference/data/*.rda: climdata, dtmcaerth, vegp, soilc — BASELINE.json configs[0]/[1]).

Run in the build container only (needs /root/reference and oracle/_ref):
    python tools/make_bundled_fixtures.py

Contents: the decoded rasters and weather table, and a `micropoint` made the way the reference's
example makes it — `runpointmodel(climdata, reqhgt, dtmcaerth, vegp, soilc)` (R/Cppwrappers.R:55) — by
oracle/pointmodel.py around the compiled reference's point model, for reqhgt = 0.05 (Tbz for -0.05 too).
The GPU box has no /root/reference, so the tests read this file instead.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from microclimf_b200 import hostmodel  # noqa: E402
from microclimf_b200.rdata import dataframe_columns, read_rda  # noqa: E402
from microclimf_b200.spatial import Raster, latlong_from_raster  # noqa: E402
from microclimf_b200.tables import SOILPARAMSP  # noqa: E402
from oracle import pointmodel  # noqa: E402

DATA = "/root/reference/data"


def main():
    dtm = Raster.from_packed(read_rda(f"{DATA}/dtmcaerth.rda")["dtmcaerth"])
    vegp_r = read_rda(f"{DATA}/vegp.rda")["vegp"]
    soilc_r = read_rda(f"{DATA}/soilc.rda")["soilc"]
    vegp = {k: Raster.from_packed(vegp_r[k]) for k in hostmodel.VEG_NAMES}
    soilc = {k: Raster.from_packed(soilc_r[k]) for k in ("soiltype", "groundr")}
    clim = dataframe_columns(read_rda(f"{DATA}/climdata.rda")["climdata"])
    ot = clim.pop("obs_time")
    tme = (np.array([f"{int(y) + 1900:04d}-{int(m) + 1:02d}-{int(d):02d}" for y, m, d in zip(ot["year"], ot["mon"], ot["mday"])],
                    dtype="datetime64[D]").astype("datetime64[s]")
           + (ot["hour"] * 3600 + ot["min"] * 60 + ot["sec"]).astype("timedelta64[s]"))
    weather = {k: np.asarray(v, dtype=np.float64) for k, v in clim.items()}
    weather["obs_time"] = tme

    # runpointmodel: checkinputs, point-model parameter vectors, then the C++ point model
    rc = hostmodel.checkinputs(weather, vegp, soilc, dtm)
    w2, vegp2, soilc2 = rc["weather"], rc["vegp"], rc["soilc"]
    vm = lambda k: float(np.nanmean(vegp2[k].values))  # noqa: E731
    vegp_p = [vm("hgt"), vm("pai"), vm("x"), vm("clump"), vm("leafr"), vm("leaft"), vm("leafd"), 0.97, vm("gsmax"), 100.0]
    sl = hostmodel._soilinit(soilc2)
    sn = int(pointmodel.getmode(soilc2["soiltype"].values))
    gm = pointmodel.getmode
    groundp_p = [gm(soilc2["groundr"].values), 0.0, 180.0, 0.97, gm(sl["rho"]), gm(sl["Vm"]), gm(sl["Vq"]), gm(sl["Mc"]),
                 gm(sl["soilb"]), gm(sl["psi_e"]), gm(sl["Smax"]), gm(sl["Smin"]), SOILPARAMSP["alpha"][sn - 1],
                 SOILPARAMSP["n"][sn - 1], SOILPARAMSP["Ksat"][sn - 1]]
    sprow = {k: SOILPARAMSP[k][sn - 1] for k in ("rmu", "mult", "pwr", "Smax", "Smin", "Ksat", "a")}
    lat, lon = latlong_from_raster(vegp2["x"])
    obstime = hostmodel._obstime(tme)
    mxhgt = float(np.nanmax(vegp2["hgt"].values))
    mp = pointmodel.runpointmodel(w2, obstime, 0.05, vegp_p, groundp_p, sprow, lat, lon, mxhgt)
    Tbz = pointmodel.soilbelowT(mp["dfo"], -0.05)

    out = dict(extent=np.array([dtm.xmin, dtm.xmax, dtm.ymin, dtm.ymax]), crs=np.array(dtm.crs), dtm=dtm.values,
               obs_time=tme.astype("int64"), mp_lat=lat, mp_long=lon, mp_zref=mp["zref"], mp_matemp=mp["matemp"],
               mp_Tbz_m005=Tbz)
    for k, v in vegp.items():
        out["vegp_" + k] = v.values
    for k, v in soilc.items():
        out["soilc_" + k] = v.values
    for k in pointmodel.WEATHER_COLS:
        out["clim_" + k] = weather[k]
        out["mpw_" + k] = mp["weather"][k]
    for k, v in mp["dfo"].items():
        out["dfo_" + k] = v
    path = os.path.join(ROOT, "tests", "golden", "bundled", "bundled_example.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; zref", mp["zref"], "lat/long", lat, lon,
          "Tc range", mp["dfo"]["Tc"].min(), mp["dfo"]["Tc"].max(), "umu", mp["dfo"]["umu"].min(), mp["dfo"]["umu"].max())


if __name__ == "__main__":
    main()
