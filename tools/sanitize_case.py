"""Small cases of every kernel for compute-sanitizer (memcheck): all modes / heights, bioclim, terrain."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from microclimf_b200 import api, synth
for mode in (1, 2, 3, 4):
    for rq in (0.05, 0.0, -0.1):
        p = synth.make_problem(13, 11, 48, reqhgt=rq, mode=mode, nlyr=2, complete=(mode % 2 == 1))
        o = api.run_problem(p)
        print(mode, rq, float(np.nanmean(o["Tz"])))
days, q = synth.bioclim_days()
p = synth.make_problem(7, 9, 336, reqhgt=0.05, mode=3, nlyr=14, day_list=days)
b = api.run_bioclim_problem(p, q["wetq"], q["dryq"], q["hotq"], q["colq"])
print("bio1", float(np.nanmean(b["bio1"])))
d = np.random.default_rng(0).uniform(0, 100, (33, 21))
h, s = api.horizon(d, 5.0)
w = api.windcoef(d, 5.0, 2.0)
print("terrain", float(h.mean()), float(s.mean()), float(w.mean()))
# packed sink, coarse-grid climate, snow (round-1 additions)
p = synth.make_problem(13, 11, 48, reqhgt=0.05, mode=1)
print("packed", int(api.run_problem_packed(p)["Tz"].max()))
p = synth.make_problem(13, 11, 48, reqhgt=-0.1, mode=1)
print("packed below", int(api.run_problem_packed(p, out=[1, 0, 0, 1, 0, 0, 0, 0, 0, 0])["Tz"].max()))
for mode, rq, ac in ((2, 0.05, 2), (4, 0.0, 1), (2, -0.1, 0)):
    p = synth.make_coarse_problem(13, 11, 48, reqhgt=rq, mode=mode, crows=3, ccols=2, altcorrect=ac, nlyr=2)
    print("coarse", mode, rq, float(np.nanmean(api.run_problem(p)["Tz"])))
from microclimf_b200 import snow
s = synth.make_snow_inputs(11, 9, 24 * 3 + 5)
r = snow.gridmodelsnow1(s["obstime"], s["climdata"], s["pointm"], s["vegp"], s["other"])
sm = dict(Tc=r["Tc"], Tg=r["Tg"], totalSWE=np.nan_to_num(r["sdepc"] * r["sden"]), groundsnowdepth=r["sdepg"], snowden=r["sden"])
mic = {n: np.zeros(r["Tc"].shape) for n in ("Tz", "tleaf", "relhum", "soilm", "windspeed", "Rdirdown", "Rdifdown", "Rlwdown", "Rswup", "Rlwup")}
m = snow.gridmicrosnow1(0.05, s["obstime"], s["climdata"], sm, mic, s["vegp"], s["other"], 3.0, [True] * 10)
print("snow", float(np.nanmax(r["sdepc"])), float(np.nanmean(m["Tz"])))
ex = lambda a: np.broadcast_to(np.asarray(a)[None, None, :], (11, 9, len(a))).copy()
c = s["climdata"]
clim = {k: (ex(v) if k != "winddir" else v) for k, v in c.items()}
pm = {k: ex(v) for k, v in s["pointm"].items()}
oth = dict(s["other"], lats=np.full((11, 9), 61.0), lons=np.full((11, 9), 10.0))
r2 = snow.gridmodelsnow2(s["obstime"], clim, pm, s["vegp"], oth)
m2 = snow.gridmicrosnow2(0.05, s["obstime"], clim, sm, mic, s["vegp"], oth, 3.0, [True] * 10)
print("snow2", float(np.nanmax(np.abs(r2["sdepc"] - r["sdepc"]))), float(np.nanmean(m2["Tz"])))
