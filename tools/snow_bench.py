"""Device time of the snow kernels (dev aid): python tools/snow_bench.py [n = 512] [hours = 240]
Runs the host-buffer entry points (the only ones) and reports, from CUDA events around each call with the transfers
excluded by a second timing of the kernels alone through ncu's launch list (tools/r02_profile.sh)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from microclimf_b200 import snow, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
T = int(sys.argv[2]) if len(sys.argv) > 2 else 240
s = synth.make_snow_inputs(n, n, T)
for rep in range(2):
    t0 = time.perf_counter()
    r = snow.gridmodelsnow1(s["obstime"], s["climdata"], s["pointm"], s["vegp"], s["other"], "Alpine")
    t1 = time.perf_counter()
with np.errstate(invalid="ignore"):
    snowm = dict(Tc=r["Tc"], Tg=r["Tg"], totalSWE=np.nan_to_num(r["sdepc"] * r["sden"]), groundsnowdepth=r["sdepg"], snowden=r["sden"])
micro = {k: np.zeros(r["Tc"].shape, order="F") for k in ("Tz", "tleaf", "relhum", "soilm", "windspeed", "Rdirdown", "Rdifdown", "Rlwdown", "Rswup", "Rlwup")}
for rep in range(2):
    t2 = time.perf_counter()
    snow.gridmicrosnow1(0.05, s["obstime"], s["climdata"], snowm, micro, s["vegp"], s["other"], 3.0, [True] * 10)
    t3 = time.perf_counter()
for rep in range(2):
    t4 = time.perf_counter()
    snow.gridmicrosnow1(0.05, s["obstime"], s["climdata"], snowm, micro, s["vegp"], s["other"], 3.0, [True] * 10, copy=False)
    t5 = time.perf_counter()
ch = n * n * T
print(json.dumps({"cells": n * n, "hours": T, "gridmodelsnow1_host_cell_hours_per_s": ch / (t1 - t0), "gridmicrosnow1_host_cell_hours_per_s": ch / (t3 - t2),
                  "gridmicrosnow1_host_inplace_cell_hours_per_s": ch / (t5 - t4),
                  "snow_covered_fraction": float((snowm["totalSWE"] > 0).mean())}))
