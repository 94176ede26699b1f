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
