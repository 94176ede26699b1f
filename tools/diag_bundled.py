"""dev aid: where does the CUDA path leave the reference on the bundled example?"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from test_bundled_example import load_example, to_problem
from microclimf_b200 import hostmodel
from oracle import pyoracle

reqhgt = float(sys.argv[1]) if len(sys.argv) > 1 else 0.05
dtm, vegp, soilc, mp, clim = load_example(reqhgt)
sub = hostmodel.subsetpointmodel(mp, tstep="month", what="tmax")
call = hostmodel.prepare_model(sub, vegp, soilc, dtm, reqhgt=reqhgt)
got = call.run()
want = pyoracle.runmicro(to_problem(call), out_mask=call.args["out"], kind="ref")
a = call.args
for name in got:
    d = np.abs(got[name] - want[name])
    d[np.isnan(d)] = 0
    bad = d > 1e-6 + 1e-6 * np.abs(want[name])
    print(name, "bad", int(bad.sum()), "of", bad.size, "max", d.max())
    if bad.any():
        cells = np.argwhere(bad.any(axis=2))
        print("  bad cells", len(cells), "hours with bad", np.unique(np.argwhere(bad)[:, 2])[:40])
        for (i, j) in cells[:6]:
            k = int(np.argmax(d[i, j]))
            lyr = k // 24
            v = {n: float(a["vegp"][n][i, j, lyr]) for n in a["vegp"]}
            print("  cell", i, j, "hour", k, "got", got[name][i, j, k], "want", want[name][i, j, k], v,
                  "slope", a["soilc"]["slope"][i, j], "gref", a["soilc"]["gref"][i, j])
hg = a["vegp"]["hgt"]
