"""dev aid: where does the FP32 build's below-ground Tz leave the FP64 reference?  (tests/test_f32_gpu.py cases)"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from microclimf_b200 import api, synth
from oracle import pyoracle
import test_f32_gpu as t

for mode, z, complete in [(1, -0.05, True), (3, -0.1, False), (1, -0.6, True), (2, -0.2, True)]:
    p = synth.make_problem(21, 17, 24 * 6, reqhgt=z, mode=mode, nlyr=2, complete=complete)
    mask = [True, False, False, True] + [False] * 6
    got = t.run_f32(p, mask)["Tz"]
    want = pyoracle.runmicro(p, out_mask=mask, kind=t.KIND)["Tz"]
    g64 = api.run_problem(p, out=mask)["Tz"]
    d = np.abs(got - want)
    d[np.isnan(d)] = 0
    percell = d.max(axis=2)
    bad = np.argwhere(percell > 0.05)
    print(f"mode {mode} z {z} complete {complete}: max err {d.max():.4f}, cells beyond 0.05: {len(bad)} of {percell.size}; "
          f"median cell err {np.median(percell):.2e}; fp64 build max err {np.nanmax(np.abs(g64 - want)):.2e}")
    for (i, j) in bad[:3]:
        print("   cell", i, j, "err by hour (first 30):", np.array2string(d[i, j, :30], precision=3, max_line_width=200))
        print("   got ", np.array2string(got[i, j, :12], precision=4, max_line_width=200))
        print("   want", np.array2string(want[i, j, :12], precision=4, max_line_width=200))
