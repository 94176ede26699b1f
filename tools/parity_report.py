"""Prints a parity table (CUDA path vs the compiled reference) for a sweep of modes and heights.
Run on the GPU box:  python tools/parity_report.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np  # noqa: E402

import parity  # noqa: E402
from microclimf_b200 import api, synth  # noqa: E402
from oracle import pyoracle  # noqa: E402

kind = "ref" if pyoracle.have_ref() else "oracle"
allok = True
worst = 0.0
for mode in (1, 2, 3, 4):
    for rq in (0.05, 0.0, -0.1, 5.0, 1.0):
        for complete in ((True, False) if rq < 0 else (True,)):
            p = synth.make_problem(40, 30, 24 * 6, reqhgt=rq, mode=mode, nlyr=3, complete=complete)
            t0 = time.time()
            want = pyoracle.runmicro(p, kind=kind)
            t1 = time.time()
            got = api.run_problem(p)
            t2 = time.time()
            ok, rows = parity.compare(got, want)
            allok &= ok
            worst = max(worst, max(r[2] for r in rows))
            print(f"mode {mode} reqhgt {rq} complete {complete}: {'OK' if ok else 'FAIL'} (cpu {t1-t0:.2f}s gpu {t2-t1:.2f}s)")
            if not ok:
                print(parity.fmt(rows))
print(f"worst error / tolerance over the sweep: {worst:.3e}")
print("ALL OK" if allok else "SOME FAILED")
