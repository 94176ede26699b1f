"""Generates the golden fixtures in this directory from the UNMODIFIED reference C++ (oracle/_ref,
built from /root/reference/src by oracle/Makefile).  Run in the build container:

    python tests/golden/make_golden.py

Each .npz stores the generator keywords (kw_*) and every output (out_*) of a small seeded problem;
tests rebuild the same problem from the keywords and compare."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from microclimf_b200 import synth  # noqa: E402

CASES = [
    dict(name="m1_below_canopy", mode=1, rows=12, cols=10, tsteps=72, reqhgt=0.05),
    dict(name="m1_surface", mode=1, rows=12, cols=10, tsteps=72, reqhgt=0.0),
    dict(name="m1_below_ground", mode=1, rows=12, cols=10, tsteps=96, reqhgt=-0.1),
    dict(name="m1_above_canopy", mode=1, rows=12, cols=10, tsteps=72, reqhgt=5.0),
    dict(name="m2_array_climate", mode=2, rows=10, cols=9, tsteps=48, reqhgt=0.05),
    dict(name="m3_layers", mode=3, rows=10, cols=9, tsteps=96, reqhgt=0.5, nlyr=2),
    dict(name="m4_layers_array", mode=4, rows=8, cols=9, tsteps=96, reqhgt=-0.05, nlyr=2),
    dict(name="bio_m1", mode=1, rows=9, cols=8, tsteps=336, reqhgt=0.05, bioclim=True),
]


def build(kw):
    kw = dict(kw)
    kw.pop("name", None)
    if kw.pop("bioclim", False):
        days, _ = synth.bioclim_days()
        return synth.make_problem(kw["rows"], kw["cols"], 336, reqhgt=kw["reqhgt"], mode=kw["mode"], nlyr=14,
                                  day_list=days)
    return synth.make_problem(kw["rows"], kw["cols"], kw["tsteps"], reqhgt=kw["reqhgt"], mode=kw["mode"],
                              nlyr=kw.get("nlyr", 1))


def main():
    from oracle import pyoracle

    for c in CASES:
        p = build(c)
        if c.get("bioclim"):
            _, q = synth.bioclim_days()
            out = pyoracle.runbioclim(p, q, air=True, kind="ref")
        else:
            out = pyoracle.runmicro(p, kind="ref")
        d = {"kw_" + k: np.array(v) for k, v in c.items() if k != "name"}
        for k, v in out.items():
            d["out_" + k] = v
        np.savez_compressed(os.path.join(HERE, c["name"] + ".npz"), **d)
        print(c["name"], {k: v.shape for k, v in out.items() if k in ("Tz", "bio1")})


if __name__ == "__main__":
    main()
