"""The two FP64 builds of the grid kernel — k_grid_pair (two threads per cell, invariants in shared memory: the default
for modes 1/3 above ground) and k_grid (one thread per cell, MCF_NO_PAIR=1) — solve the same problems to rounding level.
Both are compared with the CPU checker elsewhere (every parity test runs the default build); this test pins them
against EACH OTHER, in separate processes because the choice is read once per process."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys, numpy as np
sys.path.insert(0, %r)
from microclimf_b200 import api, synth
out = {}
for tag, kw in (("m1", dict(rows=37, cols=29, tsteps=24 * 4 + 5, reqhgt=0.05, mode=1)),
                ("m3", dict(rows=23, cols=19, tsteps=24 * 6, reqhgt=2.0, mode=3, nlyr=2)),
                ("surf", dict(rows=31, cols=17, tsteps=48, reqhgt=0.0, mode=1)),
                ("tall", dict(rows=300, cols=7, tsteps=48, reqhgt=5.0, mode=1))):
    p = synth.make_problem(kw.pop("rows"), kw.pop("cols"), kw.pop("tsteps"), **kw)
    r = api.run_problem(p)
    for k, v in r.items():
        out[tag + "_" + k] = v
    pk = api.run_problem_packed(p) if hasattr(api, "run_problem_packed") else {}
    for k, v in pk.items():
        out[tag + "_packed_" + k] = v
np.savez(sys.argv[1], **out)
"""


def _run(tmp_path, name, env_extra):
    f = str(tmp_path / (name + ".npz"))
    env = dict(os.environ, **env_extra)
    subprocess.run([sys.executable, "-c", CHILD % ROOT, f], check=True, env=env, cwd=ROOT)
    return np.load(f)


def test_pair_and_register_builds_agree(tmp_path):
    a = _run(tmp_path, "pair", {"MCF_NO_PAIR": "0"})
    b = _run(tmp_path, "grid", {"MCF_NO_PAIR": "1"})
    assert set(a.files) == set(b.files) and len(a.files) >= 40
    worst = 0.0
    for k in a.files:
        x, y = a[k], b[k]
        assert x.shape == y.shape, k
        if x.dtype.kind == "i":  # packed int16: a value within rounding of an x.5 boundary may differ by one unit
            assert np.abs(x.astype(np.int32) - y.astype(np.int32)).max() <= 1, k
            assert (x != y).mean() < 1e-4, k
            continue
        assert np.array_equal(np.isnan(x), np.isnan(y)), k
        ok = ~np.isnan(x)
        if ok.any():
            err = np.abs(x[ok] - y[ok]) / (1e-9 + 1e-9 * np.abs(y[ok]))
            worst = max(worst, float(err.max()))
    assert worst <= 1.0, worst  # three orders of magnitude inside the parity bar (1e-6)
