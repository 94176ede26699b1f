"""Randomised corner-input parity (tools/fuzz_parity.py): per-cell parameters and forcing drawn from the full ranges
`checkinputs` admits (R/dataprep.R:206-397) with special values mixed in — 0 / 1 / 1e-300 clumping, zero ground
reflectance, NA leaf reflectance, centimetre vegetation, vegetation up to zref, x from 0 to 10, -50..65 degC,
calm to 100 m/s — every mode, seven heights.  Tolerance 1e-6 abs / 1e-6 rel, identical NaN masks."""
import os
import sys

import numpy as np
import pytest

import parity

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import fuzz_parity as fz  # noqa: E402
from microclimf_b200 import api, synth  # noqa: E402
from oracle import pyoracle  # noqa: E402


def _problem(seed):
    rng = np.random.default_rng(seed)
    mode = int(rng.choice([1, 2, 3, 4]))
    reqhgt = float(rng.choice([0.05, 0.0, -0.1, 1.0, 5.0, 0.3, 20.0]))
    zref = float(rng.choice([2.0, 10.0, 30.0, 60.0]))
    if reqhgt >= zref:
        reqhgt = 0.05
    ndays = int(rng.integers(1, 4))
    p = synth.make_problem(int(rng.integers(5, 30)), int(rng.integers(5, 30)), 24 * ndays, reqhgt=reqhgt, mode=mode, seed=seed,
                           nlyr=int(rng.integers(1, ndays + 1)), zref=zref, lat=float(rng.uniform(-80, 80)),
                           lon=float(rng.uniform(-180, 180)), complete=bool(rng.random() < 0.5), start_doy=int(rng.integers(0, 360)))
    fz.mutate(p, rng)
    return p


@pytest.mark.skipif(not pyoracle.have_ref(), reason="compiled reference absent")
def test_cpu_checkers_agree_on_fuzzed_inputs():
    """The C restatement used as the fuzz checker stays within 1e-9 of the compiled reference on such inputs."""
    for seed in (7001, 7002, 7003, 7004):
        p = _problem(seed)
        ok, rows = parity.compare(pyoracle.runmicro(p, kind="oracle"), pyoracle.runmicro(p, kind="ref"), atol=1e-9, rtol=1e-9)
        assert ok, f"seed {seed}\n" + parity.fmt(rows)


@pytest.mark.gpu
@pytest.mark.parametrize("block", range(6))
def test_fuzzed_corner_inputs(block):
    for seed in range(9000 + 10 * block, 9000 + 10 * (block + 1)):
        p = _problem(seed)
        ok, rows = parity.compare(api.run_problem(p), pyoracle.runmicro(p, kind="oracle"))
        assert ok, f"seed {seed} mode {p.mode} reqhgt {p.reqhgt}\n" + parity.fmt(rows)
