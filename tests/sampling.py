"""Sampled-cell parity at raster sizes the CPU checker cannot solve whole (test infrastructure).

Cells of the grid model are independent except for ONE whole-raster number, the mean of log(twi)/tfact that
soildCppm subtracts (ref src/microclimfCpp.cpp:993-1004).  A sample of cells re-packed as an `n x 1` raster is therefore
the same problem for those cells as long as both solves subtract the same mean.  The unmodified reference computes the
mean from the raster it is given, so the big solve is run in band mode (`has_twi_mean`) with the SAMPLE's mean — any
value is a legitimate input there — and the compiled reference reproduces it from the sample on its own.
"""
import numpy as np

from microclimf_b200 import _abi
from microclimf_b200.problem import OBSTIME_FIELDS, SERIES_FIELDS


def pick_cells(p, n, seed=4):
    rng = np.random.default_rng(seed)
    return np.sort(rng.choice(p.ncells, min(n, p.ncells), replace=False))


def sample_twi_mean(p, pick):
    """mean over the sample's non-NA cells of log(twi)/tfact, summed in the reference's order (ref :993-1004)."""
    twi = np.asarray(p.arrays["twi"])[pick]
    v = np.log(twi) / p.tfact
    v = v[~np.isnan(v)]
    s = 0.0
    for x in v:  # sequential sum, as the reference's loop
        s += float(x)
    return s / len(v)


def subproblem(p, pick):
    """The cells `pick` (flat R-order indices) of host problem `p` as a len(pick) x 1 raster (fine-grid inputs only)."""
    if p.coarse:
        raise ValueError("coarse-grid climate is sampled through oracle/prep_oracle (expand, then sample)")
    nc = p.ncells
    sub = p._clone_meta()
    sub.rows, sub.cols = len(pick), 1
    sub.twi_mean = None
    for name, arr in p.arrays.items():
        a = np.asarray(arr)
        ln = p.expected_len(name)
        per_cell = not (name in OBSTIME_FIELDS or name == "winddir" or (name in SERIES_FIELDS and not p.array_climate))
        if per_cell:
            sub.arrays[name] = np.ascontiguousarray(a.reshape(ln // nc, nc)[:, pick].ravel())
        else:
            sub.arrays[name] = a
    sub.validate()
    return sub


def gather(flat, pick, nslices, ncells):
    """[nslices, ncells] flat result -> [len(pick), 1, nslices] (the sub-raster's R-shaped array)."""
    a = np.asarray(flat).reshape(nslices, ncells)[:, pick]
    return np.ascontiguousarray(a.T).reshape(len(pick), 1, nslices)
