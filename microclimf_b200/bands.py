"""Column-band sharding of one raster across ranks (one process per GPU).

Raster cells are independent, so a band needs no halo and the solve has no collective.
The single whole-raster coupling is the mean of log(twi)/tfact that the reference subtracts in
soildCppm (src/microclimfCpp.cpp:993-1004): each rank reduces its own band with `twi_partial`, the
(sum, count) pairs are all-reduced, and every rank passes the global mean via `has_twi_mean`.
In R layout (idx = i + rows*j) a column band is a contiguous slab of every [rows, cols] slice.
"""
from __future__ import annotations

import numpy as np

from .problem import GridProblem


def band_ranges(cols: int, nbands: int):
    """Split `cols` columns into `nbands` contiguous ranges, sizes differing by at most one."""
    base, extra = divmod(cols, nbands)
    out, c0 = [], 0
    for b in range(nbands):
        c1 = c0 + base + (1 if b < extra else 0)
        out.append((c0, c1))
        c0 = c1
    return out


def twi_partial_host(twi: np.ndarray, tfact: float):
    """(sum, count) of log(twi)/tfact over non-NaN cells; numpy restatement of the reduction used for
    the CPU-side (gloo) tests of the sharding logic.  The product path uses mcf_twi_partial / the
    in-kernel reduction."""
    t = np.asarray(twi, dtype=np.float64).ravel()
    l = np.log(t[~np.isnan(t)]) / tfact
    l = l[~np.isnan(l)]
    return float(l.sum()), int(l.size)


def global_twi_mean(local_sum: float, local_count: int, group=None) -> float:
    """All-reduce (sum, count) over the process group (NCCL on GPUs, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_sum / local_count
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([local_sum, float(local_count)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t[0].item() / t[1].item())


def shard(problem: GridProblem, rank: int, world: int, group=None) -> GridProblem:
    """This rank's column band of a host problem, with the whole-raster twi mean attached."""
    c0, c1 = band_ranges(problem.cols, world)[rank]
    b = problem.band(c0, c1)
    s, n = twi_partial_host(b.arrays["twi"], problem.tfact)
    b.twi_mean = global_twi_mean(s, n, group)
    return b


def gather_bands(local: np.ndarray, rows: int, cols: int, world: int, group=None):
    """Gather per-band [rows, band_cols(, n)] results (R layout, flat) on every rank, in band order."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or world == 1:
        return [local]
    outs = [None] * world
    dist.all_gather_object(outs, local, group=group)
    return outs
