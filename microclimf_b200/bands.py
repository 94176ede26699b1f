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


# ---------------------------------------------------------------------------------------------------
# Data plane of a multi-GPU run (one process per GPU, torch.distributed): NCCL moves rasters GPU to GPU over NVLink,
# gloo (CPU tensors) runs the same code in the tests.  Nothing here is on the hot path: statics are distributed once
# before the solve, summaries collected once after it (BASELINE north_star: "NCCL over NVLink is used only to
# broadcast the DTM and forcing and to gather the bioclim summaries").
# ---------------------------------------------------------------------------------------------------
def _dist():
    import torch.distributed as dist

    return dist if (dist.is_available() and dist.is_initialized()) else None


def collective_device(group=None) -> str:
    dist = _dist()
    return "cuda" if (dist is not None and dist.get_backend(group) == "nccl") else "cpu"


def rank_world(group=None):
    dist = _dist()
    return (0, 1) if dist is None else (dist.get_rank(group), dist.get_world_size(group))


def broadcast_meta(obj, src: int = 0, group=None):
    """Small Python object (problem scalars, array names and lengths, calendar columns) from `src` to every rank."""
    dist = _dist()
    if dist is None or dist.get_world_size(group) == 1:
        return obj
    box = [obj if dist.get_rank(group) == src else None]
    dist.broadcast_object_list(box, src=src, group=group, device=None)
    return box[0]


def broadcast_f64(arr, n: int, src: int = 0, group=None, device=None):
    """A flat float64 array of `n` elements (numpy on `src`, ignored elsewhere) -> a tensor on the collective's device on
    every rank (CUDA under NCCL: the transfer is GPU-to-GPU; the caller keeps rasters device-resident)."""
    import torch

    dist = _dist()
    dev = device or collective_device(group)
    rank, world = rank_world(group)
    if rank == src:
        t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64).reshape(-1)).to(dev)
        assert t.numel() == n
    else:
        t = torch.empty(n, dtype=torch.float64, device=dev)
    if dist is not None and world > 1:
        dist.broadcast(t, src=src, group=group)
    return t


def scatter_problem(root: "GridProblem | None", src: int = 0, group=None):
    """Column-band sharding of a whole-raster host problem held by rank `src`: every rank ends with ITS band as a
    GridProblem whose float64 arrays are tensors on the collective's device (device-resident under NCCL, ready for the
    *_dev entry points; CPU tensors under gloo).  Per-hour series and coarse-grid climate arrays are replicated
    (broadcast), everything per cell travels point-to-point as one contiguous band per array.  Returns
    (band problem, (c0, c1), whole-raster (rows, cols))."""
    import torch

    dist = _dist()
    rank, world = rank_world(group)
    dev = collective_device(group)
    if rank == src:
        assert root is not None
        rngs = band_ranges(root.cols, world)
        probe = root.band(*rngs[0])
        replicated = [n for n, a in probe.arrays.items() if a is root.arrays[n]]
        meta = dict(fields={k: getattr(root, k) for k in ("mode", "rows", "cols", "tsteps", "reqhgt", "zref", "lat", "lon",
                                                          "Sminp", "Smaxp", "tfact", "mat", "complete", "nlyr", "clim_rows",
                                                          "clim_cols", "clim_row0", "clim_drow", "clim_col0", "clim_dcol",
                                                          "altcorrect")},
                    lyr_st=None if root.lyr_st is None else np.asarray(root.lyr_st), lyr_ed=None if root.lyr_ed is None
                    else np.asarray(root.lyr_ed), ranges=rngs, twi_mean=root.twi_mean,
                    ints={n: np.asarray(root.arrays[n]) for n in ("year", "month", "day")},
                    replicated=[(n, int(np.asarray(root.arrays[n]).size)) for n in replicated if n not in ("year", "month", "day")],
                    banded=[(n, root.expected_len(n) // root.ncells) for n in root.arrays
                            if n not in replicated and n not in ("year", "month", "day")])
    else:
        meta = None
    meta = broadcast_meta(meta, src, group)
    f = meta["fields"]
    c0, c1 = meta["ranges"][rank]
    from .problem import GridProblem as GP

    b = GP(**{**f, "cols": c1 - c0})
    b.clim_col0 = f["clim_col0"] + f["clim_dcol"] * c0
    b.lyr_st, b.lyr_ed = meta["lyr_st"], meta["lyr_ed"]
    b.twi_mean = meta["twi_mean"]  # a caller-supplied whole-raster mean travels with the problem
    for n, a in meta["ints"].items():
        b.arrays[n] = np.ascontiguousarray(a, dtype=np.int32)
    for n, ln in meta["replicated"]:
        b.arrays[n] = broadcast_f64(root.arrays[n] if rank == src else None, ln, src, group, dev)
    R, C = f["rows"], f["cols"]
    for n, nsl in meta["banded"]:
        mine = None
        if rank == src:
            full = np.asarray(root.arrays[n]).reshape(nsl, C, R)
            for r_, (a0, a1) in enumerate(meta["ranges"]):
                part = torch.from_numpy(np.ascontiguousarray(full[:, a0:a1, :]).reshape(-1))
                if r_ == src:
                    mine = part.to(dev)
                else:
                    dist.send(part.to(dev), dst=r_, group=group)
        else:
            mine = torch.empty(nsl * (c1 - c0) * R, dtype=torch.float64, device=dev)
            dist.recv(mine, src=src, group=group)
        b.arrays[n] = mine
    return b, (c0, c1), (R, C)


def gather_rasters(local, nsl: int, rows: int, cols: int, dst: int = 0, group=None):
    """Per-band results [nsl, band_cols * rows] (a tensor on the collective's device, R layout per slice) -> on rank
    `dst` a numpy array [nsl, cols * rows] of the whole raster; None elsewhere.  Point-to-point (bands may differ in
    width), over NVLink under NCCL."""
    import torch

    dist = _dist()
    rank, world = rank_world(group)
    if dist is None or world == 1:
        return local.reshape(nsl, -1).cpu().numpy()
    rngs = band_ranges(cols, world)
    if rank != dst:
        dist.send(local.contiguous(), dst=dst, group=group)
        return None
    out = np.empty((nsl, cols, rows), dtype=np.float64)
    for r_, (a0, a1) in enumerate(rngs):
        if r_ == dst:
            part = local
        else:
            part = torch.empty(nsl * (a1 - a0) * rows, dtype=torch.float64, device=local.device)
            dist.recv(part, src=r_, group=group)
        out[:, a0:a1, :] = part.reshape(nsl, a1 - a0, rows).cpu().numpy()
    return out.reshape(nsl, cols * rows)
