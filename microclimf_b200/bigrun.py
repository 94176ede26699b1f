"""runmicro_big on N GPUs: one process per GPU, a column band of the raster each, no collective on the hot path.

The reference's `runmicro_big` (R/Cppwrappers.R:444-543) computes the whole-area terrain layers once, cuts the raster
into tiles sized for an R session's memory (`sqrt(2e7 / nt)` cells a side, :456-470), calls `runmicro` per tile and
writes one file per tile (:519-537).  On B200s the unit of work is a COLUMN BAND per GPU (cells are independent; in R
layout a column band is a contiguous slab of every slice) sized to HBM, the time axis is covered by ONE launch (reducing
sinks) or by windows (hourly sinks), and the only whole-raster coupling — the mean of log(twi)/tfact that soildCppm
subtracts (src/microclimfCpp.cpp:993-1004) — is one all-reduce of two numbers before the solve.

    rank 0 holds the whole-raster problem (terrain layers included: hostmodel.runmicro_big builds them once, as the
    reference does)  ->  bands.scatter_problem: forcing broadcast, statics point-to-point per band (NCCL over NVLink;
    every rank's band ends device-resident)  ->  twi all-reduce  ->  per-rank solve into a SINK  ->  summaries gathered
    on rank 0 (bands.gather_rasters) / files written per band.

Sinks (what the hourly results become; the full [rows, cols, 8760] arrays of BASELINE configs[3] are 4.7 TB per variable
and exist nowhere):
  "summary" : per-cell mean / min / max over the series of each requested output, reduced inside the grid kernel
              (mcf_runmicro_summary_dev) — nothing hourly is stored; 30 doubles per cell come back
  "packed"  : writetonc's int16 packing (R/dataprep.R:1063-1260) produced by the kernels, one netCDF file per band and
              time window (`area_<band>_<window>.nc`), the analogue of the reference's per-tile files; device buffers are
              double-buffered so that the copy-out and the file write of window i overlap the kernels of window i + 1
  "arrays"  : the reference's return value (FP64 [rows, cols, T] per output), gathered on rank 0 — small rasters only

`run_spmd` is the SPMD body (call it from every rank of an initialised process group, e.g. under torchrun);
`run_local` spawns the ranks itself from one Python call (what hostmodel.runmicro_big(gpus = N) uses).
"""
from __future__ import annotations

import os
import socket
import time
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _abi, bands
from .problem import GridProblem

SINKS = ("summary", "packed", "arrays")


def day_blocks(p: GridProblem) -> List[Tuple[int, int]]:
    """(first hour, layer) of every 24-hour block the solver computes, in launch order — the day-block list of the C
    library (ref :2194 for modes 1/2; :2629-2639 + :2770-2800 for modes 3/4)."""
    if p.mode <= 2:
        return [(24 * d, 0) for d in range(p.tsteps // 24)]
    out = []
    for l in range(p.nlyr):
        st, ed = int(p.lyr_st[l]), int(p.lyr_ed[l])
        for d in range((ed - st + 1) // 24):
            out.append((st + 24 * d, l))
    return out


def plan_windows(blocks: Sequence[Tuple[int, int]], window_days: int) -> List[Tuple[int, int, int]]:
    """Cut the block list into windows of at most `window_days` CONTIGUOUS days: (block0, nblocks, first hour)."""
    wins, i = [], 0
    while i < len(blocks):
        j = i + 1
        while j < len(blocks) and j - i < window_days and blocks[j][0] == blocks[j - 1][0] + 24:
            j += 1
        wins.append((i, j - i, blocks[i][0]))
        i = j
    return wins


def _twi_mean(band: GridProblem, group=None) -> float:
    import torch

    twi = band.arrays["twi"]
    if hasattr(twi, "data_ptr"):
        v = torch.log(twi) / band.tfact
        v = v[~torch.isnan(v)]
        s, n = float(v.sum().item()), int(v.numel())
    else:
        s, n = bands.twi_partial_host(twi, band.tfact)
    return bands.global_twi_mean(s, n, group)


def _solve_summary(band: GridProblem, out_mask):
    import torch

    from . import api

    nc = band.ncells
    dev = band.arrays["hgt"].device
    red = torch.empty((30, nc), dtype=torch.float64, device=dev)
    lists = [[red[k * 10 + v] if out_mask[v] else None for v in range(10)] for k in range(3)]
    hours = api.run_summary_dev(band, *lists)
    return red, hours


def _solve_arrays(band: GridProblem, out_mask):
    import torch

    from . import api

    n = band.ncells * band.tsteps
    dev = band.arrays["hgt"].device
    outs = [torch.empty(n, dtype=torch.float64, device=dev) if m else None for m in out_mask]
    api.run_problem_dev(band, outs)
    return outs


def _solve_packed(band: GridProblem, out_mask, window_days: int, write_window):
    """Windows of `window_days` days through two sets of int16 device buffers; `write_window(w, first hour, nhours,
    {name: int16 array [band cells * nhours]})` is called from a writer thread while the next window is being solved."""
    import queue
    import threading

    import torch

    from . import api

    nc = band.ncells
    wins = plan_windows(day_blocks(band), window_days)
    if not wins:
        return 0
    ring = max(w[1] for w in wins) * 24
    dev = band.arrays["hgt"].device
    dbuf = [[torch.empty(ring * nc, dtype=torch.int16, device=dev) if m else None for m in out_mask] for _ in range(2)]
    hbuf = [[torch.empty(ring * nc, dtype=torch.int16).pin_memory() if m else None for m in out_mask] for _ in range(2)]
    solve_s, copy_s = torch.cuda.Stream(), torch.cuda.Stream()
    done_k = [torch.cuda.Event() for _ in range(2)]
    done_c = [torch.cuda.Event() for _ in range(2)]
    host_free = [threading.Event() for _ in range(2)]
    for e in host_free:
        e.set()
    q: "queue.Queue" = queue.Queue()
    err: List[BaseException] = []

    def writer():
        while True:
            item = q.get()
            if item is None:
                return
            w, s = item
            try:
                done_c[s].synchronize()
                b0, nb, k0 = wins[w]
                write_window(w, k0, nb * 24, {nm: hbuf[s][v].numpy()[: nb * 24 * nc] for v, nm in enumerate(_abi.OUT_NAMES)
                                              if out_mask[v]})
            except BaseException as exc:  # surfaced by the caller
                err.append(exc)
            finally:
                host_free[s].set()

    th = threading.Thread(target=writer, daemon=True)
    th.start()
    for w, (b0, nb, k0) in enumerate(wins):
        s = w & 1
        host_free[s].wait()  # the writer is done with this slot's pinned buffers (and so is its device->host copy)
        host_free[s].clear()
        with torch.cuda.stream(solve_s):
            solve_s.wait_event(done_c[s])  # the device buffers of this slot have been copied out
            api.run_problem_packed_dev(band, dbuf[s], window=(b0, nb, k0, ring), stream=solve_s)
            done_k[s].record(solve_s)
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(done_k[s])
            for v in range(10):
                if out_mask[v]:
                    hbuf[s][v][: nb * 24 * nc].copy_(dbuf[s][v][: nb * 24 * nc], non_blocking=True)
            done_c[s].record(copy_s)
        q.put((w, s))
    q.put(None)
    th.join()
    torch.cuda.synchronize()
    if err:
        raise err[0]
    return sum(w[1] for w in wins) * 24


def run_spmd(root: Optional[GridProblem], sink: str = "summary", out: Sequence[bool] = (True,) * 10, pathout: Optional[str] = None,
             window_days: int = 5, dtm=None, group=None, gather: bool = True) -> Dict[str, object]:
    """SPMD body of a multi-GPU runmicro_big.  `root`: the whole-raster host problem on rank 0 (None elsewhere).
    Returns on rank 0 (other ranks: timings only):
      summary : {"summary": {name: {"mean" | "min" | "max": [rows, cols]}}, "hours": h}
      arrays  : {"arrays": {name: [rows, cols, T]}}
      packed  : {"files": [...]} (each rank writes its own band's windows into `pathout`)
    plus "timings" (seconds: scatter, solve, gather) and "bands"."""
    import torch

    if sink not in SINKS:
        raise ValueError(f"sink must be one of {SINKS}")
    out = [bool(o) for o in out]
    rank, world = bands.rank_world(group)
    t0 = time.perf_counter()
    band, (c0, c1), (R, C) = bands.scatter_problem(root, 0, group)
    if band.twi_mean is None:  # else: the caller's mean (e.g. this raster is itself part of a larger area)
        band.twi_mean = _twi_mean(band, group)
    on_gpu = bands.collective_device(group) == "cuda" or (world == 1 and torch.cuda.is_available())
    if world == 1 and on_gpu:
        band = band.to_device("cuda")
    if on_gpu:
        torch.cuda.synchronize()
    t1 = time.perf_counter()
    res: Dict[str, object] = {"bands": bands.band_ranges(C, world), "rank": rank}
    if not on_gpu:
        raise RuntimeError("bigrun.run_spmd needs a CUDA device per rank: the solver has no CPU fallback")
    if band.reqhgt < 0 and sink == "summary":
        raise ValueError("the summary sink covers reqhgt >= 0")
    files: List[str] = []
    if sink == "summary":
        red, hours = _solve_summary(band, out)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        full = bands.gather_rasters(red, 30, R, C, 0, group) if gather else None
        if not gather:
            res["band_summary"] = red
        if full is not None:
            summ = {}
            for v, nm in enumerate(_abi.OUT_NAMES):
                if not out[v]:
                    continue
                s = full[v].reshape(C, R).T
                with np.errstate(invalid="ignore"):
                    mean = np.where(np.isnan(s), s, s / max(hours, 1))
                summ[nm] = {"mean": mean, "min": full[10 + v].reshape(C, R).T, "max": full[20 + v].reshape(C, R).T}
            res["summary"] = summ
        res["hours"] = hours
    elif sink == "arrays":
        outs = _solve_arrays(band, out)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        arrays = {}
        T = band.tsteps
        for v, nm in enumerate(_abi.OUT_NAMES):
            if outs[v] is None:
                continue
            full = bands.gather_rasters(outs[v], T, R, C, 0, group)
            if full is not None:
                arrays[nm] = np.ascontiguousarray(full.reshape(T, C, R).transpose(2, 1, 0))
        if rank == 0:
            res["arrays"] = arrays
    else:
        if pathout is None:
            raise ValueError("the packed sink writes files: pathout is required")
        from .ncwriter import writetonc

        os.makedirs(pathout, exist_ok=True)
        tme = None if root is None else getattr(root, "tme", None)
        tme = bands.broadcast_meta(tme, 0, group)
        dtm_all = bands.broadcast_meta(dtm, 0, group)
        Rb, Cb = band.rows, band.cols

        def write_window(w, k0, nh, arrs):
            fo = os.path.join(pathout, f"area_{rank + 1:02d}_{w + 1:03d}")
            mout = {nm: a.reshape(nh, Cb, Rb).transpose(2, 1, 0) for nm, a in arrs.items()}
            if dtm_all is not None and tme is not None:
                mout["tme"] = np.asarray(tme)[k0:k0 + nh]
                writetonc(mout, fo + ".nc", dtm_all.crop(0, R, c0, c1), band.reqhgt, vars=tuple(arrs))
                files.append(fo + ".nc")
            else:  # no georeference given: the same integers as a plain .npz
                np.savez(fo + ".npz", first_hour=k0, **mout)
                files.append(fo + ".npz")

        hours = _solve_packed(band, out, window_days, write_window)
        t2 = time.perf_counter()
        res["hours"] = hours
        res["files"] = files
    t3 = time.perf_counter()
    res["timings"] = {"scatter_s": t1 - t0, "solve_s": t2 - t1, "gather_s": t3 - t2}
    return res


# ---------------------------------------------------------------------------------------------------
# one Python call -> N ranks
# ---------------------------------------------------------------------------------------------------
def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _child(rank: int, world: int, port: int, kw: dict, backend: str):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if backend == "nccl":
        from . import _lib, numa

        numa.bind_to_gpu(rank)
        torch.cuda.set_device(rank)
        _lib.lib().mcf_set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    else:
        dist.init_process_group(backend, rank=rank, world_size=world)
    try:
        run_spmd(None, **kw)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def run_local(root: GridProblem, gpus: int = 1, **kw) -> Dict[str, object]:
    """hostmodel.runmicro_big(gpus = N): the calling process becomes rank 0 (GPU 0) and spawns ranks 1 .. N-1, one per
    GPU; data reach them through NCCL (bands.scatter_problem), results come back the same way."""
    import torch
    import torch.distributed as dist
    import torch.multiprocessing as mp

    if gpus <= 1:
        return run_spmd(root, **kw)
    if dist.is_available() and dist.is_initialized():
        return run_spmd(root, **kw)  # already inside an SPMD launch (torchrun): use its ranks
    if torch.cuda.device_count() < gpus:
        raise RuntimeError(f"runmicro_big(gpus = {gpus}) needs {gpus} CUDA devices, found {torch.cuda.device_count()}")
    from . import _lib, numa

    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_child, args=(r, gpus, port, kw, "nccl")) for r in range(1, gpus)]
    for p in procs:
        p.start()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    numa.bind_to_gpu(0)
    torch.cuda.set_device(0)
    _lib.lib().mcf_set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=gpus, device_id=torch.device("cuda", 0))
    try:
        res = run_spmd(root, **kw)
        dist.barrier()
    finally:
        dist.destroy_process_group()
        for p in procs:
            p.join(timeout=600)
    bad = [p.exitcode for p in procs if p.exitcode != 0]
    if bad:
        raise RuntimeError(f"runmicro_big: worker ranks exited with {bad}")
    return res
