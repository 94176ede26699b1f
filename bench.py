#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native microclimf grid solver.

Metric (BASELINE.json): runmicro cell-hours/s.  Workload: configs[3], `runmicro_big` on a synthetic
8192 x 8192 DTM / vegp / soil raster x 8760 h, column-band sharded.  Each GPU holds ONE 8192 x 1024 band
(1/8 of the raster: static layers + the year's forcing resident in HBM) — weak scaling: at N = 8 the
job is exactly the config-4 raster.  A "step" is one pass of the hot path over a 30-day window
(720 h) of the band = 6.04e9 cell-hours per GPU; successive steps walk through the year's windows.
All 10 FP64 outputs are written, into a 24-hour HBM ring (the full [rows, cols, 8760]
result is 4.7 TB per variable and exists nowhere; the ring keeps the HBM write traffic real).

  value     : device-timed whole-job cell-hours/s, inputs resident in HBM (mcf_runmicro_dev)
  e2e       : the same metric through the host-buffer C ABI (mcf_runmicro): pinned host inputs are
              uploaded and all outputs copied back to host inside the timed region, on a bounded tile
  roofline  : dominant kernel (k_grid_pair, the pair build of the grid kernel) — algorithmic HBM bytes / CUDA-event launch time vs measured
              copy bandwidth; plus the FP64-pipe fraction (the binding roofline)
  cpu_baseline : the UNMODIFIED reference C++ (oracle/_ref) on the host cores, bounded sample
  extra keys: e2e_packed (int16 sink through the same host call), e2e_pageable (pageable result buffers, N = 1),
              fp32 (the optional FP32 build on the headline workload), clocks, gpu_launches

stdout carries exactly ONE line (the JSON); everything else, NCCL's banner included, goes to stderr.

`--impl reference` times only the reference CPU path (same metric / config), rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "runmicro cell-hours/sec"
UNIT = "cell-hours/s"
ROWS, BAND_COLS, TSTEPS, WIN_DAYS = 8192, 1024, 8760, 30
REQHGT = 0.05
# algorithmic HBM bytes: 8 B x 10 outputs per cell-hour + 440 B of static layers per cell
BYTES_PER_CELL_HOUR = 80.0
BYTES_PER_CELL_STATIC = 440.0
# FP64 flop per cell-hour of the dominant kernel: measured with ncu, read from profiles/kgrid_metrics.json


def workload_config(args):
    return {
        "workload": ("runmicro_big synthetic 8192x8192 raster x 8760 h, column-band sharded (BASELINE configs[3]); "
                     f"one {args.rows}x{args.band_cols} band per GPU resident in HBM; step = {args.win_days}-day "
                     f"window ({args.win_days * 24} h) of the band; runmicro1Cpp semantics (static veg, broadcast "
                     f"forcing), reqhgt {REQHGT}, all 10 outputs FP64 into a 24-h HBM ring sink"),
        "rows": args.rows, "band_cols_per_gpu": args.band_cols, "tsteps": TSTEPS, "window_hours": args.win_days * 24,
        "outputs": 10, "reqhgt": REQHGT,
        "l2": "per-step traffic (>= 16 GB of output + 3.7 GB static) exceeds the 126 MB L2; no flush needed",
    }


# ----------------------------------------------------------------------------------------------------
# CPU reference leg (oracle/_ref = the unmodified reference C++), N worker processes over column bands
# ----------------------------------------------------------------------------------------------------
SAMPLE_DAYS = [15, 52, 88, 125, 161, 198, 234, 271, 307, 344]  # 10 days spread over the year (240 h)


def _cpu_worker(idx, rows, cols, barrier, q, reps):
    from microclimf_b200 import synth
    from oracle import pyoracle

    kind = "ref" if pyoracle.have_ref() else "oracle"
    p = synth.make_problem(rows, cols, 240, reqhgt=REQHGT, mode=1, seed=20240321 + idx, day_list=SAMPLE_DAYS)
    p.twi_mean = 1.0
    times = []
    for _ in range(reps):
        barrier.wait()
        t0 = time.perf_counter()
        pyoracle.runmicro(p, kind=kind)
        times.append(time.perf_counter() - t0)
        barrier.wait()
    q.put((idx, kind, times))


def cpu_reference(nproc, rows, cols_per_proc, reps):
    """Runs `reps` timed passes; returns (kind, [wall seconds per pass], cell-hours per pass)."""
    import multiprocessing as mp

    ctx = mp.get_context("spawn")
    barrier = ctx.Barrier(nproc + 1)
    q = ctx.Queue()
    procs = [ctx.Process(target=_cpu_worker, args=(i, rows, cols_per_proc, barrier, q, reps)) for i in range(nproc)]
    for pr in procs:
        pr.start()
    walls = []
    for _ in range(reps):
        barrier.wait()
        t0 = time.perf_counter()
        barrier.wait()
        walls.append(time.perf_counter() - t0)
    res = [q.get() for _ in procs]
    for pr in procs:
        pr.join()
    return res[0][1], walls, float(nproc) * rows * cols_per_proc * 240


def job_sink_check(p, summary, hours, pick):
    """CPU leg (checker): the sampled cells `pick` of band problem `p` re-solved hour by hour by the compiled reference
    as a len(pick) x 1 raster; mean / min / max of its arrays against what the job's summary sink holds."""
    import numpy as np

    from microclimf_b200 import _abi
    from microclimf_b200.problem import OBSTIME_FIELDS, SERIES_FIELDS
    from oracle import pyoracle

    kind = "ref" if pyoracle.have_ref() else "oracle"
    nc = p.ncells
    sub = p._clone_meta()
    sub.rows, sub.cols = len(pick), 1
    sub.twi_mean = None if kind == "ref" else p.twi_mean
    for name, arr in p.arrays.items():
        a = np.asarray(arr)
        ln = p.expected_len(name)
        if name in OBSTIME_FIELDS or name == "winddir" or name in SERIES_FIELDS:
            sub.arrays[name] = a
        else:
            sub.arrays[name] = np.ascontiguousarray(a.reshape(ln // nc, nc)[:, pick].ravel())
    t0 = time.perf_counter()
    want = pyoracle.runmicro(sub, kind=kind)
    secs = time.perf_counter() - t0
    worst, bad = 0.0, 0
    for nm in _abi.OUT_NAMES:
        w = want[nm][:, 0, :hours]
        na = np.isnan(w[:, 0])
        with np.errstate(invalid="ignore"):
            stats = {"mean": w.sum(axis=1) / hours, "min": w.min(axis=1), "max": w.max(axis=1)}
        for st, wv in stats.items():
            g = summary[nm][st].ravel(order="F")[pick]
            bad += int((np.isnan(g) != na).sum())
            m = ~na
            if m.any():
                worst = max(worst, float((np.abs(g[m] - wv[m]) / (1e-6 + 1e-6 * np.abs(wv[m]))).max()))
    return {"cells": int(len(pick)), "hours": int(hours), "checker": "unmodified reference C++ (oracle/_ref)" if kind == "ref"
            else "C restatement (oracle/)", "max_err_over_tol": worst, "nan_mismatches": bad, "ok": bool(worst <= 1.0 and bad == 0),
            "tolerance": "1e-6 abs + 1e-6 rel on mean / min / max of each of the 10 outputs", "checker_seconds": secs}


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    nproc = host_cores()
    rows, cpp = 128, 64  # ~2e6 cell-hours per core and pass: a few seconds
    kind, walls, ch = cpu_reference(nproc, rows, cpp, args.warmup + args.steps)
    timed = walls[args.warmup:]
    total = sum(timed)
    value = ch * len(timed) / total
    sample = (f"{nproc} processes x ({rows}x{cpp} cells x 240 h: 10 days spread over the year), same synthetic recipe, "
              f"reqhgt {REQHGT}, all 10 outputs; {'unmodified reference C++ (oracle/_ref, g++ -O2)' if kind == 'ref' else 'C restatement (oracle/)'}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(timed), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nproc, "kind": "reference" if kind == "ref" else "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ----------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(prefix="mcf_clocks_", suffix=".csv")
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
                pw.append(float(parts[3]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(pw))
        return out


# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def gpu_arm(args):
    import numpy as np
    import torch

    from microclimf_b200 import _lib, api, bands, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the grid solver has no CPU fallback")
    torch.cuda.set_device(local_rank)
    # every pinned allocation below (torch's and the library's staging slots) is first-touched after this: the rank sits
    # on the CPUs, and prefers the memory, of the socket its GPU hangs off (microclimf_b200/numa.py; best effort)
    from microclimf_b200 import numa
    numa_rep = numa.bind_to_gpu(local_rank) if not args.no_numa else {"disabled": True}
    L = _lib.lib()
    if L.mcf_set_device(local_rank) != 0:
        raise SystemExit("mcf_set_device failed")
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # (its banner goes to stderr: see _reserve_stdout)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    # ------------------------------------------------------------------ the band problem (HBM-resident)
    rows, cols, T = args.rows, args.band_cols, TSTEPS
    t_gen = time.time()
    hp = synth.make_problem(rows, cols, T, reqhgt=REQHGT, mode=1, seed=20240321 + 1000 * rank)
    s, n = bands.twi_partial_host(hp.arrays["twi"], hp.tfact)
    hp.twi_mean = bands.global_twi_mean(s, n)  # the one whole-raster coupling (tiny all-reduce, off the hot path)
    dp = hp.to_device("cuda")
    ncells = rows * cols
    ring_hours = 24
    outs = [torch.empty(ring_hours * ncells, dtype=torch.float64, device="cuda") for _ in range(10)]
    t_gen = time.time() - t_gen
    nwin = (T // 24) // args.win_days
    cell_hours_step = float(ncells) * args.win_days * 24

    def step(i):
        b0 = (i % nwin) * args.win_days
        api.run_problem_dev(dp, outs, window=(b0, args.win_days, b0 * 24, ring_hours))

    for i in range(args.warmup):
        step(i)
    barrier()
    L.mcf_kernel_timing_enable(1)
    L.mcf_kernel_time_reset()
    L.mcf_launch_count_reset()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    launches = int(L.mcf_launch_count())
    kms, kn = api.kernel_time(reset=True)
    L.mcf_kernel_timing_enable(0)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = cell_hours_step * args.steps * world / (ms * 1e-3)

    # ------------------------------------------------------------------ roofline of the dominant kernel
    peaks, peak_src = measured_peaks()
    k_avg_ms = kms / max(kn, 1)
    bytes_launch = cell_hours_step * BYTES_PER_CELL_HOUR + ncells * BYTES_PER_CELL_STATIC
    achieved = bytes_launch / (k_avg_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_grid_pair<ARR=0,RQ_ABOVE,SINK_F64,ALLOUT=true>", "achieved": achieved, "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": None, "peak_source": peak_src,
                "avg_launch_ms": k_avg_ms, "launches": kn, "kernel_share_of_step": kms / ms,
                "algorithmic_bytes_per_launch": bytes_launch}
    prof = os.path.join(ROOT, "profiles", "kgrid_metrics.json")
    if os.path.exists(prof):
        with open(prof) as f:
            pm = json.load(f)
        # DRAM bytes per launch from the committed ncu capture: of a launch of exactly this size when the workload is the
        # default one (profiles/r02_dram_benchwindow.csv), else the per-cell-hour figure of the 240-hour profile window
        per_ch = pm.get("dram_bytes_per_cell_hour", 0)
        if (args.rows, args.band_cols, args.win_days) == (8192, 1024, 30):
            per_ch = pm.get("dram_bytes_per_cell_hour_bench_window", per_ch)
        roofline["traffic"] = per_ch * cell_hours_step or None
        if pm.get("dram_breakdown_bench_window"):
            roofline["traffic_breakdown_bytes_per_cell_hour"] = pm["dram_breakdown_bench_window"]
        if pm.get("fp64_flop_per_cell_hour"):
            fpk = api.fp64_peak_tflops() if rank == 0 else None
            if fpk:
                ach = pm["fp64_flop_per_cell_hour"] * cell_hours_step / (k_avg_ms * 1e-3) / 1e12
                roofline["fp64"] = {"achieved": ach, "peak": fpk, "unit": "TFLOP/s", "frac": ach / fpk,
                                    "peak_source": "mcf_fp64_peak DFMA micro-benchmark, this run",
                                    "flop_per_cell_hour": pm["fp64_flop_per_cell_hour"],
                                    # from the committed ncu capture (not live): share of cycles the FP64 pipe is busy
                                    # and the FP64 instructions behind it; fewer instructions for the same physics lower
                                    # `achieved` flop/s while raising cell-hours/s, so read `frac` together with these
                                    "pipe_active_pct_ncu": pm.get("fp64_pipe_active_pct"),
                                    "fp64_instr_per_cell_hour_ncu": pm.get("fp64_instr_per_cell_hour"),
                                    "flop_source": pm.get("source", "profiles/")}

    # ------------------------------------------------------------------ what the box can copy: all ranks at once
    # Plain cudaMemcpyAsync of a pinned 1 GiB buffer per rank, device -> host, every rank at the same time: the ceiling
    # of the host-buffer path (80 B per cell-hour of FP64 results) at this N, independent of the solver.
    def copy_ceiling(nbytes=1 << 30, reps=4):
        dbuf = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        hbuf = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        hbuf.copy_(dbuf, non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            hbuf.copy_(dbuf, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        del dbuf, hbuf
        return nbytes * reps * world / dt / 1e9

    d2h_ceiling = copy_ceiling()

    # ------------------------------------------------------------------ e2e through the host-buffer C ABI
    e2e = None
    er, ec, et = args.e2e_rows, args.e2e_cols, args.e2e_hours
    ep = synth.make_problem(er, ec, et, reqhgt=REQHGT, mode=1, seed=777 + rank, start_doy=150)
    ep.twi_mean = hp.twi_mean
    h2d = 0
    pin_keep = []
    for nme, a in list(ep.arrays.items()):  # pinned host inputs
        if nme in ("year", "month", "day"):
            continue
        tp = torch.from_numpy(a).pin_memory()
        pin_keep.append(tp)
        ep.arrays[nme] = tp.numpy()
        h2d += a.nbytes
    eouts_t = [torch.empty(er * ec * et, dtype=torch.float64).pin_memory() for _ in range(10)]
    eouts = [t_.numpy() for t_ in eouts_t]
    d2h = sum(o.nbytes for o in eouts)
    esteps = max(1, min(args.steps, 5))
    for _ in range(2):
        api.run_problem(ep, out_buffers=eouts)
    barrier()
    t0 = time.perf_counter()
    for _ in range(esteps):
        api.run_problem(ep, out_buffers=eouts)
    torch.cuda.synchronize()
    e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_s = float(t.item())
    e2e = {"value": float(er) * ec * et * esteps * world / e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "steps": esteps, "ms_per_step": 1e3 * e_s / esteps,
           # bytes over PCIe per second of the whole call: the host-buffer path is bound by the link (80 B per cell-hour
           # of FP64 results), not by the kernels
           "pcie_gb_per_s": (h2d + d2h) * esteps * world / e_s / 1e9, "pcie_gb_per_s_per_gpu": (h2d + d2h) * esteps / e_s / 1e9,
           # aggregate over all ranks of plain pinned cudaMemcpyAsync device -> host, all ranks copying at once: what the
           # box itself can move at this N — the ceiling of pcie_gb_per_s (both are whole-box figures) — and where this
           # rank's pinned memory lives.  The FP64 sink moves 80 B per cell-hour, so e2e <= ceiling / 80 B whatever the
           # kernels do; the sinks that scale are the ones that move less (e2e_packed: 20 B; job: 240 B per CELL).
           "d2h_ceiling_gb_per_s": d2h_ceiling, "frac_of_d2h_ceiling": d2h * esteps * world / e_s / 1e9 / d2h_ceiling,
           "numa": numa_rep,
           "sample": f"{er}x{ec} cells x {et} h per GPU through mcf_runmicro (pinned host buffers, all 10 outputs "
                     f"copied back; timed with the host clock around the blocking call)"}
    # the same call with the packed integer sink (writetonc's x100 / x1 int16 packing done by the kernels, SURVEY.md
    # NEXT-4): 20 instead of 80 bytes per cell-hour cross PCIe.  Reported beside, not instead of, the FP64 e2e.
    pouts_t = [torch.empty(er * ec * et, dtype=torch.int16).pin_memory() for _ in range(10)]
    pouts = [t_.numpy() for t_ in pouts_t]
    for _ in range(2):
        api.run_problem_packed(ep, out_buffers=pouts)
    barrier()
    t0 = time.perf_counter()
    for _ in range(esteps):
        api.run_problem_packed(ep, out_buffers=pouts)
    torch.cuda.synchronize()
    p_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([p_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        p_s = float(t.item())
    e2e_packed = {"value": float(er) * ec * et * esteps * world / p_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                  "d2h_bytes_per_step": sum(o.nbytes for o in pouts), "steps": esteps, "ms_per_step": 1e3 * p_s / esteps,
                  "pcie_gb_per_s": (h2d + sum(o.nbytes for o in pouts)) * esteps * world / p_s / 1e9,
                  "sample": "same tile through mcf_runmicro_packed: int16 outputs as the reference's writetonc stores them"}
    # the FP32 build through the same host path (mcf_runmicro_f32): 40 instead of 80 bytes per cell-hour over PCIe
    e2e_f32 = None
    try:
        fouts_t = [torch.empty(er * ec * et, dtype=torch.float32).pin_memory() for _ in range(10)]
        fouts = [t_.numpy() for t_ in fouts_t]
        for _ in range(2):
            api.run_problem_f32(ep, out_buffers=fouts)
        barrier()
        t0 = time.perf_counter()
        for _ in range(esteps):
            api.run_problem_f32(ep, out_buffers=fouts)
        torch.cuda.synchronize()
        f_s = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([f_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            f_s = float(t.item())
        e2e_f32 = {"value": float(er) * ec * et * esteps * world / f_s, "unit": UNIT, "dtype": "f32", "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": sum(o.nbytes for o in fouts), "steps": esteps, "ms_per_step": 1e3 * f_s / esteps,
                   "pcie_gb_per_s": (h2d + sum(o.nbytes for o in fouts)) * esteps * world / f_s / 1e9,
                   "sample": "same tile through mcf_runmicro_f32 (the optional FP32 build, float outputs)"}
        del fouts_t, fouts
    except Exception as exc:
        e2e_f32 = {"error": repr(exc)[:200]}
    # the same FP64 call into PAGEABLE result buffers — what R hands the library (its vectors are ordinary memory): served
    # by the pool of copy threads with pinned slots.  N = 1 only; reported beside the pinned e2e.
    e2e_pageable = None
    if world == 1:
        try:
            import numpy as np
            gouts = [np.empty(er * ec * et, dtype=np.float64) for _ in range(10)]
            for _ in range(2):
                api.run_problem(ep, out_buffers=gouts)
            gsteps = max(1, min(args.steps, 3))
            t0 = time.perf_counter()
            for _ in range(gsteps):
                api.run_problem(ep, out_buffers=gouts)
            torch.cuda.synchronize()
            g_s = time.perf_counter() - t0
            e2e_pageable = {"value": float(er) * ec * et * gsteps / g_s, "unit": UNIT, "steps": gsteps,
                            "ms_per_step": 1e3 * g_s / gsteps, "d2h_bytes_per_step": d2h,
                            "sample": "same tile, result buffers in pageable host memory (as R's vectors are)"}
            del gouts
        except Exception as exc:  # an extra key: never let it take the benchmark down
            e2e_pageable = {"error": repr(exc)}
    del pin_keep

    # ------------------------------------------------------------------ the whole job, end to end
    # BASELINE configs[3] to completion: every rank solves ALL 8760 hours of its 8192 x 1024 band (N = 8: the 8192 x 8192
    # raster) through the host-buffer API with the SUMMARY sink — per-cell mean / min / max of the 10 outputs over the year,
    # reduced inside the grid kernel (mcf_runmicro_summary): host statics in, 30 [rows, cols] rasters out, nothing hourly
    # is stored anywhere (the hourly arrays would be 4.7 TB per variable).  Wall clock around the blocking call, max over
    # ranks.  At N = 1 the CPU leg re-solves sampled cells of the band with the compiled reference and compares them with
    # what the sink holds (job["sink_check"]).
    job = None
    if not args.no_job:
        try:
            jp = hp
            pick = None
            if rank == 0 and world == 1 and not args.no_cpu:
                rng = np.random.default_rng(5)
                pick = np.sort(rng.choice(ncells, 256, replace=False))
                tw = np.log(np.asarray(hp.arrays["twi"])[pick]) / hp.tfact
                tw = tw[~np.isnan(tw)]
                acc = 0.0
                for x_ in tw:  # the reference sums sequentially (src/microclimfCpp.cpp:993-1004)
                    acc += float(x_)
                jp = hp.replace(twi_mean=acc / len(tw))  # any mean is a legitimate band input; the sample's lets the
                #                                          unmodified reference reproduce it from the sample alone
            barrier()
            t0 = time.perf_counter()
            jres, jhours = api.run_summary(jp)
            j_s = time.perf_counter() - t0
            if dist is not None:
                t = torch.tensor([j_s], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                j_s = float(t.item())
            job = {"workload": f"{rows}x{cols * world} raster x {jhours} h = {world} band(s) of {rows}x{cols}, one per GPU, whole "
                               f"series in one call of mcf_runmicro_summary (host statics in, per-cell mean/min/max of the 10 "
                               f"outputs out)", "wall_s": j_s, "hours": jhours, "cell_hours": float(ncells) * jhours * world,
                   "value": float(ncells) * jhours * world / j_s, "unit": UNIT, "sink": "summary (30 rasters per band)",
                   "h2d_bytes_per_gpu": int(ncells * BYTES_PER_CELL_STATIC + 25 * 8 * T), "d2h_bytes_per_gpu": int(30 * 8 * ncells)}
            if pick is not None:
                job["sink_check"] = job_sink_check(jp, jres, jhours, pick)
            del jres
        except Exception as exc:  # an extra key: never let it take the headline down
            job = {"error": repr(exc)[:300]}

    # ------------------------------------------------------------------ BASELINE configs[2] and [4] (N = 1 only)
    # Parity cases of the headline metric's definition, timed like `value` (CUDA events, inputs resident in HBM) so that
    # their numbers come from the driver's own run: runbioclim on a 2048 x 2048 raster (336 h, the 19 reductions fused into
    # the grid kernel) and runmicro with gridded climate on a 4096 x 4096 raster (climate on a 41 x 41 grid interpolated
    # in the kernel, 240 h into a 24-h ring) plus the snow-pack operator on a bounded tile through its host-buffer entry.
    configs = None
    if world == 1 and not args.no_configs:
        configs = {}
        del outs
        torch.cuda.empty_cache()

        def ev_timed(fn, reps):
            fn()
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(reps):
                fn()
            a1.record()
            torch.cuda.synchronize()
            return a0.elapsed_time(a1) / reps

        try:
            days, q = synth.bioclim_days()
            bp = synth.make_problem(2048, 2048, 336, reqhgt=REQHGT, mode=1, day_list=days, seed=77)
            bdp = bp.to_device()
            bio = [torch.empty(bp.ncells, dtype=torch.float64, device="cuda") for _ in range(19)]
            bms = ev_timed(lambda: api.run_bioclim_problem_dev(bdp, q["wetq"], q["dryq"], q["hotq"], q["colq"], True, bio), 3)
            t0 = time.perf_counter()
            api.run_bioclim_problem(bp, q["wetq"], q["dryq"], q["hotq"], q["colq"], air=True)
            bh = time.perf_counter() - t0
            configs["bioclim_2048"] = {
                "workload": "BASELINE configs[2]: runbioclim1Cpp on a synthetic 2048x2048 raster, 14 days x 24 h, air, 19 outputs",
                "value": bp.ncells * 336 / (bms * 1e-3), "unit": UNIT, "ms": bms,
                "sink": "19 [rows, cols] rasters; the reductions run inside the grid kernel (no [rows, cols, 336] scratch)",
                "e2e_value": bp.ncells * 336 / bh, "e2e_ms": bh * 1e3,
                "e2e_note": "mcf_runbioclim on pageable host buffers: 1.8 GB of statics up, 19 rasters back"}
            del bp, bdp, bio
            torch.cuda.empty_cache()
        except Exception as exc:
            configs["bioclim_2048"] = {"error": repr(exc)[:300]}
        try:
            cp = synth.make_coarse_problem(4096, 4096, 240, reqhgt=REQHGT, mode=2, crows=41, ccols=41, altcorrect=2)
            cdp = cp.to_device()
            couts = [torch.empty(24 * cp.ncells, dtype=torch.float64, device="cuda") for _ in range(10)]
            cms = ev_timed(lambda: api.run_problem_dev(cdp, couts, window=(0, 10, 0, 24)), 2)
            configs["gridded_snow_4096"] = {
                "workload": "BASELINE configs[4]: runmicro2Cpp semantics on a synthetic 4096x4096 raster x 240 h, climate and "
                            "point model on a 41x41 grid interpolated per cell-hour in the kernel (altcorrect 2), 10 outputs",
                "value": cp.ncells * 240 / (cms * 1e-3), "unit": UNIT, "ms": cms, "sink": "24-h FP64 ring in HBM"}
            del cp, cdp, couts
            torch.cuda.empty_cache()
            from microclimf_b200 import snow
            sn = synth.make_snow_inputs(1024, 1024, 120)
            snow.gridmodelsnow1(sn["obstime"], sn["climdata"], sn["pointm"], sn["vegp"], sn["other"], "Alpine")
            t0 = time.perf_counter()
            snow.gridmodelsnow1(sn["obstime"], sn["climdata"], sn["pointm"], sn["vegp"], sn["other"], "Alpine")
            ss = time.perf_counter() - t0
            configs["gridded_snow_4096"]["snow"] = {
                "workload": "gridmodelsnow1 (hourly snow-pack recurrence) on a 1024x1024 tile x 120 h through mcf_gridmodelsnow, "
                            "host buffers (5 [rows, cols, hours] arrays back)",
                "value": 1024 * 1024 * 120 / ss, "unit": UNIT, "ms": ss * 1e3}
            del sn
        except Exception as exc:
            configs.setdefault("gridded_snow_4096", {})["error"] = repr(exc)[:300]
        outs = [torch.empty(ring_hours * ncells, dtype=torch.float64, device="cuda") for _ in range(10)]

    # ------------------------------------------------------------------ the optional FP32 build, same workload
    # (north_star: within 0.05 degC / 0.5 % radiation; tests/test_f32_gpu.py).  Reported beside the FP64 headline.
    fp32 = None
    try:
        outs32 = [torch.empty(ring_hours * ncells, dtype=torch.float32, device="cuda") for _ in range(10)]

        def step32(i):
            b0 = (i % nwin) * args.win_days
            api.run_problem_f32_dev(dp, outs32, window=(b0, args.win_days, b0 * 24, ring_hours))

        for i in range(2):
            step32(i)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nf = max(1, min(args.steps, 4))
        f0.record()
        for i in range(nf):
            step32(2 + i)
        f1.record()
        barrier()
        fms = f0.elapsed_time(f1)
        if dist is not None:
            t = torch.tensor([fms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            fms = float(t.item())
        v32 = cell_hours_step * nf * world / (fms * 1e-3)
        b32 = (cell_hours_step * 40.0 + ncells * BYTES_PER_CELL_STATIC) / (fms / nf * 1e-3) / 1e9
        fp32 = {"value": v32, "unit": UNIT, "dtype": "f32", "ms_per_step": fms / nf, "steps": nf,
                "roofline": {"bound": "hbm", "achieved": b32, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": b32 / peaks["hbm_gbs"],
                             "algorithmic_bytes_per_cell_hour": 40.0},
                "note": "k_grid_f32: FP32 hour loops (SFU transcendentals), FP64 per-cell invariants, FP32 outputs"}
        del outs32
    except Exception as ex:  # the FP32 build is optional: never let it break the headline line
        fp32 = {"unavailable": str(ex)[:200]}

    # ------------------------------------------------------------------ CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        nproc = host_cores()
        kind, walls, ch = cpu_reference(nproc, 128, 128, 1)
        cpu = {"value": ch / walls[0], "unit": UNIT, "cores": nproc, "kind": "reference" if kind == "ref" else "port",
               "sample": f"{nproc} processes x (128x128 cells x 240 h: 10 days spread over the year), reqhgt {REQHGT}, "
                         f"all 10 outputs; unmodified reference C++ built with g++ -O2 (oracle/_ref)",
               "seconds": walls[0]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args), "e2e": e2e, "e2e_packed": e2e_packed, "e2e_f32": e2e_f32, "e2e_pageable": e2e_pageable,
            "fp32": fp32, "job": job, "configs": configs,
            "gpu_launches": launches,
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clk, "setup_seconds": t_gen,
        }
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


_JSON_OUT = None


def emit(line):
    """The ONE JSON line of the contract, on the process's original stdout."""
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def _reserve_stdout():
    """Rank 0 prints exactly one line on stdout.  Libraries write there too (NCCL prints its version banner on fd 1 at
    NCCL_DEBUG=WARN and above), so fd 1 is pointed at stderr for everything but emit()."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def main():
    _reserve_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--rows", type=int, default=ROWS)
    ap.add_argument("--band-cols", dest="band_cols", type=int, default=BAND_COLS)
    ap.add_argument("--win-days", dest="win_days", type=int, default=WIN_DAYS)
    ap.add_argument("--e2e-rows", dest="e2e_rows", type=int, default=1024)
    ap.add_argument("--e2e-cols", dest="e2e_cols", type=int, default=1024)
    ap.add_argument("--e2e-hours", dest="e2e_hours", type=int, default=120)
    ap.add_argument("--no-cpu", dest="no_cpu", action="store_true")
    ap.add_argument("--no-job", dest="no_job", action="store_true", help="skip the whole-year summary-sink job")
    ap.add_argument("--no-configs", dest="no_configs", action="store_true", help="skip the configs[2] / configs[4] timings")
    ap.add_argument("--no-numa", dest="no_numa", action="store_true", help="do not bind the rank next to its GPU")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "native":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        return reference_arm(args)
    return gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
