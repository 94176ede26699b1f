"""Device-resident throughput of the other BASELINE configs on one GPU (dev / documentation aid).
Prints one JSON object per case: cell-hours/s with CUDA-event timing, inputs resident in HBM."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from microclimf_b200 import api, synth

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def case_runmicro(name, rows, cols, T, mode, reqhgt, nlyr=1, ring=None):
    p = synth.make_problem(rows, cols, T, reqhgt=reqhgt, mode=mode, nlyr=nlyr)
    dp = p.to_device()
    nc = p.ncells
    hours = ring or T
    mask = [True] * 10
    if reqhgt == 0: mask = [1, 0, 0, 1, 0, 1, 1, 1, 1, 1]
    if reqhgt < 0: mask = [1, 0, 0, 1, 0, 0, 0, 0, 0, 0]
    outs = [torch.empty(hours * nc, dtype=torch.float64, device="cuda") if m else None for m in mask]
    win = None if ring is None else (0, T // 24, 0, ring)
    ms = timed(lambda: api.run_problem_dev(dp, outs, window=win))
    print(json.dumps({"case": name, "rows": rows, "cols": cols, "hours": T, "mode": mode, "reqhgt": reqhgt, "nlyr": nlyr,
                      "ms": ms, "cell_hours_per_s": nc * (T // 24) * 24 / (ms * 1e-3)}), flush=True)
    del dp, outs; torch.cuda.empty_cache()

def case_coarse(name, rows, cols, T, crows, ccols, altcorrect=2, ring=24, packed=False):
    """config 5 as this build runs it: mode 2 with the climate on the coarse grid, interpolated in the kernels."""
    p = synth.make_coarse_problem(rows, cols, T, reqhgt=0.05, mode=2, crows=crows, ccols=ccols, altcorrect=altcorrect)
    dp = p.to_device()
    nc = p.ncells
    dt = torch.int16 if packed else torch.float64
    outs = [torch.empty(ring * nc, dtype=dt, device="cuda") for _ in range(10)]
    win = (0, T // 24, 0, ring)
    run = api.run_problem_packed_dev if packed else api.run_problem_dev
    ms = timed(lambda: run(dp, outs, window=win))
    print(json.dumps({"case": name, "rows": rows, "cols": cols, "hours": T, "coarse": [crows, ccols], "altcorrect": altcorrect,
                      "packed": packed, "ms": ms, "cell_hours_per_s": nc * (T // 24) * 24 / (ms * 1e-3)}), flush=True)
    del dp, outs; torch.cuda.empty_cache()

def case_snow(rows, cols, T):
    """Snow operators through the host-buffer C ABI (uploads and copies back included; host clock)."""
    from microclimf_b200 import snow
    s = synth.make_snow_inputs(rows, cols, T)
    for rep in range(2):
        t0 = time.perf_counter()
        r = snow.gridmodelsnow1(s["obstime"], s["climdata"], s["pointm"], s["vegp"], s["other"], "Alpine")
        t1 = time.perf_counter()
    snowm = dict(Tc=r["Tc"], Tg=r["Tg"], totalSWE=np.nan_to_num(r["sdepc"] * r["sden"]), groundsnowdepth=r["sdepg"], snowden=r["sden"])
    micro = {n: np.zeros(r["Tc"].shape, order="F") for n in ("Tz", "tleaf", "relhum", "soilm", "windspeed", "Rdirdown", "Rdifdown", "Rlwdown", "Rswup", "Rlwup")}
    for rep in range(2):
        t2 = time.perf_counter()
        snow.gridmicrosnow1(0.05, s["obstime"], s["climdata"], snowm, micro, s["vegp"], s["other"], 3.0, [True] * 10)
        t3 = time.perf_counter()
    ch = rows * cols * T
    print(json.dumps({"case": "snow: gridmodelsnow1 / gridmicrosnow1 through the host C ABI", "rows": rows, "cols": cols, "hours": T,
                      "gridmodelsnow1_cell_hours_per_s": ch / (t1 - t0), "gridmicrosnow1_cell_hours_per_s": ch / (t3 - t2),
                      "snow_covered_fraction": float((snowm["totalSWE"] > 0).mean())}), flush=True)

def case_f32(rows, cols, T, ring=24):
    """The optional FP32 build on the headline workload shape (mode 1, reqhgt 0.05, 10 outputs, 24-h ring)."""
    p = synth.make_problem(rows, cols, T, reqhgt=0.05, mode=1)
    dp = p.to_device()
    nc = p.ncells
    win = (0, T // 24, 0, ring)
    o32 = [torch.empty(ring * nc, dtype=torch.float32, device="cuda") for _ in range(10)]
    o64 = [torch.empty(ring * nc, dtype=torch.float64, device="cuda") for _ in range(10)]
    ms32 = timed(lambda: api.run_problem_f32_dev(dp, o32, window=win))
    ms64 = timed(lambda: api.run_problem_dev(dp, o64, window=win))
    err = {n: float((a.double() - b).abs().nan_to_num().max()) for n, a, b in zip(("Tz", "tleaf", "relhum", "soilm", "windspeed", "Rdirdown", "Rdifdown", "Rlwdown", "Rswup", "Rlwup"), o32, o64)}
    ch = nc * (T // 24) * 24
    print(json.dumps({"case": "FP32 build vs FP64 build, mode 1, reqhgt 0.05, 10 outputs, 24-h ring", "rows": rows, "cols": cols, "hours": T,
                      "fp32_ms": ms32, "fp64_ms": ms64, "fp32_cell_hours_per_s": ch / (ms32 * 1e-3), "fp64_cell_hours_per_s": ch / (ms64 * 1e-3),
                      "fp32_algorithmic_GBps": ch * 40 / (ms32 * 1e-3) / 1e9, "max_abs_diff_last_day": err}), flush=True)

def case_bioclim(name, rows, cols, mode):
    days, q = synth.bioclim_days()
    p = synth.make_problem(rows, cols, 336, reqhgt=0.05, mode=mode, nlyr=14, day_list=days)
    dp = p.to_device()
    bio = [torch.empty(p.ncells, dtype=torch.float64, device="cuda") for _ in range(19)]
    ms = timed(lambda: api.run_bioclim_problem_dev(dp, q["wetq"], q["dryq"], q["hotq"], q["colq"], True, bio))
    print(json.dumps({"case": name, "rows": rows, "cols": cols, "hours": 336, "mode": mode, "ms": ms,
                      "cell_hours_per_s": p.ncells * 336 / (ms * 1e-3)}), flush=True)
    del dp, bio; torch.cuda.empty_cache()

if __name__ == "__main__":
    which = sys.argv[1:] or ["bio", "heights", "layers", "array", "coarse"]
    if "f32" in which:
        case_f32(4096, 1024, 240)
    if "snow" in which:
        case_snow(int(os.environ.get("SNOW_N", "1024")), int(os.environ.get("SNOW_N", "1024")), 240)
    if "coarse" in which:
        case_coarse("config5: mode 2, 4096x4096 x 240 h, climate on a 41x41 grid interpolated in-kernel", 4096, 4096, 240, 41, 41)
        case_coarse("same, packed int16 sink", 4096, 4096, 240, 41, 41, packed=True)
    if "bio" in which:
        case_bioclim("config3 runbioclim 2048x2048 (mode 1)", 2048, 2048, 1)
    if "heights" in which:
        for rq in (0.05, 0.0, 5.0, -0.1, -1.0):
            case_runmicro(f"mode1 2048x1024x240h reqhgt {rq}", 2048, 1024, 240, 1, rq)
    if "layers" in which:
        case_runmicro("config1-2 shape: mode 3, 12 layers, 1024x1024 x 8760 h, ring 24", 1024, 1024, 8760, 3, 0.05, nlyr=12, ring=24)
    if "array" in which:
        case_runmicro("config5 shape: mode 2 (array climate) 2048x1024 x 120 h", 2048, 1024, 120, 2, 0.05)
        case_runmicro("mode 4 (array climate, 3 layers) 1024x1024 x 120 h", 1024, 1024, 120, 4, 0.05, nlyr=3)
