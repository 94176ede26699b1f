"""Dev probe: where the host-buffer path (mcf_runmicro) spends its time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from microclimf_b200 import api, synth

er, ec, et = 1024, 1024, 120
ep = synth.make_problem(er, ec, et, reqhgt=0.05, mode=1, start_doy=150)
pins = []
for nme, a in list(ep.arrays.items()):
    if nme in ("year", "month", "day"): continue
    tp = torch.from_numpy(a).pin_memory(); pins.append(tp); ep.arrays[nme] = tp.numpy()
outs_t = [torch.empty(er * ec * et, dtype=torch.float64).pin_memory() for _ in range(10)]
outs = [t.numpy() for t in outs_t]
def run(mask, n=3):
    ob = [o if m else None for o, m in zip(outs, mask)]
    api.run_problem(ep, out_buffers=ob)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): api.run_problem(ep, out_buffers=ob)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n
for nout in (1, 2, 5, 10):
    mask = [i < nout for i in range(10)]
    dt = run(mask)
    print(f"outputs {nout:2d}: {dt*1e3:7.1f} ms  -> {er*ec*et/dt:.3e} c-h/s, D2H {nout*er*ec*et*8/dt/1e9:.1f} GB/s effective")
# raw D2H bandwidth
d = torch.empty(er * ec * et, dtype=torch.float64, device="cuda")
for _ in range(2): outs_t[0].copy_(d, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(10): outs_t[i].copy_(d, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"raw pinned D2H: {10*d.numel()*8/dt/1e9:.1f} GB/s")
t0 = time.perf_counter(); x = torch.empty(10 * er * ec * et, dtype=torch.float64, device="cuda"); torch.cuda.synchronize(); t1 = time.perf_counter(); del x; torch.cuda.synchronize(); torch.cuda.empty_cache(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"cudaMalloc 10 GB {1e3*(t1-t0):.1f} ms, free {1e3*(t2-t1):.1f} ms")
