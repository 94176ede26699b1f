"""Dev probe: mcf_runmicro with ordinary (pageable) numpy buffers, as an R caller would pass."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from microclimf_b200 import api, synth
er, ec, et = 1024, 1024, 120
ep = synth.make_problem(er, ec, et, reqhgt=0.05, mode=1, start_doy=150)
outs = [np.empty(er * ec * et) for _ in range(10)]
for o in outs: o[:] = 0  # touch pages
api.run_problem(ep, out_buffers=outs)
t0 = time.perf_counter()
for _ in range(3): api.run_problem(ep, out_buffers=outs)
dt = (time.perf_counter() - t0) / 3
print(f"pageable buffers: {dt*1e3:.1f} ms -> {er*ec*et/dt:.3e} c-h/s, D2H {10*er*ec*et*8/dt/1e9:.1f} GB/s effective")
