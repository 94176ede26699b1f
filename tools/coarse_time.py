"""dev aid: device-resident timing of the BASELINE configs[4] workload (gridded climate on 4096 x 4096 cells x 240 h)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from microclimf_b200 import api, synth
p = synth.make_coarse_problem(4096, 4096, 240, reqhgt=0.05, mode=2, crows=41, ccols=41, altcorrect=2)
dp = p.to_device()
o = [torch.empty(24 * p.ncells, dtype=torch.float64, device="cuda") for _ in range(10)]
for i in range(2):
    api.run_problem_dev(dp, o, window=(0, 10, 0, 24))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(3):
    api.run_problem_dev(dp, o, window=(0, 10, 0, 24))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(os.environ.get("MCF_LIB_PATH", "default"), "coarse 4096^2 x 240 h: %.1f ms = %.3e cell-hours/s" % (ms, p.ncells * 240 / ms * 1e3))
