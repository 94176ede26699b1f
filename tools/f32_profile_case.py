import sys, os
sys.path.insert(0, os.getcwd())
import torch
from microclimf_b200 import api, synth
p = synth.make_problem(2048, 512, 48, reqhgt=0.05, mode=1)
dp = p.to_device()
o32 = [torch.empty(24 * p.ncells, dtype=torch.float32, device="cuda") for _ in range(10)]
for _ in range(3):
    api.run_problem_f32_dev(dp, o32, window=(0, 2, 0, 24))
torch.cuda.synchronize()
