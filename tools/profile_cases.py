"""Small device-resident drivers for ncu captures (dev aid): python tools/profile_cases.py {headline|bio|summary}
headline: k_grid<0,RQ_ABOVE,SINK_F64,ALLOUT> on 2048 x 512 cells x 48 h into a 24-h ring (the profile workload of round 1)
bio     : k_grid<0,RQ_ABOVE,SINK_BIO> on the BASELINE configs[2] raster (2048 x 2048 x 336 h)
summary : k_grid<0,RQ_ABOVE,SINK_SUMMARY> on 2048 x 512 cells x 48 h
coarse  : k_grid<2,RQ_ABOVE,SINK_F64> (gridded climate interpolated in the kernel) on 2048 x 512 cells x 48 h"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from microclimf_b200 import api, synth  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "headline"
if which == "bio":
    days, q = synth.bioclim_days()
    p = synth.make_problem(2048, 2048, 336, reqhgt=0.05, mode=1, day_list=days, seed=77)
    dp = p.to_device()
    bio = [torch.empty(p.ncells, dtype=torch.float64, device="cuda") for _ in range(19)]
    for _ in range(3):
        api.run_bioclim_problem_dev(dp, q["wetq"], q["dryq"], q["hotq"], q["colq"], True, bio)
elif which == "coarse":
    # k_grid<2,RQ_ABOVE,SINK_F64>: gridded climate on a 41 x 41 grid interpolated in the kernel (BASELINE configs[4] shape)
    p = synth.make_coarse_problem(2048, 512, 48, reqhgt=0.05, mode=2, crows=41, ccols=41, altcorrect=2)
    dp = p.to_device()
    o = [torch.empty(24 * p.ncells, dtype=torch.float64, device="cuda") for _ in range(10)]
    for _ in range(3):
        api.run_problem_dev(dp, o, window=(0, 2, 0, 24))
else:
    hours = 240 if which == "headline10" else 48  # headline10: ten days per tile (tile set-up amortised as in the bench)
    p = synth.make_problem(2048, 512, hours, reqhgt=0.05, mode=1)
    dp = p.to_device()
    if which == "summary":
        s = [[torch.empty(p.ncells, dtype=torch.float64, device="cuda") for _ in range(10)] for _ in range(3)]
        for _ in range(3):
            api.run_summary_dev(dp, *s)
    else:
        o = [torch.empty(24 * p.ncells, dtype=torch.float64, device="cuda") for _ in range(10)]
        for _ in range(3):
            api.run_problem_dev(dp, o, window=(0, hours // 24, 0, 24))
torch.cuda.synchronize()
print("done", which)
