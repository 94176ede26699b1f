#!/bin/bash
# dev aid: the round-2 ncu captures of the final build (one GPU).  Each capture runs only after the same program has
# exited 0 without ncu.  headline / headline10: k_grid_pair<RQ_ABOVE,SINK_F64,ALLOUT> on 2048 x 512 cells x 48 h / 240 h.
set -x
mkdir -p gpurun_out
NCU="ncu --clock-control none"
tools/prof_kernel.sh k_grid_pair headline r02_kpair_headline 50331648 > /dev/null
KEEP_REP=1 tools/prof_kernel.sh k_grid_pair headline10 r02_kpair_headline10 251658240 > /dev/null
tools/prof_kernel.sh k_grid headline10 r02_kgrid_headline10 251658240 MCF_NO_PAIR=1 > /dev/null
# DRAM bytes of one launch of the bench's own size, and the launch list of the default bench
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-job --no-configs"
$B > gpurun_out/r02_bench_plain.json 2> gpurun_out/r02_bench_plain.err || exit 1
$NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:^k_grid_pair$ --launch-skip 3 --launch-count 1 --csv --log-file gpurun_out/r02_dram_benchwindow.csv $B > gpurun_out/r02_ncu_dram.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_ncu_launches.log 2>&1
ls -la gpurun_out/r02_*
