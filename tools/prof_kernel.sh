#!/bin/bash
# dev aid: one ncu --set full capture of a grid-kernel build.
# usage: tools/prof_kernel.sh <kernel name> <case> <out tag> <cell_hours> [ENV=VAL ...]     (KEEP_REP=1 keeps the .ncu-rep)
K=$1; CASE=$2; TAG=$3; CH=$4; shift 4
env "$@" python tools/profile_cases.py $CASE || exit 1
env "$@" ncu --clock-control none --set full --import-source on -k regex:^$K\$ --launch-skip 2 --launch-count 1 -f -o gpurun_out/$TAG python tools/profile_cases.py $CASE > gpurun_out/$TAG.log 2>&1
echo "ncu --set full --clock-control none --import-source on -k regex:^$K\$ --launch-skip 2 --launch-count 1 python tools/profile_cases.py $CASE   [$*]" > gpurun_out/$TAG.txt
echo "cell-hours in the profiled launch: $CH" >> gpurun_out/$TAG.txt
python tools/ncu_summary.py gpurun_out/$TAG.ncu-rep $CH >> gpurun_out/$TAG.txt 2>/dev/null
python tools/ncu_opmix.py ${MCF_LIB_PATH:-microclimf_b200/csrc/libmicroclimf_b200.so} $K gpurun_out/$TAG.ncu-rep $CH >> gpurun_out/$TAG.txt
ncu -i gpurun_out/$TAG.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h,u,v=rows[0],rows[1],rows[2]
keep=('l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__warps_eligible.avg.per_cycle_active','launch__shared_mem_per_block_dynamic','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__inst_executed_pipe_fp64.sum','smsp__sass_thread_inst_executed_op_fp64_pred_on.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum','lts__t_sectors_srcunit_tex_op_read.sum')
for a,b,c in zip(h,u,v):
    if a in keep: print('  %-72s %s %s' % (a,c,b))
" >> gpurun_out/$TAG.txt
[ -n "$KEEP_REP" ] || rm -f gpurun_out/$TAG.ncu-rep
cat gpurun_out/$TAG.txt
