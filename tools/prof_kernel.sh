#!/bin/bash
# dev aid: one ncu --set full capture of a grid-kernel build.  usage: tools/prof_kernel.sh <kernel name> <case> <out tag> <cell_hours> [ENV=VAL ...]
K=$1; CASE=$2; TAG=$3; CH=$4; shift 4
env "$@" python tools/profile_cases.py $CASE || exit 1
env "$@" ncu --clock-control none --set full --import-source on -k regex:^$K\$ --launch-skip 2 --launch-count 1 -f -o gpurun_out/$TAG python tools/profile_cases.py $CASE > gpurun_out/$TAG.log 2>&1
python tools/ncu_summary.py gpurun_out/$TAG.ncu-rep $CH > gpurun_out/$TAG.txt 2>/dev/null
python tools/ncu_opmix.py ${MCF_LIB_PATH:-microclimf_b200/csrc/libmicroclimf_b200.so} $K gpurun_out/$TAG.ncu-rep $CH >> gpurun_out/$TAG.txt
cat gpurun_out/$TAG.txt
