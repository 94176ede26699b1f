#!/bin/bash
# dev aid: the round-2 ncu captures (one GPU).  Each capture runs only after the same program has exited 0 without ncu.
set -x
mkdir -p gpurun_out
NCU="ncu --clock-control none"
for c in headline bio summary; do
  python tools/profile_cases.py $c || exit 1
  $NCU --set full --import-source on -k regex:^k_grid$ --launch-skip 2 --launch-count 1 -f -o gpurun_out/r02_kgrid_$c python tools/profile_cases.py $c > gpurun_out/r02_ncu_$c.log 2>&1
done
python tools/ncu_summary.py gpurun_out/r02_kgrid_headline.ncu-rep 50331648 > gpurun_out/r02_kgrid_headline.txt
python tools/ncu_opmix.py microclimf_b200/csrc/libmicroclimf_b200.so k_grid gpurun_out/r02_kgrid_headline.ncu-rep 50331648 >> gpurun_out/r02_kgrid_headline.txt
python tools/ncu_summary.py gpurun_out/r02_kgrid_bio.ncu-rep 1409286144 > gpurun_out/r02_kgrid_bio.txt
python tools/ncu_opmix.py microclimf_b200/csrc/libmicroclimf_b200.so k_grid gpurun_out/r02_kgrid_bio.ncu-rep 1409286144 >> gpurun_out/r02_kgrid_bio.txt
python tools/ncu_summary.py gpurun_out/r02_kgrid_summary.ncu-rep 50331648 > gpurun_out/r02_kgrid_summary.txt
# DRAM bytes of one launch of the bench's own size, and the launch list of the default bench
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-job --no-configs"
$B > gpurun_out/r02_bench_plain.json 2> gpurun_out/r02_bench_plain.err || exit 1
$NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:^k_grid$ --launch-skip 3 --launch-count 1 --csv --log-file gpurun_out/r02_dram_benchwindow.csv $B > gpurun_out/r02_ncu_dram.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_ncu_launches.log 2>&1
ls -la gpurun_out/r02_*
