set -x
python bench.py > gpurun_out/bench_v10.json 2> gpurun_out/bench_v10.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_v10_ref.json 2>/dev/null
CMD="python bench.py --rows 2048 --band-cols 512 --win-days 2 --steps 2 --warmup 3 --no-cpu --e2e-rows 256 --e2e-cols 256 --e2e-hours 48"
$CMD > gpurun_out/prof_cmd_plain.log 2>&1; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v10.csv $CMD > gpurun_out/ncu_launch_v10.log 2>&1; echo "launchlist rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:^k_grid$" --launch-skip 4 --launch-count 1 -f -o gpurun_out/prof_kgrid10 $CMD > gpurun_out/ncu_full_v10.log 2>&1; echo "full rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:^k_grid_f32$" --launch-skip 1 --launch-count 1 -f -o gpurun_out/prof_kgrid_f32_v10 $CMD > gpurun_out/ncu_full_f32_v10.log 2>&1; echo "full f32 rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
