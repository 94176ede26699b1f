set -x
python -m pytest tests -q -m gpu -x > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_final.log
python bench.py > gpurun_out/bench_v9.json 2> gpurun_out/bench_v9.err; echo "bench rc=$?"; cat gpurun_out/bench_v9.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_v9_ref.json 2>/dev/null; cat gpurun_out/bench_v9_ref.json
CMD="python bench.py --rows 2048 --band-cols 512 --win-days 2 --steps 2 --warmup 3 --no-cpu --e2e-rows 256 --e2e-cols 256 --e2e-hours 48"
$CMD > gpurun_out/prof_cmd_plain.log 2>&1; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v9.csv $CMD > gpurun_out/ncu_launch_v9.log 2>&1; echo "launchlist rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_grid --launch-skip 4 --launch-count 1 -f -o gpurun_out/prof_kgrid9 $CMD > gpurun_out/ncu_full_v9.log 2>&1; echo "full rc=$?"
ls -la gpurun_out/prof_kgrid9.ncu-rep
