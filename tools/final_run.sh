# dev aid: the round-end sequence on one box — GPU tests, smoke, both bench arms
python -m pytest tests -q -m gpu -x > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference > gpurun_out/bench_final_ref.json 2>/dev/null; echo "ref rc=$?"; wc -l gpurun_out/bench_final_ref.json
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; wc -l gpurun_out/bench_final.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_final.json')); r=json.load(open('gpurun_out/bench_final_ref.json'))
print('value %.4e  e2e %.4e (%.1f GB/s)  packed %.4e  fp32 %.4e  cpu %.3e  ref-arm %.3e' % (d['value'], d['e2e']['value'], d['e2e']['pcie_gb_per_s'], d['e2e_packed']['value'], d['fp32']['value'], d['cpu_baseline']['value'], r['value']))
print(d['clocks'], d['roofline']['frac'], d['roofline']['fp64']['frac'], d['roofline']['traffic']/d['roofline']['algorithmic_bytes_per_launch'], d['gpu_launches'])
PY
