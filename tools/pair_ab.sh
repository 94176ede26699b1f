#!/bin/bash
# dev aid: A/B of the two FP64 grid-kernel builds (k_grid via MCF_NO_PAIR=1, k_grid_pair) and pair-size variants on one box
run() { # label, env...
  local label=$1; shift
  env "$@" python bench.py --rows 4096 --band-cols 1024 --win-days 10 --steps 4 --warmup 3 --no-cpu \
     --e2e-rows 256 --e2e-cols 256 --e2e-hours 24 --no-job --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$label', '%.3e c-h/s' % d['value'], 'kernel ms %.2f' % d['roofline']['avg_launch_ms'], 'clk', d['clocks']['sm_mhz'])"
}
run k_grid MCF_NO_PAIR=1
run pair MCF_NO_PAIR=0
for v in "$@"; do n=${v%:nopair}; if [ "$n" != "$v" ]; then run $v MCF_NO_PAIR=1 MCF_LIB_PATH=$PWD/variants/lib_$n.so; else run $v MCF_LIB_PATH=$PWD/variants/lib_$v.so; fi; done
run k_grid_again MCF_NO_PAIR=1
