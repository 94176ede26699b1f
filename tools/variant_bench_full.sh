#!/bin/bash
# usage: tools/variant_bench_full.sh lib.so[:persist_mb] ...   (dev aid: the default bench workload, kernel leg only)
for spec in "$@"; do
  lib=${spec%%:*}; mb=${spec#*:}; [ "$mb" = "$spec" ] && mb=""
  if [ -n "$mb" ]; then export MCF_L2_PERSIST_MB=$mb; else unset MCF_L2_PERSIST_MB; fi
  MCF_LIB_PATH=$PWD/$lib python bench.py --steps 4 --warmup 3 --no-cpu --e2e-rows 256 --e2e-cols 256 --e2e-hours 24 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$spec', '%.3e c-h/s' % d['value'], 'kernel ms %.2f' % d['roofline']['avg_launch_ms'], 'fp32 %.3e' % d['fp32']['value'], 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'], d['clocks'].get('power_w_max'))"
done
