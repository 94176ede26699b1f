#!/bin/bash
# usage: tools/variant_bench.sh lib1.so lib2.so ...   (dev aid: compares kernel-build variants on one box)
for lib in "$@"; do
  MCF_LIB_PATH=$PWD/$lib python bench.py --rows 4096 --band-cols 1024 --win-days 10 --steps 4 --warmup 3 --no-cpu \
     --e2e-rows 256 --e2e-cols 256 --e2e-hours 24 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib', '%.3e c-h/s' % d['value'], 'kernel ms %.2f' % d['roofline']['avg_launch_ms'], 'clk', d['clocks']['sm_mhz'])"
done
