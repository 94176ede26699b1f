#!/bin/bash
# usage: tools/loop_sizes.sh [kernel-name-substring]  — static SASS size and loop-body sizes of a kernel in the built library
K=${1:-_ZN3mcf6k_gridILb0ELi0}
cuobjdump -sass $(dirname $0)/../microclimf_b200/csrc/libmicroclimf_b200.so > /tmp/lib.sass
awk -v k="$K" '/Function :/{f=(index($0,k)>0);next} f' /tmp/lib.sass | grep -E "^ +/\*[0-9a-f]{4,}\*/" > /tmp/kg.sass
echo "SASS instructions: $(wc -l < /tmp/kg.sass)"
python3 - <<'PY'
import re, collections
ops = collections.Counter()
for ln in open('/tmp/kg.sass'):
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", ln)
    if not m:
        continue
    addr = int(m.group(1), 16); ins = m.group(2)
    t = ins.split(); op = t[1] if t[0].startswith('@') else t[0]
    ops[op.split('.')[0]] += 1
    mb = re.search(r"\bBRA\S*\s+(?:\S+,\s*)?(0x[0-9a-f]+)", ins)
    if mb:
        tgt = int(mb.group(1), 16)
        if tgt < addr and addr - tgt > 0x400:
            print(f"  loop back-edge {addr:#x} -> {tgt:#x}: body {(addr-tgt)//16} instr ({(addr-tgt)/1024:.1f} KB)")
print("  static mix:", ", ".join(f"{o} {n}" for o, n in ops.most_common(14)))
PY
