"""Attribute an ncu SASS-level source page to CUDA source lines.

usage: ncu_by_line.py <lib.so> <kernel mangled-name substring> <ncu --page source --csv file> [topN]

nvdisasm -g gives file:line per SASS instruction (libdevice math inherits the call-site line); ncu's
CSV gives per-instruction executed counts and stall samples in the same order.  Prints, per source
line, warp-level instructions executed and stall samples (share of total)."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

so, kname, csvpath = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL, check=True)
lines_of = []
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin"):
        continue
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], stdout=subprocess.PIPE, text=True).stdout
    cur, loc, on = None, ("?", 0), False
    for ln in txt.splitlines():
        m = re.match(r"^\.text\.(\S+):", ln)
        if m:
            on = kname in m.group(1)
            continue
        if ln.startswith("//---------------------") and on and lines_of:
            on = False
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            loc = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"^\s+/\*[0-9a-f]{4,}\*/", ln):
            lines_of.append((loc, ln.split("*/", 1)[1].strip()))
rows = list(csv.reader(open(csvpath)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
print(f"nvdisasm instructions: {len(lines_of)}   ncu rows: {len(data)}")
n = min(len(lines_of), len(data))
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
tot_i = tot_s = 0
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for (loc, _), r in zip(lines_of[:n], data[:n]):
    ie = int(r[ix["Instructions Executed"]] or 0)
    ss = int(r[ix["# Samples"]] or 0)
    a = agg[loc]
    a[0] += ie
    a[1] += ss
    a[2] += 1
    for c in stall_cols:
        v = int(r[ix[c]] or 0)
        if v:
            a[3][c] += v
    tot_i += ie
    tot_s += ss
print(f"total warp-instructions {tot_i}, samples {tot_s}")
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    top = ", ".join(f"{k[6:]}:{v}" for k, v in a[3].most_common(3))
    print(f"{loc[0]:18s}:{loc[1]:4d}  sass {a[2]:4d}  inst {100*a[0]/tot_i:5.2f}%  samples {100*a[1]/tot_s:5.2f}%  [{top}]")
