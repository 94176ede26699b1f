"""Dynamic opcode mix of a kernel from an ncu SASS source page + the built library.
usage: ncu_opmix.py <lib.so> <kernel substring> <report.ncu-rep> <cell_hours>"""
import collections, csv, io, os, re, subprocess, sys, tempfile
so, kname, rep, ch = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
agg = collections.Counter()
for r in data:
    ins = r[ix["Source"]].strip(); n = int(r[ix["Instructions Executed"]] or 0)
    t = ins.split(); op = t[1] if t[0].startswith("@") else t[0]
    key = op.split(".")[0]
    if op.startswith("IMAD") and "MOV" in op:
        key = "IMAD.MOV imm" if re.search(r"RZ, RZ, (0x|-?[0-9])", ins) else "IMAD.MOV reg"
    agg[key] += n
tot = sum(agg.values())
print(f"total {tot*32/ch:.0f} thread-instr per cell-hour")
for k, v in agg.most_common(24):
    print(f"  {k:14s} {v*32/ch:8.1f}  {100*v/tot:5.1f}%")
