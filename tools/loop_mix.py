"""dev aid: static instruction mix of the two hour loops of a kernel.  usage: loop_mix.py <obj or lib> <mangled substring>"""
import re, collections, subprocess, sys
txt = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], stdout=subprocess.PIPE, text=True).stdout
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0]
    if sys.argv[2] not in name:
        continue
    ins = []
    for ln in f.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", ln)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    print(name, len(ins))
    loops = []
    for addr, i in ins:
        mb = re.search(r"\bBRA\S*\s+(?:\S+,\s*)?(0x[0-9a-f]+)", i)
        if mb:
            tgt = int(mb.group(1), 16)
            if tgt < addr and 0x2000 < addr - tgt < 0x7000:
                loops.append((tgt, addr))
    for tgt, addr in loops[:2]:
        body = [i for a_, i in ins if tgt <= a_ <= addr]
        ops = collections.Counter()
        for i in body:
            t = i.split(); op = t[1] if t[0].startswith("@") else t[0]
            ops[op if op.startswith("LDS") else op.split(".")[0]] += 1
        print(f"  loop {tgt:#x}-{addr:#x}: {len(body)} instr;", ", ".join(f"{o} {n}" for o, n in ops.most_common(22)))
