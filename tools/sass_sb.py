"""dev aid: scoreboard aliasing in a kernel's SASS.  usage: sass_sb.py <lib.so> <mangled-name substring> [addr_lo addr_hi]
Decodes the control bits of every instruction (stall, write / read barrier, wait mask) and reports waits on a scoreboard
whose pending producers (long-latency loads armed earlier in straight-line order) do not feed the waiting instruction:
such a wait serialises the instruction behind loads it does not need (e.g. a software prefetch sharing a scoreboard with
the prologue loads of the same loop)."""
import re, subprocess, sys
so, kname = sys.argv[1:3]
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
txt = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True).stdout
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0]
    if kname not in name:
        continue
    lines = f.split("\n"); ins = []; i = 0
    while i < len(lines):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", lines[i])
        if m and i + 1 < len(lines):
            w1 = int(re.match(r"\s+/\* (0x[0-9a-f]+) \*/", lines[i + 1]).group(1), 16)
            c = (w1 >> 41) & 0x1FFFFF
            ins.append((int(m.group(1), 16), m.group(2), c & 0xF, (c >> 5) & 7, (c >> 8) & 7, (c >> 11) & 0x3F)); i += 2
        else:
            i += 1
    print(name, len(ins), "instructions")
    def regs(tok, width):
        m = re.match(r"-?\|?~?R(\d+)", tok)
        if not m: return set()
        r = int(m.group(1)); return set(range(r, r + width))
    pending = {s: [] for s in range(6)}
    for a, s, stall, wr, rd, wait in ins:
        t = s.split(None, 1)
        if t[0].startswith("@"):
            t = t[1].split(None, 1)
        op = t[0]; ops = [x.strip() for x in (t[1] if len(t) > 1 else "").split(",")]
        width = 4 if ".128" in op else (2 if (".64" in op or op[0] == "D" or "WIDE" in op) else 1)
        srcs = set()
        for o in ops[1:] if not op.startswith(("ST", "BRA", "CCTL")) else ops:
            for mm in re.finditer(r"R(\d+)(\.64)?", o):
                r = int(mm.group(1)); srcs |= {r, r + 1}
        if lo <= a <= hi and wait:
            for sb in range(6):
                if wait >> sb & 1 and pending[sb]:
                    need = [p for p in pending[sb] if p[2] & srcs]
                    extra = [p for p in pending[sb] if not (p[2] & srcs)]
                    longextra = [p for p in extra if p[1].split()[-1 if False else 0].startswith(("LDG", "LDL", "LD.")) or "LDG" in p[1] or "LDL" in p[1]]
                    if longextra and not any("LDG" in p[1] or "LDL" in p[1] for p in need):
                        print(f"  {a:05x} {s[:60]:60s} waits SB{sb} for " + "; ".join(f"{p[0]:05x} {p[1][:40]}" for p in longextra[:3]) + (f" (+{len(longextra)-3})" if len(longextra) > 3 else ""))
            for sb in range(6):
                if wait >> sb & 1: pending[sb] = []
        elif wait:
            for sb in range(6):
                if wait >> sb & 1: pending[sb] = []
        if wr != 7:
            pending[wr].append((a, s, regs(ops[0], width) if ops else set()))
