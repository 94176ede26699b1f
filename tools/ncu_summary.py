"""Short summary of one ncu report: python tools/ncu_summary.py <report.ncu-rep> [cell_hours]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
ch = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    g = lambda k: float(d[k][0].replace(",", ""))
    print("kernel:", d["Kernel Name"][0], " grid", d["launch__grid_size"][0], "block", d["launch__block_size"][0],
          "regs", d["launch__registers_per_thread"][0])
    dur = g("gpu__time_duration.sum")
    print(f"  duration {dur} {d['gpu__time_duration.sum'][1]}")
    for k in ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
              "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
              "dram__bytes_read.sum", "dram__bytes_write.sum", "sass__inst_executed_register_spilling",
              "smsp__inst_executed.sum", "sm__sass_thread_inst_executed_op_fp64_pred_on.sum",
              "sm__sass_thread_inst_executed_op_dfma_pred_on.sum", "sm__sass_thread_inst_executed_op_dadd_pred_on.sum",
              "sm__sass_thread_inst_executed_op_dmul_pred_on.sum"):
        if k in d:
            print(f"  {k:72s} {d[k][0]} {d[k][1]}")
    st = {k[len("smsp__pcsamp_warps_issue_stalled_"):]: g(k) for k in d if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued")}
    tot = sum(st.values())
    print("  stall samples:", ", ".join(f"{k} {100*v/tot:.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
    if ch and "sm__sass_thread_inst_executed_op_dfma_pred_on.sum" in d:
        wi = g("smsp__inst_executed.sum")
        f = {k: g(f"sm__sass_thread_inst_executed_op_{k}_pred_on.sum") for k in ("dfma", "dadd", "dmul", "fp64")}
        print(f"  per cell-hour: warp-level instr x32 = {wi*32/ch:.0f}, fp64 instr {f['fp64']/ch:.0f} (dfma {f['dfma']/ch:.0f} dadd {f['dadd']/ch:.0f} dmul {f['dmul']/ch:.0f}), "
              f"flop {(2*f['dfma']+f['dadd']+f['dmul'])/ch:.0f}, dram B {(g('dram__bytes_read.sum')+g('dram__bytes_write.sum'))*(1e9 if d['dram__bytes_read.sum'][1]=='Gbyte' else 1e6)/ch:.1f}")
