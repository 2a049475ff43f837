#!/usr/bin/env python3
"""Turn an .ncu-rep (ncu --set full) of the step kernel into the text summary committed under profiles/.
usage: summarize_profile.py <report.ncu-rep> <cubin> <kernel-substring> <out.md> [title]"""
import csv
import subprocess
import sys

rep, cubin, kname, out = sys.argv[1:5]
title = sys.argv[5] if len(sys.argv) > 5 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
lines = [f"# {title}", "", f"source: `{rep}` (ncu --set full --clock-control none, one launch after warm-up; values under a profiler are",
         "for shares and counters only, never a bench number)", "", "| metric | value | unit |", "|---|---|---|"]
for k in want:
    if k in hdr:
        i = hdr.index(k)
        lines.append(f"| {k} | {vals[i]} | {units[i]} |")
reg = subprocess.run([sys.executable, __file__.replace("summarize_profile.py", "ncu_lines.py"), rep, cubin, kname, "25"],
                     capture_output=True, text=True).stdout
lines += ["", "## warp-instructions and stall samples by source line / phase (nvdisasm line info joined to ncu SASS counters)", "", "```", reg.rstrip(), "```"]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = rows[1]
cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
tot = {h[i]: 0 for i in cols}
for r in rows[2:]:
    if len(r) < len(h):
        continue
    for i in cols:
        try:
            tot[h[i]] += int(r[i])
        except ValueError:
            pass
s = sum(tot.values()) or 1
lines += ["", "## warp stall reasons (all samples)", "", "| reason | share |", "|---|---|"]
lines += [f"| {k} | {100 * v / s:.1f}% |" for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v * 200 > s]
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out)
