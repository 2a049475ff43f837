#!/usr/bin/env python3
"""Attribute ncu per-SASS-instruction counts to source lines (nvdisasm -g line info joined by instruction order).
usage: ncu_lines.py <report.ncu-rep> <cubin> <kernel-substring> [top]"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

rep, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# walk the disassembly of the wanted function, record (file,line) per instruction
func, cur, lines, infunc = None, None, [], False
for ln in dis:
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        infunc = kname in m.group(1)
        continue
    if not infunc:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        lines.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# find the kernel block
start = next(i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and True)
hdr = rows[start + 1]
ci, ct, cs = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
body = []
for r in rows[start + 2:]:
    if not r or r[0] == "Kernel Name":
        break
    body.append(r)
print(f"sass instrs: ncu {len(body)} nvdisasm {len(lines)}")
agg = defaultdict(lambda: [0, 0, 0])
n = min(len(body), len(lines))
for k in range(n):
    a = agg[lines[k]]
    a[0] += int(body[k][ci]); a[1] += int(body[k][ct]); a[2] += int(body[k][cs])
tot = sum(a[0] for a in agg.values()); tots = sum(a[2] for a in agg.values())
print(f"total warp-instr {tot}  samples {tots}")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{str(key):36s} inst {a[0]:>12d} {100*a[0]/tot:5.1f}%  lanes {a[1]/max(1,a[0]):5.1f}  samples {100*a[2]/max(1,tots):5.1f}%")

# ---- coarse regions of sim_core.h
REGIONS = [(54, 71, "math wrappers"), (72, 84, "warp_sum"), (85, 100, "philox"), (181, 235, "vec helpers (cross/dot6/mv3/inert_mul/rsqrt/contact_u)"),
           (236, 254, "impedance"), (255, 387, "arrow_factor_solve"), (388, 410, "arrow_row_dot"), (411, 471, "P1 FK"), (472, 516, "P2 S+inertia"),
           (517, 542, "P3 comp+V"), (543, 569, "P4 rootcomp+velprod"), (570, 605, "P5 CRBA+A"), (606, 633, "P6 F+corner candidates"),
           (634, 655, "P7 subtree+slots"), (656, 675, "P7b contact params"), (676, 718, "P8 aref+qfs+limits"), (719, 751, "P9 newton init+Pm"),
           (752, 775, "P10 a cF/cW"), (776, 789, "P10 c Ff"), (790, 805, "P10 d grad"), (806, 838, "P10 e Af,T"), (839, 859, "P10 f H"),
           (860, 886, "P10 g images"), (887, 933, "P10 h linesearch+update"), (934, 971, "P11 lagged"), (972, 1031, "P12 euler+integrate"),
           (1032, 1400, "env level")]
reg = defaultdict(lambda: [0, 0, 0])
for (f, ln), a in agg.items() if all(k is not None for k in agg) else [(k, v) for k, v in agg.items() if k is not None]:
    name = f if f != "sim_core.h" else next((n for lo, hi, n in REGIONS if lo <= ln <= hi), "other")
    r = reg[name]
    r[0] += a[0]; r[1] += a[1]; r[2] += a[2]
print("---- by region")
for name, a in sorted(reg.items(), key=lambda kv: -kv[1][0]):
    print(f"{name:40s} inst {100*a[0]/tot:5.1f}%  lanes {a[1]/max(1,a[0]):5.1f}  samples {100*a[2]/max(1,tots):5.1f}%")
