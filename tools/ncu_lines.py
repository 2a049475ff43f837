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
REGIONS = [(60, 82, "math wrappers (out of line)"), (83, 100, "warp_sum"), (101, 120, "philox"), (210, 267, "vec helpers (cross/dot6/mv3/inert_mul/rsqrt/contact_u)"),
           (268, 283, "impedance"), (284, 313, "seg_seg_dist2"), (314, 448, "arrow_factor_solve"), (449, 472, "arrow_row_dot"),
           (473, 510, "constraint_images"), (511, 571, "P1 FK"), (572, 616, "P2 S+inertia"), (617, 642, "P3 comp+V"),
           (643, 669, "P4 rootcomp+velprod"), (670, 705, "P5 CRBA+A"), (706, 733, "P6 F+corner candidates"), (734, 755, "P7 subtree+slots"),
           (756, 775, "P7b contact params"), (776, 819, "P8 aref+qfs+limits"), (820, 835, "P9 Pm + newton init"), (836, 856, "P10 a cF/cW"),
           (857, 870, "P10 c Ff"), (871, 886, "P10 d grad"), (887, 919, "P10 e Af,T"), (920, 940, "P10 f H"), (941, 948, "P10 g images call"),
           (949, 992, "P10 h linesearch+update"), (993, 1057, "P11 lagged+selfcol"), (1058, 1117, "P12 euler+integrate"), (1118, 1500, "env level")]
reg = defaultdict(lambda: [0, 0, 0])
for (f, ln), a in agg.items() if all(k is not None for k in agg) else [(k, v) for k, v in agg.items() if k is not None]:
    name = f if f != "sim_core.h" else next((n for lo, hi, n in REGIONS if lo <= ln <= hi), "other")
    r = reg[name]
    r[0] += a[0]; r[1] += a[1]; r[2] += a[2]
print("---- by region")
for name, a in sorted(reg.items(), key=lambda kv: -kv[1][0]):
    print(f"{name:40s} inst {100*a[0]/tot:5.1f}%  lanes {a[1]/max(1,a[0]):5.1f}  samples {100*a[2]/max(1,tots):5.1f}%")
