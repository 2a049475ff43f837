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

# ---- coarse regions of sim_core.h, derived from anchor text in the header itself (no hand-kept line numbers)
import os
ANCHORS = [("LHW_DEV float m_sqrt", "math wrappers (out of line)"), ("LHW_DEV real warp_sum", "warp_sum"),
           ("LHW_DEV unsigned warp_ballot", "ballot / bit helpers"), ("LHW_DEV void philox", "philox"),
           ("struct Model {", "struct definitions"), ("LHW_DEV void cross(", "vec helpers (cross/dot6/mv3/inert_mul/rsqrt/contact_u)"),
           ("LHW_DEV void pmap_sel(", "pmap_sel / sflip"), ("// stepping stones, broad phase", "slab broad phase"),
           ("// ---------------- P7x ", "P7x slab-edge crossings"), ("LHW_DEV void slab_frames(", "env level"),
           ("LHW_DEVNI real impedance(", "impedance"), ("LHW_DEV real seg_seg_dist2(", "seg_seg_dist2"),
           ("LHW_DEVNI void arrow_factor_solve(", "arrow_factor_solve"), ("LHW_DEV real arrow_row_dot(", "arrow_row_dot"),
           ("LHW_DEVNI void constraint_images(", "constraint_images"), ("LHW_DEV real floss_force(", "floss_force"),
           ("LHW_DEVNI void substep(", "substep prologue"), ("// ---------------- P1 ", "P1 FK"), ("// ---------------- P2 ", "P2 S+inertia"),
           ("// ---------------- P3 ", "P3 comp+V"), ("// ---------------- P4 ", "P4 rootcomp+velprod"), ("// ---------------- P5 ", "P5 CRBA+A"),
           ("// ---------------- P6 ", "P6 F+contact candidates"), ("// ---------------- P7 ", "P7 subtree+slots"),
           ("// ---------------- P7b ", "P7b contact params"), ("// ---------------- P8 ", "P8 aref+qfs+limits+floss"),
           ("// ---------------- P9 ", "P9 Pm + newton init"), ("// ---------------- P10 ", "P10 loop head"), ("    // (a) per contact", "P10 a cF/cW"),
           ("    // (c) per foot", "P10 c Ff"), ("    // (d) gradient", "P10 d grad"), ("    // (e) a Newton step", "P10 e Af,T"),
           ("    // (f) H = ", "P10 f H"), ("    // (g) images", "P10 g images call"), ("    // (h) exact line search", "P10 h linesearch+update"),
           ("// ---------------- P11 ", "P11 lagged+selfcol"), ("// ---------------- P12 ", "P12 euler+integrate"),
           ("LHW_DEV void load_state(", "env level")]
hdr_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "learninghumanoidwalking_b200", "csrc", "sim_core.h")
src_lines = open(hdr_path).read().splitlines()
marks = []
for text, name in ANCHORS:
    ln = next((k + 1 for k, l in enumerate(src_lines) if text in l), None)
    if ln is not None:
        marks.append((ln, name))
marks.sort()
REGIONS = [(lo, (marks[k + 1][0] - 1) if k + 1 < len(marks) else 10 ** 6, name) for k, (lo, name) in enumerate(marks)]
reg = defaultdict(lambda: [0, 0, 0])
for (f, ln), a in agg.items() if all(k is not None for k in agg) else [(k, v) for k, v in agg.items() if k is not None]:
    name = f if f != "sim_core.h" else next((n for lo, hi, n in REGIONS if lo <= ln <= hi), "other")
    r = reg[name]
    r[0] += a[0]; r[1] += a[1]; r[2] += a[2]
print("---- by region")
for name, a in sorted(reg.items(), key=lambda kv: -kv[1][0]):
    print(f"{name:40s} inst {100*a[0]/tot:5.1f}%  lanes {a[1]/max(1,a[0]):5.1f}  samples {100*a[2]/max(1,tots):5.1f}%")
