#!/usr/bin/env python3
"""Turn an `ncu --metrics gpu__time_duration.sum --csv --log-file X.csv` launch list into the markdown table under profiles/.
usage: launch_table.py <launches.csv> <out.md> "<title>" """
import csv
import sys
from collections import defaultdict

src, out, title = sys.argv[1:4]
rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 10]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1e-3)
    a = agg[r[ik]]
    a[0] += 1
    a[1] += v * scale
tot = sum(a[1] for a in agg.values()) or 1.0
n = sum(a[0] for a in agg.values())
lines = [f"# {title}", "", f"`ncu --metrics gpu__time_duration.sum --clock-control none --csv` ({src}). Per-launch times under ncu are cold-cache and",
         "serialised: only the SHARES are meaningful.", "", "| launches | total us | share | kernel |", "|---|---|---|---|"]
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    lines.append(f"| {c} | {t:.1f} | {100 * t / tot:.1f}% | `{k[:110]}` |")
lines += ["", f"total {tot / 1e3:.2f} ms over {n} launches."]
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out)
