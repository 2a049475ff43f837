#!/usr/bin/env python3
"""Static SASS instruction count per source line range (nvdisasm -g) for one kernel: where the code bytes are."""
import re, subprocess, sys
from collections import Counter
cubin, kname = sys.argv[1:3]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
cur, infunc, cnt = None, False, Counter()
for ln in dis:
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        infunc = kname in m.group(1); continue
    if not infunc: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln): cnt[cur] += 1
tot = sum(cnt.values())
print("total", tot)
bins = Counter()
for (f, l), n in cnt.items():
    bins[(f, l // 50 * 50)] += n
for (f, l), n in sorted(bins.items(), key=lambda kv: -kv[1])[:25]:
    print(f"{f}:{l}-{l+49}  {n}  {100*n/tot:.1f}%")
