#!/usr/bin/env python3
"""profiles/ncu_counters.json from `ncu --set full` reports of one 4096-env step launch per workload / precision:
DRAM bytes and executed warp instructions, tied to the md5 of the step kernels' SASS of the CURRENT build record
(bench.py uses the counters only when its library has the same SASS).
usage: make_ncu_counters.py jvrc_walk/fp64=<rep> [jvrc_walk/fp32=<rep> ...]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from learninghumanoidwalking_b200.build import step_kernel_sass_md5  # noqa: E402

out = {"step_kernel_sass_md5": step_kernel_sass_md5(), "source": "profiles/ncu_counters.json <- " + ", ".join(a.split("=")[1] for a in sys.argv[1:]),
       "launch_4096_envs": {}}
assert out["step_kernel_sass_md5"], "build first (build_record.json)"
for arg in sys.argv[1:]:
    key, rep = arg.split("=")
    rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    def get(name):
        i = hdr.index(name)
        v = float(vals[i].replace(",", ""))
        return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}.get(units[i], 1)
    out["launch_4096_envs"][key] = {"dram_bytes": get("dram__bytes_read.sum") + get("dram__bytes_write.sum"),
                                    "warp_instructions": get("smsp__inst_executed.sum"), "duration_ms_under_ncu": get("gpu__time_duration.sum") / 1e6
                                    if units[hdr.index("gpu__time_duration.sum")] in ("ns", "nsecond") else get("gpu__time_duration.sum"),
                                    "grid": get("launch__grid_size")}
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_counters.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
