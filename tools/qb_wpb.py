"""Sweep helper: one (precision, model) at the LHW_WARPS_PER_BLOCK / LHW_BLOCK_SYNC_MODE of the environment."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.quick_bench import run
prec = int(sys.argv[1])
model = sys.argv[2] if len(sys.argv) > 2 else "jvrc_walk"
sizes = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [4096, 32768]
for n in sizes:
    for sigma in (0.223,):
        ms, sps, it = run(n, prec, sigma=sigma, model=model)
        print(f"{model} wpb={os.environ.get('LHW_WARPS_PER_BLOCK','default')} sync={os.environ.get('LHW_BLOCK_SYNC_MODE','default')} fp{prec} N={n} sigma={sigma}: {ms:.3f} ms/step {sps/1e6:.3f} M/s iters={it:.1f}", flush=True)
