import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.quick_bench import run
prec = int(sys.argv[1])
for n in (4096, 32768):
    for sigma in (0.0, 0.223):
        ms, sps, it = run(n, prec, sigma=sigma)
        print(f"wpb={os.environ.get('LHW_WARPS_PER_BLOCK','1')} fp{prec} N={n} sigma={sigma}: {ms:.3f} ms/step {sps/1e6:.3f} M/s iters={it:.1f}", flush=True)
