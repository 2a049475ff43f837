import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.quick_bench import run
for precision in (64, 32):
    for n in (4096, 32768):
        for sigma in (0.0, 0.223):
            ms, sps, it = run(n, precision, sigma=sigma)
            print(f"fp{precision} N={n} sigma={sigma}: {ms:.3f} ms/step {sps/1e6:.3f} M/s iters={it:.1f}", flush=True)
