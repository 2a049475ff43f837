#!/usr/bin/env python3
"""Quick device-resident step-kernel timing sweep (development tool; bench.py is the contract)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learninghumanoidwalking_b200.envs import BatchedHumanoidEnv  # noqa: E402


def run(n, precision, steps=40, warm=10, sigma=0.0, model="jvrc_walk"):
    env = BatchedHumanoidEnv(n, model=model, precision=precision, seed=0)
    env.reset()
    A = env.act_dim
    a = torch.zeros(n, A, device="cuda", dtype=env.dtype)
    g = torch.Generator(device="cuda").manual_seed(0)
    for _ in range(warm):
        if sigma:
            a = torch.randn(n, A, device="cuda", generator=g, dtype=env.dtype) * sigma
        env.step(a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    acts = [torch.randn(n, A, device="cuda", generator=g, dtype=env.dtype) * sigma for _ in range(steps)]
    e0.record()
    for k in range(steps):
        env.step(acts[k])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    iters = env.solver_iterations().float().mean().item()
    env.close()
    return ms, n / ms * 1e3, iters


if __name__ == "__main__":
    wpb = os.environ.get("LHW_WARPS_PER_BLOCK", "default")
    models = sys.argv[1].split(",") if len(sys.argv) > 1 else ["jvrc_walk"]
    sizes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [4096, 16384, 32768]
    for model in models:
      for precision in (32, 64):
        for n in sizes:
            for sigma in (0.0, 0.223):
                ms, sps, it = run(n, precision, sigma=sigma, model=model)
                print(f"{model} wpb={wpb} fp{precision} N={n} sigma={sigma}: {ms:.3f} ms/step  {sps/1e6:.3f} M env-steps/s  newton-iters/step/env={it:.1f}", flush=True)
