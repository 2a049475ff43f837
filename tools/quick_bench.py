#!/usr/bin/env python3
"""Quick device-resident step-kernel timing sweep (development tool; bench.py is the contract)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learninghumanoidwalking_b200.envs import BatchedHumanoidEnv  # noqa: E402


def run(n, precision, steps=40, warm=10, sigma=0.0):
    env = BatchedHumanoidEnv(n, precision=precision, seed=0)
    env.reset()
    a = torch.zeros(n, 12, device="cuda", dtype=env.dtype)
    g = torch.Generator(device="cuda").manual_seed(0)
    for _ in range(warm):
        if sigma:
            a = torch.randn(n, 12, device="cuda", generator=g, dtype=env.dtype) * sigma
        env.step(a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    acts = [torch.randn(n, 12, device="cuda", generator=g, dtype=env.dtype) * sigma for _ in range(steps)]
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
    for precision in (32, 64):
        for n in (4096, 16384, 32768):
            for sigma in (0.0, 0.223):
                ms, sps, it = run(n, precision, sigma=sigma)
                print(f"wpb={wpb} fp{precision} N={n} sigma={sigma}: {ms:.3f} ms/step  {sps/1e6:.3f} M env-steps/s  newton-iters/step/env={it:.1f}", flush=True)
