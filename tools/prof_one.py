#!/usr/bin/env python3
"""Run a few step launches of one configuration (for ncu captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learninghumanoidwalking_b200.envs import BatchedHumanoidEnv  # noqa: E402

precision, n, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
sigma = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
model = sys.argv[5] if len(sys.argv) > 5 else "jvrc_walk"
env = BatchedHumanoidEnv(n, model=model, precision=precision, seed=0)
env.reset()
g = torch.Generator(device="cuda").manual_seed(0)
for _ in range(steps):
    env.step(torch.randn(n, env.act_dim, device="cuda", generator=g, dtype=env.dtype) * sigma)
torch.cuda.synchronize()
print("done", env.solver_iterations().float().mean().item())
