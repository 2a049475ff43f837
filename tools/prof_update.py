#!/usr/bin/env python3
"""Is the PPO optimisation phase launch/host bound?  Wall time vs summed CUDA kernel time of N minibatch updates."""
import os, sys, time
from types import SimpleNamespace
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learninghumanoidwalking_b200.envs import BatchedHumanoidEnv
from learninghumanoidwalking_b200.rl import PPO
from learninghumanoidwalking_b200.rl.symmetric import SymmetricEnv

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
base = lambda: BatchedHumanoidEnv(4096, precision=32, seed=0)
probe = base(); r = probe.robot; probe.close()
env_fn = lambda: SymmetricEnv(base, mirrored_obs=r.mirrored_obs, mirrored_act=r.mirrored_acts, clock_inds=r.clock_inds)
args = SimpleNamespace(gamma=0.99, lam=0.95, lr=3e-4, eps=1e-5, entropy_coeff=0.0, clip=0.2, minibatch_size=mb, epochs=1,
                       max_traj_len=400, num_procs=4096, max_grad_norm=0.05, mirror_coeff=0.4, eval_freq=1000, recurrent=False,
                       imitate_coeff=0.0, std_dev=0.223, learn_std=False, logdir="/tmp/lhw_prof", steps_per_env=16)
ppo = PPO(env_fn, args, seed=0)
ppo.make_optimizers()
batch = ppo.sample_parallel_with_workers()
adv = ppo.normalize_advantages(batch.returns.contiguous(), batch.values.contiguous())
env = ppo.env
idx = torch.arange(mb, device=ppo.device)
def one():
    ob, ab, rb, db = ppo.gather_minibatch(batch.states, batch.actions, batch.returns.contiguous(), adv, idx)
    return ppo.update_actor_critic(ob, ab, rb, db, 1, mirror_observation=env.mirror_clock_observation, mirror_action=env.mirror_action)
for _ in range(5): one()
torch.cuda.synchronize()
N = 30
t0 = time.time()
for _ in range(N): one()
torch.cuda.synchronize()
wall = (time.time() - t0) / N
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(N): one()
    torch.cuda.synchronize()
ev = prof.key_averages()
cuda_us = sum(getattr(e, "self_device_time_total", getattr(e, "self_cuda_time_total", 0)) for e in ev)
nk = sum(e.count for e in ev if getattr(e, "self_device_time_total", getattr(e, "self_cuda_time_total", 0)) > 0)
def one_g():
    ob, ab, rb, db = ppo.gather_minibatch(batch.states, batch.actions, batch.returns.contiguous(), adv, idx)
    return ppo._update_step(ob, ab, rb, db, env.mirror_clock_observation, env.mirror_action)
for _ in range(6): one_g()
torch.cuda.synchronize()
t0 = time.time()
for _ in range(N): one_g()
torch.cuda.synchronize()
wall_g = (time.time() - t0) / N
print(f"minibatch {mb}: graph-replayed update {wall_g*1e3:.3f} ms/update (captured: {ppo._ug is not None})")
print(f"minibatch {mb}: wall {wall*1e3:.3f} ms/update, summed CUDA kernel time {cuda_us/N/1e3:.3f} ms/update, ~{nk/N:.0f} kernels/update")
rows = sorted(((getattr(e, "self_device_time_total", getattr(e, "self_cuda_time_total", 0)), e.count, e.key) for e in ev), reverse=True)
print("top kernels by device time (us per update, launches per update, name):")
for us, cnt, key in rows[:14]:
    if us > 0:
        print(f"  {us / N:9.1f} {cnt / N:6.1f}  {key[:110]}")
