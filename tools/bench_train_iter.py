#!/usr/bin/env python3
"""One full PPO iteration the way the reference defines "fps" (rl/algos/ppo.py:587-595: samples / (sampling +
optimisation time)): N envs x T steps rollout on device, GAE, advantage normalisation, epochs x minibatches of
clip+Adam updates.  Prints one JSON object.  usage: bench_train_iter.py [envs] [steps_per_env] [minibatch] [precision]"""
import json
import os
import sys
import time
from types import SimpleNamespace

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learninghumanoidwalking_b200.envs import BatchedHumanoidEnv  # noqa: E402
from learninghumanoidwalking_b200.rl import PPO  # noqa: E402
from learninghumanoidwalking_b200.rl.symmetric import SymmetricEnv  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 400
mb = int(sys.argv[3]) if len(sys.argv) > 3 else 32768
prec = int(sys.argv[4]) if len(sys.argv) > 4 else 64
tf32 = len(sys.argv) > 5 and sys.argv[5] == "tf32"
base = lambda: BatchedHumanoidEnv(n, precision=prec, seed=0)
probe = base(); r = probe.robot; probe.close()
env_fn = lambda: SymmetricEnv(base, mirrored_obs=r.mirrored_obs, mirrored_act=r.mirrored_acts, clock_inds=r.clock_inds)
args = SimpleNamespace(gamma=0.99, lam=0.95, lr=3e-4, eps=1e-5, entropy_coeff=0.0, clip=0.2, minibatch_size=mb, epochs=3,
                       max_traj_len=400, num_procs=n, max_grad_norm=0.05, mirror_coeff=0.4, eval_freq=10**9, recurrent=False,
                       imitate_coeff=0.0, std_dev=0.223, learn_std=False, logdir="/tmp/lhw_bench_train", steps_per_env=T,
                       eval_batches=0, tf32=tf32, eval_at_start=False)     # time sampling + optimisation only: the evaluation pass of iteration 0 is five more batches
ppo = PPO(env_fn, args, seed=0)
ppo.train(None, 1, verbose=False)          # warm-up iteration (graph capture, cuBLAS heuristics)
torch.cuda.synchronize()
t0 = time.perf_counter()
log = ppo.train(None, 2, verbose=False)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(json.dumps({"tf32": tf32, "envs": n, "steps_per_env": T, "minibatch": mb, "epochs": 3, "precision": prec, "samples_per_iter": n * T,
                  "iter_s": dt / 2, "fps_sampling_plus_optimisation": 2 * n * T / dt, "sample_s": log[-1]["sample_time"],
                  "optimize_s": log[-1]["optimize_time"], "mean_ep_len": log[-1]["ep_len"], "mean_ep_rew": log[-1]["ep_rew"]}))
