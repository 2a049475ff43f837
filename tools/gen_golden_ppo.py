#!/usr/bin/env python3
"""Generate tests/golden/ppo_update.json by RUNNING the reference's PPO.update_actor_critic (rl/algos/ppo.py:299-406) on the CPU.

`rl/algos/ppo.py` imports `ray` at module level (not installable here); a stub module object lets the import go through — the
method itself touches no Ray API.  The method is called unbound on a small attribute holder with the reference's own
Gaussian_FF_Actor / FF_V (37 -> 256 -> 256 -> 12 / 1), torch Adam optimisers created as rl/algos/ppo.py:429-430 does, the jvrc
mirror functions of the reference's SymmetricEnv (rl/envs/wrappers.py), and a fixed minibatch.  Network weights are written
from a closed formula (so the test can rebuild them without shipping 154 k floats).  Recorded: the 7 returned scalars of two
consecutive updates and checksums / leading entries of every parameter tensor afterwards.
"""
import json
import os
import sys
import types
from copy import deepcopy

import numpy as np
import torch

REF = os.environ.get("LHW_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
MIRRORED_OBS = [-0.1, 1, -2, 3, -4, 11, -12, -13, 14, -15, 16, 5, -6, -7, 8, -9, 10,
                23, -24, -25, 26, -27, 28, 17, -18, -19, 20, -21, 22] + list(range(29, 37))
MIRRORED_ACT = [6, -7, -8, 9, -10, 11, 0.1, -1, -2, 3, -4, 5]


def formula_weights_(module, salt):
    """p.flat[i] = scale * sin(0.37 i + k + salt), scale ~ 1/sqrt(fan_in): deterministic, well conditioned, reproducible."""
    with torch.no_grad():
        for k, p in enumerate(module.parameters()):
            n = p.numel()
            scale = 1.0 / np.sqrt(p.shape[-1]) if p.dim() > 1 else 0.05
            p.copy_((scale * torch.sin(0.37 * torch.arange(n, dtype=torch.float64) + k + salt)).float().view_as(p))


def batch(B=64, obs_dim=37, act_dim=12):
    i = torch.arange(B, dtype=torch.float64).unsqueeze(1)
    obs = (0.8 * torch.sin(0.11 * i + 0.7 * torch.arange(obs_dim, dtype=torch.float64))).float()
    obs[:, 29] = torch.sin(0.3 * i[:, 0]).float()
    obs[:, 30] = torch.cos(0.3 * i[:, 0]).float()
    act = (0.4 * torch.cos(0.13 * i + 0.5 * torch.arange(act_dim, dtype=torch.float64))).float()
    ret = (1.5 + torch.sin(0.21 * i)).float()
    adv = (torch.cos(0.17 * i) * 1.2).float()
    return obs, act, ret, adv


def main():
    sys.path.insert(0, REF)
    ray = types.ModuleType("ray")
    ray.remote = lambda *a, **k: (a[0] if a and callable(a[0]) else (lambda c: c))
    ray.get = ray.put = ray.init = ray.is_initialized = lambda *a, **k: None
    sys.modules["ray"] = ray
    import rl.algos.ppo as refppo
    from rl.envs.wrappers import SymmetricEnv
    from rl.policies.actor import Gaussian_FF_Actor
    from rl.policies.critic import FF_V

    torch.manual_seed(0)
    policy = Gaussian_FF_Actor(37, 12, init_std=0.223, learn_std=False, bounded=False)
    critic = FF_V(37)
    formula_weights_(policy, 0.0)
    formula_weights_(critic, 3.0)
    old_policy = deepcopy(policy)
    formula_weights_(policy, 0.02)          # the current policy has moved a little away from the old one: ratio != 1
    obs_mean = 0.1 * torch.sin(torch.arange(37, dtype=torch.float32))
    obs_std = 1.0 + 0.5 * torch.cos(torch.arange(37, dtype=torch.float32)) ** 2
    for net in (policy, old_policy, critic):
        net.obs_mean, net.obs_std = obs_mean, obs_std

    sym = SymmetricEnv(lambda: types.SimpleNamespace(base_obs_len=37), mirrored_obs=MIRRORED_OBS, mirrored_act=MIRRORED_ACT,
                       clock_inds=[29, 30])       # the env itself is only asked for base_obs_len

    holder = types.SimpleNamespace(policy=policy, old_policy=old_policy, critic=critic, clip=0.2, ent_coeff=0.01,
                                   mirror_coeff=0.4, imitate_coeff=0.0, imitation_projector=None, base_policy=None,
                                   recurrent=False, grad_clip=0.05,
                                   actor_optimizer=torch.optim.Adam(policy.parameters(), lr=3e-4, eps=1e-5),
                                   critic_optimizer=torch.optim.Adam(critic.parameters(), lr=3e-4, eps=1e-5))
    obs, act, ret, adv = batch()
    steps = []
    for _ in range(2):
        out = refppo.PPO.update_actor_critic(holder, obs, act, ret, adv, 1, mirror_observation=sym.mirror_clock_observation,
                                             mirror_action=sym.mirror_action)
        steps.append([float(x) for x in out])

    def digest(module):
        return [dict(shape=list(p.shape), sum=float(p.double().sum()), abs=float(p.double().abs().sum()),
                     head=p.detach().reshape(-1)[:8].double().tolist()) for p in module.parameters()]
    json.dump(dict(names=["actor_loss", "entropy_penalty", "critic_loss", "approx_kl_div", "mirror_loss", "imitation_loss",
                          "clip_fraction"], steps=steps, actor=digest(policy), critic=digest(critic),
                   hyper=dict(clip=0.2, ent_coeff=0.01, mirror_coeff=0.4, grad_clip=0.05, lr=3e-4, eps=1e-5, B=64)),
              open(os.path.join(OUT, "ppo_update.json"), "w"))
    print("wrote ppo_update.json", steps)


if __name__ == "__main__":
    main()
