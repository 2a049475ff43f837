"""The learner seam against vectors produced by RUNNING the reference's PPO.update_actor_critic
(rl/algos/ppo.py:299-406) on the CPU (tools/gen_golden_ppo.py): same networks (weights from a closed formula), same minibatch,
same mirror functions, torch Adam optimisers made by hand as tests/test_training.py:140-141 does -> the 7 returned scalars of
two consecutive updates and every parameter tensor afterwards.  Runs on the CPU: the loss graph is torch autograd; what is
CUDA-only in this build (fused clip+Adam, the graph replay) is compared with THIS path by tests/test_gpu_ppo.py."""
import json
import os
from copy import deepcopy

import numpy as np
import torch

G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ppo_update.json")))


def _formula_weights_(module, salt):
    with torch.no_grad():
        for k, p in enumerate(module.parameters()):
            n = p.numel()
            scale = 1.0 / np.sqrt(p.shape[-1]) if p.dim() > 1 else 0.05
            p.copy_((scale * torch.sin(0.37 * torch.arange(n, dtype=torch.float64) + k + salt)).float().view_as(p))


def _batch(B=64, obs_dim=37, act_dim=12):
    i = torch.arange(B, dtype=torch.float64).unsqueeze(1)
    obs = (0.8 * torch.sin(0.11 * i + 0.7 * torch.arange(obs_dim, dtype=torch.float64))).float()
    obs[:, 29] = torch.sin(0.3 * i[:, 0]).float()
    obs[:, 30] = torch.cos(0.3 * i[:, 0]).float()
    act = (0.4 * torch.cos(0.13 * i + 0.5 * torch.arange(act_dim, dtype=torch.float64))).float()
    return obs, act, (1.5 + torch.sin(0.21 * i)).float(), (torch.cos(0.17 * i) * 1.2).float()


def test_update_actor_critic_matches_the_reference_method():
    from learninghumanoidwalking_b200.envs.batched_env import BatchedHumanoidEnv  # noqa: F401  (import only: no CUDA needed)
    from learninghumanoidwalking_b200.rl.policies import FF_V, Gaussian_FF_Actor
    from learninghumanoidwalking_b200.rl.ppo import PPO
    from learninghumanoidwalking_b200.rl.symmetric import SymmetricEnv
    h = G["hyper"]
    policy, critic = Gaussian_FF_Actor(37, 12, init_std=0.223, learn_std=False), FF_V(37)
    assert [list(p.shape) for p in policy.parameters()] == [d["shape"] for d in G["actor"]]     # same layer order as the reference
    assert [list(p.shape) for p in critic.parameters()] == [d["shape"] for d in G["critic"]]
    _formula_weights_(policy, 0.0)
    _formula_weights_(critic, 3.0)
    old_policy = deepcopy(policy)
    _formula_weights_(policy, 0.02)
    obs_mean = 0.1 * torch.sin(torch.arange(37, dtype=torch.float32))
    obs_std = 1.0 + 0.5 * torch.cos(torch.arange(37, dtype=torch.float32)) ** 2
    for net in (policy, old_policy, critic):
        net.obs_mean, net.obs_std = obs_mean, obs_std
    mirrored_obs = [-0.1, 1, -2, 3, -4, 11, -12, -13, 14, -15, 16, 5, -6, -7, 8, -9, 10,
                    23, -24, -25, 26, -27, 28, 17, -18, -19, 20, -21, 22] + list(range(29, 37))
    sym = SymmetricEnv(lambda: object(), mirrored_obs=mirrored_obs, mirrored_act=[6, -7, -8, 9, -10, 11, 0.1, -1, -2, 3, -4, 5],
                       clock_inds=[29, 30])
    ppo = PPO.__new__(PPO)      # the method under test only needs these attributes (hand-made torch optimisers: the reference's sequence)
    ppo.__dict__.update(policy=policy, old_policy=old_policy, critic=critic, clip=h["clip"], ent_coeff=h["ent_coeff"],
                        mirror_coeff=h["mirror_coeff"], imitate_coeff=0.0, grad_clip=h["grad_clip"], _comm=None, world=1,
                        actor_optimizer=torch.optim.Adam(policy.parameters(), lr=h["lr"], eps=h["eps"]),
                        critic_optimizer=torch.optim.Adam(critic.parameters(), lr=h["lr"], eps=h["eps"]))
    obs, act, ret, adv = _batch(h["B"])
    for ref in G["steps"]:
        out = ppo.update_actor_critic(obs, act, ret, adv, 1, mirror_observation=sym.mirror_clock_observation,
                                      mirror_action=sym.mirror_action)
        got = [float(x) for x in out]
        assert len(got) == 7
        for name, a, b in zip(G["names"], got, ref):
            assert abs(a - b) <= 2e-5 * max(1.0, abs(b)), (name, a, b)
    for module, digest in ((policy, G["actor"]), (critic, G["critic"])):
        for p, d in zip(module.parameters(), digest):
            assert abs(float(p.double().sum()) - d["sum"]) < 1e-4 * max(1.0, d["abs"] * 1e-2)
            assert abs(float(p.double().abs().sum()) - d["abs"]) < 1e-5 * max(1.0, d["abs"])
            assert np.abs(p.detach().reshape(-1)[:8].double().numpy() - np.array(d["head"])).max() < 2e-6


def test_minibatch_indices_are_the_reference_sampler():
    """rl/algos/ppo.py:504-517: SubsetRandomSampler(range(n), generator seeded seed + itr * epochs + epoch) under
    BatchSampler(drop_last=True).  PPO.minibatch_indices draws the same permutation in one call."""
    from torch.utils.data.sampler import BatchSampler, SubsetRandomSampler
    from learninghumanoidwalking_b200.rl.ppo import PPO
    ppo = PPO.__new__(PPO)
    ppo.__dict__.update(seed=11, epochs=3, minibatch_size=64, device=torch.device("cpu"))
    for itr, epoch, n in ((0, 0, 1000), (5, 2, 777), (40, 1, 64), (3, 0, 63)):
        g = torch.Generator()
        g.manual_seed(11 + itr * 3 + epoch)
        ref = [list(b) for b in BatchSampler(SubsetRandomSampler(range(n), generator=g), 64, drop_last=True)]
        got = ppo.minibatch_indices(n, itr, epoch)
        assert got.shape == (n // 64, 64) and got.tolist() == ref
