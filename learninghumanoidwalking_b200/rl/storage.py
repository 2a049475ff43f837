"""Rollout containers.

`BatchData` is the typed hand-off between sampler and learner with the reference's nine fields
(rl/storage/rollout_storage.py:7-22); here every tensor is a CUDA float32 tensor.
`DeviceRolloutBuffer` replaces PPOBuffer (rl/storage/rollout_storage.py:25-107): time-major [T, N, .] device
tensors written one row per control step, GAE(lambda) by one CUDA launch (lhw_gae) instead of a Python loop
per trajectory.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from .. import _lib


@dataclass
class BatchData:
    states: torch.Tensor
    actions: torch.Tensor
    rewards: torch.Tensor
    values: torch.Tensor
    returns: torch.Tensor
    dones: torch.Tensor
    traj_idx: torch.Tensor
    ep_lens: torch.Tensor
    ep_rewards: torch.Tensor


class DeviceRolloutBuffer:
    def __init__(self, T: int, N: int, obs_dim: int, act_dim: int, device, gamma: float = 0.99, lam: float = 0.95):
        f32 = dict(dtype=torch.float32, device=device)
        self.T, self.N, self.gamma, self.lam = T, N, gamma, lam
        self.states = torch.zeros(T, N, obs_dim, **f32)
        self.actions = torch.zeros(T, N, act_dim, **f32)
        self.rewards = torch.zeros(T, N, **f32)
        self.values = torch.zeros(T, N, **f32)
        self.returns = torch.zeros(T, N, **f32)
        self.boot = torch.zeros(T, N, **f32)          # (not done) * critic(next_state) where the episode ended
        self.ended = torch.zeros(T, N, dtype=torch.int32, device=device)   # done or truncated ("dones" in the reference)
        self.ep_len = torch.zeros(T, N, dtype=torch.int32, device=device)
        self.ep_rew = torch.zeros(T, N, **f32)
        self.last_val = torch.zeros(N, **f32)
        # per-block (sum, sumsq) of returns - values, left behind by the GAE launch for the advantage normalisation
        self.adv_partials = torch.zeros(_lib.lib().lhw_gae_partial_words(N), dtype=torch.float64, device=device)
        self.generation = 0     # bumped by every finish(): tells the learner whether adv_partials belong to the batch it holds

    def finish(self):
        """PPOBuffer.finish_path for every path of every env in one launch."""
        _lib.ops().gae(self.rewards, self.values, self.ended, self.boot, self.last_val, self.returns, self.gamma, self.lam,
                       self.adv_partials)
        self.generation += 1

    def get_data(self, env_major: bool = True) -> BatchData:
        """Flatten to [N*T, .].  env_major=True reproduces the reference's ordering (torch.cat over workers:
        env index major, time minor, rl/algos/ppo.py:263-270)."""
        def flat(x):
            if env_major:
                x = x.transpose(0, 1)
            return x.reshape(self.T * self.N, *x.shape[2:])
        mask = flat(self.ended).bool()
        return BatchData(states=flat(self.states), actions=flat(self.actions), rewards=flat(self.rewards).unsqueeze(-1),
                         values=flat(self.values).unsqueeze(-1), returns=flat(self.returns).unsqueeze(-1),
                         dones=flat(self.ended).float().unsqueeze(-1), traj_idx=torch.empty(0, dtype=torch.long),
                         ep_lens=flat(self.ep_len)[mask], ep_rewards=flat(self.ep_rew)[mask])
