"""Mirror-symmetry helpers (rl/envs/wrappers.py:26-85 SymmetricEnv): signed permutation matrices built from
the env's mirrored_obs / mirrored_acts index lists, and the clock-aware observation mirror used by the PPO
mirror loss (rl/algos/ppo.py:347-358)."""
from __future__ import annotations

import numpy as np
import torch


def symmetry_matrix(mirrored) -> torch.Tensor:
    m = np.asarray(mirrored, dtype=float)
    n = len(m)
    mat = np.zeros((n, n))
    mat[np.arange(n), np.abs(m).astype(int)] = np.sign(m)
    return torch.tensor(mat, dtype=torch.float32)


class SymmetricEnv:
    def __init__(self, env_fn, mirrored_obs=None, mirrored_act=None, clock_inds=None):
        assert mirrored_obs and mirrored_act, "mirror index lists are required"
        self.act_mirror_matrix = symmetry_matrix(mirrored_act)
        self.obs_mirror_matrix = symmetry_matrix(mirrored_obs)
        self.clock_inds = list(clock_inds or [])
        self.env = env_fn()

    def __getattr__(self, attr):
        return getattr(self.env, attr)

    def mirror_action(self, action):
        return action @ self.act_mirror_matrix.to(action.device)

    def mirror_observation(self, obs):
        return obs @ self.obs_mirror_matrix.to(obs.device)

    def mirror_clock_observation(self, obs):
        """obs @ P then shift the phase clock by pi.  The reference writes sin(arcsin(c) + pi) per clock entry
        (rl/envs/wrappers.py:64-75), which is -c; the negation is used directly (also NaN-free for |c| = 1 + ulp)."""
        out = obs @ self.obs_mirror_matrix.to(obs.device)
        if self.clock_inds:
            out[:, self.clock_inds] = -out[:, self.clock_inds]
        return out
