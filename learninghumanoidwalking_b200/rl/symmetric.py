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
        # obs @ P followed by the clock negation is obs @ (P diag(s)), s = -1 on the clock columns: P has one +-1 per column,
        # so folding the sign into the matrix gives bit-identical results with one GEMM and no index kernels
        sign = torch.ones(self.obs_mirror_matrix.shape[1])
        sign[self.clock_inds] = -1.0
        self.obs_mirror_clock_matrix = self.obs_mirror_matrix * sign
        self._dev = {}
        self.env = env_fn()

    def __getattr__(self, attr):
        return getattr(self.env, attr)

    def _on(self, name: str, device) -> torch.Tensor:
        """The matrix `name` on `device`, uploaded once (the reference re-uploads it in every loss evaluation)."""
        key = (name, str(device))
        m = self._dev.get(key)
        if m is None:
            m = self._dev[key] = getattr(self, name).to(device)
        return m

    def mirror_action(self, action):
        return action @ self._on("act_mirror_matrix", action.device)

    def mirror_observation(self, obs):
        return obs @ self._on("obs_mirror_matrix", obs.device)

    def mirror_clock_observation(self, obs):
        """obs @ P then shift the phase clock by pi.  The reference writes sin(arcsin(c) + pi) per clock entry
        (rl/envs/wrappers.py:64-75), which is -c; the negation is folded into the matrix (also NaN-free for |c| = 1 + ulp)."""
        return obs @ self._on("obs_mirror_clock_matrix", obs.device)
