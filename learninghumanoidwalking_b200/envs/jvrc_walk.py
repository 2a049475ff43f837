"""JvrcWalkEnv — single-environment view with the exact reference protocol
(envs/jvrc/jvrc_walk.py, tests/test_environments.py): numpy in, (obs f64, float, bool, dict) out.
It is a 1-env BatchedHumanoidEnv, i.e. it runs the same CUDA kernel; there is no CPU implementation."""
from __future__ import annotations

import numpy as np
import torch

from .batched_env import BatchedHumanoidEnv


class JvrcWalkEnv:
    MODEL = "jvrc_walk"

    def __init__(self, path_to_yaml: str | None = None, precision: int = 64, seed: int = 0, env_id: int = 0, device=None,
                 **env_kwargs):
        self._b = BatchedHumanoidEnv(1, self.MODEL, precision=precision, seed=seed, first_env_id=env_id, device=device,
                                     path_to_yaml=path_to_yaml, **env_kwargs)
        for k in ("observation_space", "action_space", "obs_mean", "obs_std", "robot", "history_len", "base_obs_len", "dt",
                  "interface", "task", "model", "data"):
            setattr(self, k, getattr(self._b, k))

    def reset(self) -> np.ndarray:
        return self._b.reset()[0].double().cpu().numpy()

    def step(self, action: np.ndarray):
        if not isinstance(action, np.ndarray):
            raise TypeError("Expected action to be a numpy array")  # robots/robot_base.py:65-66
        assert action.shape == (self._b.act_dim,), f"Action vector length expected to be: {self._b.act_dim} but is {action.shape}"
        a = torch.as_tensor(np.copy(action), dtype=self._b.dtype).reshape(1, -1)
        obs, rew, done, _ = self._b.step(a, autoreset=False)
        terms = self._b.rew_terms[0].double().cpu().numpy()
        info = {k: float(v) for k, v in zip(self._b.reward_names, terms)}
        return obs[0].double().cpu().numpy(), float(sum(info.values())), bool(done[0].item()), info

    def close(self):
        self._b.close()
