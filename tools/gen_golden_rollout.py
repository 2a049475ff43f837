#!/usr/bin/env python3
"""Generate tests/golden/rollout_worker.json by RUNNING the reference's RolloutWorker.sample (rl/workers/rollout_worker.py:98-199,
imported behind a `ray` stub: `ray.remote` becomes the identity) on a scripted duck-typed environment.

What it pins: the sampling semantics that the device rollout (`rl/workers.py`, `lhw_gae`) and its oracle
(`oracle/ppo_oracle.py:gae_rollout`) restate — an episode ends on done OR when the trajectory reaches max_traj_len; the path
is closed with (not done) * critic(next_state) taken BEFORE the reset; a path still open when the buffer fills is closed with
critic(current state) and the episode continues in the next call; only completed episodes report ep_lens / ep_rewards.
The environment is a script (rewards and terminations from closed formulas), the actor / critic tiny linear maps.
"""
import json
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("LHW_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
DONE_AT = {6, 19, 20, 37}          # global step numbers (1-based) whose transition terminates the episode


class ScriptEnv:
    """Observation = [k / 50, sin(0.3 k), episode / 10] with k the global step counter."""

    def __init__(self):
        self.k, self.ep = 0, 0
        self.robot = types.SimpleNamespace(iteration_count=0)
        self.log = []

    def _obs(self):
        return np.array([self.k / 50.0, np.sin(0.3 * self.k), self.ep / 10.0])

    def reset(self):
        self.ep += 1
        return self._obs()

    def step(self, action):
        self.k += 1
        rew = 0.1 * (self.k % 7) + 0.01 * float(np.sum(action))
        done = self.k in DONE_AT
        obs = self._obs()
        self.log.append(dict(k=self.k, reward=rew, done=bool(done), next_obs=obs.tolist()))
        return obs, rew, done, {}


class Lin(torch.nn.Module):
    def __init__(self, w, b, out):
        super().__init__()
        self.w, self.b, self.state_dim, self.action_dim = torch.tensor(w, dtype=torch.float32), b, 3, out

    def forward(self, state, deterministic=True):
        return (state.float() @ self.w + self.b).reshape(-1)


def main():
    sys.path.insert(0, REF)
    ray = types.ModuleType("ray")
    ray.remote = lambda *a, **k: (a[0] if a and callable(a[0]) else (lambda c: c))
    ray.get = ray.put = ray.init = ray.is_initialized = lambda *a, **k: None
    sys.modules["ray"] = ray
    from rl.workers.rollout_worker import RolloutWorker
    policy = Lin([[0.3, -0.2], [0.1, 0.4], [-0.5, 0.2]], 0.05, 2)
    critic = Lin([[0.7], [-0.3], [0.9]], 0.2, 1)
    worker = RolloutWorker(ScriptEnv, policy, critic, seed=3, worker_id=0)
    gamma, lam, T, max_traj_len = 0.99, 0.95, 16, 9
    calls = []
    for _ in range(3):
        n0 = len(worker.env.log)
        data = worker.sample(gamma, lam, T, max_traj_len, deterministic=True)
        steps = worker.env.log[n0:]
        assert len(steps) == T
        calls.append(dict(states=data.states.tolist(), rewards=data.rewards.reshape(-1).tolist(),
                          values=data.values.reshape(-1).tolist(), returns=data.returns.reshape(-1).tolist(),
                          dones=data.dones.reshape(-1).tolist(), ep_lens=[int(x) for x in data.ep_lens],
                          ep_rewards=[float(x) for x in data.ep_rewards], traj_idx=[int(x) for x in data.traj_idx],
                          env_done=[s["done"] for s in steps], next_obs=[s["next_obs"] for s in steps],
                          open_state=None if worker.current_state is None else worker.current_state.tolist()))
    json.dump(dict(gamma=gamma, lam=lam, T=T, max_traj_len=max_traj_len, critic_w=[[0.7], [-0.3], [0.9]], critic_b=0.2,
                   calls=calls), open(os.path.join(OUT, "rollout_worker.json"), "w"))
    print("wrote rollout_worker.json;", [c["ep_lens"] for c in calls], [c["traj_idx"] for c in calls])


if __name__ == "__main__":
    main()
