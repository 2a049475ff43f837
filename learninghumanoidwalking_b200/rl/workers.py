"""DeviceRolloutWorker — replaces the N Ray actors of rl/workers/rollout_worker.py with one device-resident
vectorised sampler: N environments advance together, one CUDA launch per control step, policy / critic
inference batched on the same device (cuBLAS), transitions written straight into device rollout buffers.

Semantics kept from RolloutWorker.sample (rl/workers/rollout_worker.py:98-199), per environment:
  * exactly `max_steps` transitions per call; episodes persist across calls (state, traj_len, episode stats);
  * an episode ends on done or when traj_len reaches max_traj_len (truncation); `dones` stores done OR truncated;
  * bootstrap on an ended episode = (not done) * critic(next_state) with next_state the PRE-reset observation;
  * a path still open when the buffer fills bootstraps with critic(current state);
  * only completed episodes contribute ep_lens / ep_rewards.
"""
from __future__ import annotations

import torch

from .storage import BatchData, DeviceRolloutBuffer


class DeviceRolloutWorker:
    def __init__(self, env, policy, critic, seed: int | None = None, worker_id: int = 0):
        self.env, self.policy, self.critic = env, policy, critic
        self.worker_id = worker_id
        self.device = env.device
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(0 if seed is None else int(seed))
        self.current_state = None      # [N, obs_dim] float32; None before the first reset
        self._buf = None
        self.total_steps = 0

    def sync_state(self, policy_state_dict=None, critic_state_dict=None, obs_mean=None, obs_std=None, iteration_count=0):
        """The learner and the sampler share the same modules on the same device: nothing to ship.  Kept so the
        call site of rl/algos/ppo.py:238-242 works unchanged; state dicts are loaded if given."""
        if policy_state_dict is not None:
            self.policy.load_state_dict(policy_state_dict)
        if critic_state_dict is not None:
            self.critic.load_state_dict(critic_state_dict)
        if obs_mean is not None:
            self.policy.obs_mean = self.critic.obs_mean = obs_mean
        if obs_std is not None:
            self.policy.obs_std = self.critic.obs_std = obs_std
        self.env.robot.iteration_count = iteration_count

    @torch.no_grad()
    def sample(self, gamma, lam, max_steps, max_traj_len, deterministic=False, env_major=True) -> BatchData:
        env = self.env
        N, T = env.num_envs, int(max_steps)
        env.max_traj_len = int(max_traj_len)
        if self._buf is None or self._buf.T != T or self._buf.N != N:
            self._buf = DeviceRolloutBuffer(T, N, env.obs_dim, env.act_dim, self.device, gamma, lam)
        buf = self._buf
        buf.gamma, buf.lam = gamma, lam
        if self.current_state is None:
            self.current_state = env.reset().float().clone()
        state = self.current_state
        std = self.policy.stds
        for t in range(T):
            mu = self.policy(state, deterministic=True)
            if deterministic:
                action = mu
            else:
                action = mu + std * torch.randn(mu.shape, device=self.device, generator=self.gen)
            buf.states[t] = state
            buf.actions[t] = action
            buf.values[t] = self.critic(state).squeeze(-1)
            obs, reward, done, ended = env.step(action if env.dtype == torch.float32 else action.double())
            buf.rewards[t] = reward
            buf.ended[t] = ended
            buf.ep_len[t] = env.ep_len
            buf.ep_rew[t] = env.ep_rew
            # truncation bootstrap: (not done) * critic(pre-reset next_state) where the episode ended
            v_term = self.critic(env.term_obs.float()).squeeze(-1)
            buf.boot[t] = torch.where((ended != 0) & (done == 0), v_term, torch.zeros_like(v_term))
            state = obs.float().clone()
        buf.last_val.copy_(self.critic(state).squeeze(-1))
        buf.finish()
        self.current_state = state
        self.total_steps += T * N
        return buf.get_data(env_major=env_major)
