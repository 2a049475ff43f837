"""BatchedHumanoidEnv — N device-resident copies of a reference humanoid env stepped by one CUDA launch.

Mirrors, per environment, the reference's env protocol (envs/common/base_humanoid_env.py:177-276,
envs/jvrc/jvrc_walk.py): `reset() -> obs`, `step(actions) -> (obs, reward, done, info)`, `observation_space`,
`action_space`, `obs_mean`, `obs_std`, `robot.{mirrored_obs, mirrored_acts, clock_inds, iteration_count}`.
All state lives in two torch CUDA tensors (the HBM records the kernel streams); nothing touches the host
inside a control step.
"""
from __future__ import annotations

import ctypes
from types import SimpleNamespace

import numpy as np
import torch

from .. import _lib
from ..model import load_model, pack_model
from ..model.loader import curriculum_height

REWARD_NAMES = ("foot_frc_score", "foot_vel_score", "root_accel", "height_error", "com_vel_error", "yaw_vel_error",
                "upper_body_reward", "posture_error", "torque_penalty", "action_penalty")  # tasks/walking_task.py:131-146
STAND_REWARD_NAMES = ("com_vel_error", "yaw_vel_error", "height", "upperbody", "joint_torque_reward",
                      "posture")                                                          # tasks/standing_task.py:97-104
STEP_REWARD_NAMES = ("foot_frc_score", "foot_vel_score", "orient_cost", "height_error", "step_reward",
                     "upper_body_reward")                                                 # tasks/stepping_task.py:107-120
N_REWARD_SLOTS = 10   # width of the kernel's reward-term record (csrc/sim_core.h NREW); unused slots are 0


class _Robot(SimpleNamespace):
    """The attributes the trainer reads / writes on `env.robot` (robots/robot_base.py:35, run_experiment.py:123-125).
    Writing `iteration_count` (rl/workers/rollout_worker.py:95) drives the SteppingTask's height curriculum."""

    def __setattr__(self, name, value):
        super().__setattr__(name, value)
        if name == "iteration_count" and getattr(self, "_on_iteration", None) is not None:
            self._on_iteration(value)


class BatchedHumanoidEnv:
    def __init__(self, num_envs: int, model: str = "jvrc_walk", precision: int = 32, seed: int = 0,
                 first_env_id: int = 0, device: int | torch.device | None = None, max_traj_len: int = 400,
                 tolerance: float | None = None, max_iter: int | None = None, observation_noise: bool = True,
                 domain_randomization: bool = True, init_noise: bool = True, pd_gain_randomization: float = 0.0,
                 iteration_count: float = float("inf"), path_to_yaml=None):
        if not torch.cuda.is_available():
            raise _lib.LhwError("BatchedHumanoidEnv needs a CUDA device (no CPU fallback on the rollout path)")
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self.num_envs, self.precision, self.seed, self.first_env_id = int(num_envs), int(precision), int(seed), int(first_env_id)
        self.max_traj_len = int(max_traj_len)
        self.dtype = torch.float64 if precision == 64 else torch.float32
        self.model_name = model
        self.mj = load_model(model)
        if path_to_yaml is not None:
            # partial(Env, path_to_yaml) (run_experiment.py:115): the user's YAML over the compiled model's cfg block
            from .config import apply_config, load_yaml
            self.mj = apply_config(self.mj, load_yaml(path_to_yaml))
        self._iteration_count = float(iteration_count)
        if tolerance is None and precision == 32:
            tolerance = 1e-6  # fp32 cannot reach the reference's 1e-10; gradient floor is ~1e-6 of the force scale
        flat = pack_model(self.mj, tolerance=tolerance, max_iter=max_iter, observation_noise=observation_noise,
                          domain_randomization=domain_randomization, init_noise=init_noise,
                          pd_gain_randomization=pd_gain_randomization, iteration_count=iteration_count)
        L = _lib.lib()
        h = ctypes.c_void_p()
        _lib.check(L.lhw_sim_create(ctypes.byref(h), flat.ctypes.data_as(ctypes.c_void_p), len(flat), self.precision,
                                    self.device.index), "lhw_sim_create")
        self._h = h
        self.obs_dim, self.act_dim = L.lhw_sim_obs_dim(h), L.lhw_sim_act_dim(h)
        n, dev = self.num_envs, self.device
        self.state_r = torch.zeros(n, L.lhw_sim_state_reals(h), dtype=self.dtype, device=dev)
        self.state_i = torch.zeros(n, L.lhw_sim_state_ints(h), dtype=torch.int32, device=dev)
        self.obs = torch.zeros(n, self.obs_dim, dtype=self.dtype, device=dev)
        self.term_obs = torch.zeros(n, self.obs_dim, dtype=self.dtype, device=dev)
        self.reward = torch.zeros(n, dtype=self.dtype, device=dev)
        self.rew_terms = torch.zeros(n, N_REWARD_SLOTS, dtype=self.dtype, device=dev)
        self.done = torch.zeros(n, dtype=torch.int32, device=dev)
        self.ended = torch.zeros(n, dtype=torch.int32, device=dev)
        self.ep_len = torch.zeros(n, dtype=torch.int32, device=dev)
        self.ep_rew = torch.zeros(n, dtype=self.dtype, device=dev)
        self._fresh = True
        # ---- reference-facing attributes (envs/jvrc/jvrc_base.py:69-131, envs/jvrc/jvrc_walk.py:43-63)
        cfg = self.mj["cfg"]
        self.history_len = cfg["obs_history_len"]
        self.base_obs_len = self.obs_dim
        self.dt = cfg["control_dt"]
        self.action_space = np.zeros(self.act_dim)
        self.observation_space = np.zeros(self.obs_dim * self.history_len)
        self.nq, self.nv = 7 + self.act_dim, 6 + self.act_dim
        self._setup_reference_attributes(model, cfg)
        # env.interface / env.task / env.model / env.data of the reference's env protocol, as views of env 0
        from ..tasks.descriptors import DataView, DeviceRobotInterface, make_task, model_view
        self.interface = DeviceRobotInterface(self, 0)
        self.task = make_task(self, 0)
        self.model, self.data = model_view(self), DataView(self, 0)

    def _setup_reference_attributes(self, model: str, cfg: dict) -> None:
        """obs_mean / obs_std, reward names and the robot's mirror lists exactly as the reference env classes set them
        (pure numpy; pinned to the reference's own method bodies by tests/golden/env_attributes.json)."""
        if model == "h1":
            # envs/h1/h1_env.py:37-55 (normalisation), envs/h1/h1_base.py:66-76 ; no mirror lists on the H1 robot
            self.reward_names = STAND_REWARD_NAMES
            half = np.asarray(cfg["half_sitting_pose"], dtype=float)
            self.obs_mean = np.concatenate((np.zeros(5), half, np.zeros(10), np.zeros(10)))
            self.obs_std = np.concatenate(([0.2, 0.2, 1, 1, 1], 0.5 * np.ones(10), 4 * np.ones(10), 100 * np.ones(10)))
            self.robot = SimpleNamespace(iteration_count=np.inf)
            return
        self.reward_names = REWARD_NAMES
        half = np.deg2rad(cfg["half_sitting_pose_deg"])
        self.obs_mean = np.concatenate((np.zeros(5), half, np.zeros(12), [0, 0, 0.5, 0.5, 0.5, 0, 0, 0]))
        self.obs_std = np.concatenate(([0.2, 0.2, 1, 1, 1], 0.5 * np.ones(12), 4 * np.ones(12), [1, 1, 1, 1, 1, 0.5, 0.5, 0.5]))
        base_mir_obs = [-0.1, 1, -2, 3, -4, 11, -12, -13, 14, -15, 16, 5, -6, -7, 8, -9, 10,
                        23, -24, -25, 26, -27, 28, 17, -18, -19, 20, -21, 22]
        append_obs = [len(base_mir_obs) + i for i in range(8)]
        if model not in ("jvrc_walk", "jvrc_walk_terrain", "jvrc_step"):
            raise ValueError(f"unknown model {model!r}")
        if model == "jvrc_step":
            # envs/jvrc/jvrc_step.py:41-63: clock(2) + goal steps x(2) y(2) z(2) theta(2)
            self.reward_names = STEP_REWARD_NAMES
            self.obs_mean = np.concatenate((np.zeros(5), half, np.zeros(12), [0.5, 0.5], np.zeros(8)))
            self.obs_std = np.concatenate(([0.2, 0.2, 1, 1, 1], 0.5 * np.ones(12), 4 * np.ones(12), [1, 1], np.ones(8)))
            append_obs = [len(base_mir_obs) + i for i in range(10)]
        self.robot = _Robot(mirrored_obs=base_mir_obs + append_obs,
                            mirrored_acts=[6, -7, -8, 9, -10, 11, 0.1, -1, -2, 3, -4, 5],
                            clock_inds=append_obs[0:2], iteration_count=self._iteration_count, _on_iteration=None)
        if model == "jvrc_step":
            self.robot._on_iteration = self._set_iteration_count

    def _set_iteration_count(self, iteration_count: float) -> None:
        """SteppingTask.reset(iter_count) (tasks/stepping_task.py:262-264, 312): the curriculum's step height for the
        episodes that start from now on."""
        _lib.check(_lib.lib().lhw_sim_set_step_height(self._h, float(curriculum_height(iteration_count))),
                   "lhw_sim_set_step_height")

    # ------------------------------------------------------------------ device API (torch tensors in / out)
    def reset(self, mask: torch.Tensor | None = None) -> torch.Tensor:
        """Reset all envs (or those where mask != 0); returns the [N, obs_dim] observation tensor."""
        L = _lib.lib()
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.int32).contiguous()
        if _lib.use_torch_ops():
            _lib.ops().sim_reset(self._h.value, self.state_r, self.state_i, self.seed, self.first_env_id, mask,
                                 bool(self._fresh and mask is None), self.obs)
        else:
            with torch.cuda.device(self.device):
                _lib.check(L.lhw_sim_reset(self._h, self.state_r.data_ptr(), self.state_i.data_ptr(), self.num_envs, self.seed,
                                           self.first_env_id, _lib.ptr(mask), int(self._fresh and mask is None),
                                           self.obs.data_ptr(), _lib.current_stream_ptr()), "lhw_sim_reset")
        if mask is None:
            self._fresh = False
        return self.obs

    def step(self, actions: torch.Tensor, autoreset: bool = True):
        """actions [N, act_dim] (device, env dtype). Returns (obs, reward, done, ended) device tensors.
        With autoreset the RolloutWorker semantics apply (see include/lhw_b200.h: lhw_sim_step)."""
        if actions.dtype != self.dtype or not actions.is_contiguous() or actions.device != self.device:
            actions = actions.to(device=self.device, dtype=self.dtype).contiguous()
        if _lib.use_torch_ops():     # torch.ops.lhw.sim_step: device / dtype / shape checks, then the C-ABI launch
            _lib.ops().sim_step(self._h.value, self.state_r, self.state_i, self.seed, self.first_env_id, actions, self.max_traj_len,
                                bool(autoreset), self.obs, self.term_obs, self.reward, self.rew_terms, self.done, self.ended,
                                self.ep_len, self.ep_rew)
            return self.obs, self.reward, self.done, self.ended
        assert actions.shape == (self.num_envs, self.act_dim), actions.shape
        L = _lib.lib()
        with torch.cuda.device(self.device):
            _lib.check(L.lhw_sim_step(self._h, self.state_r.data_ptr(), self.state_i.data_ptr(), self.num_envs, self.seed,
                                      self.first_env_id, actions.data_ptr(), self.max_traj_len, int(autoreset),
                                      self.obs.data_ptr(), self.term_obs.data_ptr(), self.reward.data_ptr(),
                                      self.rew_terms.data_ptr(), self.done.data_ptr(), self.ended.data_ptr(),
                                      self.ep_len.data_ptr(), self.ep_rew.data_ptr(), _lib.current_stream_ptr()),
                       "lhw_sim_step")
        return self.obs, self.reward, self.done, self.ended

    def step_slice(self, lo: int, hi: int, actions: torch.Tensor, autoreset: bool = True):
        """One control step of environments [lo, hi) only (their records, outputs and Philox ids are the same as in a full-batch
        step: an environment's trajectory does not depend on what it is launched with).  The rollout worker advances two halves
        of the batch on two streams with this, so that the tail of one half's launch overlaps the other half's work."""
        sl = slice(lo, hi)
        if actions.dtype != self.dtype or not actions.is_contiguous() or actions.device != self.device:
            actions = actions.to(device=self.device, dtype=self.dtype).contiguous()
        out = (self.obs[sl], self.reward[sl], self.done[sl], self.ended[sl])
        if _lib.use_torch_ops():
            _lib.ops().sim_step(self._h.value, self.state_r[sl], self.state_i[sl], self.seed, self.first_env_id + lo, actions,
                                self.max_traj_len, bool(autoreset), out[0], self.term_obs[sl], out[1], self.rew_terms[sl], out[2], out[3],
                                self.ep_len[sl], self.ep_rew[sl])
            return out
        assert actions.shape == (hi - lo, self.act_dim), actions.shape
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().lhw_sim_step(self._h, self.state_r[sl].data_ptr(), self.state_i[sl].data_ptr(), hi - lo, self.seed,
                                               self.first_env_id + lo, actions.data_ptr(), self.max_traj_len, int(autoreset),
                                               out[0].data_ptr(), self.term_obs[sl].data_ptr(), out[1].data_ptr(),
                                               self.rew_terms[sl].data_ptr(), out[2].data_ptr(), out[3].data_ptr(),
                                               self.ep_len[sl].data_ptr(), self.ep_rew[sl].data_ptr(), _lib.current_stream_ptr()),
                       "lhw_sim_step")
        return out

    def bind(self):
        """Upload this env's model constants if another env of the same precision used the constant bank last
        (needed before replaying CUDA graphs that contain lhw_sim_step launches)."""
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().lhw_sim_bind(self._h, _lib.current_stream_ptr()), "lhw_sim_bind")

    # ------------------------------------------------------------------ host API (numpy in / out, copies inside)
    def step_host(self, actions: np.ndarray, autoreset: bool = True):
        """Reference-facing batched call with HOST buffers: H2D actions, one launch, D2H obs/reward/done."""
        a = torch.from_numpy(np.ascontiguousarray(actions, dtype=np.float64 if self.precision == 64 else np.float32))
        obs, rew, done, ended = self.step(a.to(self.device, non_blocking=True), autoreset)
        return obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy().astype(bool), ended.cpu().numpy().astype(bool)

    # ------------------------------------------------------------------ state access (tests / checkpointing)
    @property
    def qpos(self) -> torch.Tensor:
        return self.state_r[:, 0:self.nq]

    @property
    def qvel(self) -> torch.Tensor:
        return self.state_r[:, self.nq:self.nq + self.nv]

    def solver_iterations(self) -> torch.Tensor:
        """Newton iterations spent in the last launch, per env."""
        return self.state_i[:, 7]

    def status_flags(self) -> torch.Tensor:
        """Per-env status word of the last launch (1: non-finite / diverging acceleration seen, forced reset — the analogue
        of MuJoCo's mj_checkAcc auto-reset; 2: a Cholesky pivot had to be clamped)."""
        return self.state_i[:, 6]

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().lhw_sim_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
