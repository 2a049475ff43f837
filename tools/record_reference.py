#!/usr/bin/env python3
"""record_reference.py — to be run on a host that HAS `mujoco==3.4.0` (+ the reference's other dependencies) installed.

This container and the GPU box do not, which is why DESIGN.md calls the dynamics oracle "parity unpinned" at the
MuJoCo boundary.  This script is the other half of the pinning protocol of SURVEY §8c: it drives the UNMODIFIED reference
environment open-loop with a seeded action sequence and dumps what the oracle needs to replay the same trajectory:

    python tools/record_reference.py --reference /path/to/LearningHumanoidWalking --env jvrc_walk --steps 1000 --seed 0 \
        --out tests/golden/mujoco_jvrc_walk.npz

  * the state right after `env.reset()` (qpos, qvel, qacc_warmstart, actuator_length / velocity, task phase / mode / mode_ref),
  * the action sequence a_t ~ N(0, 0.223^2) from numpy's RandomState(seed) (the action distribution of a fresh policy),
  * per control step: qpos, qvel, observation, the reward dict (in insertion order), done,
  * for the first `--substeps` control steps additionally per physics substep: qpos, qvel, ctrl, ncon, efc_force norm.

While recording, the task's random mode switches are disabled (np.random.randint is patched to never return 0 inside
task.step), because numpy's MT19937 draw order cannot be reproduced by the device's counter-based streams; everything else is
the stock code path.  `tests/test_mujoco_recording.py` replays any `tests/golden/mujoco_*.npz` found through the oracle and the
CUDA path and applies the north-star's bar (<= 1e-4 relative on qpos/qvel over the recorded horizon); without a recording it
skips with the reason.  NOT EXERCISED HERE (no mujoco wheel in this image) — written against the reference's API as read from
its sources: envs/common/mujoco_env.py:113-127, envs/common/base_humanoid_env.py:199-276, robots/robot_base.py:49-98,
envs/common/robot_interface.py:535-546.
"""
import argparse
import os
import sys

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True, help="checkout of rohanpsingh/LearningHumanoidWalking")
    ap.add_argument("--env", default="jvrc_walk", choices=["jvrc_walk", "jvrc_step", "h1"])
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--substeps", type=int, default=4, help="control steps recorded at substep resolution")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--sigma", type=float, default=0.223)
    ap.add_argument("--out", required=True)
    args = ap.parse_args()

    sys.path.insert(0, os.path.abspath(args.reference))
    os.chdir(args.reference)                       # tasks open data files relative to the repo root
    import mujoco  # noqa: F401  (fails loudly where it is missing)
    if args.env == "jvrc_walk":
        from envs.jvrc import JvrcWalkEnv as Env
    elif args.env == "jvrc_step":
        from envs.jvrc import JvrcStepEnv as Env
    else:
        from envs.h1 import H1Env as Env

    np.random.seed(args.seed)
    env = Env()
    if args.env == "h1":                            # randomisation / noise off: they draw from the global MT19937 stream
        env.dynrand_interval = 0
        env.perturb_interval = 0
        if hasattr(env.cfg, "observation_noise"):
            env.cfg.observation_noise.enabled = False
        if hasattr(env.cfg, "init_noise"):
            env.cfg.init_noise = 0
    obs0 = env.reset()
    task, d, m = env.task, env.data, env.model
    init = dict(qpos=d.qpos.copy(), qvel=d.qvel.copy(), qacc_warmstart=d.qacc_warmstart.copy(),
                actuator_length=d.actuator_length.copy(), actuator_velocity=d.actuator_velocity.copy(), obs=np.asarray(obs0))
    for name in ("_phase", "_period", "mode_ref", "t1", "t2"):
        if hasattr(task, name):
            init["task" + name] = np.asarray(getattr(task, name), dtype=float)
    if hasattr(task, "mode"):
        init["task_mode"] = np.asarray(task.mode.value if hasattr(task.mode, "value") else task.mode)
    if hasattr(task, "sequence"):
        init["task_sequence"] = np.asarray(task.sequence, dtype=float)
    model_facts = dict(nq=m.nq, nv=m.nv, nu=m.nu, timestep=m.opt.timestep, total_mass=float(np.sum(m.body_mass)),
                       dof_invweight0=m.dof_invweight0.copy(), body_invweight0=m.body_invweight0.copy(),
                       meaninertia=float(m.stat.meaninertia), solver=int(m.opt.solver), cone=int(m.opt.cone),
                       iterations=int(m.opt.iterations), tolerance=float(m.opt.tolerance), impratio=float(m.opt.impratio))

    # freeze the task's random switches during the replay (they would consume MT19937 draws)
    task_step = task.step
    real_randint = np.random.randint

    def frozen_step():
        np.random.randint = lambda *a, **k: 1
        try:
            return task_step()
        finally:
            np.random.randint = real_randint
    task.step = frozen_step

    # substep tap
    sub = dict(qpos=[], qvel=[], ctrl=[], ncon=[], efc_force_norm=[])
    iface_step = env.interface.step
    tap = {"on": True}

    def tapped_step(*a, **k):
        if tap["on"]:
            sub["ctrl"].append(d.ctrl.copy())
        out = iface_step(*a, **k)
        if tap["on"]:
            sub["qpos"].append(d.qpos.copy())
            sub["qvel"].append(d.qvel.copy())
            sub["ncon"].append(int(d.ncon))
            sub["efc_force_norm"].append(float(np.linalg.norm(d.efc_force)) if d.nefc else 0.0)
        return out
    env.interface.step = tapped_step

    rng = np.random.RandomState(args.seed)
    actions = rng.normal(size=(args.steps, m.nu)) * args.sigma
    rec = dict(qpos=[], qvel=[], obs=[], reward=[], terms=[], done=[])
    names = None
    n_done = args.steps
    for t in range(args.steps):
        tap["on"] = t < args.substeps
        obs, rew, done, info = env.step(actions[t].copy())
        names = list(info.keys())
        rec["qpos"].append(d.qpos.copy()); rec["qvel"].append(d.qvel.copy()); rec["obs"].append(np.asarray(obs))
        rec["reward"].append(float(rew)); rec["terms"].append([float(v) for v in info.values()]); rec["done"].append(bool(done))
        if done:                                   # the replay is open loop: stop at the first termination
            n_done = t + 1
            break
    out = {("init_" + k): v for k, v in init.items()}
    out.update({("model_" + k): np.asarray(v) for k, v in model_facts.items()})
    out.update({k: np.asarray(v) for k, v in rec.items()})
    out.update({("sub_" + k): np.asarray(v) for k, v in sub.items()})
    out.update(actions=actions[:n_done], reward_names=np.asarray(names), env=np.asarray(args.env), seed=np.asarray(args.seed),
               mujoco_version=np.asarray(mujoco.__version__))
    np.savez_compressed(args.out, **out)
    print(f"wrote {args.out}: {n_done} control steps of {args.env}, {len(sub['qpos'])} substeps, mujoco {mujoco.__version__}")


if __name__ == "__main__":
    main()
