#!/usr/bin/env python3
"""Generate tests/golden/*.json by RUNNING the reference's own importable code in this container.

/root/reference does not exist on the GPU box, so the vectors are committed; this script is the
provenance.  Only modules that import with torch/numpy/scipy are used:
  tasks/rewards.py (loaded by file path; tasks/__init__ pulls transforms3d),
  rl/storage/rollout_storage.py, rl/policies/{actor,critic}.py, rl/envs/wrappers.py.
The roll/pitch vectors use scipy's Rotation as a stand-in for transforms3d.quat2euler (not
installed): 'sxyz' static-frame euler == scipy extrinsic 'xyz'.
"""
import importlib.util
import json
import os
import sys

import numpy as np
import torch

REF = os.environ.get("LHW_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")


def load_by_path(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    os.makedirs(OUT, exist_ok=True)
    sys.path.insert(0, REF)
    rewards = load_by_path("ref_rewards", "tasks/rewards.py")
    rng = np.random.RandomState(1234)

    # ---- gait clocks: create_phase_reward(0.75, 0.35, 0.1, "grounded", 40) at integer phases (walking_task.py:198-200)
    right, left = rewards.create_phase_reward(0.75, 0.35, 0.1, "grounded", 40)
    period = int(np.floor(2 * 1.1 * 40))
    ph = np.arange(period)
    clocks = dict(period=period, swing=0.75, stance=0.35, relax=0.1, freq=40,
                  r_frc=[float(right[0](p)) for p in ph], r_vel=[float(right[1](p)) for p in ph],
                  l_frc=[float(left[0](p)) for p in ph], l_vel=[float(left[1](p)) for p in ph])
    json.dump(clocks, open(os.path.join(OUT, "gait_clocks.json"), "w"), indent=0)

    # ---- scalar reward terms on random inputs
    cases = []
    for _ in range(64):
        c = {}
        rv, gv = rng.uniform(-1, 1, 2), rng.uniform(-0.5, 0.5, 2)
        c["fwd_vel"] = dict(root_vel=rv.tolist(), goal=gv.tolist(), out=float(rewards.calc_fwd_vel_reward(rv, gv)))
        yv, yr = rng.uniform(-2, 2), rng.uniform(-0.5, 0.5)
        c["yaw_vel"] = dict(yaw_vel=yv, ref=yr, out=float(rewards.calc_yaw_vel_reward(yv, yr)))
        a, pa = rng.uniform(-1, 1, 12), rng.uniform(-1, 1, 12)
        c["action"] = dict(a=a.tolist(), prev=pa.tolist(), out=float(rewards.calc_action_reward(a, pa)))
        t, pt = rng.uniform(-80, 80, 12), rng.uniform(-80, 80, 12)
        c["torque"] = dict(t=t.tolist(), prev=pt.tolist(), out=float(rewards.calc_torque_reward(t, pt)))
        h, gs, cz = rng.uniform(0.6, 1.0), rng.uniform(0, 0.4), rng.uniform(-0.01, 0.0)
        c["height"] = dict(h=h, goal=0.8, speed=gs, cz=cz, out=float(rewards.calc_height_reward(h, 0.8, gs, cz)))
        qv, qa = rng.uniform(-3, 3, 18), rng.uniform(-10, 10, 18)
        c["root_accel"] = dict(qvel=qv.tolist(), qacc=qa.tolist(), out=float(rewards.calc_root_accel_reward(qv, qa)))
        lf, rf, p = rng.uniform(0, 700), rng.uniform(0, 700), int(rng.randint(period))
        c["foot_frc"] = dict(l=lf, r=rf, phase=p, mass=62.4,
                             out=float(rewards.calc_foot_frc_clock_reward(lf, rf, p, left[0], right[0], 62.4)))
        lv, rvv = rng.uniform(-0.3, 0.3, 3), rng.uniform(-0.3, 0.3, 3)
        c["foot_vel"] = dict(l=lv.tolist(), r=rvv.tolist(), phase=p,
                             out=float(rewards.calc_foot_vel_clock_reward(lv, rvv, p, left[1], right[1])))
        cases.append(c)
    json.dump(cases, open(os.path.join(OUT, "reward_terms.json"), "w"))

    # ---- roll/pitch from quaternions (tasks/observations.py:22), scipy stand-in for transforms3d
    from scipy.spatial.transform import Rotation
    rp = []
    for _ in range(64):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        e = Rotation.from_quat([q[1], q[2], q[3], q[0]]).as_euler("xyz")
        rp.append(dict(quat=q.tolist(), roll=float(e[0]), pitch=float(e[1])))
    json.dump(rp, open(os.path.join(OUT, "roll_pitch.json"), "w"))

    # ---- GAE: PPOBuffer.finish_path (rl/storage/rollout_storage.py:53-85), several paths per buffer
    from rl.storage.rollout_storage import PPOBuffer
    gae = []
    for case in range(8):
        T = int(rng.randint(5, 60))
        gamma, lam = float(rng.choice([0.99, 0.95])), float(rng.choice([0.95, 0.9, 1.0]))
        buf = PPOBuffer(3, 2, gamma=gamma, lam=lam, size=T)
        rew, val = rng.uniform(-1, 1, T), rng.uniform(-2, 2, T)
        ends = sorted(set(rng.randint(1, T, size=3).tolist() + [T]))
        last_vals, dones = [], np.zeros(T)
        t0 = 0
        for e_ in ends:
            for t in range(t0, e_):
                buf.store(torch.zeros(3), torch.zeros(2), torch.tensor(rew[t]), torch.tensor(val[t]), t == e_ - 1)
            lv = float(rng.uniform(-2, 2)) if rng.rand() < 0.5 else 0.0
            last_vals.append(lv)
            dones[e_ - 1] = 1
            buf.finish_path(last_val=torch.tensor([lv], dtype=torch.float64))
            t0 = e_
        data = buf.get_data()
        gae.append(dict(gamma=gamma, lam=lam, rewards=rew.tolist(), values=val.tolist(), path_ends=ends,
                        last_vals=last_vals, returns=data.returns[:, 0].tolist(), traj_idx=data.traj_idx.tolist()))
    # SURVEY Appendix B known answer
    buf = PPOBuffer(1, 1, gamma=0.99, lam=0.95, size=5)
    for t in range(5):
        buf.store(torch.zeros(1), torch.zeros(1), torch.tensor(float(t + 1)), torch.tensor(0.5 * t), t == 4)
    buf.finish_path(last_val=torch.tensor([2.0], dtype=torch.float64))
    gae.append(dict(gamma=0.99, lam=0.95, rewards=[1, 2, 3, 4, 5], values=[0, .5, 1, 1.5, 2], path_ends=[5],
                    last_vals=[2.0], returns=buf.get_data().returns[:, 0].tolist(), traj_idx=[0, 5]))
    json.dump(gae, open(os.path.join(OUT, "gae.json"), "w"))

    # ---- policy / critic forward (rl/policies/actor.py:122-189, critic.py:15-49) and mirror matrices (rl/envs/wrappers.py)
    from rl.envs.wrappers import _get_symmetry_matrix
    from rl.policies.actor import Gaussian_FF_Actor
    from rl.policies.critic import FF_V
    torch.manual_seed(7)
    # small hidden layers keep the committed fixture small; parameter counts use the real (256,256) nets
    actor = Gaussian_FF_Actor(37, 12, layers=(16, 16), init_std=0.223, learn_std=False, bounded=False)
    critic = FF_V(37, layers=(16, 16))
    full_actor = Gaussian_FF_Actor(37, 12, init_std=0.223, learn_std=False, bounded=False)
    full_critic = FF_V(37)
    obs_mean = torch.tensor(rng.uniform(-0.5, 0.5, 37), dtype=torch.float32)
    obs_std = torch.tensor(rng.uniform(0.5, 2.0, 37), dtype=torch.float32)
    actor.obs_mean, actor.obs_std, critic.obs_mean, critic.obs_std = obs_mean, obs_std, obs_mean, obs_std
    x = torch.tensor(rng.uniform(-1, 1, (5, 37)), dtype=torch.float32)
    with torch.no_grad():
        mu = actor(x, deterministic=True)
        v = critic(x)
    sd = {k: v_.tolist() for k, v_ in actor.state_dict().items()}
    sdc = {k: v_.tolist() for k, v_ in critic.state_dict().items()}
    mirrored_obs = [-0.1, 1, -2, 3, -4, 11, -12, -13, 14, -15, 16, 5, -6, -7, 8, -9, 10,
                    23, -24, -25, 26, -27, 28, 17, -18, -19, 20, -21, 22] + list(range(29, 37))
    mirrored_act = [6, -7, -8, 9, -10, 11, 0.1, -1, -2, 3, -4, 5]
    json.dump(dict(actor=sd, critic=sdc, obs_mean=obs_mean.tolist(), obs_std=obs_std.tolist(), x=x.tolist(),
                   mu=mu.tolist(), v=v.tolist(), n_actor=sum(p.numel() for p in full_actor.parameters()),
                   n_critic=sum(p.numel() for p in full_critic.parameters()),
                   out_layer_norm=float(full_actor.means.weight.norm(dim=1).mean()),
                   mirrored_obs=mirrored_obs, mirrored_act=mirrored_act,
                   obs_mirror_matrix=_get_symmetry_matrix(mirrored_obs).tolist(),
                   act_mirror_matrix=_get_symmetry_matrix(mirrored_act).tolist()),
              open(os.path.join(OUT, "policy.json"), "w"))
    print("golden vectors written to", os.path.abspath(OUT))


def seeding_vectors():
    """rl/utils/seeding.py:get_worker_seed (the rank plays the worker's role in this build) on a grid of arguments."""
    seeding = load_by_path("ref_seeding", "rl/utils/seeding.py")
    grid = [(m, w, o) for m in (0, 1, 7, 12345, 2 ** 31 - 1, 4294967295) for w in (0, 1, 7, 255) for o in (0, 1, 3)]
    out = [dict(master_seed=m, worker_id=w, offset=o, seed=int(seeding.get_worker_seed(m, w, o))) for m, w, o in grid]
    json.dump(out, open(os.path.join(OUT, "seeding.json"), "w"))
    return out


if __name__ == "__main__":
    main()
    seeding_vectors()
