"""Pin the oracle (and the product's host-side tables) to vectors produced by the reference's own code
(tools/gen_golden.py ran tasks/rewards.py, rl/storage/rollout_storage.py, rl/policies, rl/envs/wrappers)."""
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gold(name):
    return json.load(open(os.path.join(GOLD, name)))


def test_gait_clock_table_matches_reference_splines():
    from learninghumanoidwalking_b200.tasks.gait_clock import phase_clock_table
    g = gold("gait_clocks.json")
    period, table = phase_clock_table(0.75, 0.35, 0.1, "grounded", 40, total_duration=1.1)
    assert period == g["period"] == 88
    for row, key in enumerate(("r_frc", "r_vel", "l_frc", "l_vel")):
        assert np.abs(table[row] - np.array(g[key])).max() < 1e-12
    # WalkingTask.step tests `clock == 1` exactly (tasks/walking_task.py:156): double-support phases must be exact
    dbl = [p for p in range(88) if table[0][p] == 1.0 and table[2][p] == 1.0]
    assert dbl == [p for p in range(88) if g["r_frc"][p] == 1 and g["l_frc"][p] == 1] == list(range(32, 43)) + list(range(76, 87))


def test_roll_pitch_matches_static_xyz_euler(oracle_tight):
    for c in gold("roll_pitch.json"):
        r, p = oracle_tight.quat2rp(c["quat"])
        assert abs(r - c["roll"]) < 1e-12 and abs(p - c["pitch"]) < 1e-12


def test_reward_terms_match_reference(oracle_tight):
    """Drive oracle calc_reward through its env fields so that each term sees exactly the golden inputs."""
    o = oracle_tight
    envs = o.make_envs(1)
    W = dict(fwd_vel=0.15, yaw_vel=0.15, action=0.025, torque=0.025, height=0.05, root_accel=0.05, foot_frc=0.225, foot_vel=0.225)
    for c in gold("reward_terms.json"):
        ph = c["foot_frc"]["phase"]
        o.set_field(envs, 0, "phase", [ph])
        o.set_field(envs, 0, "mode", [2])  # FORWARD: refs (yaw 0, vx, vy=0)
        # com velocity: identity root frame, goal (vx, 0)
        o.set_field(envs, 0, "root_xmat", np.eye(3).reshape(-1))
        rv, gv = c["fwd_vel"]["root_vel"], c["fwd_vel"]["goal"]
        o.set_field(envs, 0, "mode_ref", [0.0, gv[0], 0.0])
        o.set_field(envs, 0, "root_vlin", [rv[0], rv[1] - gv[1], 0.0])
        qv = np.array(c["root_accel"]["qvel"])
        qv[5] = c["yaw_vel"]["yaw_vel"] - c["yaw_vel"]["ref"]  # yaw ref is 0 in FORWARD
        o.set_field(envs, 0, "qvel", qv)
        o.set_field(envs, 0, "qacc", c["root_accel"]["qacc"])
        o.set_field(envs, 0, "lfoot_grf", [c["foot_frc"]["l"]])
        o.set_field(envs, 0, "rfoot_grf", [c["foot_frc"]["r"]])
        o.set_field(envs, 0, "lfoot_vel", c["foot_vel"]["l"])
        o.set_field(envs, 0, "rfoot_vel", c["foot_vel"]["r"])
        o.set_field(envs, 0, "prev_torque", c["torque"]["prev"])
        o.set_field(envs, 0, "act_force", c["torque"]["t"])
        o.set_field(envs, 0, "prev_action", c["action"]["prev"])
        t = o.calc_reward(envs, 0, c["action"]["a"])
        assert abs(t[0] - W["foot_frc"] * c["foot_frc"]["out"]) < 1e-12
        assert abs(t[1] - W["foot_vel"] * c["foot_vel"]["out"]) < 1e-12
        qv_ref = np.array(c["root_accel"]["qvel"])
        exp_acc = np.exp(-0.25 * (np.abs(qv[3:6]).sum() + np.abs(np.array(c["root_accel"]["qacc"])[0:3]).sum()))
        assert abs(t[2] - W["root_accel"] * exp_acc) < 1e-12
        assert abs(t[4] - W["fwd_vel"] * c["fwd_vel"]["out"]) < 1e-12
        assert abs(t[5] - W["yaw_vel"] * c["yaw_vel"]["out"]) < 1e-12
        assert abs(t[8] - W["torque"] * c["torque"]["out"]) < 1e-12
        assert abs(t[9] - W["action"] * c["action"]["out"]) < 1e-12
        del qv_ref
    # root_accel and height with their own golden inputs (FORWARD speed couples height's dead zone to vx)
    for c in gold("reward_terms.json"):
        o.set_field(envs, 0, "mode", [2])
        o.set_field(envs, 0, "mode_ref", [0.0, c["height"]["speed"], 0.0])
        o.set_field(envs, 0, "root_xpos", [0.0, 0.0, c["height"]["h"]])
        o.set_field(envs, 0, "contact_z_min", [c["height"]["cz"]])
        o.set_field(envs, 0, "ncon_r", [1])
        o.set_field(envs, 0, "qvel", c["root_accel"]["qvel"])
        o.set_field(envs, 0, "qacc", c["root_accel"]["qacc"])
        t = o.calc_reward(envs, 0, np.zeros(12))
        assert abs(t[3] - 0.05 * c["height"]["out"]) < 1e-12
        assert abs(t[2] - 0.05 * c["root_accel"]["out"]) < 1e-12


def test_known_answers_from_survey_appendix_b():
    from oracle.ppo_oracle import gae_path
    r = gae_path([1, 2, 3, 4, 5], [0, .5, 1, 1.5, 2], 2.0, 0.99, 0.95)
    assert np.allclose(r, [14.542834082482912, 14.3732951435225, 13.103450445, 10.66369, 6.98], atol=1e-12)


def test_gae_oracle_matches_reference_buffer():
    from oracle.ppo_oracle import gae_rollout
    for c in gold("gae.json"):
        T = len(c["rewards"])
        ended = np.zeros((T, 1), dtype=int)
        boot = np.zeros((T, 1))
        for e_, lv in zip(c["path_ends"], c["last_vals"]):
            ended[e_ - 1, 0] = 1
            boot[e_ - 1, 0] = lv
        ret = gae_rollout(np.array(c["rewards"])[:, None], np.array(c["values"])[:, None], ended, boot, np.zeros(1),
                          c["gamma"], c["lam"])
        assert np.abs(ret[:, 0] - np.array(c["returns"])).max() < 1e-12


def test_model_constants():
    from learninghumanoidwalking_b200.model import load_model
    m = load_model("jvrc_walk")
    assert abs(m["total_mass"] - 62.4) < 1e-9
    assert len(m["links"]) == 13 and m["cfg"]["frame_skip"] == 25
    names = [lk["joint"]["name"] for lk in m["links"][1:]]
    assert names == ["R_HIP_P", "R_HIP_R", "R_HIP_Y", "R_KNEE", "R_ANKLE_R", "R_ANKLE_P",
                     "L_HIP_P", "L_HIP_R", "L_HIP_Y", "L_KNEE", "L_ANKLE_R", "L_ANKLE_P"]  # envs/jvrc/gen_xml.py:42-55
    assert abs(m["total_mass"] * 9.8 * 0.5 - 305.76) < 1e-9


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_compiled_model_is_reproducible_from_reference():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(__file__)), "tools"))
    import compile_model
    from learninghumanoidwalking_b200.model import load_model
    fresh, stored = compile_model.compile_jvrc(), load_model("jvrc_walk")
    for a, b in zip(fresh["links"], stored["links"]):
        assert a["name"] == b["name"] and np.allclose(a["inertia"], b["inertia"]) and np.allclose(a["com"], b["com"])
    assert np.allclose(fresh["dof_invweight0"], stored["dof_invweight0"])


def test_rollout_bootstrap_semantics_match_the_reference_worker():
    """tests/golden/rollout_worker.json = the reference's RolloutWorker.sample run on a scripted env for three consecutive
    calls (tools/gen_golden_rollout.py).  The oracle's time-major GAE (ppo_oracle.gae_rollout: what the CUDA GAE kernel and the
    device rollout worker are tested against) must give the reference's returns from the same per-step rewards / values /
    episode-end flags with boot = (not done) * critic(pre-reset next state) and the open tail closed by critic(current state)."""
    import torch
    from oracle.ppo_oracle import gae_rollout
    Gr = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rollout_worker.json")))
    W, b = torch.tensor(Gr["critic_w"], dtype=torch.float32), Gr["critic_b"]
    V = lambda s: float((torch.tensor(s, dtype=torch.float32) @ W + b).reshape(-1)[0])
    T, carried = Gr["T"], 0
    for call in Gr["calls"]:
        ended = np.array(call["dones"]).astype(bool)
        done = np.array(call["env_done"])
        assert (ended >= done).all()                                  # every true termination ends the episode ...
        boot = np.array([(0.0 if done[t] else V(call["next_obs"][t])) if ended[t] else 0.0 for t in range(T)])
        last_val = V(call["open_state"]) if call["open_state"] is not None else 0.0
        assert (call["open_state"] is None) == bool(ended[-1])
        ret = gae_rollout(np.array(call["rewards"])[:, None], np.array(call["values"])[:, None], ended[:, None], boot[:, None],
                          np.array([last_val]), Gr["gamma"], Gr["lam"])[:, 0]
        assert np.abs(ret - np.array(call["returns"])).max() < 1e-6
        # ... and truncation at max_traj_len does too; episode lengths carry over between calls; only completed episodes report
        lens, run = [], carried
        for t in range(T):
            run += 1
            if ended[t]:
                assert done[t] or run == Gr["max_traj_len"]
                lens.append(run)
                run = 0
        carried = run
        assert lens == call["ep_lens"] and len(call["ep_rewards"]) == len(lens)
