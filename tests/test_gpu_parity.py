"""GPU parity: the CUDA path (through torch.ops.lhw -> the C-ABI) against oracle/ on identical seeded inputs, with the Newton
tolerance tightened to 1e-14 on both sides so that the two formulations can be held to 1e-7.

Bar (BASELINE.json north_star): qpos/qvel within 1e-4 relative over 1000 control steps — asserted at the SHIPPED solver settings for
all four configs in tests/test_gpu_parity_shipped.py; this file is the tight-tolerance companion.  The fp32 kernel is an optional
fast path, not the parity path: here it only has to stay inside a 5e-3 envelope for 30 steps; how long it stays inside 1e-4 is
MEASURED in test_gpu_parity_shipped.py (400 steps standing jvrc_walk, 7 steps H1) and repeated in bench.py's extras.
The oracle itself is "parity unpinned" against MuJoCo (see oracle/sim_oracle.h).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


def _run_pair(oracle, precision, n, steps, act_fn, max_traj_len, tol):
    from learninghumanoidwalking_b200.envs import BatchedHumanoidEnv
    env = BatchedHumanoidEnv(n, precision=precision, seed=11, first_env_id=3, max_traj_len=max_traj_len,
                             tolerance=1e-14 if precision == 64 else None)
    envs = oracle.make_envs(n, seed=11, first_id=3)
    o_obs = oracle.batch_reset(envs, n)
    g_obs = env.reset().double().cpu().numpy()
    assert _rel(g_obs, o_obs) < tol
    worst = dict(obs=0.0, rew=0.0, qpos=0.0, qvel=0.0)
    n_end = 0
    for k in range(steps):
        a = act_fn(k, n)
        o_obs, o_tobs, o_terms, o_rew, o_done, o_end = oracle.batch_step(envs, n, a, max_traj_len=max_traj_len)
        g_obs, g_rew, g_done, g_end = env.step(torch.as_tensor(a, device="cuda", dtype=env.dtype))
        g_done, g_end = g_done.cpu().numpy(), g_end.cpu().numpy()
        assert (g_done == o_done).all() and (g_end == o_end).all(), f"done/ended mismatch at step {k}"
        n_end += int(o_end.sum())
        worst["obs"] = max(worst["obs"], _rel(g_obs.double().cpu().numpy(), o_obs))
        worst["rew"] = max(worst["rew"], _rel(g_rew.double().cpu().numpy(), o_rew))
        q = env.qpos.double().cpu().numpy()
        v = env.qvel.double().cpu().numpy()
        oq = np.stack([oracle.field(envs, i, "qpos") for i in range(n)])
        ov = np.stack([oracle.field(envs, i, "qvel") for i in range(n)])
        worst["qpos"] = max(worst["qpos"], _rel(q, oq))
        worst["qvel"] = max(worst["qvel"], _rel(v, ov))
        m = o_end.astype(bool)
        if m.any():
            assert _rel(env.term_obs.double().cpu().numpy()[m], o_tobs[m]) < tol
    env.close()
    return worst, n_end


def test_fp64_closed_loop_random_actions_with_resets(oracle_tight):
    rng = np.random.RandomState(0)
    worst, n_end = _run_pair(oracle_tight, 64, 8, 250, lambda k, n: rng.normal(size=(n, 12)) * 0.3, 80, 1e-7)
    assert n_end > 10, "expected falls / truncations to exercise the auto-reset path"
    assert max(worst.values()) < 1e-7, worst


def test_fp64_1000_steps_open_loop(oracle_tight):
    """north_star bar: <= 1e-4 rel on qpos/qvel over 1000 control steps from identical seeds."""
    worst, n_end = _run_pair(oracle_tight, 64, 2, 1000, lambda k, n: np.zeros((n, 12)), 400, 1e-6)
    assert worst["qpos"] < 1e-4 and worst["qvel"] < 1e-4, worst
    assert max(worst.values()) < 1e-6, worst


def test_fp32_tracks_oracle_within_an_episode(oracle_tight):
    """fp32 production kernel: same episode boundaries, observations within 5e-3 while an episode lasts
    (fp32 round-off is amplified by the falling robot; this is a sanity envelope, not the parity bar)."""
    from learninghumanoidwalking_b200.envs import BatchedHumanoidEnv
    n = 4
    env = BatchedHumanoidEnv(n, precision=32, seed=11, first_env_id=3, max_traj_len=400)
    envs = oracle_tight.make_envs(n, seed=11, first_id=3)
    o_obs = oracle_tight.batch_reset(envs, n)
    g_obs = env.reset().double().cpu().numpy()
    assert _rel(g_obs, o_obs) < 1e-5
    worst = 0.0
    for k in range(30):
        a = np.zeros((n, 12))
        o_obs, _, _, o_rew, o_done, o_end = oracle_tight.batch_step(envs, n, a, 400)
        g_obs, g_rew, g_done, g_end = env.step(torch.as_tensor(a, device="cuda", dtype=env.dtype))
        assert (g_end.cpu().numpy() == o_end).all()
        worst = max(worst, _rel(g_obs.double().cpu().numpy(), o_obs))
    assert worst < 5e-3, worst
    env.close()


def test_single_env_reference_protocol():
    """tests/test_environments.py contract: shapes, types, reward dict sums to the scalar, finite under +-10."""
    from learninghumanoidwalking_b200.envs import JvrcWalkEnv
    env = JvrcWalkEnv()
    obs = env.reset()
    assert isinstance(obs, np.ndarray) and obs.shape == (37,) and obs.dtype == np.float64
    assert env.observation_space.shape[0] == 37 and env.action_space.shape[0] == 12
    for a in (np.zeros(12), 10 * np.ones(12), -10 * np.ones(12)):
        obs, r, d, info = env.step(a)
        assert isinstance(r, float) and isinstance(d, bool) and isinstance(info, dict) and len(info) == 10
        assert abs(r - sum(info.values())) < 1e-6
        assert np.isfinite(obs).all()
    with pytest.raises(TypeError):
        env.step([0.0] * 12)
    # env.task / env.interface / env.model / env.data of the env protocol (envs/jvrc/jvrc_walk.py:24-40)
    from learninghumanoidwalking_b200.tasks.base_task import BaseTask
    assert isinstance(env.task, BaseTask) and env.task._client is env.interface
    assert env.task.calc_reward(None, None, None) == info and env.task.done() == d
    assert env.interface.get_qpos().shape == (19,) and env.interface.get_qvel().shape == (18,) and env.interface.nu() == 12
    assert np.array_equal(env.data.qpos, env.interface.get_qpos()) and env.model.opt.timestep == 0.001
    assert env.interface.get_act_joint_positions().shape == (12,) and 0 <= env.task._phase < 88
    env.close()


def test_path_to_yaml_changes_the_device_constants(tmp_path):
    """partial(Env, path_to_yaml) (run_experiment.py:115): softer PD gains from a user YAML give a different trajectory from the
    same seed; the reference's own values in a YAML give the identical one."""
    from learninghumanoidwalking_b200.envs import JvrcWalkEnv
    same, soft = tmp_path / "same.yaml", tmp_path / "soft.yaml"
    same.write_text("kp: [200, 200, 200, 250, 80, 80, 200, 200, 200, 250, 80, 80]\naction_smoothing: 0.5\n")
    soft.write_text("kp: [100, 100, 100, 125, 40, 40, 100, 100, 100, 125, 40, 40]\n")
    outs = []
    for y in (None, same, soft):
        env = JvrcWalkEnv(y, seed=4)
        env.reset()
        for _ in range(5):
            obs, _, _, _ = env.step(0.1 * np.ones(12))
        outs.append(obs)
        env.close()
    assert np.array_equal(outs[0], outs[1]) and np.abs(outs[0] - outs[2]).max() > 1e-3


def test_determinism_same_seed_bitwise():
    from learninghumanoidwalking_b200.envs import BatchedHumanoidEnv
    outs = []
    for _ in range(2):
        env = BatchedHumanoidEnv(64, precision=32, seed=5)
        env.reset()
        g = torch.Generator(device="cuda").manual_seed(1)
        for _k in range(20):
            a = torch.randn(64, 12, device="cuda", generator=g) * 0.3
            env.step(a)
        outs.append((env.state_r.clone(), env.state_i.clone()))
        env.close()
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
