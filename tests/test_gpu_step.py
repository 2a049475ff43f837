"""GPU parity for the JVRC-1 stepping environment (BASELINE configs[2], "jvrc_step footstep-plan task"): the CUDA path
through the C-ABI against oracle/ on identical seeds — footstep sequences, stepping-stone slab contacts (the oracle lists
one contact per supporting surface, the kernel merges coplanar surfaces into multiplicities), floor removal in FORWARD
mode, target tracking, goal-step observations, the height curriculum."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


@pytest.fixture(scope="module")
def step_oracle():
    from oracle.oracle import Oracle
    return Oracle("jvrc_step", tolerance=1e-14)


def test_step_fp64_closed_loop_all_modes_with_resets(step_oracle):
    from learninghumanoidwalking_b200.envs import BatchedHumanoidEnv
    o, n = step_oracle, 32
    env = BatchedHumanoidEnv(n, model="jvrc_step", precision=64, seed=21, first_env_id=7, max_traj_len=60, tolerance=1e-14)
    assert env.obs_dim == 39 and env.act_dim == 12 and env.state_r.shape[1] == 204
    envs = o.make_envs(n, seed=21, first_id=7)
    assert _rel(env.reset().cpu().numpy(), o.batch_reset(envs, n)) < 1e-9
    rng = np.random.RandomState(0)
    n_end, worst, modes, ncons = 0, 0.0, set(), set()
    for k in range(300):
        a = rng.normal(size=(n, 12)) * 0.25
        o_obs, o_tobs, o_terms, o_rew, o_done, o_end = o.batch_step(envs, n, a, max_traj_len=60)
        g_obs, g_rew, g_done, g_end = env.step(torch.as_tensor(a, device="cuda", dtype=env.dtype))
        assert (g_done.cpu().numpy() == o_done).all() and (g_end.cpu().numpy() == o_end).all(), f"step {k}"
        worst = max(worst, _rel(g_obs.cpu().numpy(), o_obs), _rel(g_rew.cpu().numpy(), o_rew),
                    _rel(env.rew_terms.cpu().numpy(), o_terms))
        oq = np.stack([o.field(envs, i, "qpos") for i in range(n)])
        ov = np.stack([o.field(envs, i, "qvel") for i in range(n)])
        worst = max(worst, _rel(env.qpos.cpu().numpy(), oq), _rel(env.qvel.cpu().numpy(), ov))
        m = o_end.astype(bool)
        if m.any():
            assert _rel(env.term_obs.cpu().numpy()[m], o_tobs[m]) < 1e-7
            n_end += int(m.sum())
        for i in range(n):
            modes.add(int(o.field(envs, i, "mode")[0]))
            ncons.add(int(o.field(envs, i, "ncon")[0]))
    assert n_end > 100 and modes == {0, 1, 2, 3, 4} and max(ncons) > 40
    assert worst < 1e-7, worst
    seq_o = np.stack([o.field(envs, i, "seq") for i in range(n)])
    assert np.abs(env.state_r[:, 119:199].cpu().numpy() - seq_o).max() < 1e-12      # same footstep sequences / slabs
    tk = env.state_r[:, 199:204].cpu().numpy()
    assert (tk[:, 1] == [int(o.field(envs, i, "t1")[0]) for i in range(n)]).all()
    assert (tk[:, 4] == [int(o.field(envs, i, "target_reached_frames")[0]) for i in range(n)]).all()
    assert (env.status_flags().cpu().numpy() == 0).all()
    env.close()


def test_step_curriculum_through_iteration_count(step_oracle):
    """env.robot.iteration_count = itr (rl/workers/rollout_worker.py:95) -> step heights of the next episodes."""
    from learninghumanoidwalking_b200.envs import BatchedHumanoidEnv
    from oracle.oracle import Oracle
    n = 64
    env = BatchedHumanoidEnv(n, model="jvrc_step", precision=64, seed=3, tolerance=1e-14)
    for it, h in ((0, 0.0), (7000, 0.05), (np.inf, 0.1)):
        env.robot.iteration_count = it
        env._fresh = True
        env.reset()
        o = Oracle("jvrc_step", tolerance=1e-14, iteration_count=it)
        envs = o.make_envs(n, seed=3)
        o.batch_reset(envs, n)
        seq = env.state_r[:, 119:199].cpu().numpy().reshape(n, 20, 4)
        seq_o = np.stack([o.field(envs, i, "seq") for i in range(n)]).reshape(n, 20, 4)
        assert np.abs(seq - seq_o).max() < 1e-12
        fwd = env.state_i[:, 1].cpu().numpy() == 4
        assert fwd.any() and abs(np.abs(seq[fwd][:, :, 2]).max() - (16 * h if h else 0)) <= h + 1e-12
    env.close()


def test_step_fp32_tracks_oracle_over_a_short_horizon(step_oracle):
    from learninghumanoidwalking_b200.envs import BatchedHumanoidEnv
    o, n = step_oracle, 8
    env = BatchedHumanoidEnv(n, model="jvrc_step", precision=32, seed=5, first_env_id=0)
    envs = o.make_envs(n, seed=5, first_id=0)
    assert np.abs(env.reset().double().cpu().numpy() - o.batch_reset(envs, n)).max() < 2e-3
    for k in range(8):
        a = np.zeros((n, 12))
        o_obs, _, _, o_rew, o_done, o_end = o.batch_step(envs, n, a)
        g_obs, g_rew, g_done, g_end = env.step(torch.as_tensor(a, device="cuda", dtype=env.dtype))
        if o_end.any() or g_end.any():
            break
        assert (np.abs(g_obs.double().cpu().numpy() - o_obs) / env.obs_std).max() < 5e-2
        assert np.abs(g_rew.double().cpu().numpy() - o_rew).max() < 5e-3
    env.close()


def test_step_env_protocol_and_ppo_iteration():
    """Reference test strategy (tests/test_environments.py:38-114, 174-188) on the JvrcStepEnv view, then one PPO
    iteration on the batched env (sampling + update) with the curriculum hook in the loop."""
    from learninghumanoidwalking_b200.envs import JvrcStepEnv
    env = JvrcStepEnv(seed=3)
    obs = env.reset()
    assert obs.shape == (39,) and obs.dtype == np.float64 and np.isfinite(obs).all()
    assert env.observation_space.shape == (39,) and env.action_space.shape == (12,)
    assert env.obs_mean.shape == env.obs_std.shape == (39,)
    assert len(env.robot.mirrored_obs) == 39 and env.robot.clock_inds == [29, 30]
    for a in (np.zeros(12), np.full(12, 10.0), np.full(12, -10.0)):
        obs, rew, done, info = env.step(a)
        assert obs.shape == (39,) and np.isfinite(obs).all() and isinstance(rew, float) and isinstance(done, bool)
        assert list(info) == ["foot_frc_score", "foot_vel_score", "orient_cost", "height_error", "step_reward",
                              "upper_body_reward"]
        assert abs(rew - sum(info.values())) < 1e-6
    with pytest.raises(TypeError):
        env.step([0.0] * 12)
    env.close()



def test_ppo_trains_on_the_stepping_environment_with_mirror_loss(tmp_path):
    """PPO data path on jvrc_step: obs 39, the robot's mirror lists extended over the 10 external observations
    (envs/jvrc/jvrc_base.py:69-110), the trainer's iteration counter driving the curriculum hook."""
    from types import SimpleNamespace
    from learninghumanoidwalking_b200.envs import BatchedHumanoidEnv
    from learninghumanoidwalking_b200.rl import PPO
    from learninghumanoidwalking_b200.rl.symmetric import SymmetricEnv
    args = SimpleNamespace(gamma=0.99, lam=0.95, lr=3e-4, eps=1e-5, entropy_coeff=0.0, clip=0.2, minibatch_size=256,
                           epochs=1, max_traj_len=50, num_procs=64, max_grad_norm=0.05, mirror_coeff=0.4, eval_freq=100,
                           recurrent=False, imitate_coeff=0.0, std_dev=0.223, learn_std=False, logdir=str(tmp_path),
                           steps_per_env=20)
    base = lambda: BatchedHumanoidEnv(64, model="jvrc_step", precision=32, seed=2)
    probe = base()
    r = probe.robot
    probe.close()
    env_fn = lambda: SymmetricEnv(base, mirrored_obs=r.mirrored_obs, mirrored_act=r.mirrored_acts, clock_inds=r.clock_inds)
    finals = []
    for _ in range(2):
        ppo = PPO(env_fn, args, seed=2)
        log = ppo.train(None, 2, verbose=False)
        assert np.isfinite(log[-1]["critic_loss"]) and float(log[-1]["mirror"]) > 0.0
        finals.append(ppo._flat_param.clone())
        batch = ppo.sample_parallel_with_workers()
        assert batch.states.shape == (64 * 20, 39) and batch.actions.shape == (64 * 20, 12)
        assert ppo.env.robot.iteration_count == 1      # rl/workers/rollout_worker.py:95
        ppo.env.close()
    assert torch.equal(finals[0], finals[1])     # same seed, bit-identical weights


def test_terrain_extension_fp64_closed_loop_against_oracle():
    """BASELINE configs[4] (uneven / compliant terrain; an extension): CUDA path vs oracle, with re-poses."""
    from learninghumanoidwalking_b200.envs import BatchedHumanoidEnv
    from oracle.oracle import Oracle
    o, n = Oracle("jvrc_walk_terrain", tolerance=1e-14), 32
    env = BatchedHumanoidEnv(n, model="jvrc_walk_terrain", precision=64, seed=9, first_env_id=3, max_traj_len=80, tolerance=1e-14)
    assert env.obs_dim == 37 and env.act_dim == 12 and env.state_r.shape[1] == 204
    envs = o.make_envs(n, seed=9, first_id=3)
    assert _rel(env.reset().cpu().numpy(), o.batch_reset(envs, n)) < 1e-9
    rng = np.random.RandomState(0)
    n_end, worst, reposed = 0, 0.0, 0
    prev = env.state_r[:, 119:199].clone()
    for k in range(300):
        a = rng.normal(size=(n, 12)) * 0.2
        o_obs, o_tobs, o_terms, o_rew, o_done, o_end = o.batch_step(envs, n, a, max_traj_len=80)
        g_obs, g_rew, g_done, g_end = env.step(torch.as_tensor(a, device="cuda", dtype=env.dtype))
        assert (g_done.cpu().numpy() == o_done).all() and (g_end.cpu().numpy() == o_end).all(), f"step {k}"
        worst = max(worst, _rel(g_obs.cpu().numpy(), o_obs), _rel(g_rew.cpu().numpy(), o_rew),
                    _rel(env.rew_terms.cpu().numpy(), o_terms))
        n_end += int(o_end.sum())
        cur = env.state_r[:, 119:199]
        reposed += int(((cur - prev).abs().amax(dim=1) > 0).sum().item())
        prev = cur.clone()
    assert n_end > 40 and reposed >= 5 and worst < 1e-7, (n_end, reposed, worst)
    seq_o = np.stack([o.field(envs, i, "seq") for i in range(n)])
    assert np.abs(env.state_r[:, 119:199].cpu().numpy() - seq_o).max() < 1e-12      # same terrain poses (FMA-level differences)
    env.close()
    env32 = BatchedHumanoidEnv(256, model="jvrc_walk_terrain", precision=32, seed=1)
    env32.reset()
    for _ in range(20):
        obs, rew, done, ended = env32.step(torch.randn(256, 12, device="cuda") * 0.2)
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all() and (env32.status_flags() == 0).all()
    env32.close()
