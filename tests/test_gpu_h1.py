"""GPU parity for the H1 standing environment (BASELINE config "h1 standing task, domain-randomised PD gains/mass"):
the CUDA path through the C-ABI against oracle/ on identical seeds — observation noise, per-episode dynamics
randomisation, random pushes and initial-pose noise included (all drawn from the same counter-based Philox streams)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


@pytest.fixture(scope="module")
def h1_oracle():
    from oracle.oracle import Oracle
    return Oracle("h1", tolerance=1e-14)


@pytest.mark.parametrize("sigma,steps,min_ends", [(0.3, 300, 40), (1.0, 120, 60)])   # sigma 1.0: legs cross -> self-collision flag
def test_h1_fp64_closed_loop_with_randomisation_and_resets(h1_oracle, sigma, steps, min_ends):
    from learninghumanoidwalking_b200.envs import BatchedHumanoidEnv
    o, n = h1_oracle, 16
    env = BatchedHumanoidEnv(n, model="h1", precision=64, seed=21, first_env_id=7, max_traj_len=60, tolerance=1e-14)
    assert env.obs_dim == 35 and env.act_dim == 10
    envs = o.make_envs(n, seed=21, first_id=7)
    assert _rel(env.reset().cpu().numpy(), o.batch_reset(envs, n)) < 1e-9
    rng = np.random.RandomState(0)
    n_end, pushed, worst = 0, False, 0.0
    for k in range(steps):
        a = rng.normal(size=(n, 10)) * sigma
        o_obs, o_tobs, o_terms, o_rew, o_done, o_end = o.batch_step(envs, n, a, max_traj_len=60)
        g_obs, g_rew, g_done, g_end = env.step(torch.as_tensor(a, device="cuda", dtype=env.dtype))
        assert (g_done.cpu().numpy() == o_done).all() and (g_end.cpu().numpy() == o_end).all(), f"step {k}"
        worst = max(worst, _rel(g_obs.cpu().numpy(), o_obs), _rel(g_rew.cpu().numpy(), o_rew),
                    _rel(env.rew_terms.cpu().numpy(), o_terms))
        oq = np.stack([o.field(envs, i, "qpos")[:17] for i in range(n)])
        ov = np.stack([o.field(envs, i, "qvel")[:16] for i in range(n)])
        worst = max(worst, _rel(env.qpos.cpu().numpy(), oq), _rel(env.qvel.cpu().numpy(), ov))
        x = np.stack([o.field(envs, i, "xfrc") for i in range(n)])
        pushed |= bool(np.any(x))
        assert np.abs(env.state_r[:, -12:].cpu().numpy() - x).max() < 1e-12
        m = o_end.astype(bool)
        if m.any():
            assert _rel(env.term_obs.cpu().numpy()[m], o_tobs[m]) < 1e-7
            n_end += int(m.sum())
    assert n_end > min_ends and (pushed or steps < 200)
    assert worst < 1e-7, worst
    fl = np.stack([o.field(envs, i, "P_frictionloss")[6:16] for i in range(n)])
    assert np.abs(env.state_r[:, 163:173].cpu().numpy() - fl).max() < 1e-14
    env.close()


def test_h1_fp32_tracks_oracle_over_a_short_horizon(h1_oracle):
    from learninghumanoidwalking_b200.envs import BatchedHumanoidEnv
    o, n = h1_oracle, 8
    env = BatchedHumanoidEnv(n, model="h1", precision=32, seed=5, first_env_id=0)
    envs = o.make_envs(n, seed=5, first_id=0)
    std = np.concatenate(([0.2, 0.2, 1, 1, 1], 0.5 * np.ones(10), 4 * np.ones(10), 100 * np.ones(10)))
    assert (np.abs(env.reset().double().cpu().numpy() - o.batch_reset(envs, n)) / std).max() < 2e-3
    for k in range(8):
        a = np.zeros((n, 10))
        o_obs, _, _, o_rew, o_done, o_end = o.batch_step(envs, n, a)
        g_obs, g_rew, g_done, g_end = env.step(torch.as_tensor(a, device="cuda", dtype=env.dtype))
        if o_end.any() or g_end.any():
            break
        assert (np.abs(g_obs.double().cpu().numpy() - o_obs) / std).max() < 5e-2
        assert np.abs(g_rew.double().cpu().numpy() - o_rew).max() < 5e-3
    env.close()


def test_h1_env_protocol_and_determinism():
    """Reference test strategy (tests/test_environments.py:38-114) on the H1Env view + batch-composition independence."""
    from learninghumanoidwalking_b200.envs import BatchedHumanoidEnv, H1Env
    env = H1Env(seed=3)
    obs = env.reset()
    assert obs.shape == (35,) and obs.dtype == np.float64 and np.isfinite(obs).all()
    assert env.observation_space.shape == (35,) and env.action_space.shape == (10,)
    assert env.obs_mean.shape == env.obs_std.shape == (35,)
    o2, r, d, info = env.step(np.zeros(10))
    assert o2.shape == (35,) and isinstance(r, float) and isinstance(d, bool)
    assert list(info) == ["com_vel_error", "yaw_vel_error", "height", "upperbody", "joint_torque_reward", "posture"]
    assert abs(r - sum(info.values())) < 1e-6 and 0.0 < r <= 1.0 + 1e-9
    with pytest.raises(TypeError):
        env.step([0.0] * 10)
    env.close()
    # an env's trajectory depends only on (seed, env id): slices of a big batch equal a small batch
    big = BatchedHumanoidEnv(64, model="h1", precision=64, seed=9, first_env_id=0)
    small = BatchedHumanoidEnv(8, model="h1", precision=64, seed=9, first_env_id=24)
    ob, os_ = big.reset(), small.reset()
    assert torch.equal(ob[24:32], os_)
    rng = torch.Generator(device="cuda").manual_seed(0)
    for _ in range(30):
        a = torch.randn(64, 10, device="cuda", dtype=torch.float64, generator=rng) * 0.3
        ob, rb, db, eb = big.step(a)
        os_, rs, ds, es = small.step(a[24:32].contiguous())
        assert torch.equal(ob[24:32], os_) and torch.equal(rb[24:32], rs) and torch.equal(eb[24:32], es)
    # the randomised parameters differ between envs and stay inside the reference's ranges
    floss, damp = big.state_r[:, 163:173], big.state_r[:, 153:163]
    assert float(floss.min()) >= 0 and float(floss.max()) <= 2 and float(damp.min()) >= 0.02 and float(damp.max()) <= 2
    assert float(floss.std()) > 0.3
    mass = big.state_r[:, 103:114]
    nominal = torch.tensor([lk["mass"] for lk in big.mj["links"]], device="cuda", dtype=torch.float64)
    ratio = mass[:, 1:] / nominal[1:]
    assert float(ratio.min()) >= 0.95 - 1e-12 and float(ratio.max()) <= 1.05 + 1e-12
    big.close(); small.close()


def test_h1_switches_turn_the_randomisation_off():
    from learninghumanoidwalking_b200.envs import BatchedHumanoidEnv
    env = BatchedHumanoidEnv(4, model="h1", precision=64, seed=1, observation_noise=False, domain_randomization=False,
                             init_noise=False)
    obs = env.reset()
    assert torch.equal(obs[0], obs[1]) and torch.equal(obs[0], obs[3])       # nothing random left
    assert float(env.state_r[:, 163:173].abs().max()) == 0.0                   # no friction loss
    for _ in range(5):
        obs, *_ = env.step(torch.zeros(4, 10, device="cuda", dtype=torch.float64))
    assert torch.equal(obs[0], obs[2])
    env.close()


def test_ppo_trains_on_the_h1_environment(tmp_path):
    """The PPO data path (device rollout worker, GAE / adv-norm / gather / fused clip+Adam kernels) on the H1 standing
    env: obs 35, act 10, no mirror lists on the robot (envs/h1/h1_env.py has none) -> mirror loss is 0."""
    from types import SimpleNamespace
    from learninghumanoidwalking_b200.envs import BatchedHumanoidEnv
    from learninghumanoidwalking_b200.rl import PPO
    args = SimpleNamespace(gamma=0.99, lam=0.95, lr=3e-4, eps=1e-5, entropy_coeff=0.0, clip=0.2, minibatch_size=256,
                           epochs=1, max_traj_len=50, num_procs=64, max_grad_norm=0.05, mirror_coeff=0.4, eval_freq=100,
                           recurrent=False, imitate_coeff=0.0, std_dev=0.223, learn_std=False, logdir=str(tmp_path),
                           steps_per_env=20)
    finals = []
    for _ in range(2):
        ppo = PPO(lambda: BatchedHumanoidEnv(64, model="h1", precision=32, seed=2), args, seed=2)
        log = ppo.train(None, 2, verbose=False)
        assert np.isfinite(log[-1]["critic_loss"]) and float(log[-1]["mirror"]) == 0.0
        finals.append(ppo._flat_param.clone())
        batch = ppo.sample_parallel_with_workers()
        assert batch.states.shape == (64 * 20, 35) and batch.actions.shape == (64 * 20, 10)
        ppo.env.close()
    assert torch.equal(finals[0], finals[1])     # same seed, bit-identical weights
    actor = torch.load(tmp_path / "actor_0.pt", weights_only=False)   # saved at eval_freq boundaries (itr 0)
    assert actor.cuda()(batch.states[:4]).shape == (4, 10)


def test_h1_pd_gain_randomisation_matches_oracle():
    """BASELINE configs[3] names domain-randomised PD gains: RobotBase(pdrand_k) (robots/robot_base.py:41-47, off in every
    reference env) is available as pd_gain_randomization=k; CUDA path vs oracle with k = 0.3."""
    from learninghumanoidwalking_b200.envs import BatchedHumanoidEnv
    from oracle.oracle import Oracle
    o, n = Oracle("h1", tolerance=1e-14, pdrand_k=0.3), 8
    env = BatchedHumanoidEnv(n, model="h1", precision=64, seed=4, first_env_id=1, max_traj_len=50, tolerance=1e-14,
                             pd_gain_randomization=0.3)
    envs = o.make_envs(n, seed=4, first_id=1)
    assert _rel(env.reset().cpu().numpy(), o.batch_reset(envs, n)) < 1e-9
    rng = np.random.RandomState(0)
    for k in range(60):
        a = rng.normal(size=(n, 10)) * 0.3
        o_obs, _, _, o_rew, o_done, o_end = o.batch_step(envs, n, a, max_traj_len=50)
        g_obs, g_rew, g_done, g_end = env.step(torch.as_tensor(a, device="cuda", dtype=env.dtype))
        assert (g_done.cpu().numpy() == o_done).all() and (g_end.cpu().numpy() == o_end).all()
        assert _rel(g_obs.cpu().numpy(), o_obs) < 1e-7 and _rel(g_rew.cpu().numpy(), o_rew) < 1e-7
    env.close()
