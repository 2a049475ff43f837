"""Size-independent properties of the CUDA path at BASELINE's full batch sizes (the oracle cannot run these sizes in
seconds): per-env independence from batch composition, odd batch sizes, truncation / auto-reset semantics, the
reference's reward-sum and finiteness contracts (tests/test_environments.py:105-114, 174-188)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _env(n, **kw):
    from learninghumanoidwalking_b200.envs import BatchedHumanoidEnv
    return BatchedHumanoidEnv(n, **kw)


def _actions(n, steps, seed=0, sigma=0.223, dtype=torch.float64):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(steps, n, 12, device="cuda", generator=g, dtype=dtype) * sigma


@pytest.mark.parametrize("precision", [64, 32])
def test_env_trajectory_does_not_depend_on_batch_size_or_block_assignment(precision):
    """Env i is keyed by (seed, env_id) only: the first 50 envs of a 4096 batch evolve bit-identically in a 50-env batch
    (different grid, different warp-to-block assignment, different neighbours)."""
    dt = torch.float64 if precision == 64 else torch.float32
    big, small = _env(4096, precision=precision, seed=9, max_traj_len=30), _env(50, precision=precision, seed=9, max_traj_len=30)
    big.reset(); small.reset()
    acts = _actions(4096, 40, dtype=dt)
    for k in range(40):
        big.step(acts[k]); small.step(acts[k, :50].contiguous())
    assert torch.equal(big.state_r[:50], small.state_r) and torch.equal(big.state_i[:50, :7], small.state_i[:, :7])
    assert torch.equal(big.obs[:50], small.obs) and torch.equal(big.reward[:50], small.reward)
    big.close(); small.close()


def test_odd_batch_sizes_match_the_oracle_on_the_last_env(oracle_tight):
    """n not a multiple of the block's warp count: the tail block has idle warps that must still reach the barriers."""
    for n in (1, 13, 4097):
        env = _env(n, precision=64, seed=3, tolerance=1e-14)
        obs = env.reset()
        envs = oracle_tight.make_envs(1, seed=3, first_id=n - 1)
        o_obs = oracle_tight.batch_reset(envs, 1)
        assert np.abs(obs[-1].cpu().numpy() - o_obs[0]).max() < 1e-9
        rng = np.random.RandomState(n)
        for _ in range(5):
            a = rng.normal(size=(n, 12)) * 0.2
            oo, _, _, orew, odone, _ = oracle_tight.batch_step(envs, 1, a[-1:], 400)
            go, grew, gdone, _ = env.step(torch.as_tensor(a, device="cuda"))
            assert np.abs(go[-1].cpu().numpy() - oo[0]).max() < 1e-8 and abs(grew[-1].item() - orew[0]) < 1e-9
        env.close()


def test_reward_sum_finiteness_and_extreme_actions_at_full_size():
    env = _env(4096, precision=64, seed=1)
    env.reset()
    acts = _actions(4096, 6, seed=2)
    for k in range(6):
        a = acts[k]
        if k == 3:
            a = torch.full_like(a, 10.0)       # tests/test_environments.py:105-114 (+-10 stays finite)
        if k == 4:
            a = torch.full_like(a, -10.0)
        obs, rew, done, ended = env.step(a)
        assert torch.isfinite(obs).all() and torch.isfinite(rew).all() and torch.isfinite(env.state_r).all()
        assert (env.rew_terms.sum(1) - rew).abs().max() < 1e-9     # abs(total - sum(dict)) < 1e-6 in the reference's test
        assert (env.status_flags() == 0).all()
    env.close()


def test_truncation_autoreset_and_episode_bookkeeping():
    """rl/workers/rollout_worker.py:146-176: ended = done or traj_len >= max_traj_len; obs is post-reset, term_obs pre-reset;
    episode statistics are reported once, for completed episodes only."""
    n, L = 2048, 7
    env = _env(n, precision=64, seed=5, max_traj_len=L)
    env.reset()
    zeros = torch.zeros(n, 12, device="cuda", dtype=torch.float64)
    ep_rew = torch.zeros(n, device="cuda", dtype=torch.float64)
    for k in range(1, 2 * L + 1):
        obs, rew, done, ended = env.step(zeros)
        ep_rew += rew
        if k % L:
            assert int(ended.sum()) == int(done.sum())            # nothing but (rare) falls ends before the limit
        else:
            alive = done == 0
            assert bool((ended[alive] == 1).all())                 # everyone still standing is truncated at the limit
            assert bool((env.ep_len[alive] == L).all())
            assert (env.ep_rew[alive] - ep_rew[alive]).abs().max() < 1e-9
            assert (env.term_obs[alive] - obs[alive]).abs().max() > 1e-3   # pre-reset vs post-reset observation differ
            assert bool((env.state_i[:, 2] == 0).all())            # traj_len restarted
            ep_rew.zero_()
    env.close()


def test_fp32_and_fp64_agree_statistically_at_full_size():
    e64, e32 = _env(4096, precision=64, seed=2), _env(4096, precision=32, seed=2)
    e64.reset(); e32.reset()
    acts = _actions(4096, 4, seed=4)
    for k in range(4):
        o64, r64, d64, _ = e64.step(acts[k])
        o32, r32, d32, _ = e32.step(acts[k].float())
    assert (d64 == d32).float().mean() > 0.999
    assert (o64 - o32.double()).abs().mean() < 1e-4 and (r64 - r32.double()).abs().mean() < 1e-4
    e64.close(); e32.close()
