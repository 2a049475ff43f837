"""Edge shapes of the PPO data-path kernels through the C-ABI against oracle/ppo_oracle.py: single-step / single-env rollouts, batch
sizes that are not multiples of a warp, a block or a float4, unaligned buffers, every step ending an episode, repeated minibatch
indices, tiny parameter vectors, and empty inputs (the reference's PPOBuffer takes any path length >= 1,
rl/storage/rollout_storage.py:53-85; its sampler may hand out any index list, rl/algos/ppo.py:504-539)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _L():
    from learninghumanoidwalking_b200 import _lib
    return _lib


def _gae(rew, val, ended, boot, last, gamma=0.99, lam=0.95):
    L = _L()
    T, N = rew.shape
    d = lambda a, dt=torch.float32: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device="cuda")
    r, v, e, b, lv = d(rew), d(val), d(ended, torch.int32), d(boot), d(last)
    ret = torch.full_like(r, float("nan"))
    part = torch.zeros(max(2, L.lib().lhw_gae_partial_words(N)), dtype=torch.float64, device="cuda")
    L.check(L.lib().lhw_gae(r.data_ptr(), v.data_ptr(), e.data_ptr(), b.data_ptr(), lv.data_ptr(), ret.data_ptr(), T, N, gamma, lam,
                            part.data_ptr(), L.current_stream_ptr()))
    return ret.cpu().numpy(), part.cpu().numpy()


@pytest.mark.parametrize("T,N", [(1, 1), (1, 33), (7, 1), (3, 513), (5, 31), (401, 257)])
@pytest.mark.parametrize("ends", ["none", "all", "some"])
def test_gae_any_rollout_shape(T, N, ends):
    from oracle.ppo_oracle import gae_rollout
    rng = np.random.RandomState(T * 1000 + N)
    rew, val = rng.uniform(-1, 1, (T, N)), rng.uniform(-2, 2, (T, N))
    ended = {"none": np.zeros((T, N)), "all": np.ones((T, N)), "some": rng.rand(T, N) < 0.3}[ends].astype(np.int32)
    boot = rng.uniform(-2, 2, (T, N)) * (rng.rand(T, N) < 0.5)
    last = rng.uniform(-2, 2, N)
    f = lambda a: a.astype(np.float32).astype(np.float64)
    ret, part = _gae(rew, val, ended, boot, last)
    exp = gae_rollout(f(rew), f(val), ended, f(boot), f(last), 0.99, 0.95)
    assert np.isfinite(ret).all() and np.abs(ret - exp).max() < 1e-4
    adv = exp - f(val)
    nb = len(part) // 2
    assert abs(part[:nb].sum() - adv.sum()) < 1e-4 * max(1.0, np.abs(adv).sum())
    assert abs(part[nb:].sum() - (adv * adv).sum()) < 1e-4 * max(1.0, (adv * adv).sum())


def test_gae_empty_rollout_is_a_no_op():
    L = _L()
    z = torch.zeros(4, device="cuda")
    zi = torch.zeros(4, dtype=torch.int32, device="cuda")
    for T, N in ((0, 4), (4, 0), (0, 0)):
        assert L.lib().lhw_gae(z.data_ptr(), z.data_ptr(), zi.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(), T, N, 0.99, 0.95,
                               None, L.current_stream_ptr()) == 0
    torch.cuda.synchronize()
    assert float(z.abs().sum()) == 0.0


@pytest.mark.parametrize("n", [2, 3, 5, 255, 1023, 4097, 65537])
@pytest.mark.parametrize("offset", [0, 1])
def test_advantage_normalisation_any_count_and_alignment(n, offset):
    """count not a multiple of 4 and buffers that start 4 bytes off a 16-byte boundary take the scalar path of adv_apply_kernel;
    the unbiased std needs n >= 2 (torch's .std() of one sample is NaN in the reference too, rl/algos/ppo.py:484-485)."""
    from oracle.ppo_oracle import adv_normalize
    L = _L()
    rng = np.random.RandomState(n + offset)
    ret, val = rng.normal(size=n).astype(np.float32), (rng.normal(size=n) * 0.5 + 0.1).astype(np.float32)
    rbuf, vbuf, abuf = (torch.zeros(n + 8, device="cuda") for _ in range(3))
    r, v, a = rbuf[offset:offset + n], vbuf[offset:offset + n], abuf[offset:offset + n]
    r.copy_(torch.as_tensor(ret))
    v.copy_(torch.as_tensor(val))
    stats = torch.zeros(L.lib().lhw_adv_stats_words(), dtype=torch.float64, device="cuda")
    st = L.current_stream_ptr()
    L.check(L.lib().lhw_adv_stats(r.data_ptr(), v.data_ptr(), stats.data_ptr(), n, st))
    L.check(L.lib().lhw_adv_apply(r.data_ptr(), v.data_ptr(), a.data_ptr(), stats.data_ptr(), n, n, 1e-5, st))
    exp = adv_normalize(ret, val, 1e-5)
    assert np.abs(a.cpu().numpy() - exp).max() < 2e-5 * max(1.0, np.abs(exp).max())
    assert float(abuf[:offset].abs().sum()) == 0.0 and float(abuf[offset + n:].abs().sum()) == 0.0     # nothing written outside


@pytest.mark.parametrize("B,obs_dim,act_dim", [(1, 37, 12), (3, 39, 12), (257, 35, 10), (4096, 37, 12)])
def test_gather_any_index_list(B, obs_dim, act_dim):
    L = _L()
    g = torch.Generator(device="cuda").manual_seed(B)
    n = 1000
    obs, act = torch.randn(n, obs_dim, device="cuda", generator=g), torch.randn(n, act_dim, device="cuda", generator=g)
    ret, adv = torch.randn(n, 1, device="cuda", generator=g), torch.randn(n, 1, device="cuda", generator=g)
    idx = torch.randint(0, n, (B,), device="cuda", generator=g)            # with repeats (B may exceed n)
    idx[0], idx[-1] = n - 1, 0                                              # both ends of the buffer
    o, a, r, d = (torch.empty(B, obs_dim, device="cuda"), torch.empty(B, act_dim, device="cuda"), torch.empty(B, 1, device="cuda"),
                  torch.empty(B, 1, device="cuda"))
    L.check(L.lib().lhw_gather_minibatch(obs.data_ptr(), act.data_ptr(), ret.data_ptr(), adv.data_ptr(), idx.data_ptr(), o.data_ptr(),
                                         a.data_ptr(), r.data_ptr(), d.data_ptr(), B, obs_dim, act_dim, L.current_stream_ptr()))
    assert torch.equal(o, obs[idx]) and torch.equal(a, act[idx]) and torch.equal(r, ret[idx]) and torch.equal(d, adv[idx])
    assert L.lib().lhw_gather_minibatch(obs.data_ptr(), act.data_ptr(), ret.data_ptr(), adv.data_ptr(), idx.data_ptr(), o.data_ptr(),
                                        a.data_ptr(), r.data_ptr(), d.data_ptr(), 0, obs_dim, act_dim, L.current_stream_ptr()) == 0


@pytest.mark.parametrize("n", [1, 5, 33, 1025])
def test_clip_adam_any_parameter_count(n):
    from oracle.ppo_oracle import clip_adam
    L = _L()
    rng = np.random.RandomState(n)
    p0, g0 = rng.normal(size=n).astype(np.float32), (rng.normal(size=n) * 3).astype(np.float32)
    p, g = torch.as_tensor(p0, device="cuda"), torch.as_tensor(g0, device="cuda")
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    scratch = torch.zeros(4, device="cuda")
    pe, me, ve = p0.astype(np.float64), np.zeros(n), np.zeros(n)
    st = L.current_stream_ptr()
    for step in (1, 2, 3):
        L.check(L.lib().lhw_grad_sumsq(g.data_ptr(), scratch.data_ptr(), n, 1.0, st))
        L.check(L.lib().lhw_clip_adam(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), scratch.data_ptr(), n, step, 3e-4,
                                      0.9, 0.999, 1e-5, 0.05, 1.0, st))
        pe, me, ve, tn = clip_adam(pe, g0.astype(np.float64), me, ve, step, 3e-4, 1e-5, 0.05)
        assert np.abs(p.double().cpu().numpy() - pe).max() < 2e-6
