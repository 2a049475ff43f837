"""PPO.evaluate (rl/algos/ppo.py:408-426) + ModelCheckpointer.save_if_best (rl/utils/checkpointer.py:54-83) on the CPU: the method
is run unbound on a stand-in object whose sampler returns scripted batches, so the averaging over completed episodes, the
"always suffixed, un-suffixed only when improved" rule and the no-episode case are checked without a device.  The device
sampler it calls in production is covered by tests/test_gpu_ppo.py / test_gpu_entrypoint.py."""
import math
from types import SimpleNamespace

import torch

from learninghumanoidwalking_b200.rl.ppo import PPO


class _Net:
    def __init__(self):
        self.mode = "train"

    def eval(self):
        self.mode = "eval"


def _stand_in(batches):
    calls, saved = [], []
    it = iter(batches)

    def sample(deterministic=False):
        calls.append(deterministic)
        rew, lens = next(it)
        return SimpleNamespace(ep_rewards=torch.tensor(rew, dtype=torch.float32), ep_lens=torch.tensor(lens, dtype=torch.int64))

    me = SimpleNamespace(sample_parallel_with_workers=sample, world=1, rank=0, device=torch.device("cpu"), _best_eval=float("-inf"),
                         save=lambda itr: saved.append(itr))
    return me, calls, saved


def test_evaluate_averages_completed_episodes_and_keeps_the_best_pair():
    nets = {"actor": _Net(), "critic": _Net()}
    me, calls, saved = _stand_in([([10.0, 20.0], [40, 50]), ([], []), ([30.0], [30]), ([], []), ([20.0], [40])] +
                                 [([1.0], [5])] * 5 + [([], [])] * 5 + [([50.0, 70.0], [10, 20])] + [([], [])] * 4)
    batches, rew, ln = PPO.evaluate(me, None, nets, 0)
    assert len(batches) == 5 and calls == [True] * 5                      # five deterministic batches
    assert all(n.mode == "eval" for n in nets.values())
    assert rew == 20.0 and ln == 40.0                                     # means over the 4 completed episodes, not over batches
    assert saved == [0, None] and me._best_eval == 20.0                   # suffixed pair always; first result is the best so far
    _, rew, _ = PPO.evaluate(me, None, nets, 99)
    assert rew == 1.0 and saved == [0, None, 99] and me._best_eval == 20.0     # worse: only the suffixed pair
    _, rew, ln = PPO.evaluate(me, None, nets, 199)
    assert math.isnan(rew) and math.isnan(ln) and saved[-1] == 199 and me._best_eval == 20.0   # no episode: never "best"
    _, rew, _ = PPO.evaluate(me, None, nets, 299)
    assert rew == 60.0 and saved[-2:] == [299, None] and me._best_eval == 60.0


def test_evaluate_on_other_ranks_never_writes():
    me, _, saved = _stand_in([([5.0], [7])] * 5)
    me.rank = 1
    _, rew, ln = PPO.evaluate(me, None, {}, 0)
    assert (rew, ln) == (5.0, 7.0) and saved == []


def test_train_loop_host_logic_runs_end_to_end_with_the_device_operations_stubbed(tmp_path, monkeypatch, capsys):
    """PPO.train (rl/algos/ppo.py:428-641) on the CPU with everything that needs the CUDA library replaced by stand-ins (sampler,
    advantage kernel, gather kernel, optimiser step): the loop itself, its stdout lines, its log records, the evaluation pass at
    iteration 0 and every eval_freq, and the checkpoint files are the real code.  The device operations are tested where they
    can run (tests/test_gpu_ppo.py)."""
    from learninghumanoidwalking_b200.rl.policies import FF_V, Gaussian_FF_Actor
    from learninghumanoidwalking_b200.rl.storage import BatchData
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setenv("LHW_TENSORBOARD", "0")
    N, T, OBS, ACT = 6, 8, 37, 12
    ppo = object.__new__(PPO)
    ppo.policy, ppo.critic = Gaussian_FF_Actor(OBS, ACT, init_std=0.223, learn_std=False), FF_V(OBS)
    for net in (ppo.policy, ppo.critic):
        net.obs_mean, net.obs_std = torch.zeros(OBS), torch.ones(OBS)
    ppo.old_policy = Gaussian_FF_Actor(OBS, ACT, init_std=0.223, learn_std=False)
    ppo.old_policy.obs_mean, ppo.old_policy.obs_std = torch.zeros(OBS), torch.ones(OBS)
    ppo.env = SimpleNamespace()
    ppo.device, ppo.world, ppo.rank, ppo.seed = torch.device("cpu"), 1, 0, 5
    ppo.epochs, ppo.minibatch_size, ppo.mirror_coeff, ppo.eval_freq, ppo.eval_batches = 2, 16, 0.0, 2, 5
    ppo.total_steps, ppo.iteration_count, ppo.save_path, ppo._best_eval = 0, 0, tmp_path, float("-inf")
    ppo.actor_optimizer = ppo.critic_optimizer = object()
    calls = dict(det=0, train=0, updates=0)
    g = torch.Generator().manual_seed(0)

    def sample(deterministic=False):
        calls["det" if deterministic else "train"] += 1
        k = calls["det"] + calls["train"]
        z = lambda *s: torch.randn(*s, generator=g)
        return BatchData(states=z(N * T, OBS), actions=z(N * T, ACT), rewards=z(N * T, 1), values=z(N * T, 1), returns=z(N * T, 1),
                         dones=torch.zeros(N * T, 1), traj_idx=torch.zeros(1), ep_lens=torch.tensor([8, 4 + k % 3]),
                         ep_rewards=torch.tensor([1.0 * k, 2.0]))

    def update(ob, ab, rb, db, om, am):
        calls["updates"] += 1
        return torch.tensor([0.1, 0.2, 0.3, 0.01, 0.0, 0.0, 0.05])

    ppo.sample_parallel_with_workers = sample
    ppo.normalize_advantages = lambda ret, val, from_rollout=False: (ret - val - (ret - val).mean()) / ((ret - val).std() + 1e-5)
    ppo.gather_minibatch = lambda o, a, r, d, idx: (o[idx], a[idx], r[idx], d[idx])
    ppo._update_step = update
    log = ppo.train(None, 4, verbose=True)
    out = capsys.readouterr().out
    assert len(log) == 4 and calls == dict(det=15, train=4, updates=4 * 2 * (N * T // 16))   # evaluation at iterations 0, 1, 3
    assert [("eval_rew" in r) for r in log] == [True, True, False, True]
    assert out.count("====EVALUATE EPISODE====") == 3 and out.count("********** Iteration") == 4
    assert f"Sampling took" in out and f"for {N * T} steps." in out and "| " in out and "Mean Eprew" in out
    names = sorted(p.name for p in tmp_path.iterdir())
    assert names == ["actor.pt", "actor_0.pt", "actor_1.pt", "actor_3.pt", "critic.pt", "critic_0.pt", "critic_1.pt", "critic_3.pt"]
    assert abs(log[0]["actor_loss"] - 0.1) < 1e-6 and abs(log[0]["clip_frac"] - 0.05) < 1e-6 and log[-1]["fps"] > 0
    assert ppo._best_eval == max(r["eval_rew"] for r in log if "eval_rew" in r)
    best = torch.load(tmp_path / "actor.pt", weights_only=False)
    assert isinstance(best, Gaussian_FF_Actor)
