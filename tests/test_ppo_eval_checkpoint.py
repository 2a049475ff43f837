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
