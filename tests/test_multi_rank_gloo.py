"""world_size-2 gloo test (CPU) of the host-side logic of the N > 1 path: env sharding, identical minibatch index
streams on every rank, summed flat gradient / world == mean gradient, global advantage statistics, evaluation means over all
ranks with rank 0 as the only checkpoint writer."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from learninghumanoidwalking_b200.rl.dist_utils import allreduce_sum_, env_shard, global_mean_std
    first, n = env_shard(rank, world, 10)
    g = torch.Generator().manual_seed(123)
    full = torch.randn(10, 7, generator=g, dtype=torch.float64)     # same on every rank
    local = full[first:first + n]
    # gradient of a sum-loss over local samples; averaged across ranks must equal the global mean of per-rank grads
    grad = local.sum(0).float().clone()
    allreduce_sum_(grad)
    grad /= world
    mean, std = global_mean_std(local.sum(), (local * local).sum(), local.numel())
    # identical permutation stream (seed + itr*epochs + epoch)
    gi = torch.Generator().manual_seed(7 + 3 * 4 + 1)
    perm = torch.randperm(50, generator=gi)
    # the evaluation pass (PPO.evaluate): every rank samples its own shard, the mean is over ALL ranks' completed episodes and
    # only rank 0 writes checkpoints; rank 1 completes no episode at all here
    from types import SimpleNamespace
    from learninghumanoidwalking_b200.rl.ppo import PPO
    saved = []
    ep = ([10.0, 30.0], [40, 20]) if rank == 0 else ([], [])
    me = SimpleNamespace(world=world, rank=rank, device=torch.device("cpu"), _best_eval=float("-inf"), save=saved.append,
                         sample_parallel_with_workers=lambda deterministic=False: SimpleNamespace(
                             ep_rewards=torch.tensor(ep[0], dtype=torch.float32), ep_lens=torch.tensor(ep[1], dtype=torch.int64)))
    _, ev_rew, ev_len = PPO.evaluate(me, None, {}, 7, num_batches=2)
    out[rank] = dict(first=first, n=n, grad=grad.numpy(), mean=float(mean), std=float(std), perm=perm.numpy(),
                     eval=(ev_rew, ev_len, list(saved)),
                     expect_grad=(full[:5].sum(0) + full[5:].sum(0)).float().numpy() / 2,
                     expect_mean=float(full.mean()), expect_std=float(full.std()))
    dist.destroy_process_group()


def test_two_rank_host_logic():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, 29731, out), nprocs=world, join=True)
    assert (out[0]["first"], out[0]["n"], out[1]["first"], out[1]["n"]) == (0, 5, 5, 5)
    for r in range(world):
        assert np.allclose(out[r]["grad"], out[r]["expect_grad"], atol=1e-6)
        assert abs(out[r]["mean"] - out[r]["expect_mean"]) < 1e-12 and abs(out[r]["std"] - out[r]["expect_std"]) < 1e-12
    assert (out[0]["perm"] == out[1]["perm"]).all()
    assert out[0]["eval"] == (20.0, 30.0, [7, None]) and out[1]["eval"] == (20.0, 30.0, [])


def test_env_shard_covers_everything_once():
    from learninghumanoidwalking_b200.rl.dist_utils import env_shard
    for world in (1, 2, 3, 8):
        for n in (8, 4096, 4099):
            ids = []
            for r in range(world):
                f, k = env_shard(r, world, n)
                ids += list(range(f, f + k))
            assert ids == list(range(n))
