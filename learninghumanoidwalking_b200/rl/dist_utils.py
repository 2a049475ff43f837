"""Host-side logic of the multi-GPU path (one process per GPU, torch.distributed; NCCL on the box, gloo in the
CPU tests).  Env copies shard by index across ranks, the only exchange per optimiser step is one all-reduce of
the flat gradient; advantage statistics are global (2 doubles per iteration)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def env_shard(rank: int, world: int, global_envs: int) -> tuple[int, int]:
    """(first_env_id, n_local): contiguous blocks, remainder to the lowest ranks; ids key the per-env RNG streams,
    so a run with W ranks simulates exactly the same environments as a single-rank run with global_envs."""
    base, rem = divmod(global_envs, world)
    n = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, n


def allreduce_sum_(flat: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def global_mean_std(local_sum: torch.Tensor, local_sumsq: torch.Tensor, local_count: int, group=None):
    """Unbiased global mean/std from per-rank (sum, sumsq, count) — rl/algos/ppo.py:484-485 is a global statistic."""
    t = torch.stack([local_sum.double(), local_sumsq.double(), torch.tensor(float(local_count), dtype=torch.float64, device=local_sum.device)])
    allreduce_sum_(t, group)
    n = t[2]
    mean = t[0] / n
    var = (t[1] - n * mean * mean) / (n - 1)
    return mean, var.clamp_min(0).sqrt()
