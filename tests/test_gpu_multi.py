"""N > 1 on real GPUs (skipped where the box has fewer): torchrun + NCCL, env sharding, one gradient exchange per optimiser
step (fused peer-memory kernels, replayed from the CUDA graph of the update; NCCL baseline), replicas bit-identical.
The host logic is covered on CPU by test_multi_rank_gloo.py."""
import os
import re
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, port, **env):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "ppo_dist_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, **env))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert "identical_weights=True" in out.stdout and "ranks_simulate_different_envs=True" in out.stdout
    return out.stdout


@pytest.mark.parametrize("world", [2, 4, 8])
def test_n_rank_ppo_replicas_stay_identical(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs >= {world} GPUs")
    sums = {}
    for fused in ("1", "0"):
        # two training iterations: sampling on sharded envs, optimiser steps replayed from the captured graph on every rank
        if fused == "1" or world == 2:
            out = _run(world, 29641 + world, LHW_FUSED_EXCHANGE=fused, LHW_CHECK_ONE_STEP="0")
            # the fused exchange is replayed from the captured update graph on every rank; the NCCL baseline runs eagerly
            assert f"fused_exchange={fused == '1'}" in out and f"update_graph={fused == '1'}" in out
        out = _run(world, 29641 + world, LHW_FUSED_EXCHANGE=fused, LHW_CHECK_ONE_STEP="1")
        sums[fused] = [float(x) for x in re.search(r"wsum=(\S+) wabs=(\S+)", out).groups()]
    if world == 2:
        # BASELINE configs[2]: the stepping-stone task sharded over the ranks, fused NVLink exchange
        out = _run(world, 29641 + world, LHW_FUSED_EXCHANGE="1", LHW_CHECK_ONE_STEP="0", LHW_MODEL="jvrc_step")
        assert "model=jvrc_step" in out
        # the eager loop (no graph) gives the same replicas
        out = _run(world, 29641 + world, LHW_FUSED_EXCHANGE="1", LHW_CHECK_ONE_STEP="0", LHW_UPDATE_GRAPH="0")
        assert "update_graph=False" in out
    # after ONE optimiser step on identical data the fused peer-memory kernels and the NCCL + clip/Adam baseline agree to
    # rounding (different summation orders over the ranks and in the norm reductions)
    assert abs(sums["1"][0] - sums["0"][0]) < 1e-4 and abs(sums["1"][1] - sums["0"][1]) < 1e-4, sums


def test_uneven_shards_are_rejected():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29655", os.path.join(ROOT, "tools", "ppo_dist_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, LHW_ENVS="257"))
    assert out.returncode != 0 and "multiple of the world size" in (out.stdout + out.stderr)
