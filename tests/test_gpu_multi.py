"""N > 1 on real GPUs (skipped on a single-GPU box): torchrun + NCCL, env sharding, one gradient all-reduce per
optimiser step, replicas bit-identical.  The host logic is covered on CPU by test_multi_rank_gloo.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_rank_nccl_ppo_replicas_stay_identical():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29641", os.path.join(ROOT, "tools", "ppo_dist_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "identical_weights=True" in out.stdout and "ranks_simulate_different_envs=True" in out.stdout
