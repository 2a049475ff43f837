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
    import re
    sums = {}
    for fused in ("1", "0"):
        for one_step in ("0", "1"):
            env = dict(os.environ, LHW_FUSED_EXCHANGE=fused, LHW_CHECK_ONE_STEP=one_step)
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
            assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
            assert "identical_weights=True" in out.stdout and "ranks_simulate_different_envs=True" in out.stdout
            assert f"fused_exchange={fused == '1'}" in out.stdout
            if one_step == "1":
                sums[fused] = [float(x) for x in re.search(r"wsum=(\S+) wabs=(\S+)", out.stdout).groups()]
    # BASELINE configs[2]: the stepping-stone task sharded over the ranks, fused NVLink exchange
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, LHW_FUSED_EXCHANGE="1", LHW_CHECK_ONE_STEP="0", LHW_MODEL="jvrc_step"))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "model=jvrc_step" in out.stdout and "identical_weights=True" in out.stdout
    assert "ranks_simulate_different_envs=True" in out.stdout
    # after ONE optimiser step on identical data the fused NVLink kernel and the NCCL + clip/Adam baseline agree to rounding
    # (2 ranks: a + b is order-free; only the norm reductions differ in the last bits)
    assert abs(sums["1"][0] - sums["0"][0]) < 1e-4 and abs(sums["1"][1] - sums["0"][1]) < 1e-4, sums
