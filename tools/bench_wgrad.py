#!/usr/bin/env python3
"""lhw_linear_wgrad against what autograd runs for the same Linear layer (mm(gy^T, x) + sum(gy, 0)) at the shapes of one PPO update
(`run_experiment.py` defaults at 4096 envs: minibatch 21 845, actor batch doubled by the mirror pass).  CUDA events, L2 flushed
before every timed call, median of 20; fp32 FFMA peak 148 SMs x 128 lanes x 2 x 1.965 GHz = 74.4 TFLOP/s.  One JSON object."""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from learninghumanoidwalking_b200 import _lib  # noqa: E402

PEAK_TF = 148 * 128 * 2 * 1.965e9 / 1e12


def timed(fn, flush, reps=20, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def main():
    L, ops = _lib.lib(), _lib.ops()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    out, tot_k, tot_t = {}, 0.0, 0.0
    for name, M, N, K in (("actor_l2", 43690, 256, 256), ("actor_l1", 43690, 256, 37), ("actor_out", 43690, 12, 256),
                          ("critic_l2", 21845, 256, 256), ("critic_l1", 21845, 256, 37), ("critic_out", 21845, 1, 256)):
        gy, x = torch.randn(M, N, device="cuda", generator=g), torch.randn(M, K, device="cuda", generator=g)
        gw, gb = torch.empty(N, K, device="cuda"), torch.empty(N, device="cuda")
        ws = torch.empty(L.lhw_linear_wgrad_workspace_floats(M, N, K), device="cuda")
        mk = timed(lambda: ops.linear_wgrad(gy, x, gw, gb, ws), flush)
        mt = timed(lambda: (gy.t().mm(x), gy.sum(0)), flush)
        ref = gy.double().t().mm(x.double())
        err = float((gw.double() - ref).abs().max() / gy.abs().double().t().mm(x.abs().double()).max())
        fl = 2.0 * M * N * K
        out[name] = {"M": M, "N": N, "K": K, "kernel_ms": mk, "torch_mm_plus_sum_ms": mt, "kernel_tflops": fl / mk / 1e9,
                     "frac_of_fp32_peak": fl / mk / 1e9 / PEAK_TF, "rel_err_vs_fp64": err}
        tot_k += mk
        tot_t += mt
    out["sum_kernel_ms"], out["sum_torch_ms"], out["fp32_peak_tflops"] = tot_k, tot_t, PEAK_TF
    print(json.dumps(out))


if __name__ == "__main__":
    main()
