#!/usr/bin/env python3
"""Achieved HBM GB/s of the PPO data-path kernels (GAE 20 B/sample, advantage normalisation
12 B/sample, minibatch gather, clip+Adam 28 B/param) at the BASELINE batch geometry (T=400 x N=4096 samples),
against the measured copy bandwidth in MEASURED_PEAKS.json.  CUDA events on the launching stream, warm-up, L2 flushed
(256 MiB write) before every timed launch; prints one JSON object."""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from learninghumanoidwalking_b200 import _lib  # noqa: E402


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
    L, st = _lib.lib(), _lib.current_stream_ptr()
    dev = "cuda"
    T, N = 400, 4096
    n = T * N
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    rew, val, boot = (torch.randn(T, N, device=dev, generator=g) for _ in range(3))
    ended = (torch.rand(T, N, device=dev, generator=g) < 0.01).int()
    last, ret, adv = torch.randn(N, device=dev, generator=g), torch.empty(T, N, device=dev), torch.empty(T, N, device=dev)
    out = {"peak_gbs": peak, "T": T, "N": N}

    part = torch.zeros(L.lhw_gae_partial_words(N), dtype=torch.float64, device=dev)
    ms = timed(lambda: L.lhw_gae(rew.data_ptr(), val.data_ptr(), ended.data_ptr(), boot.data_ptr(), last.data_ptr(), ret.data_ptr(), T, N, 0.99, 0.95,
                                 part.data_ptr(), st), flush)
    out["gae"] = {"ms": ms, "bytes": 20 * n, "gbs": 20 * n / ms / 1e6, "frac": 20 * n / ms / 1e6 / peak,
                  "note": "incl. the advantage statistics (sum, sumsq per block) it leaves behind for the normalisation"}

    stats = torch.zeros(L.lhw_adv_stats_words(), dtype=torch.float64, device=dev)
    def advnorm():
        L.lhw_adv_stats(ret.data_ptr(), val.data_ptr(), stats.data_ptr(), n, st)
        L.lhw_adv_apply(ret.data_ptr(), val.data_ptr(), adv.data_ptr(), stats.data_ptr(), n, n, 1e-5, st)
    ms = timed(advnorm, flush)
    # two passes over (returns, values) + one write of adv = 20 B/sample for this two-pass implementation (12 B is the one-pass ideal)
    out["adv_norm_two_pass"] = {"ms": ms, "bytes": 20 * n, "gbs": 20 * n / ms / 1e6, "frac": 20 * n / ms / 1e6 / peak,
                                "note": "stand-alone statistics pass + apply (2 x 8 B read + 4 B write per sample): the path for batches that did not come from the rollout"}

    def advnorm_fused():      # what PPO.train runs: statistics from the GAE launch, one 12 B/sample pass
        L.lhw_adv_stats_from_gae(part.data_ptr(), N, stats.data_ptr(), st)
        L.lhw_adv_apply(ret.data_ptr(), val.data_ptr(), adv.data_ptr(), stats.data_ptr(), n, n, 1e-5, st)
    ms = timed(advnorm_fused, flush)
    out["adv_norm"] = {"ms": ms, "bytes": 12 * n, "gbs": 12 * n / ms / 1e6, "frac": 12 * n / ms / 1e6 / peak,
                       "note": "statistics from the GAE launch + one pass: 8 B read + 4 B write per sample (SURVEY 8d's 12 B/sample)"}

    obs, act = torch.randn(n, 37, device=dev, generator=g), torch.randn(n, 12, device=dev, generator=g)
    r1, a1 = ret.reshape(n, 1), adv.reshape(n, 1)
    for B in (64, 32768):
        idx = torch.randperm(n, device=dev)[:B]
        o, a, r, d = (torch.empty(B, 37, device=dev), torch.empty(B, 12, device=dev), torch.empty(B, 1, device=dev), torch.empty(B, 1, device=dev))
        ms = timed(lambda: L.lhw_gather_minibatch(obs.data_ptr(), act.data_ptr(), r1.data_ptr(), a1.data_ptr(), idx.data_ptr(), o.data_ptr(),
                                                  a.data_ptr(), r.data_ptr(), d.data_ptr(), B, 37, 12, st), flush)
        by = B * (51 * 4 * 2 + 8)
        out[f"gather_B{B}"] = {"ms": ms, "bytes": by, "gbs": by / ms / 1e6, "frac": by / ms / 1e6 / peak}

    npar = 154381
    p, gr, m, v = (torch.randn(npar, device=dev, generator=g) for _ in range(4))
    v.abs_()
    norm = torch.zeros(1, device=dev)
    def clip_adam():
        L.lhw_grad_sumsq(gr.data_ptr(), norm.data_ptr(), npar, 1.0, st)
        L.lhw_clip_adam(p.data_ptr(), gr.data_ptr(), m.data_ptr(), v.data_ptr(), norm.data_ptr(), npar, 5, 3e-4, 0.9, 0.999, 1e-5, 0.05, 1.0, st)
    ms = timed(clip_adam, flush)
    out["clip_adam_2launch"] = {"ms": ms, "bytes": 32 * npar, "gbs": 32 * npar / ms / 1e6, "frac": 32 * npar / ms / 1e6 / peak,
                                "note": "617 KB working set: launch-latency bound, not bandwidth bound"}
    from learninghumanoidwalking_b200.rl.comm import PeerComm
    comm = PeerComm(npar, torch.device("cuda", 0))
    comm.grad.copy_(gr)
    ms = timed(lambda: comm.fused_step(p, m, v, 78604, 3e-4, (0.9, 0.999), 1e-5, 0.05), flush)
    out["fused_exchange_clip_adam_3launch_world1"] = {"ms": ms, "bytes": 36 * npar, "gbs": 36 * npar / ms / 1e6, "frac": 36 * npar / ms / 1e6 / peak}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
