#!/usr/bin/env python3
"""Timeline model of a lock-step block (development tool, no GPU): how much of a warp's time goes to the per-substep rendez-vous.

A warp's substep = pre-solver phases (a) + n Newton iterations (c each) + post-solver phases (b), with the per-substep iteration
counts drawn from the distribution recorded from the oracle on the bench workload (63 % one, 27 % two, 7 % three, 3 % four;
independent per substep, DESIGN.md §4.1 finding 6).  Modes:
  lock   __syncthreads() at the end of every substep (the shipped kernel; measured stall_barrier share 21-25 %)
  split  arrive at the end of the substep, wait before the solver of the next one for everybody's arrival (LHW_X_SPLITBAR, mode 4)
  late   the same with the wait after the solver (mode 8)
  none   no rendez-vous (only the end of the control step)
Prints the share of warp time spent waiting, the time of a 25-substep control step in units of a one-iteration substep, and the
spread between the first and the last warp of a block (what instruction-cache sharing depends on)."""
import sys

import numpy as np


def run(W, mode, a=0.5, b=0.25, c=0.25, S=25, blocks=3000, seed=0):
    rng = np.random.RandomState(seed)
    p, iters = np.array([0.63, 0.27, 0.07, 0.03]), np.arange(1, 5)
    wait = time = 0.0
    spread = []
    for _ in range(blocks):
        n = rng.choice(iters, size=(W, S), p=p)
        t, fin_prev = np.zeros(W), None
        for s in range(S):
            t = t + a
            if mode == "split" and fin_prev is not None:
                w_ = np.maximum(fin_prev.max() - t, 0)
                wait += w_.sum(); t = t + w_
            t = t + c * n[:, s]
            if mode == "late" and fin_prev is not None:
                w_ = np.maximum(fin_prev.max() - t, 0)
                wait += w_.sum(); t = t + w_
            t = t + b
            if mode == "lock":
                wait += (t.max() - t).sum(); t[:] = t.max()
            fin_prev = t.copy()
            spread.append(t.max() - t.min())
        wait += (t.max() - t).sum()
        time += t.max() * W
    return wait / time, time / blocks / W, float(np.mean(spread))


if __name__ == "__main__":
    widths = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [4, 5, 8]
    for W in widths:
        for mode in ("lock", "split", "late", "none"):
            share, T, sp = run(W, mode)
            print(f"{W} warps/block  {mode:5s}  waiting {100 * share:5.1f} %   control step {T:6.2f}   spread before the barrier {sp:.2f} substeps")
