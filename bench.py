#!/usr/bin/env python3
"""bench.py — env-steps/sec of jvrc_walk (BASELINE.json metric) on N x B200, one JSON line on rank 0.

A "step" is one pass of the rollout hot path over one batch: one control step (25 physics substeps + reward +
observation + termination + auto-reset) for every environment of the batch — the work of
BaseHumanoidEnv.step x num_envs in the reference.  Workload: BASELINE.json configs[1], jvrc_walk, 4096
environments per GPU (weak scaling: envs shard by index, no data-path collective), actions ~ N(0, 0.223^2)
(the action distribution of the reference's freshly initialised Gaussian_FF_Actor: output layer x0.01,
std_dev 0.223), synthetic, generated up front.

  value   device-resident: actions already in HBM, one lhw_sim_step launch per step.
  e2e     the same steps through the host-facing API: actions from pinned host memory (H2D every step),
          observation / reward / done read back to pinned host memory (D2H every step).
  roofline  the step kernel: algorithmic HBM bytes per env-step x envs / CUDA-event time of
          the launches, against the measured HBM peak (MEASURED_PEAKS.json).  The kernel is ALU/latency bound;
          the honest secondary bound is reported beside it as roofline.issue (warp-instruction issue rate, instruction
          count per env-step from the committed ncu capture); roofline.traffic = DRAM bytes of that capture.
  cpu_baseline / --impl reference   the CPU restatement (oracle/, "port": the reference's own MuJoCo path is
          not installable here) on the box's host cores, same workload, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env-steps/sec jvrc_walk"
UNIT = "env-steps/s"
SIGMA = 0.223
# per-launch numbers of the committed ncu captures (profiles/r01_step_kernel_*.md; 4096 envs, the bench's action distribution):
# (DRAM bytes read + written, warp instructions executed).  Under ncu the state record is L2 resident when the launch starts
# (no flush between replays), so the DRAM traffic is BELOW the algorithmic bytes; nothing is re-read.
NCU_PER_LAUNCH_4096 = {("jvrc_walk", 64): (4.712192e6 + 0.103424e6, 761598198), ("jvrc_walk", 32): (2.526976e6 + 0.082432e6, 709974877),
                       ("jvrc_step", 64): (7.463424e6 + 0.455168e6, 1238509046)}   # end-of-round captures (profiles/*_end_of_round.md)
ALG_BYTES = {32: 1220, 64: 2288}   # state read+write, action read, obs/reward/done write


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 8 and r[4 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def effective_cpus() -> int:
    """Host cores this process may really use: affinity mask capped by the cgroup CPU quota (the GPU boxes expose 128
    logical CPUs but a container quota of 16)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            n = min(n, max(1, int(round(int(q) / int(per)))))
    except Exception:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, int(round(q / per))))
        except Exception:
            pass
    return n


WORKLOADS = {
    "jvrc_walk": dict(model="jvrc_walk", metric="env-steps/sec jvrc_walk",
                      desc="jvrc_walk {n} envs/GPU (BASELINE configs[1]), JVRC-1 sim_dt=0.001 control_dt=0.025 flat terrain"),
    "jvrc_step": dict(model="jvrc_step", metric="env-steps/sec jvrc_step",
                      desc="jvrc_step footstep-plan task {n} envs/GPU (BASELINE configs[2]), JVRC-1 sim_dt=0.001 control_dt=0.025, "
                           "20 stepping-stone slabs per env (footstep sequences, floor dropped in FORWARD mode, 0.1 m stairs: "
                           "iteration_count = inf) in the kernel"),
    "jvrc_walk_terrain": dict(model="jvrc_walk_terrain", metric="env-steps/sec jvrc_walk uneven/compliant terrain",
                              desc="jvrc_walk on uneven / compliant terrain {n} envs/GPU (BASELINE configs[4]; an EXTENSION — the reference "
                                   "has only the unused manip_hfield hook): 20 terraces re-posed with the hook's ranges, contact solref 0.04 s"),
    "h1": dict(model="h1", metric="env-steps/sec h1 standing",
               desc="h1 standing task {n} envs/GPU (BASELINE configs[3]), Unitree H1 sim_dt=0.001 control_dt=0.025, observation "
                    "noise + dynamics randomisation (damping, frictionloss, mass, CoM) + random pushes in the kernel"),
}


def cpu_reference(n_envs: int, seconds: float, warmup: int, seed: int, nthreads: int = 0, model: str = "jvrc_walk"):
    """The oracle (CPU port of the reference path) on the host cores, run for about `seconds` of wall time
    (a bounded sample of the same workload): env-steps/s, threads used, elapsed, control steps done."""
    import numpy as np
    from oracle.oracle import Oracle
    o = Oracle(model)
    nthreads = nthreads or effective_cpus()
    envs = o.make_envs(n_envs, seed=seed)
    o.batch_reset(envs, n_envs, nthreads)
    rng = np.random.RandomState(seed)
    for _ in range(warmup):
        o.batch_step(envs, n_envs, rng.normal(size=(n_envs, o.nu)) * SIGMA, 400, nthreads)
    steps, t0 = 0, time.perf_counter()
    while True:
        a = rng.normal(size=(n_envs, o.nu)) * SIGMA
        o.batch_step(envs, n_envs, a, 400, nthreads)
        steps += 1
        dt = time.perf_counter() - t0
        if dt >= seconds or steps >= 400:
            break
    return n_envs * steps / dt, nthreads, dt, steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=4096, help="environments per GPU")
    ap.add_argument("--precision", type=int, default=int(os.environ.get("LHW_BENCH_PRECISION", "64")), choices=[32, 64])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--workload", default="jvrc_walk", choices=sorted(WORKLOADS),
                    help="jvrc_walk: the configuration BASELINE.json's metric is quoted on (default); jvrc_step: configs[2]; h1: configs[3]; jvrc_walk_terrain: configs[4] (extension)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    K, W = args.steps, max(3, args.warmup)
    wl = WORKLOADS[args.workload]
    metric = wl["metric"]
    config = {"workload": wl["desc"].format(n=args.envs),
              "envs_per_gpu": args.envs, "global_envs": args.envs * world, "actions": f"N(0,{SIGMA}^2) synthetic, pre-generated",
              "parallelism": f"env-sharded x{world} (no data-path collective)"}

    if args.impl == "reference":
        # the reference's own Ray+MuJoCo path cannot be installed here (mujoco/ray absent, no network):
        # this arm times the CPU port (oracle/) on all host cores, rank 0 only.
        if rank != 0:
            return
        ncores = effective_cpus()
        n_sample = max(256, min(args.envs * world, 64 * ncores))
        sps, threads, dt, steps_ref = cpu_reference(n_sample, 15.0, 2, args.seed, model=wl["model"])
        print(json.dumps({"metric": metric, "value": sps, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
                          "ms_per_step": 1e3 * dt / steps_ref, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f64", "data": "synthetic", "impl": "reference", "config": config,
                          "cpu_baseline": {"value": sps, "unit": UNIT, "cores": threads, "kind": "port",
                                           "sample": f"{n_sample} envs x {steps_ref} control steps ({dt:.1f} s) after 2 warm-up steps, OpenMP over envs; "
                                                     f"threads = cgroup CPU quota ({threads} of {os.cpu_count()} logical CPUs); "
                                                     "reference Ray+MuJoCo path not runnable on this box (mujoco/ray not installable)"},
                          "e2e": {"value": sps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    from learninghumanoidwalking_b200 import _lib
    from learninghumanoidwalking_b200.envs import BatchedHumanoidEnv

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner on the process's stdout (fd 1) when NCCL_DEBUG=VERSION is set in the environment;
        # stdout carries exactly ONE JSON line, so fd 1 points at stderr while the communicator comes up
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    n = args.envs
    env = BatchedHumanoidEnv(n, model=wl["model"], precision=args.precision, seed=args.seed, first_env_id=rank * n,
                             device=local_rank)
    env.reset()
    A = env.act_dim
    g = torch.Generator(device=dev).manual_seed(args.seed * 1000 + rank)
    acts = torch.randn(K + W, n, A, device=dev, generator=g, dtype=env.dtype) * SIGMA
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for k in range(W):
        env.step(acts[k])
    launches0 = _lib.lib().lhw_launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- value: device resident, per-step CUDA events (L2 flushed before every step, flush not timed)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    t0 = time.perf_counter()
    for k in range(K):
        flush.fill_(k & 0xFF)
        ev[k][0].record()
        env.step(acts[W + k])
        ev[k][1].record()
    barrier()
    wall = time.perf_counter() - t0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = sum(step_ms)
    launches = _lib.lib().lhw_launch_count() - launches0
    # ---- e2e: pinned host actions in, pinned host obs/reward/done out, every step
    h_acts = torch.empty(K, n, A, dtype=env.dtype).pin_memory()
    h_acts.copy_(acts[W:W + K].cpu())
    h_obs = torch.empty(n, env.obs_dim, dtype=env.dtype).pin_memory()
    h_rew = torch.empty(n, dtype=env.dtype).pin_memory()
    h_done = torch.empty(n, dtype=torch.int32).pin_memory()
    d_act = torch.empty(n, A, dtype=env.dtype, device=dev)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(K):
        d_act.copy_(h_acts[k], non_blocking=True)
        obs, rew, done, _ = env.step(d_act)
        h_obs.copy_(obs, non_blocking=True)
        h_rew.copy_(rew, non_blocking=True)
        h_done.copy_(done, non_blocking=True)
        torch.cuda.current_stream().synchronize()   # the host consumer needs this step's result before the next action
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    # ---- extras (not the headline): the fp32 build of the same kernel, and the full rollout loop with the policy /
    # critic MLPs (cuBLAS) and buffer writes in the loop (DeviceRolloutWorker.sample)
    extras = {}
    if not args.no_extras:
        env32 = BatchedHumanoidEnv(n, model=wl["model"], precision=32, seed=args.seed, first_env_id=rank * n, device=local_rank)
        env32.reset()
        a32 = acts.float()
        for k in range(W):
            env32.step(a32[k])
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for k in range(K):
            env32.step(a32[W + k])
        f1.record()
        barrier()
        extras["fp32_kernel_env_steps_per_s_per_gpu"] = n * K / (f0.elapsed_time(f1) * 1e-3)
        from learninghumanoidwalking_b200.rl import FF_V, DeviceRolloutWorker, Gaussian_FF_Actor
        torch.manual_seed(args.seed)
        pol = Gaussian_FF_Actor(env.obs_dim, env.act_dim, init_std=SIGMA).to(dev)
        cri = FF_V(env.obs_dim).to(dev)
        pol.obs_mean = cri.obs_mean = torch.tensor(env.obs_mean, dtype=torch.float32, device=dev)
        pol.obs_std = cri.obs_std = torch.tensor(env.obs_std, dtype=torch.float32, device=dev)
        worker = DeviceRolloutWorker(env, pol, cri, seed=args.seed)
        T = max(4, min(K, 32))
        worker.sample(0.99, 0.95, T, 400)   # warm-up with the same horizon: captures the per-step CUDA graph
        barrier()
        f0.record()
        worker.sample(0.99, 0.95, T, 400)
        f1.record()
        barrier()
        extras["rollout_with_policy_env_steps_per_s_per_gpu"] = n * T / (f0.elapsed_time(f1) * 1e-3)
        extras["rollout_note"] = f"DeviceRolloutWorker.sample: {T} control steps incl. actor+critic forward, sampling, buffer writes, GAE"
        env32.close()
    esz = 8 if args.precision == 64 else 4
    # max over ranks
    t = torch.tensor([total_ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = t.tolist()
    iters = env.solver_iterations().float().mean().item()
    if rank == 0:
        value = n * world * K / (total_ms * 1e-3)
        e2e = n * world * K / (e2e_ms * 1e-3)
        peak, peak_src = peaks()
        kernel_ms = statistics.mean(step_ms)   # one launch per step: the event pair brackets exactly the step kernel
        esz_ = 8 if args.precision == 64 else 4
        # algorithmic bytes per env-step: state record read + written, actions in, obs / reward / flags out
        alg_bytes = ALG_BYTES[args.precision] if args.workload == "jvrc_walk" else \
            (2 * env.state_r.shape[1] + A + env.obs_dim + 2) * esz_ + 2 * 8 * 4 + 2 * 4
        achieved = n * alg_bytes / (kernel_ms * 1e-3) / 1e9
        # secondary, honest bound: warp-instruction issue rate (instructions per env-step from the ncu capture of this
        # workload / precision) against 148 SMs x 4 schedulers x 1 warp-instruction per clock at the sampled SM clock
        cap = NCU_PER_LAUNCH_4096.get((args.workload, args.precision))
        traffic, issue = None, None
        if cap is not None:
            traffic = cap[0] * n / 4096.0
            inst_per_env_step = cap[1] / 4096.0
            sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
            issue_peak = 148 * 4 * sm_mhz * 1e6
            issue_ach = n * inst_per_env_step / (kernel_ms * 1e-3)
            issue = {"bound": "warp-issue", "achieved": issue_ach / 1e9, "peak": issue_peak / 1e9, "unit": "Gwarp-inst/s",
                     "frac": issue_ach / issue_peak, "warp_inst_per_env_step": inst_per_env_step,
                     "source": "smsp__inst_executed.sum of the committed ncu capture (profiles/)"}
        out = {"metric": metric, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
               "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f64" if args.precision == 64 else "f32", "data": "synthetic",
               "config": dict(config, l2="flushed (256 MiB write) before every timed step; CUDA events bracket the step only",
                              wall_s_incl_flush=wall, newton_iters_per_env_step=iters),
               "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": n * A * esz,
                       "d2h_bytes_per_step": n * (env.obs_dim * esz + esz + 4)},
               "gpu_launches": int(launches),
               "clocks": clocks,
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                            "traffic": traffic, "peak_source": peak_src,
                            "algorithmic_bytes_per_env_step": alg_bytes,
                            "note": "the step kernel is ALU/latency bound (25 substeps of O(nv^3) work per ~2 KB of state); "
                                    "traffic = dram bytes of one 4096-env launch in the committed ncu capture (state L2 resident "
                                    "under ncu, hence below the algorithmic bytes), scaled to this batch; `issue` is the bound that "
                                    "actually applies",
                            "issue": issue}}
        out["extras"] = extras
        if not args.no_cpu_baseline and world == 1:
            ncores = effective_cpus()
            n_cpu = max(256, min(n, 64 * ncores))
            sps, threads, dt, nst = cpu_reference(n_cpu, 10.0, 2, args.seed, model=wl["model"])
            out["cpu_baseline"] = {"value": sps, "unit": UNIT, "cores": threads, "kind": "port",
                                   "sample": f"{n_cpu} envs x {nst} control steps ({dt:.1f} s) after 2 warm-up steps, same action distribution; "
                                             "oracle/ C port with OpenMP (reference Ray+MuJoCo path not installable here)"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
