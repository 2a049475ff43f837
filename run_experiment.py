#!/usr/bin/env python3
"""run_experiment.py — the reference's entry point (run_experiment.py:105-330) on the B200 path.

    python run_experiment.py train --env jvrc_walk --logdir /tmp/logs --num-procs 4096 --n-itr 100 --seed 0
    python -m torch.distributed.run --nproc-per-node 8 run_experiment.py train --env jvrc_step --num-procs 8192 ...
    python run_experiment.py eval --logdir /tmp/logs [--ep-len 10]

Same sub-commands and flags as the reference.  What differs underneath: `--num-procs` is the number of parallel environment
copies (device resident, sharded by index across the ranks of a torchrun launch) instead of Ray worker processes; there is
no Ray, no MuJoCo and no viewer — `eval` rolls the saved actor out deterministically on one device environment and prints
the per-term reward means (the reference's EvaluateEnv renders a video with MuJoCo's GL context, out of scope here).
Artefacts keep the reference's names and format: `<logdir>/<timestamp>_<env>/experiment.pkl`, `actor_<itr>.pt`,
`critic_<itr>.pt` (whole pickled modules, rl/utils/checkpointer.py:51-83).
"""
import argparse
import os
import pickle
import platform
import re
import sys
from datetime import datetime
from functools import partial
from pathlib import Path

import torch

_F, _I, _P, _S = float, int, Path, str
# the reference's command line (run_experiment.py:152-260), flag for flag, as data: (flag, argparse keywords)
TRAIN_FLAGS = [
    ("--env", dict(required=True, type=_S)), ("--logdir", dict(default=Path("/tmp/logs"), type=_P, help="run directory root")),
    ("--input-norm-steps", dict(type=_I, default=100000)), ("--n-itr", dict(type=_I, default=20000, help="PPO iterations")),
    ("--lr", dict(type=_F, default=3e-4)), ("--eps", dict(type=_F, default=1e-5, help="Adam epsilon")),
    ("--gamma", dict(type=_F, default=0.99)), ("--lam", dict(type=_F, default=0.95, help="GAE lambda")),
    ("--std-dev", dict(type=_F, default=0.223, help="exploration noise")), ("--learn-std", dict(action="store_true")),
    ("--entropy-coeff", dict(type=_F, default=0.0)), ("--clip", dict(type=_F, default=0.2, help="PPO clip range")),
    ("--minibatch-size", dict(type=_I, default=64)), ("--epochs", dict(type=_I, default=3)),
    ("--num-procs", dict(type=_I, default=4096, help="parallel environment copies over all ranks (the reference: Ray workers)")),
    ("--max-grad-norm", dict(type=_F, default=0.5)), ("--max-traj-len", dict(type=_I, default=400, help="episode horizon")),
    ("--no-mirror", dict(action="store_true", help="train without the SymmetricEnv mirror loss")),
    ("--mirror-coeff", dict(default=0.4, type=_F)), ("--eval-freq", dict(default=100, type=_I, help="checkpoint every N iterations")),
    ("--continued", dict(type=_P, help="actor checkpoint to continue from")), ("--recurrent", dict(action="store_true")),
    ("--imitate", dict(type=_S, default=None)), ("--imitate-coeff", dict(type=_F, default=0.0)),
    ("--yaml", dict(type=_S, default=None)), ("--device", dict(type=_S, default="cuda", choices=["auto", "cuda"])),
    ("--seed", dict(type=_I, default=None)),
    # additions of this build
    ("--precision", dict(type=_I, default=32, choices=[32, 64], help="simulator arithmetic (64 = the reference's float64)")),
    ("--tf32", dict(action="store_true", help="TF32 tensor-core GEMMs in the MLPs (off: fp32 like the reference)")),
    ("--steps-per-env", dict(type=_I, default=None, help="transitions per env per iteration (default: max-traj-len)")),
    ("--minibatch-scale", dict(type=_S, default="auto", choices=["auto", "off"],
                               help="auto: --minibatch-size is scaled by (this rank's batch / 4800) so that an iteration keeps the reference's "
                                    "225 optimiser steps at its default flags; off: --minibatch-size samples per rank, literally")),
]
EVAL_FLAGS = [("--path", dict(type=_P, default=None)), ("--logdir", dict(type=_P, default=None)),
              ("--out-dir", dict(type=_P, default=None, help="(videos are not produced by this build)")),
              ("--ep-len", dict(type=_I, default=10, help="seconds to play")), ("--seed", dict(type=_I, default=None))]
ENVS = ("jvrc_walk", "jvrc_step", "h1", "jvrc_walk_terrain")   # the last one is an extension (BASELINE configs[4])


def print_system_info(args, training=True):
    print("=" * 60)
    print("System Information")
    print("=" * 60)
    print(f"PyTorch version: {torch.__version__}")
    print(f"Platform: {platform.system()} {platform.release()}")
    print(f"CUDA devices: {torch.cuda.device_count()}"
          + (f" ({torch.cuda.get_device_name(0)})" if torch.cuda.is_available() else ""))
    if training:
        print("-" * 60)
        print("Training Configuration")
        print("-" * 60)
        for k, v in (("Environment", args.env), ("Log directory", args.logdir), ("Parallel envs", args.num_procs),
                     ("Learning rate", args.lr), ("Max trajectory length", args.max_traj_len), ("Iterations", args.n_itr),
                     ("Seed", args.seed)):
            print(f"{k}: {v}")
    print("=" * 60)


def get_latest_run_dir(logdir):
    logdir = Path(logdir)
    subdirs = [d for d in logdir.iterdir() if d.is_dir()] if logdir.exists() else []
    return max(subdirs, key=lambda d: d.stat().st_mtime) if subdirs else None


def get_latest_actor(run_dir):
    files = list(Path(run_dir).glob("actor_*.pt"))
    if not files:
        return Path(run_dir) / "actor.pt"
    return max(files, key=lambda f: int(m.group(1)) if (m := re.search(r"actor_(\d+)\.pt$", f.name)) else -1)


def import_env(env_name_str):
    """run_experiment.py:87-101: env name -> factory of a device-resident batch of that environment."""
    if env_name_str not in ENVS:
        raise Exception("Check env name! (this build has: " + ", ".join(ENVS) + ")")
    from learninghumanoidwalking_b200.envs import BatchedHumanoidEnv
    return partial(BatchedHumanoidEnv, model=env_name_str)


def _dist_setup():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        return dist.get_rank(), world, local
    return 0, 1, 0


def run_experiment(args):
    from learninghumanoidwalking_b200.rl import PPO
    from learninghumanoidwalking_b200.rl.dist_utils import env_shard
    from learninghumanoidwalking_b200.rl.symmetric import SymmetricEnv
    rank, world, local = _dist_setup()
    timestamp = datetime.now().strftime("%y-%m-%d-%H-%M-%S-%f")[:-3]
    args.logdir = Path(args.logdir) / f"{timestamp}_{args.env}"
    if rank == 0:
        print_system_info(args)
    if args.recurrent or args.imitate:
        raise NotImplementedError("--recurrent / --imitate are outside the accelerated path (SURVEY.md §2)")
    Env = import_env(args.env)
    if args.num_procs % world:
        raise SystemExit(f"--num-procs {args.num_procs} is not a multiple of the {world} ranks: every rank must hold the same number "
                         "of environments (each optimiser step is one gradient exchange on every rank)")
    first, n_local = env_shard(rank, world, args.num_procs)
    seed = args.seed if args.seed is not None else 0
    env_fn = partial(Env, n_local, precision=args.precision, seed=seed, first_env_id=first, device=local,
                     max_traj_len=args.max_traj_len, path_to_yaml=args.yaml)
    _env = env_fn()
    if not args.no_mirror:
        try:
            r = _env.robot
            env_fn = partial(SymmetricEnv, env_fn, mirrored_obs=r.mirrored_obs, mirrored_act=r.mirrored_acts,
                             clock_inds=r.clock_inds)
            if rank == 0:
                print("Wrapping in SymmetricEnv.")
        except AttributeError as e:
            print("Warning! Cannot use SymmetricEnv.", e)
    _env.close()
    if rank == 0:
        Path.mkdir(args.logdir, parents=True, exist_ok=True)
        with open(Path(args.logdir, "experiment.pkl"), "wb") as f:
            pickle.dump(args, f)
        if args.yaml is not None:      # run_experiment.py:141-144: the YAML rides along with the run
            import shutil
            shutil.copyfile(args.yaml, Path(args.logdir, Path(args.yaml).name))
    algo = PPO(env_fn, args, seed=args.seed)
    if args.continued is not None:
        actor = torch.load(args.continued, weights_only=False)
        critic = torch.load(Path(args.continued.parent, "critic" + str(args.continued).split("actor")[1]), weights_only=False)
        algo.load_pretrained(actor, critic)
        if rank == 0:
            print("Loaded (pre-trained) actor from: ", args.continued)
    algo.train(env_fn, args.n_itr, verbose=rank == 0)


def evaluate(args):
    """Deterministic roll-out of a saved actor on one device environment (the numbers EvaluateEnv would overlay)."""
    if args.path is not None:
        if args.path.is_file() and args.path.suffix == ".pt":
            path_to_actor = args.path
        elif args.path.is_dir():
            path_to_actor = get_latest_actor(args.path)
        else:
            raise Exception("Invalid path to actor module: ", args.path)
    elif args.logdir is not None:
        latest = get_latest_run_dir(args.logdir)
        if latest is None:
            raise Exception(f"No run directories found under: {args.logdir}")
        path_to_actor = get_latest_actor(latest)
    else:
        raise Exception("Must provide either --path or --logdir")
    print(f"Loading model: {path_to_actor}")
    with open(Path(path_to_actor.parent, "experiment.pkl"), "rb") as f:
        run_args = pickle.load(f)
    policy = torch.load(path_to_actor, weights_only=False).cuda().eval()
    print_system_info(args, training=False)
    env = import_env(run_args.env)(1, precision=64, seed=args.seed or 0)
    obs = env.reset().float()
    n_steps = int(args.ep_len / env.dt)
    totals, ep_ret, ep_len, episodes = torch.zeros(10, dtype=env.dtype, device=env.device), 0.0, 0, []
    with torch.no_grad():
        for _ in range(n_steps):
            obs, rew, done, _ = env.step(policy(obs, deterministic=True).to(env.dtype), autoreset=False)
            totals += env.rew_terms[0]
            ep_ret, ep_len = ep_ret + float(rew[0]), ep_len + 1
            if bool(done[0]):
                episodes.append((ep_len, ep_ret))
                ep_ret, ep_len = 0.0, 0
                obs = env.reset()
            obs = obs.float()
    episodes.append((ep_len, ep_ret))
    print(f"{len(episodes)} episode(s) in {n_steps} control steps: " + ", ".join(f"len {l} return {r:.2f}" for l, r in episodes))
    for name, v in zip(env.reward_names, (totals / n_steps).tolist()):
        print(f"  mean {name:>20s} {v:.4f}")
    env.close()
    return episodes


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    parser = argparse.ArgumentParser()
    if argv and argv[0] == "train":
        for flag, kw in TRAIN_FLAGS:
            parser.add_argument(flag, **kw)
        args = parser.parse_args(argv[1:])
        if args.seed is not None:
            torch.manual_seed(args.seed)
            print(f"Deterministic mode enabled with seed: {args.seed}")
        return run_experiment(args)
    elif argv and argv[0] == "eval":
        for flag, kw in EVAL_FLAGS:
            parser.add_argument(flag, **kw)
        args = parser.parse_args(argv[1:])
        return evaluate(args)
    else:
        parser.error("usage: run_experiment.py {train,eval} ...")


if __name__ == "__main__":
    main()
