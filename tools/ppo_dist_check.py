#!/usr/bin/env python3
"""torchrun entry: PPO with env copies sharded across ranks and ONE NCCL all-reduce of the flat gradient per
optimiser step.  Checks that every rank ends with bit-identical weights (replicas stay in lock-step) and that
different ranks really simulated different env ids."""
import os
import sys
from types import SimpleNamespace

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from learninghumanoidwalking_b200.envs import BatchedHumanoidEnv
    from learninghumanoidwalking_b200.rl import PPO
    from learninghumanoidwalking_b200.rl.dist_utils import env_shard
    from learninghumanoidwalking_b200.rl.symmetric import SymmetricEnv
    n_global = int(os.environ.get("LHW_ENVS", "256"))
    first, n = env_shard(rank, world, n_global)
    model = os.environ.get("LHW_MODEL", "jvrc_walk")     # jvrc_step: BASELINE configs[2] (env-sharded, gradient exchange)
    base = lambda: BatchedHumanoidEnv(n, model=model, precision=32, seed=0, first_env_id=first, device=local)
    probe = base()
    r = probe.robot
    probe.close()
    env_fn = lambda: SymmetricEnv(base, mirrored_obs=r.mirrored_obs, mirrored_act=r.mirrored_acts, clock_inds=r.clock_inds)
    args = SimpleNamespace(gamma=0.99, lam=0.95, lr=3e-4, eps=1e-5, entropy_coeff=0.0, clip=0.2, minibatch_size=128, epochs=2,
                           max_traj_len=50, num_procs=n, max_grad_norm=0.05, mirror_coeff=0.4, eval_freq=1000, recurrent=False,
                           imitate_coeff=0.0, std_dev=0.223, learn_std=False, logdir=f"/tmp/lhw_dist_{rank}", steps_per_env=16)
    ppo = PPO(env_fn, args, seed=0)
    if os.environ.get("LHW_CHECK_ONE_STEP") == "1":
        # exactly one optimiser step on identical data: the two exchange implementations must agree to rounding
        ppo.make_optimizers()
        batch = ppo.sample_parallel_with_workers()
        adv = ppo.normalize_advantages(batch.returns.contiguous(), batch.values.contiguous())
        ppo.update_actor_critic(batch.states[:128], batch.actions[:128], batch.returns[:128], adv[:128], 1,
                                mirror_observation=ppo.env.mirror_clock_observation, mirror_action=ppo.env.mirror_action)
        log = [dict(critic_loss=0.0, fps=0.0)]
    else:
        log = ppo.train(None, 2, verbose=False)
    flat = ppo._flat_param
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    obs0 = ppo.env.obs[0, :5].clone()
    obs_all = [torch.empty_like(obs0) for _ in range(world)]
    dist.all_gather(obs_all, obs0)
    differ = world == 1 or not torch.equal(obs_all[0], obs_all[-1])
    if rank == 0:
        fused = ppo._comm is not None
        print(f"DIST_CHECK model={model} world={world} envs/rank={n} fused_exchange={fused} update_graph={ppo._ug is not None} "
              f"identical_weights={same} "
              f"ranks_simulate_different_envs={differ} critic_loss={log[-1]['critic_loss']:.4f} fps={log[-1]['fps']:.0f} "
              f"wsum={flat.double().sum().item():.10f} wabs={flat.double().abs().sum().item():.10f}")
    dist.destroy_process_group()
    if not (same and differ):
        sys.exit(1)


if __name__ == "__main__":
    main()
