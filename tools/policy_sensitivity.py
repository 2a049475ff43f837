#!/usr/bin/env python3
"""Does a policy trained on this build's simulator still walk if the unverifiable MuJoCo details were different?  (CPU, oracle only.)

tests/test_assumption_switches.py measures how far each assumption of SURVEY.md Appendix A moves a TRAJECTORY.  This script asks the
question a user of the trained policy cares about: the actor of a finished training run (tests/golden/trained_actor_jvrc_walk.pt,
trained on the fp32 CUDA simulator) is rolled out in oracle variants that each flip one assumption — explicit instead of implicit
joint damping, impratio 2, contact / limit diagApprox +10 %, model numbers rounded one digit coarser, solimp d0 0.85 — for whole
400-step episodes (deterministic mean + N(0, 0.05^2), 16 environments), and the episode length and return are compared with the
unmodified oracle.  Output: tests/golden/policy_sensitivity.json."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def rollout(o, actor, n=16, steps=400, seed=31):
    envs = o.make_envs(n, seed=seed, first_id=5)
    obs = o.batch_reset(envs, n)
    rng = np.random.RandomState(3)
    ret, length, alive = np.zeros(n), np.zeros(n, int), np.ones(n, bool)
    for k in range(steps):
        with torch.no_grad():
            act = actor(torch.from_numpy(obs), deterministic=True).numpy()
        obs, _, _, rew, done, end = o.batch_step(envs, n, act + 0.05 * rng.normal(size=(n, o.nu)), max_traj_len=steps)
        ret += rew * alive
        length += alive
        alive &= ~np.asarray(end, bool)
    return dict(mean_episode_length=float(length.mean()), min_episode_length=int(length.min()), mean_episode_return=float(ret.mean()),
                falls=int((length < steps).sum()))


def main():
    from learninghumanoidwalking_b200.rl.policies import install_reference_aliases
    from oracle.oracle import Oracle
    from test_assumption_switches import SWITCHES, _variant
    install_reference_aliases()
    actor = torch.load(os.path.join(ROOT, "tests", "golden", "trained_actor_jvrc_walk.pt"), map_location="cpu", weights_only=False).double().eval()
    out = {"_doc": "tools/policy_sensitivity.py: 16 envs x one 400-step episode of the CPU oracle under tests/golden/trained_actor_jvrc_walk.pt",
           "baseline": rollout(Oracle("jvrc_walk"), actor)}
    for name, (mut, what) in SWITCHES.items():
        r = rollout(_variant(mut), actor)
        r["stands_for"] = what
        out[name] = r
        print(name, r, flush=True)
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "policy_sensitivity.json"), "w"), indent=1)
    print(json.dumps(out["baseline"]))


if __name__ == "__main__":
    main()
