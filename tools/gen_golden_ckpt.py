#!/usr/bin/env python3
"""Generate tests/golden/ref_actor.pt / ref_critic.pt / ref_ckpt.json: a checkpoint pair written by the REFERENCE's own
classes exactly the way its trainer does (`torch.save(module)`, rl/utils/checkpointer.py:51), plus the outputs those modules
give on a fixed input.  tests/test_checkpoint_interchange.py loads the pair with this build (no reference on the path).
Run here only (/root/reference is not on the GPU box); this process must never import this repo's package (it would
install the `rl.policies` aliases)."""
import json
import os
import sys

import torch

REF = os.environ.get("LHW_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def main():
    sys.path.insert(0, REF)
    from rl.policies.actor import Gaussian_FF_Actor
    from rl.policies.critic import FF_V
    assert Gaussian_FF_Actor.__module__ == "rl.policies.actor" and "learninghumanoidwalking_b200" not in sys.modules
    torch.manual_seed(2024)
    actor = Gaussian_FF_Actor(37, 12, layers=(24, 24), init_std=0.3, learn_std=False, bounded=False)
    critic = FF_V(37, layers=(24, 24))
    with torch.no_grad():     # rl/algos/ppo.py:109-113
        actor.obs_mean = torch.linspace(-0.5, 0.5, 37)
        actor.obs_std = torch.linspace(0.5, 2.0, 37)
        critic.obs_mean, critic.obs_std = actor.obs_mean, actor.obs_std
        for p in list(actor.parameters()) + list(critic.parameters()):
            p.add_(0.05 * torch.randn_like(p))        # not the init any more: "trained" weights
    x = torch.randn(5, 37)
    torch.save(actor, os.path.join(OUT, "ref_actor.pt"))
    torch.save(critic, os.path.join(OUT, "ref_critic.pt"))
    json.dump(dict(x=x.tolist(), mu=actor(x).tolist(), v=critic(x).tolist(), stds=actor.stds.tolist(),
                   actor_keys=sorted(actor.state_dict()), critic_keys=sorted(critic.state_dict())),
              open(os.path.join(OUT, "ref_ckpt.json"), "w"))
    print("wrote ref_actor.pt ref_critic.pt ref_ckpt.json")


if __name__ == "__main__":
    main()
