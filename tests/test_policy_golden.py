"""Host-side policy / mirror helpers against vectors produced by the reference's own classes
(tools/gen_golden.py: rl/policies/actor.py, critic.py, rl/envs/wrappers.py)."""
import json
import os

import numpy as np
import torch

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "policy.json")))


def _load(net, sd):
    """the golden file holds the reference modules' state dicts: the keys must be ours, one to one"""
    assert set(sd) == set(net.state_dict())
    net.load_state_dict({k: torch.tensor(v) for k, v in sd.items()})


def test_actor_and_critic_forward_match_reference_modules():
    from learninghumanoidwalking_b200.rl.policies import FF_V, Gaussian_FF_Actor
    a, c = Gaussian_FF_Actor(37, 12, layers=(16, 16)), FF_V(37, layers=(16, 16))
    _load(a, G["actor"])
    _load(c, G["critic"])
    a.obs_mean = c.obs_mean = torch.tensor(G["obs_mean"])
    a.obs_std = c.obs_std = torch.tensor(G["obs_std"])
    x = torch.tensor(G["x"])
    assert (a(x) - torch.tensor(G["mu"])).abs().max() < 1e-6
    assert (c(x) - torch.tensor(G["v"])).abs().max() < 1e-6


def test_parameter_counts_and_normc_init():
    from learninghumanoidwalking_b200.rl.policies import FF_V, Gaussian_FF_Actor
    a, c = Gaussian_FF_Actor(37, 12), FF_V(37)
    assert sum(p.numel() for p in a.parameters()) == G["n_actor"] == 78604
    assert sum(p.numel() for p in c.parameters()) == G["n_critic"] == 75777
    assert abs(float(a.means.weight.norm(dim=1).mean()) - G["out_layer_norm"]) < 1e-6   # output layer x 0.01
    assert torch.allclose(a.actor_layers[0].weight.norm(dim=1), torch.ones(256), atol=1e-5)    # normc rows
    assert float(a.actor_layers[0].bias.abs().max()) == 0.0


def test_mirror_matrices_match_reference():
    from learninghumanoidwalking_b200.rl.symmetric import symmetry_matrix
    assert np.array_equal(symmetry_matrix(G["mirrored_obs"]).numpy(), np.array(G["obs_mirror_matrix"], dtype=np.float32))
    assert np.array_equal(symmetry_matrix(G["mirrored_act"]).numpy(), np.array(G["act_mirror_matrix"], dtype=np.float32))


def test_clock_mirror_is_the_reference_formula():
    """sin(arcsin(c) + pi) == -c on the clock entries (rl/envs/wrappers.py:64-75)."""
    from learninghumanoidwalking_b200.rl.symmetric import SymmetricEnv
    env = SymmetricEnv(lambda: object(), mirrored_obs=G["mirrored_obs"], mirrored_act=G["mirrored_act"], clock_inds=[29, 30])
    obs = torch.rand(6, 37) * 1.6 - 0.8
    ref = obs @ torch.tensor(G["obs_mirror_matrix"], dtype=torch.float32)
    for i in (29, 30):
        ref[:, i] = torch.sin(torch.arcsin(ref[:, i]) + np.pi)
    assert (env.mirror_clock_observation(obs) - ref).abs().max() < 1e-6


def test_folded_clock_mirror_matrix_is_bit_identical_to_matmul_then_negate():
    """SymmetricEnv folds the clock negation into the mirror matrix (one GEMM, no index kernels, capturable in a CUDA graph):
    every column of P has a single +-1, so obs @ (P diag(s)) equals (obs @ P) with the clock columns negated, bit for bit."""
    from learninghumanoidwalking_b200.rl.symmetric import SymmetricEnv
    env = SymmetricEnv(lambda: object(), mirrored_obs=G["mirrored_obs"], mirrored_act=G["mirrored_act"], clock_inds=[29, 30])
    obs = torch.randn(64, 37)
    ref = obs @ env.obs_mirror_matrix
    ref[:, [29, 30]] = -ref[:, [29, 30]]
    assert torch.equal(env.mirror_clock_observation(obs), ref)
    assert torch.equal(env.mirror_observation(obs), obs @ env.obs_mirror_matrix)
    a = torch.randn(5, 12)
    assert torch.equal(env.mirror_action(env.mirror_action(a)), a)            # an involution
    assert env._on("obs_mirror_matrix", "cpu") is env._on("obs_mirror_matrix", "cpu")   # uploaded once, then cached


def test_height_curriculum_formula():
    """tasks/stepping_task.py:312: h = clip((iteration_count - 3000) / 8000, 0, 1) * 0.1, iteration_count = inf by default."""
    from learninghumanoidwalking_b200.model.loader import curriculum_height
    for it, h in ((0, 0.0), (3000, 0.0), (3800, 0.01), (7000, 0.05), (11000, 0.1), (10 ** 9, 0.1), (float("inf"), 0.1)):
        assert abs(curriculum_height(it) - h) < 1e-15
        assert curriculum_height(it) == float(np.clip((it - 3000) / 8000, 0, 1) * 0.1)


def test_worker_seed_formula_matches_reference():
    """rl/utils/seeding.py:34-52 run from the reference's file (tests/golden/seeding.json)."""
    from learninghumanoidwalking_b200.rl.ppo import get_worker_seed
    cases = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "seeding.json")))
    assert len(cases) == 72
    for c in cases:
        assert get_worker_seed(c["master_seed"], c["worker_id"], c["offset"]) == c["seed"]
    assert len({c["seed"] for c in cases if c["master_seed"] == 12345}) == 12     # collision-free on the grid


def test_env_attributes_match_the_reference_env_classes():
    """obs_mean / obs_std and the mirror index lists of BatchedHumanoidEnv against the reference's own method bodies, lifted out of
    envs/jvrc/jvrc_walk.py, jvrc_step.py, jvrc_base.py and envs/h1/h1_env.py and executed (tools/gen_golden_env_attrs.py)."""
    from learninghumanoidwalking_b200.envs.batched_env import BatchedHumanoidEnv
    from learninghumanoidwalking_b200.model import load_model
    ref = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "env_attributes.json")))
    for model, key in (("jvrc_walk", "jvrc_walk"), ("jvrc_step", "jvrc_step"), ("h1", "h1"), ("jvrc_walk_terrain", "jvrc_walk")):
        env = BatchedHumanoidEnv.__new__(BatchedHumanoidEnv)      # attribute setup only: no CUDA, no library handle
        env._iteration_count = float("inf")
        env._setup_reference_attributes(model, load_model(model)["cfg"])
        r = ref[key]
        assert np.array_equal(env.obs_mean, np.array(r["obs_mean"])) and np.array_equal(env.obs_std, np.array(r["obs_std"]))
        if key != "h1":
            assert [float(x) for x in env.robot.mirrored_obs] == r["mirrored_obs"]
            assert [float(x) for x in env.robot.mirrored_acts] == r["mirrored_acts"]
            assert list(env.robot.clock_inds) == r["clock_inds"]
            assert env.robot.iteration_count == float("inf")                 # robots/robot_base.py:35
        else:
            assert not hasattr(env.robot, "mirrored_obs")


def test_linear_helper_is_the_plain_layer_off_the_cuda_training_path():
    """rl/policies/base.py:linear routes ONLY CUDA float32 training forwards to the weight-gradient kernel; on the CPU (these tests,
    the reference's own checkpoints in a CPU process) it is `layer(x)` with torch's own backward, bit for bit."""
    import torch
    from learninghumanoidwalking_b200.rl.policies.base import _use_wgrad_kernel, linear
    torch.manual_seed(0)
    layer = torch.nn.Linear(37, 256)
    x = torch.randn(16, 37)
    assert not _use_wgrad_kernel(x, layer)
    y = linear(layer, x)
    assert torch.equal(y, layer(x)) and type(y.grad_fn).__name__ == "AddmmBackward0"
    with torch.no_grad():
        assert linear(layer, x).grad_fn is None
