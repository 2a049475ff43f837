"""GPU parity AT THE SHIPPED SOLVER SETTINGS, for every BASELINE config, at the north-star bar.

The other GPU parity files tighten the Newton tolerance to 1e-14 on both sides so that kernel and oracle can be held to
1e-7.  This file runs what `bench.py` times and `run_experiment.py` trains on: the model's own solver options
(`models/jvrc_mj_description/xml/jvrc1.xml:8`: Newton, 50 iterations, tolerance 1e-10) in fp64, for jvrc_walk, jvrc_step,
h1 and the terrain extension, 64 environments (8+ lock-step blocks), 1000 control steps, in two regimes:

  zero     a = 0 (standing; the contact-heavy steady state of SURVEY.md §8d regime ii), truncation + auto-reset at 400
  policy   closed loop: each side feeds ITS OWN observation through the same fixed tanh policy + the same exploration noise
           (the early-training regime: falls, terminations, resets), so any disagreement is fed back
  trained  (jvrc_walk) closed loop through the actor of a finished training run: 400-step walking episodes

Bar (BASELINE.json north_star): qpos / qvel within 1e-4 relative over the 1000 steps, identical done / ended flags.
The achieved figures go to gpurun_out/parity_shipped.json (DESIGN.md §2 quotes them).
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_ENVS, STEPS, BAR = 64, 1000, 1e-4
_results = {}


def _rel(a, b):
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


def _policy(obs_dim, act_dim, mean, std, seed):
    rng = np.random.RandomState(seed)
    W = rng.normal(size=(obs_dim, act_dim)) / np.sqrt(obs_dim)
    return lambda obs: 0.3 * np.tanh(((obs - mean) / std) @ W)


def _trained_policy(model):
    """The actor of a 40-iteration `run_experiment.py train --env jvrc_walk --num-procs 4096 --seed 0` run of this build
    (tests/golden/trained_actor_jvrc_walk.pt; mean episode length 398 of 400: it walks), evaluated in float64 on each side's own
    observation: a smooth function, so a 1e-13 difference in the observation is not blown up by a float32 rounding boundary."""
    from learninghumanoidwalking_b200.rl.policies import install_reference_aliases
    install_reference_aliases()
    actor = torch.load(os.path.join(ROOT, "tests", "golden", f"trained_actor_{model}.pt"), map_location="cpu", weights_only=False).double()
    actor.eval()

    def pol(obs):
        with torch.no_grad():
            return actor(torch.from_numpy(np.ascontiguousarray(obs, dtype=np.float64)), deterministic=True).numpy()
    return pol


def _run(model, regime, precision=64, n=N_ENVS, steps=STEPS, stop_outside=None):
    """Kernel (C-ABI, shipped tolerance) and oracle (the model's own tolerance) side by side; returns the worst relative
    errors, the number of ended episodes and, for `stop_outside`, the first step at which qpos / qvel left that band."""
    from learninghumanoidwalking_b200.envs import BatchedHumanoidEnv
    from oracle.oracle import Oracle
    o = Oracle(model)                                    # tolerance = the model's (1e-10), Newton, 50 iterations
    env = BatchedHumanoidEnv(n, model=model, precision=precision, seed=31, first_env_id=5, max_traj_len=400)
    tol, its = env.mj["opt"]["tolerance"], env.mj["opt"]["iterations"]     # jvrc1.xml:8: 1e-10 / 50 ; h1.xml has no <option>: MuJoCo defaults 1e-8 / 100
    assert (tol, its) == ((1e-8, 100) if model == "h1" else (1e-10, 50))
    envs = o.make_envs(n, seed=31, first_id=5)
    o_obs = o.batch_reset(envs, n)
    g_obs = env.reset().double().cpu().numpy()
    A, nq, nv = env.act_dim, env.nq, env.nv
    pol = _trained_policy(model) if regime == "trained" else _policy(env.obs_dim, A, env.obs_mean, env.obs_std, seed=7)
    sigma = 0.05 if regime == "trained" else 0.223
    rng = np.random.RandomState(3)
    worst = dict(qpos=0.0, qvel=0.0, obs=0.0, reward=0.0)
    n_end, first_out = 0, None
    for k in range(steps):
        if regime == "zero":
            a_o = a_g = np.zeros((n, A))
        else:
            noise = rng.normal(size=(n, A)) * sigma
            a_o, a_g = pol(o_obs) + noise, pol(g_obs) + noise
        o_obs, o_tobs, o_terms, o_rew, o_done, o_end = o.batch_step(envs, n, a_o, max_traj_len=400)
        t_obs, t_rew, t_done, t_end = env.step(torch.as_tensor(a_g, device="cuda", dtype=env.dtype))
        g_obs = t_obs.double().cpu().numpy()
        if stop_outside is None:
            assert (t_done.cpu().numpy() == o_done).all() and (t_end.cpu().numpy() == o_end).all(), \
                f"{model}/{regime}: done / ended flags differ at control step {k}"
        n_end += int(o_end.sum())
        oq = np.stack([o.field(envs, i, "qpos")[:nq] for i in range(n)])
        ov = np.stack([o.field(envs, i, "qvel")[:nv] for i in range(n)])
        eq, ev = _rel(env.qpos.double().cpu().numpy(), oq), _rel(env.qvel.double().cpu().numpy(), ov)
        if stop_outside is not None and max(eq, ev) > stop_outside:
            first_out = k
            break
        worst["qpos"], worst["qvel"] = max(worst["qpos"], eq), max(worst["qvel"], ev)
        worst["obs"] = max(worst["obs"], _rel(g_obs, o_obs))
        worst["reward"] = max(worst["reward"], _rel(t_rew.double().cpu().numpy(), o_rew))
    iters = float(env.solver_iterations().float().mean())
    env.close()
    return worst, n_end, first_out, iters, tol


def _save():
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        path = os.path.join(out, "parity_shipped.json")
        try:
            merged = json.load(open(path))       # a partial run (-k ...) adds to what an earlier run left, it does not erase it
        except Exception:
            merged = {}
        merged.update(_results)
        json.dump(merged, open(path, "w"), indent=1, sort_keys=True)


@pytest.mark.parametrize("regime", ["zero", "policy"])
@pytest.mark.parametrize("model", ["jvrc_walk", "jvrc_step", "h1", "jvrc_walk_terrain"])
def test_fp64_shipped_tolerance_1000_steps_64_envs(model, regime):
    worst, n_end, _, iters, tol = _run(model, regime)
    _results[f"{model}/{regime}/fp64"] = dict(worst, ended_episodes=n_end, envs=N_ENVS, control_steps=STEPS,
                                              newton_iters_last_step=iters, tolerance=tol)
    _save()
    print(f"\n[parity shipped] {model:18s} {regime:6s} fp64 tol {tol:g}: " + "  ".join(f"{k} {v:.2e}" for k, v in worst.items())
          + f"  ended {n_end}")
    assert n_end >= N_ENVS, "every env should at least hit one truncation / termination in 1000 steps"
    assert worst["qpos"] < BAR and worst["qvel"] < BAR, worst
    assert worst["obs"] < BAR and worst["reward"] < BAR, worst


def test_fp64_shipped_tolerance_under_a_trained_walking_policy():
    """The regime the trainer converges to: a policy that WALKS for the whole 400-step horizon (heel strikes, double-support
    phases, mode switches) instead of falling after 50 steps.  Same bar, same settings; episodes end by truncation only."""
    worst, n_end, _, iters, tol = _run("jvrc_walk", "trained")
    _results["jvrc_walk/trained/fp64"] = dict(worst, ended_episodes=n_end, envs=N_ENVS, control_steps=STEPS, newton_iters_last_step=iters,
                                              tolerance=tol)
    _save()
    print("\n[parity shipped] jvrc_walk trained policy fp64: " + "  ".join(f"{k} {v:.2e}" for k, v in worst.items()) + f"  ended {n_end}")
    assert N_ENVS * 2 <= n_end <= N_ENVS * 2 + N_ENVS // 2, n_end      # truncations at 400 and 800, (almost) no falls
    assert worst["qpos"] < BAR and worst["qvel"] < BAR and worst["obs"] < BAR and worst["reward"] < BAR, worst


@pytest.mark.parametrize("model", ["jvrc_walk", "h1"])
def test_fp32_kernel_how_long_it_stays_inside_the_bar(model):
    """The optional fp32 kernel (solver tolerance 1e-6) is NOT the parity path.  This measures, it does not promise: the
    number of control steps the standing regime stays inside 1e-4 of the fp64 oracle, and asserts only the envelope the
    trainer relies on (no divergence inside an episode: 1e-2 over the first 100 steps)."""
    _, _, first_out, _, _ = _run(model, "zero", precision=32, n=16, steps=400, stop_outside=BAR)
    _, _, first_out2, _, _ = _run(model, "zero", precision=32, n=16, steps=100, stop_outside=1e-2)
    _results[f"{model}/zero/fp32"] = {"control_steps_inside_1e-4": first_out if first_out is not None else 400,
                                      "control_steps_inside_1e-2": first_out2 if first_out2 is not None else 100}
    _save()
    print(f"\n[parity shipped] {model} fp32 kernel: inside 1e-4 for {first_out} control steps, inside 1e-2 for {first_out2 or '>=100'}")
    assert first_out2 is None, f"fp32 left the 1e-2 envelope at control step {first_out2}"
