"""The modelling assumptions SURVEY.md Appendix A marks as unverifiable here (MuJoCo is not installable) are entries of
model/*.json ("assumptions" + the fields that carry them).  This file
  * shows that each one is a live switch on BOTH sides (oracle and kernel source) and how far it moves a trajectory — so the
    day a MuJoCo recording exists (tools/record_reference.py) a mismatch of a given size points at a short list of suspects;
  * pins the soft-constraint law with closed-form known answers that do not depend on any simulator: spring / damper
    constants from solref, the impedance sigmoid from solimp, the static penetration of a stance that carries m g.
CPU tier: oracle + the host emulation of the kernel source."""
import copy
import math

import numpy as np
import pytest

from oracle import oracle as O


def _traj(o, steps=30, n=2, sigma=0.1, seed=3):
    envs = o.make_envs(n, seed=seed)
    o.batch_reset(envs, n)
    rng = np.random.RandomState(seed)
    out = []
    for _ in range(steps):
        o.batch_step(envs, n, rng.normal(size=(n, o.nu)) * sigma, 400)
        out.append(np.stack([np.concatenate((o.field(envs, i, "qpos")[:o.nq], o.field(envs, i, "qvel")[:o.nv])) for i in range(n)]))
    return np.stack(out)


def _variant(mut):
    mj = copy.deepcopy(O.load_model_json("jvrc_walk"))
    mut(mj)
    return O.Oracle("jvrc_walk", model_dict=mj)


def _round_sig(x, digits):
    return float(f"%.{digits}g" % x)


def _reround(mj, digits=4):
    for lk in mj["links"]:
        for k in ("pos", "com", "inertia"):
            lk[k] = [_round_sig(v, digits) for v in lk[k]]
        lk["mass"] = _round_sig(lk["mass"], digits)


SWITCHES = {
    # name: (mutation, what it stands for)
    "explicit_euler": (lambda mj: mj["assumptions"].update(implicit_damping=False), "A.1 implicit joint damping in mj_Euler"),
    "impratio_x2": (lambda mj: mj["opt"].update(impratio=2.0), "A.3 pyramid regulariser R = 2 mu^2 R_n / impratio"),
    "foot_invweight_x1.1": (lambda mj: [mj["link_invweight0"][k].__setitem__(0, mj["link_invweight0"][k][0] * 1.1)
                                        for k in (mj["rfoot_link"], mj["lfoot_link"])], "A.3 diagApprox of contact rows"),
    "dof_invweight_x1.1": (lambda mj: mj.update(dof_invweight0=[1.1 * v for v in mj["dof_invweight0"]]), "A.3 diagApprox of limit rows"),
    "round_4_digits": (lambda mj: _reround(mj, 4), "A.0 '%.5g' export rounding (one digit coarser: an upper bound of its effect)"),
    "solimp_d0_0.85": (lambda mj: mj["opt"]["solimp"].__setitem__(0, 0.85), "A.0 default solimp"),
}


def test_every_assumption_is_a_live_switch_and_its_trajectory_sensitivity_is_known():
    sens = {}
    bases = {}
    for name, (mut, _) in SWITCHES.items():
        # joint-limit rows only exist when a joint is driven into its range: that switch is measured under large actions
        sigma = 1.5 if name.startswith("dof_invweight") else 0.1
        if sigma not in bases:
            bases[sigma] = _traj(O.Oracle("jvrc_walk"), sigma=sigma)
        base = bases[sigma]
        sens[name] = float(np.abs(_traj(_variant(mut), sigma=sigma) - base).max() / np.abs(base).max())
    print("\n[assumption sensitivity] max |d(qpos, qvel)| / scale over 30 control steps, jvrc_walk: "
          + "  ".join(f"{k} {v:.1e}" for k, v in sens.items()))
    for name, v in sens.items():
        assert v > 1e-9, f"{name} changed nothing: it is not wired to the dynamics"
        assert v < 0.5, f"{name}: the trajectory is unrecognisable ({v:.2f}) — the switch breaks the model rather than perturbing it"
    # orders of magnitude the day a recording exists: damping treatment and contact regularisation are first-order suspects
    # (1e-3 .. 1e-1 in 30 steps), rounding an order of magnitude below
    assert sens["round_4_digits"] < sens["explicit_euler"]
    # the parity bar is 1e-4: every one of these would be VISIBLE against a MuJoCo recording
    assert min(sens.values()) > 1e-6


def test_kernel_source_honours_the_explicit_euler_switch():
    """The flags word at the end of the packed model reaches the kernel source: emulation == oracle with the switch on."""
    from emu import Emu
    from learninghumanoidwalking_b200.model import load_model, pack_model
    mj = load_model("jvrc_walk")
    mj["assumptions"] = dict(mj["assumptions"], implicit_damping=False)
    assert pack_model(mj, tolerance=1e-14)[-1] == 1.0 and pack_model(load_model("jvrc_walk"))[-1] == 0.0
    o = O.Oracle("jvrc_walk", tolerance=1e-14, model_dict=mj)
    e = Emu(pack_model(mj, tolerance=1e-14), 64, 2, seed=2, first_id=0)
    envs = o.make_envs(2, seed=2)
    assert np.abs(e.reset() - o.batch_reset(envs, 2)).max() < 1e-9
    rng = np.random.RandomState(0)
    for _ in range(6):
        a = rng.normal(size=(2, 12)) * 0.2
        oo = o.batch_step(envs, 2, a)[0]
        assert np.abs(e.step(a)[0] - oo).max() < 1e-9
    # ... and it is not the default trajectory
    d = O.Oracle("jvrc_walk", tolerance=1e-14)
    envs_d = d.make_envs(2, seed=2)
    d.batch_reset(envs_d, 2)
    rng = np.random.RandomState(0)
    for _ in range(6):
        od = d.batch_step(envs_d, 2, rng.normal(size=(2, 12)) * 0.2)[0]
    assert np.abs(od - oo).max() > 1e-6


def test_planebox_never_has_more_than_four_candidate_corners():
    """A.5: mjc_PlaneBox's selection rule (first four in index order) can only matter if more than four corners qualify; a
    corner qualifies when it is below the plane AND on the plane side of the box centre, which holds for at most four of the
    eight unless the box stands exactly on an edge.  Checked over tumbling rollouts."""
    o = O.Oracle("jvrc_walk")
    n = 8
    envs = o.make_envs(n, seed=9)
    o.batch_reset(envs, n)
    rng = np.random.RandomState(1)
    worst = 0
    for _ in range(120):
        o.batch_step(envs, n, rng.normal(size=(n, 12)) * 0.6, 400)
        for i in range(n):
            _, dist, foot, _ = o.contacts(envs, i)
            for f in np.unique(foot):
                worst = max(worst, int((foot == f).sum()))
    assert 1 <= worst <= 4


# ------------------------------------------------------------------ closed-form known answers of the soft-constraint law
def _law(mj):
    tc, zeta = mj["opt"]["solref"]
    d0, dw, width, mid, power = mj["opt"]["solimp"]
    tc = max(tc, 2 * mj["opt"]["timestep"])
    K = 1.0 / (dw * dw * tc * tc * zeta * zeta)        # mj_makeImpedance: stiffness of the reference acceleration
    B = 2.0 / (dw * tc)
    return K, B, (d0, dw, width, mid, power)


def _imp(solimp, r):
    d0, dw, width, mid, power = solimp
    x = min(1.0, abs(r) / width)
    y = (x / mid) ** power * mid if x <= mid else 1 - ((1 - x) / (1 - mid)) ** power * (1 - mid)
    return d0 + y * (dw - d0)


def test_solref_constants_are_a_critically_damped_20ms_response():
    """solref = (0.02, 1): a_ref = -B v - K imp r with K = 1 / (dmax^2 tc^2 zeta^2) = 2770.08 and B = 2 / (dmax tc) = 105.263.
    For a fully engaged constraint (imp = dmax) r'' + B r' + K dmax r = 0 has the roots -(1 -+ sqrt(1 - dmax)) / (dmax tc):
    time constants 24.5 ms and 15.5 ms around the nominal 20 ms, no oscillation (damping ratio 1 / sqrt(dmax) > 1)."""
    K, B, solimp = _law(O.load_model_json("jvrc_walk"))
    assert abs(K - 2770.0831) < 1e-3 and abs(B - 105.26316) < 1e-4
    dw = solimp[1]
    disc = B * B - 4 * K * dw
    assert disc > 0
    s1, s2 = (-B + math.sqrt(disc)) / 2, (-B - math.sqrt(disc)) / 2
    assert abs(-1 / s1 - 0.02447) < 1e-4 and abs(-1 / s2 - 0.01553) < 1e-4
    assert abs(math.sqrt(s1 * s2) - 1 / (math.sqrt(dw) * 0.02)) < 1e-9        # geometric mean of the rates = 1 / (sqrt(dmax) tc)
    assert abs(_imp(solimp, 0.0) - 0.9) < 1e-15 and abs(_imp(solimp, 0.001) - 0.95) < 1e-15 and abs(_imp(solimp, 0.0005) - 0.925) < 1e-15


def test_static_stance_penetration_matches_the_closed_form_law():
    """At rest every active pyramid edge carries f = D (-K imp(r) r) with D = 1 / (2 mu^2 R_n / impratio),
    R_n = (1 - imp) / imp * invweight0 * (1 + mu^2): a contact at penetration d pushes with N(d) = 4 D K imp d along the
    normal.  The oracle's settled stance must satisfy sum_i N(d_i) = m g with ITS contact depths — the closed-form law, the
    collision depths and the solver meet in one number — and the mean depth must be the root of 8 N(d) = m g."""
    mj = copy.deepcopy(O.load_model_json("jvrc_walk"))
    mj["cfg"]["kp"] = [40 * k for k in mj["cfg"]["kp"]]
    mj["cfg"]["kd"] = [5 * k for k in mj["cfg"]["kd"]]
    o = O.Oracle("jvrc_walk", tolerance=1e-14, model_dict=mj)
    envs = o.make_envs(1)
    o.reset(envs)
    for _ in range(240):
        o.step(envs, 0, np.zeros(12))
    K, _, solimp = _law(mj)
    mu, impratio = mj["opt"]["friction"][0], mj["opt"]["impratio"]
    invw = mj["link_invweight0"][mj["rfoot_link"]][0]
    assert abs(invw - mj["link_invweight0"][mj["lfoot_link"]][0]) < 1e-12

    def N(d):
        imp = _imp(solimp, d)
        Rn = (1 - imp) / imp * invw * (1 + mu * mu)
        return 4 * (1.0 / (2 * mu * mu * Rn / impratio)) * K * imp * d
    _, dist, _, _ = o.contacts(envs, 0)
    assert len(dist) == 8 and (dist < 0).all()
    weight = mj["total_mass"] * 9.81
    total = sum(N(-d) for d in dist)
    assert abs(total - weight) < 5e-3 * weight, (total, weight)            # residual motion of the quasi-static stance
    grf = o.field(envs, 0, "rfoot_grf")[0] + o.field(envs, 0, "lfoot_grf")[0]
    assert abs(total - grf) < 1e-2 * weight
    # root of 8 N(d) = m g by bisection: the depth a perfectly even stance would have; the measured mean is within 15 % of it
    lo, hi = 0.0, 1e-3
    for _ in range(80):
        md = 0.5 * (lo + hi)
        lo, hi = (md, hi) if 8 * N(md) < weight else (lo, md)
    assert 2e-5 < lo < 5e-4 and abs(np.mean(-dist) - lo) < 0.15 * lo, (lo, np.mean(-dist))


def test_trained_policy_walks_whatever_the_unverifiable_details_are():
    """tools/policy_sensitivity.py -> tests/golden/policy_sensitivity.json: the actor of a finished training run of this build walks the
    whole 400-step horizon in the oracle under EVERY assumption switch above, and its return moves by less than 1 % — whatever MuJoCo
    really does at those points, the behaviour a user trains is the same.  One variant is recomputed here."""
    import json
    import os
    import sys
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ev = json.load(open(os.path.join(root, "tests", "golden", "policy_sensitivity.json")))
    base = ev["baseline"]
    assert base["falls"] == 0 and base["mean_episode_length"] == 400.0
    for name in SWITCHES:
        assert ev[name]["falls"] == 0 and ev[name]["min_episode_length"] == 400, name
        assert abs(ev[name]["mean_episode_return"] / base["mean_episode_return"] - 1.0) < 0.01, name
    sys.path.insert(0, os.path.join(root, "tools"))
    from policy_sensitivity import rollout
    from learninghumanoidwalking_b200.rl.policies import install_reference_aliases
    install_reference_aliases()
    actor = torch.load(os.path.join(root, "tests", "golden", "trained_actor_jvrc_walk.pt"), map_location="cpu", weights_only=False).double().eval()
    r = rollout(_variant(SWITCHES["impratio_x2"][0]), actor)
    assert r["falls"] == 0 and abs(r["mean_episode_return"] - ev["impratio_x2"]["mean_episode_return"]) < 1e-6
