#!/usr/bin/env python3
"""How good are the self-collision capsule proxies?  (robot_interface.py:472-484 terminates an
episode on ANY contact between two robot geoms; the geoms that can touch are the convex hulls of {R,L}_{HIP_R,HIP_Y,KNEE}_S
(envs/jvrc/gen_xml.py:104-119) and the two foot boxes.)

Ground truth here = exact intersection of the convex hulls (a feasibility LP per pair: a point that is a convex combination
of both vertex sets), at sampled joint configurations; proxies = the capsules of model/jvrc_walk.json with the kernel's
segment-segment test.  Pairs MuJoCo would test: every geom of one leg against every geom of the other (16) and the same-leg
pairs that are neither parent-child nor excluded (HIP_R-KNEE, HIP_R-foot, HIP_Y-foot per leg: 6).  Reports, per pair set
and per sampling distribution, P(hulls intersect), false positives (proxy says contact, hulls do not) and false negatives.
Runs here only (needs the reference's STL files); writes tests/golden/self_collision_eval.json.
usage: eval_collision_proxies.py [n_samples]"""
import json
import os
import sys

import numpy as np
from scipy.optimize import linprog

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.compile_model import kinematics  # noqa: E402
from tools.fit_collision_proxies import GEOM_QUAT, MESH_DIR, load_stl, quat2mat  # noqa: E402


def hulls_intersect(A, B):
    """A [n,3], B [m,3] vertex sets: is there a point in conv(A) and conv(B)?  LP feasibility."""
    ca, cb = A.mean(0), B.mean(0)
    if np.linalg.norm(ca - cb) > np.linalg.norm(A - ca, axis=1).max() + np.linalg.norm(B - cb, axis=1).max():
        return False
    n, m = len(A), len(B)
    Aeq = np.zeros((5, n + m))
    Aeq[:3, :n], Aeq[:3, n:] = A.T, -B.T
    Aeq[3, :n] = 1
    Aeq[4, n:] = 1
    r = linprog(np.zeros(n + m), A_eq=Aeq, b_eq=[0, 0, 0, 1, 1], bounds=(0, None), method="highs")
    return r.status == 0


def seg_seg_dist2(p1, q1, p2, q2):
    d1, d2, r = q1 - p1, q2 - p2, p1 - p2
    a, e, f = d1 @ d1, d2 @ d2, d2 @ r
    EPS = 1e-12
    if a <= EPS and e <= EPS:
        s = t = 0.0
    elif a <= EPS:
        s, t = 0.0, min(max(f / e, 0.0), 1.0)
    else:
        c = d1 @ r
        if e <= EPS:
            t, s = 0.0, min(max(-c / a, 0.0), 1.0)
        else:
            b = d1 @ d2
            den = a * e - b * b
            s = min(max((b * f - c * e) / den, 0.0), 1.0) if den > EPS else 0.0
            t = (b * s + f) / e
            if t < 0:
                t, s = 0.0, min(max(-c / a, 0.0), 1.0)
            elif t > 1:
                t, s = 1.0, min(max((b - c) / a, 0.0), 1.0)
    w = r + d1 * s - d2 * t
    return w @ w


def main():
    n_samples = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
    mj = json.load(open(os.path.join(ROOT, "learninghumanoidwalking_b200", "model", "jvrc_walk.json")))
    li = {lk["name"]: i for i, lk in enumerate(mj["links"])}
    geoms = []       # (name, link, vertices in the link frame)
    for side in ("R", "L"):
        for part in ("HIP_R", "HIP_Y", "KNEE"):
            name = f"{side}_{part}_S"
            geoms.append((name, li[name], load_stl(os.path.join(MESH_DIR, name + ".stl")) @ quat2mat(GEOM_QUAT[part]).T))
    for g in mj["geoms"]:
        sx, sy, sz = g["size"]
        v = np.array([[a * sx, b * sy, c * sz] for a in (-1, 1) for b in (-1, 1) for c in (-1, 1)]) + np.array(g["pos"])
        geoms.append((g["name"], g["link"], v))
    caps = mj["self_collision"]["capsules"]
    assert sorted({c["name"] for c in caps}) == sorted(g[0] for g in geoms)
    cap_pairs = [tuple(p) for p in mj["self_collision"]["pairs"]]
    cap_geom = [[g[0] for g in geoms].index(c["name"]) for c in caps]
    nj = (len(mj["links"]) - 1) // 2
    leg = lambda k: 0 if geoms[k][1] <= nj else 1
    cross = [(a, b) for a in range(len(geoms)) for b in range(len(geoms)) if leg(a) == 0 and leg(b) == 1]
    # same leg: not parent-child (HIP_R-HIP_Y, HIP_Y-KNEE), not the explicit exclude KNEE-foot (gen_xml.py:133-134)
    part = lambda k: geoms[k][0].split("_", 1)[1].rsplit("_S", 1)[0] if "foot" not in geoms[k][0] else "foot"
    same = [(a, b) for a in range(len(geoms)) for b in range(a + 1, len(geoms)) if leg(a) == leg(b)
            and {part(a), part(b)} in ({"HIP_R", "KNEE"}, {"HIP_R", "foot"}, {"HIP_Y", "foot"})]
    lo = np.array([lk["joint"]["range"][0] for lk in mj["links"][1:]])
    hi = np.array([lk["joint"]["range"][1] for lk in mj["links"][1:]])
    nominal = np.array(mj["cfg"]["nominal_qpos"])
    rng = np.random.RandomState(0)
    out = {"n_samples": n_samples, "pairs": {"cross_leg": len(cross), "same_leg": len(same)}, "distributions": {}}
    def gait_poses():
        """states the CPU oracle visits under the actor of a finished training run (tests/golden/trained_actor_jvrc_walk.pt,
        deterministic mean + N(0, 0.05^2)): the poses of an actual gait.  The policy was trained WITH the proxies ending its episodes,
        so proxy hits are absent by selection; what this distribution measures is whether the hulls touch where the proxies do not."""
        import torch
        from learninghumanoidwalking_b200.rl.policies import install_reference_aliases
        from oracle.oracle import Oracle
        install_reference_aliases()
        actor = torch.load(os.path.join(ROOT, "tests", "golden", "trained_actor_jvrc_walk.pt"), map_location="cpu", weights_only=False).double().eval()
        o = Oracle("jvrc_walk")
        n = 16
        envs = o.make_envs(n, seed=31, first_id=5)
        obs = o.batch_reset(envs, n)
        r2 = np.random.RandomState(3)
        poses = []
        while len(poses) < n_samples:
            with torch.no_grad():
                act = actor(torch.from_numpy(obs), deterministic=True).numpy()
            obs = o.batch_step(envs, n, act + 0.05 * r2.normal(size=(n, 12)), max_traj_len=400)[0]
            poses += [np.asarray(o.field(envs, i, "qpos")).copy() for i in range(n)]
        return poses[:n_samples]

    dists = ["nominal + N(0, 0.35 rad) (what an untrained / exploring policy visits)", "uniform inside the joint ranges"]
    if os.path.exists(os.path.join(ROOT, "tests", "golden", "trained_actor_jvrc_walk.pt")):
        dists.append("states of a trained walking policy in the oracle (400-step episodes, no falls)")
    for dist_name in dists:
        stats = {k: dict(hull=0, proxy=0, fp=0, fn=0) for k in ("cross_leg", "same_leg")}
        gait = gait_poses() if dist_name.startswith("states of") else None
        for si in range(n_samples):
            q = nominal.copy()
            if dist_name.startswith("nominal"):
                q[7:] = np.clip(nominal[7:] + rng.normal(size=12) * 0.35, lo, hi)
            elif gait is not None:
                q = gait[si]
            else:
                q[7:] = rng.uniform(lo, hi)
            xpos, xmat = kinematics(mj, q)
            W = [xpos[lk] + v @ np.asarray(xmat[lk]).T for _, lk, v in geoms]
            E = [(xpos[c["link"]] + xmat[c["link"]] @ np.array(c["p0"]), xpos[c["link"]] + xmat[c["link"]] @ np.array(c["p1"]), c["radius"]) for c in caps]
            for key, pairs in (("cross_leg", cross), ("same_leg", same)):
                hull = any(hulls_intersect(W[a], W[b]) for a, b in pairs)
                # the proxy side: the model's own pair list (cross-leg); same-leg geom pairs map to every capsule pair of the two geoms
                cps = cap_pairs if key == "cross_leg" else [(i, j) for i in range(len(caps)) for j in range(len(caps))
                                                            if (cap_geom[i], cap_geom[j]) in pairs]
                proxy = any(seg_seg_dist2(E[a][0], E[a][1], E[b][0], E[b][1]) < (E[a][2] + E[b][2]) ** 2 for a, b in cps)
                s = stats[key]
                s["hull"] += hull; s["proxy"] += proxy; s["fp"] += (proxy and not hull); s["fn"] += (hull and not proxy)
        out["distributions"][dist_name] = {k: {a: b / n_samples for a, b in v.items()} for k, v in stats.items()}
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "self_collision_eval.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
