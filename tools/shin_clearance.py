"""How often would the leg collision meshes of jvrc_step touch a stepping stone?  (CPU study, oracle only.)

The reference keeps convex collision meshes on {R,L}_{HIP_R,HIP_Y,KNEE}_S (envs/jvrc/gen_xml.py:104-119) and the 20 stepping
stones are ordinary world boxes (tasks/stepping_task.py:318-334), so MuJoCo would also collide thighs and shins with the stones.
The step kernel and the oracle model the foot boxes against the stones (top faces + risers) and nothing else.  This script puts a
number on what that omission is worth BEFORE termination: it rolls the oracle's jvrc_step out under the early-training action
distribution (N(0, 0.223^2), the freshly initialised policy's spread) at the top of the height curriculum, places the fitted leg
capsules of model/jvrc_step.json["self_collision"] (tools/fit_collision_proxies.py: they enclose the hulls, so a capsule that
clears a stone means the mesh clears it) by forward kinematics of the model's link table, and measures the signed distance of every
thigh / shin capsule to every stone (box signed-distance of the segment, minus the radius; sampled along the segment).

Output: tests/golden/shin_clearance.json — fraction of (env, control step) pairs with a capsule inside a stone, split by capsule,
the clearance quantiles, and how many of the penetrating steps are the last step of an episode (the fall that terminates it).

    python tools/shin_clearance.py [n_envs] [steps] [out.json] [action sigma]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def quat2mat(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def axis_rot(axis, a):
    x, y, z = axis
    c, s = np.cos(a), np.sin(a)
    K = np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]])
    return np.eye(3) + s * K + (1 - c) * (K @ K)


def link_poses(links, qpos):
    """World (R, p) of every link frame: free root from qpos[0:7], hinge k about its axis at the link origin (the model table of
    model/*.json: `pos` / `rot` relative to the parent link, joints in link order)."""
    R, p = [None] * len(links), [None] * len(links)
    hinge = 7
    for i, l in enumerate(links):
        if l["joint"]["type"] == "free":
            q = np.asarray(qpos[3:7], float)
            R[i], p[i] = quat2mat(q / np.linalg.norm(q)), np.asarray(qpos[0:3], float)
            continue
        pr, pp = R[l["parent"]], p[l["parent"]]
        Rl = pr @ np.asarray(l["rot"])
        R[i] = Rl @ axis_rot(l["joint"]["axis"], qpos[hinge])
        p[i] = pp + pr @ np.asarray(l["pos"])
        hinge += 1
    return R, p


def box_sdf(pts, slab, half):
    """Signed distance of world points to the stone `slab` = (x, y, z_top, yaw) of half sizes `half` (top face through z_top)."""
    c, s = np.cos(slab[3]), np.sin(slab[3])
    d = pts - np.array([slab[0], slab[1], slab[2] - half[2]])
    loc = np.stack([c * d[:, 0] + s * d[:, 1], -s * d[:, 0] + c * d[:, 1], d[:, 2]], axis=1)
    q = np.abs(loc) - np.asarray(half)
    return np.linalg.norm(np.maximum(q, 0), axis=1) + np.minimum(q.max(axis=1), 0)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
    out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "tests", "golden", "shin_clearance.json")
    sigma = float(sys.argv[4]) if len(sys.argv) > 4 else 0.223
    from oracle.oracle import Oracle, load_model_json
    mj = load_model_json("jvrc_step")
    links, half = mj["links"], mj["stepping"]["slab_half"]
    caps = [c for c in mj["self_collision"]["capsules"] if "foot" not in c["name"]]
    names = sorted({c["name"] for c in caps})
    o = Oracle("jvrc_step", iteration_count=float("inf"))
    envs = o.make_envs(n, seed=0)
    o.batch_reset(envs, n)
    rng = np.random.default_rng(0)
    samples = np.linspace(0.0, 1.0, 9)
    clear = {k: [] for k in names}           # per (env, step): min over the capsule's samples and the stones
    pen_total = pen_last = total = 0
    fk_err = 0.0
    for t in range(steps):
        a = sigma * rng.standard_normal((n, 12))
        res = o.batch_step(envs, n, a)
        ended = np.asarray(res["ended"] if isinstance(res, dict) else res[-1]).astype(bool).reshape(-1)
        for i in range(n):
            if ended[i]:
                continue                      # the env was just reset; the step that fell is counted through `last` below
            qpos = np.asarray(o.field(envs, i, "qpos"))
            seq = np.asarray(o.field(envs, i, "seq")).reshape(20, 4)
            nseq = int(np.asarray(o.field(envs, i, "seq_len")).reshape(-1)[0])
            R, p = link_poses(links, qpos)
            worst = np.inf
            for c in caps:
                p0 = p[c["link"]] + R[c["link"]] @ np.asarray(c["p0"])
                p1 = p[c["link"]] + R[c["link"]] @ np.asarray(c["p1"])
                pts = p0[None] + samples[:, None] * (p1 - p0)[None]
                d = min(box_sdf(pts, seq[k], half).min() for k in range(nseq)) - c["radius"] if nseq else np.inf
                clear[c["name"]].append(float(d))
                worst = min(worst, d)
            total += 1
            if worst < 0:
                pen_total += 1
                z = qpos[2]
                pen_last += int(z < 0.65)     # root about to cross the 0.6 m termination height (tasks/stepping_task.py:300-306)
    rep = {"n_envs": n, "steps": steps, "pairs": total, "action_sigma": sigma, "iteration_count": "inf (top of the height curriculum)",
           "fraction_of_steps_with_a_leg_capsule_inside_a_stone": pen_total / max(1, total),
           "of_those_root_below_0.65m": pen_last / max(1, pen_total),
           "per_capsule": {k: {"fraction_inside": float(np.mean(np.asarray(v) < 0)),
                               "clearance_m_quantiles_0_1_5_50": [float(x) for x in np.quantile(v, [0, 0.01, 0.05, 0.5])]}
                           for k, v in clear.items()}}
    json.dump(rep, open(out, "w"), indent=1)
    print(json.dumps(rep, indent=1))


if __name__ == "__main__":
    main()
