"""H1: what the leg-vs-leg self-collision flag misses.  (CPU study; reads the reference's h1.xml, so it runs in the build container only.)

`StandingTask.done` ends an episode on ANY robot-robot contact (envs/common/robot_interface.py:472-484).  The kernel and the oracle
flag the leg-vs-leg pairs of h1.xml's own capsule / sphere primitives (model/h1.json["self_collision"]); the hip cylinders
(h1.xml:70,76,84), the torso box / hip capsule / head (h1.xml:151-154) and the welded arms (h1.xml:166-179) are not modelled.  This
script measures what that omission is worth: every collision geom of h1.xml is placed by forward kinematics of model/h1.json's link
table (the waist and arm joints are removed by envs/h1/gen_xml.py:49-61, so torso and arms ride on the pelvis), MuJoCo's pair filter
is restated (same weld body and weld parent-child pairs are skipped; the two <exclude> entries are moot once the arms are welded),
and each candidate pair is tested on sphere-swept convex cores: capsule = segment + r, sphere = point + r, cylinder = two 48-gon rims,
box = 8 vertices; distance between the cores by Gilbert's (Frank-Wolfe) iteration with an exact segment-segment shortcut for
capsule / sphere pairs.  Poses: (A) states an oracle rollout visits before termination under sigma = 1.0 actions (the regime that
exercises the flag), (B) leg joints uniform over their ranges.

    python tools/eval_h1_self_collision.py [out.json]
"""
import json
import os
import sys
import xml.etree.ElementTree as ET

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from shin_clearance import link_poses, quat2mat   # noqa: E402

H1_XML = "/root/reference/models/mujoco_menagerie/unitree_h1/h1.xml"
FOOT = {"foot1": ([-.035, 0, -0.056], [.02, 0, -0.045]), "foot2": ([.02, 0, -0.045], [.115, 0, -0.056]),
        "foot3": ([.14, -.03, -0.056], [.14, .03, -0.056])}     # h1.xml:14-25 (defaults)


def vec(s, n=None):
    v = np.array([float(x) for x in s.split()])
    return v if n is None else v[:n]


def rims(r, h, n=48):
    a = np.linspace(0, 2 * np.pi, n, endpoint=False)
    ring = np.stack([r * np.cos(a), r * np.sin(a), np.zeros(n)], 1)
    return np.concatenate([ring + [0, 0, h], ring - [0, 0, h]])


def collect_geoms(link_names):
    """[(link index, xml body name, geom label, core points in the LINK frame, radius)] for every collision geom of the robot."""
    root = ET.parse(H1_XML).getroot()
    out = []

    def walk(body, link, R, p):
        name = body.get("name")
        if name in link_names:
            link, R, p = link_names.index(name), np.eye(3), np.zeros(3)
        else:     # welded body: accumulate its pose in the frame of the link it rides on
            bp = vec(body.get("pos", "0 0 0"))
            bq = vec(body.get("quat", "1 0 0 0"))
            p, R = p + R @ bp, R @ quat2mat(bq / np.linalg.norm(bq))
        k = 0
        for g in body.findall("geom"):
            cls = g.get("class", "")
            if cls == "visual":
                continue
            label = g.get("name") or f"{name}:{g.get('type', cls)}{k}"
            k += 1
            if cls in FOOT:
                core, r = np.array(FOOT[cls], float), 0.014
            else:
                typ, size = g.get("type"), vec(g.get("size"))
                gp = vec(g.get("pos", "0 0 0"))
                gq = vec(g.get("quat", "1 0 0 0"))
                gR = quat2mat(gq / np.linalg.norm(gq))
                if typ == "capsule":
                    ft = vec(g.get("fromto"))
                    core, r = np.array([ft[:3], ft[3:]]), size[0]
                elif typ == "sphere":
                    core, r = np.array([gp]), size[0]
                elif typ == "cylinder":
                    core, r = gp + rims(size[0], size[1]) @ gR.T, 0.0
                elif typ == "box":
                    s = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], float) * size
                    core, r = gp + s @ gR.T, 0.0
                else:
                    raise ValueError(typ)
            out.append((link, name, label, p + core @ R.T, float(r)))
        for b in body.findall("body"):
            walk(b, link, R, p)

    pelvis = root.find("worldbody").find("body")
    walk(pelvis, 0, np.eye(3), np.zeros(3))
    return out


def seg_seg(p1, q1, p2, q2):
    """Distance between two segments (points allowed)."""
    d1, d2, r = q1 - p1, q2 - p2, p1 - p2
    a, e, f = d1 @ d1, d2 @ d2, d2 @ r
    if a < 1e-14 and e < 1e-14:
        return np.linalg.norm(r)
    if a < 1e-14:
        s, t = 0.0, np.clip(f / e, 0, 1)
    else:
        c = d1 @ r
        if e < 1e-14:
            t, s = 0.0, np.clip(-c / a, 0, 1)
        else:
            b = d1 @ d2
            den = a * e - b * b
            s = np.clip((b * f - c * e) / den, 0, 1) if den > 1e-14 else 0.0
            t = (b * s + f) / e
            if t < 0:
                t, s = 0.0, np.clip(-c / a, 0, 1)
            elif t > 1:
                t, s = 1.0, np.clip((b - c) / a, 0, 1)
    return np.linalg.norm(p1 + s * d1 - p2 - t * d2)


def hull_dist(A, B, stop):
    """Distance between conv(A) and conv(B): Gilbert / Frank-Wolfe with exact line search on the Minkowski difference; stops
    early once the lower bound exceeds `stop` (no contact) or the upper bound is below it (contact)."""
    x = A.mean(0) - B.mean(0)
    for _ in range(200):
        n = np.linalg.norm(x)
        if n < 1e-9:
            return 0.0
        s = A[np.argmin(A @ x)] - B[np.argmax(B @ x)]       # support point of A - B towards -x
        lower = (x @ s) / n                                 # n is an upper bound of the distance, lower a lower bound
        if lower > stop:
            return lower
        if n < stop or n - lower < 1e-6:
            return n
        d = s - x
        t = np.clip(-(x @ d) / (d @ d), 0, 1)
        x = x + t * d
    return np.linalg.norm(x)


def in_contact(ga, gb):
    (A, ra), (B, rb) = ga, gb
    rr = ra + rb
    ca, cb = A.mean(0), B.mean(0)
    if np.linalg.norm(ca - cb) - np.abs(A - ca).sum(1).max() - np.abs(B - cb).sum(1).max() > rr:
        return False
    if len(A) <= 2 and len(B) <= 2:
        return seg_seg(A[0], A[-1], B[0], B[-1]) < rr
    return hull_dist(A, B, rr) < rr


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "h1_self_collision_eval.json")
    from oracle.oracle import Oracle, load_model_json
    mj = load_model_json("h1")
    links = mj["links"]
    names = [l["name"] for l in links]
    geoms = collect_geoms(names)
    parent = [l["parent"] for l in links]
    # the flag's pair set, mapped onto the XML geoms by link + end points
    def geom_of(c):
        for gi, (link, _, _, core, r) in enumerate(geoms):
            if link == c["link"] and len(core) <= 2 and abs(r - c["radius"]) < 1e-6 and \
                    np.allclose(core[0], c["p0"], atol=1e-4) and np.allclose(core[-1], c["p1"], atol=1e-4):
                return gi
        raise KeyError(c["name"])
    gidx = [geom_of(c) for c in mj["self_collision"]["capsules"]]
    modelled = {frozenset((gidx[a], gidx[b])) for a, b in mj["self_collision"]["pairs"]}

    def is_modelled(i, j):
        return frozenset((i, j)) in modelled

    pairs = []
    for i in range(len(geoms)):
        for j in range(i + 1, len(geoms)):
            la, lb = geoms[i][0], geoms[j][0]
            if la == lb or parent[la] == lb or parent[lb] == la:
                continue              # same weld body / weld parent-child: MuJoCo's default filter
            pairs.append((i, j, is_modelled(i, j)))

    def group(g):
        link, body, label, _, _ = g
        if link == 0:
            return "arm" if ("shoulder" in body or "elbow" in body) else "torso/head/hip-capsule"
        if "cylinder" in label:
            return "hip cylinder"
        return "leg primitive"

    rng = np.random.default_rng(0)
    # (A) states visited before termination under sigma = 1.0 actions
    o = Oracle("h1")
    n = 48
    envs = o.make_envs(n, seed=0)
    o.batch_reset(envs, n)
    posesA = []
    for t in range(120):
        res = o.batch_step(envs, n, rng.standard_normal((n, o.nu)))
        ended = np.asarray(res[-1]).astype(bool)
        for i in range(n):
            if not ended[i] and (t * n + i) % 3 == 0:
                posesA.append(np.asarray(o.field(envs, i, "qpos")).copy())
    # (B) leg joints uniform over their ranges
    lo = np.array([l["joint"]["range"][0] for l in links[1:]])
    hi = np.array([l["joint"]["range"][1] for l in links[1:]])
    posesB = [np.concatenate([[0, 0, 1, 1, 0, 0, 0], lo + (hi - lo) * rng.random(len(lo))]) for _ in range(1500)]

    def study(poses):
        stats = {"poses": len(poses), "any_contact": 0, "modelled_flag": 0, "missed": 0, "missed_by_group": {}, "examples": {}}
        for q in poses:
            R, p = link_poses(links, q)
            world = [(p[g[0]] + g[3] @ R[g[0]].T, g[4]) for g in geoms]
            hit_mod, hit_un = False, []
            for i, j, mod in pairs:
                if mod and hit_mod:
                    continue
                if in_contact(world[i], world[j]):
                    if mod:
                        hit_mod = True
                    else:
                        hit_un.append((i, j))
            stats["any_contact"] += bool(hit_mod or hit_un)
            stats["modelled_flag"] += hit_mod
            if hit_un and not hit_mod:
                stats["missed"] += 1
                seen = set()
                for i, j in hit_un:
                    key = " vs ".join(sorted((group(geoms[i]), group(geoms[j]))))
                    if key not in seen:
                        seen.add(key)
                        stats["missed_by_group"][key] = stats["missed_by_group"].get(key, 0) + 1
                        stats["examples"].setdefault(key, f"{geoms[i][2]} x {geoms[j][2]}")
        k = max(1, stats["poses"])
        stats["fraction_any_contact"] = stats["any_contact"] / k
        stats["fraction_flagged_by_the_modelled_pairs"] = stats["modelled_flag"] / k
        stats["fraction_missed"] = stats["missed"] / k
        stats["false_negative_rate"] = stats["missed"] / max(1, stats["any_contact"])
        return stats

    rep = {"geoms": len(geoms), "candidate_pairs": len(pairs), "modelled_pairs": sum(m for _, _, m in pairs),
           "rollout_sigma_1.0": study(posesA), "uniform_joint_ranges": study(posesB)}
    json.dump(rep, open(out, "w"), indent=1)
    print(json.dumps(rep, indent=1))


if __name__ == "__main__":
    main()
