#!/usr/bin/env python3
"""Fit capsule proxies to the convex collision meshes the reference keeps on the legs.

The reference terminates an episode on ANY robot-robot contact (`RobotInterface.check_self_collisions`,
envs/common/robot_interface.py:472-484) and the geoms that can produce one are the convex hulls of
{R,L}_{HIP_R,HIP_Y,KNEE}_S (envs/jvrc/gen_xml.py:104-119) plus the two foot boxes.  Mesh-mesh narrow phase is
not worth a kernel; the device path tests capsule proxies instead (termination flag only, no contact force —
the episode ends at that control step anyway).  This script derives the capsules from the STL hulls (numbers
only — no mesh data is copied into the repo) and writes them into the compiled model JSON.

Capsule = segment p0-p1 (link frame) + radius: axis = first principal axis of the hull vertices, radius = the
smaller lateral half-extent, half-length so that the end caps reach the extreme vertices.  The thigh (HIP_Y) and shin
(KNEE) hulls are visibly wider in one lateral direction than in the other: they get TWO such capsules, offset along the wide
lateral axis so that together they span it ("stadium" cross-section).  tools/eval_collision_proxies.py measures the proxies
against exact hull intersection on sampled poses: one capsule per hull missed 5.1 % of the poses in which hulls touch (false
positives 0.08 %); with the split thigh / shin proxies 1.2 % / 0.75 %.  Same-leg pairs never intersect inside the joint
ranges, so only cross-leg pairs are listed.
"""
import json
import os
import struct

import numpy as np

REF = os.environ.get("LHW_REFERENCE", "/root/reference")
MESH_DIR = os.path.join(REF, "models/jvrc_mj_description/meshes/convex")
MODEL = os.path.join(os.path.dirname(__file__), "..", "learninghumanoidwalking_b200", "model", "jvrc_walk.json")
# geom quat of each kept mesh in jvrc1.xml (w x y z); KNEE has none
GEOM_QUAT = {"HIP_R": (0.707105, 0, 0, 0.707108), "HIP_Y": (0.707105, 0, 0, 0.707108), "KNEE": (1, 0, 0, 0)}


def load_stl(path):
    b = open(path, "rb").read()
    n = struct.unpack("<I", b[80:84])[0]
    if 84 + 50 * n == len(b):
        a = np.frombuffer(b, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]), count=n, offset=84)
        return np.unique(np.round(a["v"].reshape(-1, 3).astype(np.float64), 7), axis=0)
    v = [[float(x) for x in ln.split()[1:4]] for ln in b.decode(errors="ignore").splitlines() if ln.strip().startswith("vertex")]
    return np.unique(np.round(np.array(v), 7), axis=0)


def quat2mat(q):
    q = np.asarray(q, float)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def fit_capsules(v, split):
    """One capsule along the first principal axis (radius = the smaller lateral half-extent), or — `split` — two of them
    offset along the wider lateral axis so that their union spans it."""
    c = v.mean(0)
    _, _, vt = np.linalg.svd(v - c)
    ax = vt[0]
    t = (v - c) @ ax
    lat = (v - c) @ vt[1:].T
    ext, mid = 0.5 * (lat.max(0) - lat.min(0)), 0.5 * (lat.max(0) + lat.min(0))
    k = int(np.argmin(ext))
    r, wide = float(ext[k]), 1 - k
    lo, hi = t.min() + r, t.max() - r
    if hi < lo:
        lo = hi = 0.5 * (t.min() + t.max())
    cc = c + vt[1 + wide] * mid[wide] + vt[1 + k] * mid[k]        # centre of the lateral bounding rectangle
    off = max(0.0, float(ext[wide]) - r)
    if not split or off < 1e-4:
        return [((cc + lo * ax), (cc + hi * ax), r)]
    return [((cc + s * off * vt[1 + wide] + lo * ax), (cc + s * off * vt[1 + wide] + hi * ax), r) for s in (-1, 1)]


def main():
    m = json.load(open(MODEL))
    li = {lk["name"]: i for i, lk in enumerate(m["links"])}
    caps = []
    for side in ("R", "L"):
        for part in ("HIP_R", "HIP_Y", "KNEE"):
            name = f"{side}_{part}_S"
            v = load_stl(os.path.join(MESH_DIR, name + ".stl")) @ quat2mat(GEOM_QUAT[part]).T
            for p0, p1, r in fit_capsules(v, split=part in ("HIP_Y", "KNEE")):
                caps.append(dict(name=name, link=li[name], p0=[float("%.5g" % x) for x in p0], p1=[float("%.5g" % x) for x in p1],
                                 radius=float("%.4g" % r)))
    # foot boxes (envs/jvrc/gen_xml.py:125-130): a capsule along the box's long (x) axis, radius = half the box height +
    # half of the remaining half-width, so the proxy is between the inscribed and the circumscribed one
    for g in m["geoms"]:
        sx, sy, sz = g["size"]
        r = 0.5 * (sz + sy)
        px, py, pz = g["pos"]
        caps.append(dict(name=g["name"], link=g["link"], p0=[px - (sx - r), py, pz], p1=[px + (sx - r), py, pz], radius=r))
    # pairs that can touch: everything of one leg against everything of the other (same-leg pairs are either
    # parent-child filtered, explicitly excluded (KNEE-ANKLE_P, gen_xml.py:133-134) or need extreme flexion)
    nl = len(m["links"])
    nj = (nl - 1) // 2
    right = [i for i, c in enumerate(caps) if 1 <= c["link"] <= nj]
    left = [i for i, c in enumerate(caps) if c["link"] > nj]
    pairs = [[a, b] for a in right for b in left]
    block = dict(capsules=caps, pairs=pairs,
                 note="capsule proxies of the convex leg hulls (two per thigh / shin hull) + foot boxes; termination flag only "
                      "(no contact force); accuracy against exact hull intersection: tests/golden/self_collision_eval.json")
    for name in ("jvrc_walk", "jvrc_step", "jvrc_walk_terrain"):      # the three JVRC-1 models share the leg geometry
        path = os.path.join(os.path.dirname(MODEL), name + ".json")
        mm = json.load(open(path))
        mm["self_collision"] = block
        json.dump(mm, open(path, "w"), indent=1)
        open(path, "a").write("\n")
    for c in caps:
        print(c)
    print(len(pairs), "pairs")


if __name__ == "__main__":
    main()
