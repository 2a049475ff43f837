#!/usr/bin/env python3
"""Generate tests/golden/step_task.json by RUNNING the reference's tasks/stepping_task.py:SteppingTask in this container.

/root/reference does not exist on the GPU box, so the vectors are committed; this script is the provenance.
mujoco / transforms3d are not installable here, so the real class runs against stand-ins:

  * `transforms3d` — a module object holding the six functions SteppingTask calls (euler2quat, quat2euler, euler2mat,
    mat2euler in the default 'sxyz' convention, affines.compose, quaternions.quat2mat), restated from the package's
    published formulas and cross-checked against scipy.spatial.transform here;
  * RobotInterface — a recorder object with the accessors the task calls (body / site poses, foot velocities, ground
    reaction forces, contacts, `model.body(name).pos` ... for the 20 boxes and the floor);
  * numpy's / python's global RNGs — `np.random.choice/uniform/randint` and `random.choice` consume, in the reference's
    call order, the Philox words the oracle draws for the same (seed, env, event counter) key
    (oracle/sim_oracle.c:task_reset_step documents the stream / lane of each draw), so the reference code fed the same
    uniforms must produce exactly the oracle's mode, phase, footstep sequence and box poses.
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np
from scipy.spatial.transform import Rotation

REF = os.environ.get("LHW_REFERENCE", "/root/reference")
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
EPS4 = 4 * np.finfo(float).eps
TASK_MASS = None


# ---------------------------------------------------------------- transforms3d stand-in ('sxyz' = static x, y, z)
def quat2mat(q):
    w, x, y, z = q
    Nq = w * w + x * x + y * y + z * z
    if Nq < np.finfo(float).eps:
        return np.eye(3)
    s = 2.0 / Nq
    X, Y, Z = x * s, y * s, z * s
    wX, wY, wZ, xX, xY, xZ, yY, yZ, zZ = w * X, w * Y, w * Z, x * X, x * Y, x * Z, y * Y, y * Z, z * Z
    return np.array([[1.0 - (yY + zZ), xY - wZ, xZ + wY], [xY + wZ, 1.0 - (xX + zZ), yZ - wX], [xZ - wY, yZ + wX, 1.0 - (xX + yY)]])


def euler2mat(ai, aj, ak):
    si, sj, sk, ci, cj, ck = np.sin(ai), np.sin(aj), np.sin(ak), np.cos(ai), np.cos(aj), np.cos(ak)
    cc, cs, sc, ss = ci * ck, ci * sk, si * ck, si * sk
    return np.array([[cj * ck, sj * sc - cs, sj * cc + ss], [cj * sk, sj * ss + cc, sj * cs - sc], [-sj, cj * si, cj * ci]])


def mat2euler(M):
    M = np.asarray(M, dtype=float)[:3, :3]
    cy = np.sqrt(M[0, 0] * M[0, 0] + M[1, 0] * M[1, 0])
    if cy > EPS4:
        return np.arctan2(M[2, 1], M[2, 2]), np.arctan2(-M[2, 0], cy), np.arctan2(M[1, 0], M[0, 0])
    return np.arctan2(-M[1, 2], M[1, 1]), np.arctan2(-M[2, 0], cy), 0.0


def euler2quat(ai, aj, ak):
    ai, aj, ak = ai / 2.0, aj / 2.0, ak / 2.0
    ci, si, cj, sj, ck, sk = np.cos(ai), np.sin(ai), np.cos(aj), np.sin(aj), np.cos(ak), np.sin(ak)
    cc, cs, sc, ss = ci * ck, ci * sk, si * ck, si * sk
    return np.array([cj * cc + sj * ss, cj * sc - sj * cs, cj * ss + sj * cc, cj * cs - sj * sc])


def compose(T, R, Z):
    A = np.eye(4)
    A[:3, :3] = np.asarray(R) @ np.diag(Z)
    A[:3, 3] = T
    return A


def install_transforms3d():
    for _ in range(50):   # cross-check the stand-in against scipy ('sxyz' static == scipy extrinsic 'xyz')
        e = np.random.RandomState(_).uniform(-1.2, 1.2, 3)
        Rs = Rotation.from_euler("xyz", e).as_matrix()
        assert np.abs(euler2mat(*e) - Rs).max() < 1e-14
        assert np.abs(np.array(mat2euler(Rs)) - e).max() < 1e-12
        q = euler2quat(*e)
        assert np.abs(quat2mat(q) - Rs).max() < 1e-14
    tf3 = types.ModuleType("transforms3d")
    tf3.euler = types.SimpleNamespace(euler2quat=euler2quat, quat2euler=lambda q: mat2euler(quat2mat(q)), euler2mat=euler2mat,
                                      mat2euler=mat2euler)
    tf3.affines = types.SimpleNamespace(compose=compose)
    tf3.quaternions = types.SimpleNamespace(quat2mat=quat2mat)
    sys.modules["transforms3d"] = tf3


# ---------------------------------------------------------------- RNG stand-ins fed with the oracle's Philox words
class WordFeed:
    def __init__(self):
        self.words = []

    def load(self, words):
        self.words = list(words)

    def _w(self):
        return self.words.pop(0)

    def choice(self, a, p=None):
        w = self._w()
        if p is None:
            return a[(w * len(a)) >> 32]                       # numpy: a[randint(0, len(a))]
        cdf = np.cumsum(np.asarray(p, dtype=float))
        cdf /= cdf[-1]
        return a[int(cdf.searchsorted((w >> 8) * (1.0 / 16777216.0), side="right"))]

    def uniform(self, lo, hi):
        return lo + (hi - lo) * ((self._w() >> 8) * (1.0 / 16777216.0))

    def randint(self, lo, hi):
        return lo + ((self._w() * (hi - lo)) >> 32)


class Named:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class Client:
    """The slice of RobotInterface that SteppingTask touches."""

    def __init__(self):
        self.boxes = {"box" + repr(i + 1).zfill(2): Named(pos=np.array([0.0, 0.0, -0.2]), quat=np.array([1.0, 0, 0, 0]))
                      for i in range(20)}
        self.boxes["floor"] = Named(pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]))
        self.geoms = {k: Named(size=np.array([1.0, 1.0, 0.1]), rgba=np.ones(4)) for k in self.boxes}
        self.model = Named(body=lambda n: self.boxes[n], geom=lambda n: self.geoms[n])
        self.pose = {}
        self.contacts_r, self.contacts_l = [], []
        self.selfcol = False

    def get_robot_mass(self): return TASK_MASS      # mj_getTotalmass incl. the 20 static boxes
    def get_object_xpos_by_name(self, name, typ): return self.pose[name][0].copy()
    def get_object_xquat_by_name(self, name, typ): return self.pose[name][1].copy()
    def get_lfoot_body_pos(self): return self.pose["lfoot"][0].copy()
    def get_rfoot_body_pos(self): return self.pose["rfoot"][0].copy()
    def get_lfoot_body_vel(self): return [self.lvel.copy(), np.zeros(3)]
    def get_rfoot_body_vel(self): return [self.rvel.copy(), np.zeros(3)]
    def get_lfoot_grf(self): return self.lgrf
    def get_rfoot_grf(self): return self.rgrf
    def check_rfoot_floor_collision(self): return len(self.contacts_r) > 0
    def check_lfoot_floor_collision(self): return len(self.contacts_l) > 0
    def get_rfoot_floor_contacts(self): return [(i, Named(pos=p)) for i, p in enumerate(self.contacts_r)]
    def get_lfoot_floor_contacts(self): return [(i, Named(pos=p)) for i, p in enumerate(self.contacts_l)]
    def check_self_collisions(self): return self.selfcol


def main():
    from oracle.oracle import Oracle, curriculum_height
    install_transforms3d()
    os.makedirs(OUT, exist_ok=True)
    os.chdir(REF)                      # SteppingTask opens "utils/footstep_plans.txt" relative to the cwd
    pkg = types.ModuleType("tasks")    # bare package: tasks/__init__ would import every task
    pkg.__path__ = [os.path.join(REF, "tasks")]
    sys.modules["tasks"] = pkg
    spec = importlib.util.spec_from_file_location("tasks.stepping_task", os.path.join(REF, "tasks/stepping_task.py"))
    stm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(stm)
    feed = WordFeed()
    stm.np.random.choice, stm.np.random.uniform, stm.np.random.randint = feed.choice, feed.uniform, feed.randint
    stm.random.choice = lambda seq: seq[(feed._w() * len(seq)) >> 32]
    MODES = {stm.WalkModes.CURVED: 0, stm.WalkModes.STANDING: 1, stm.WalkModes.BACKWARD: 2, stm.WalkModes.LATERAL: 3,
             stm.WalkModes.FORWARD: 4}

    o = Oracle("jvrc_step")
    global TASK_MASS
    TASK_MASS = o.mj["stepping"]["task_mass"]
    rng = np.random.RandomState(20260923)
    cases = []
    for case in range(12):
        seed, env_id, ctr = int(rng.randint(1 << 30)), int(rng.randint(1 << 20)), int(rng.randint(1, 1 << 16))
        itc = [np.inf, 0, 5000, 12000][case % 4]
        # make sure every walk mode is drawn at least once
        while True:
            u = o.philox(seed, env_id, ctr, 3)
            cm = (u[0] >> 8) / 16777216.0
            md = 0 if cm < 0.15 else 1 if cm < 0.2 else 2 if cm < 0.4 else 3 if cm < 0.7 else 4
            if case >= 5 or md == case:
                break
            ctr += 1
        u, v = o.philox(seed, env_id, ctr, 3), o.philox(seed, env_id, ctr, 4)
        cm = (u[0] >> 8) / 16777216.0
        mode = 0 if cm < 0.15 else 1 if cm < 0.2 else 2 if cm < 0.4 else 3 if cm < 0.7 else 4
        words = [u[1], u[0]] + ([v[0]] if mode == 4 else []) + ([u[2]] if mode in (0, 3) else [u[2], u[3]])
        feed.load(words)
        c = Client()
        yaw0 = rng.uniform(-0.5, 0.5)
        quat = euler2quat(rng.normal() * 0.05, rng.normal() * 0.05, yaw0)
        root = np.array([rng.normal() * 0.1, rng.normal() * 0.1, 0.8 + rng.normal() * 0.02])
        Rr = quat2mat(quat)
        lf = root + Rr @ np.array([0.09, 0.096, -0.75]) + rng.normal(size=3) * 0.01
        rf = root + Rr @ np.array([0.09, -0.096, -0.75]) + rng.normal(size=3) * 0.01
        c.pose = {"pelvis": (root, quat), "lfoot": (lf, None), "rfoot": (rf, None), "head": (root + Rr @ np.array([0, 0, 0.6]), None)}
        task = stm.SteppingTask(client=c, dt=0.025, neutral_foot_orient=np.array([1, 0, 0, 0]), root_body="pelvis",
                                lfoot_body="lfoot", rfoot_body="rfoot", head_body="head")
        task._goal_height_ref, task._total_duration, task._swing_duration, task._stance_duration = 0.80, 1.1, 0.75, 0.35
        task.reset(iter_count=itc)
        assert not feed.words, "draw order mismatch"
        rec = dict(seed=seed, env_id=env_id, rng_ctr=ctr, iteration_count=(None if np.isinf(itc) else itc),
                   step_height=curriculum_height(itc), root_xpos=root.tolist(), root_quat=quat.tolist(),
                   root_xmat=Rr.reshape(-1).tolist(), lfoot_xpos=lf.tolist(), rfoot_xpos=rf.tolist(),
                   mode=MODES[task.mode], phase=int(task._phase), period=float(task._period), delay_frames=int(task.delay_frames),
                   seq=[list(map(float, s)) for s in task.sequence], t1=int(task.t1), t2=int(task.t2),
                   box_pos=[c.boxes["box" + repr(i + 1).zfill(2)].pos.tolist() for i in range(20)],
                   box_yaw=[float(mat2euler(quat2mat(c.boxes["box" + repr(i + 1).zfill(2)].quat))[2]) for i in range(20)],
                   box_size=c.geoms["box01"].size.tolist(), floor_z=float(c.boxes["floor"].pos[2]), steps=[])
        # ---- a few control steps: the feet hover around the current target so that targets get reached and advance
        for k in range(36):
            tgt = np.array(task.sequence[task.t1][0:3])
            near = k < 33
            lsite = tgt + rng.normal(size=3) * (0.03 if near else 0.4)
            rsite = tgt + rng.normal(size=3) * 0.3 + np.array([0, -0.2, 0])
            root = root + rng.normal(size=3) * 0.01
            root[2] = max(root[2], tgt[2] + 0.7)
            quat = euler2quat(rng.normal() * 0.05, rng.normal() * 0.05, yaw0 + rng.normal() * 0.1)
            Rr = quat2mat(quat)
            head = root + Rr @ np.array([0.0, 0.0, 0.6]) + rng.normal(size=3) * 0.02
            c.pose.update({"pelvis": (root, quat), "head": (head, None), "lf_force": (lsite, np.array([1.0, 0, 0, 0])),
                           "rf_force": (rsite, np.array([1.0, 0, 0, 0]))})
            c.lvel, c.rvel = rng.normal(size=3) * 0.2, rng.normal(size=3) * 0.2
            c.lgrf, c.rgrf = float(abs(rng.normal()) * 300), float(abs(rng.normal()) * 300)
            c.contacts_r = [rng.normal(size=3) * 0.05 for _ in range(rng.randint(0, 3))]
            c.contacts_l = [rng.normal(size=3) * 0.05 for _ in range(rng.randint(0, 3))]
            if k == 35:   # last recorded step: drop the pelvis so that done() fires on the relative height
                root = np.array([root[0], root[1], min(lsite[2], rsite[2]) + 0.55])
                c.pose["pelvis"] = (root, quat)
            task.step()
            r = task.calc_reward(None, None, None)
            cz = min([p[2] for p in c.contacts_r + c.contacts_l]) if (c.contacts_r or c.contacts_l) else 0.0
            rec["steps"].append(dict(
                root_xpos=root.tolist(), root_quat=quat.tolist(), head_xpos=head.tolist(),
                lsite=lsite.tolist(), rsite=rsite.tolist(), lvel=c.lvel.tolist(), rvel=c.rvel.tolist(), lgrf=c.lgrf, rgrf=c.rgrf,
                ncon_r=len(c.contacts_r), ncon_l=len(c.contacts_l), contact_z_min=float(cz),
                phase=int(task._phase), t1=int(task.t1), t2=int(task.t2), target_reached=bool(task.target_reached),
                frames=int(task.target_reached_frames),
                goal_steps=[float(x) for x in (list(task._goal_steps_x) + list(task._goal_steps_y) + list(task._goal_steps_z)
                                               + list(task._goal_steps_theta))],
                terms=[float(x) for x in r.values()], done=bool(task.done())))
            rec["names"] = list(r.keys())
        cases.append(rec)
    json.dump(cases, open(os.path.join(OUT, "step_task.json"), "w"))
    modes = [c["mode"] for c in cases]
    print("wrote step_task.json:", len(cases), "cases, modes", np.bincount(modes, minlength=5).tolist(),
          "advances", sum(s["t1"] > 0 for c in cases for s in c["steps"][-1:]),
          "dones", sum(s["done"] for c in cases for s in c["steps"]))


if __name__ == "__main__":
    main()
