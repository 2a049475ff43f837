#!/usr/bin/env python3
"""Generate tests/golden/walk_task.json by RUNNING the reference's tasks/walking_task.py:WalkingTask (reset / step / calc_reward /
done, with manip_hfield=True so that the otherwise unused terrain hook :172-179 is exercised too) in this container.

The class only needs numpy, scipy (gait clocks) and a RobotInterface; the interface is a recorder object here.  numpy's global RNG
is replaced by a feed of the Philox words the oracle draws for the same (seed, env, event counter) key, in the reference's call
order (oracle/sim_oracle.c: task_reset -> stream 3 lane 0 mode, lane 1 phase, stream 4 mode_ref; task_step -> stream 0 lane 0
`randint(100)`, lane 1 `randint(200)`, lane 2 the hook's `randint(200)`, streams 1 / 2 the re-sampled mode_ref, stream 5 the hook's
pose), so the reference code fed the same uniforms must produce the oracle's task state.  Event counters are searched so that every
random switch actually fires in the recorded cases.
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

REF = os.environ.get("LHW_REFERENCE", "/root/reference")
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


class WordFeed:
    def __init__(self):
        self.words = []

    def _w(self):
        return self.words.pop(0)

    def _u(self):
        return (self._w() >> 8) * (1.0 / 16777216.0)

    def choice(self, a, p=None):
        if p is None:
            return a[(self._w() * len(a)) >> 32]
        cdf = np.cumsum(np.asarray(p, dtype=float))
        cdf /= cdf[-1]
        return a[int(cdf.searchsorted(self._u(), side="right"))]

    def uniform(self, lo, hi, size=None):
        if size is None:
            return lo + (hi - lo) * self._u()
        return np.array([lo + (hi - lo) * self._u() for _ in range(size)])

    def randint(self, lo, hi=None):
        if hi is None:
            lo, hi = 0, lo
        return int(lo + ((self._w() * int(hi - lo)) >> 32))


class Named:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class Client:
    def __init__(self):
        self.hfield = Named(pos=np.zeros(3))
        self.model = Named(geom=lambda n: self.hfield)
        self.s = {}

    def get_robot_mass(self): return 62.4
    def get_lfoot_body_vel(self, frame=0): return [self.s["lvel"].copy(), np.zeros(3)]
    def get_rfoot_body_vel(self, frame=0): return [self.s["rvel"].copy(), np.zeros(3)]
    def get_lfoot_grf(self): return self.s["lgrf"]
    def get_rfoot_grf(self): return self.s["rgrf"]
    def get_object_xpos_by_name(self, name, typ): return self.s[name].copy()
    def get_body_vel(self, name, frame=0): return [self.s["root_vloc"].copy(), np.zeros(3)]
    def get_qvel(self): return self.s["qvel"].copy()
    def get_qacc(self): return self.s["qacc"].copy()
    def get_qpos(self): return self.s["qpos"].copy()
    def get_act_joint_torques(self): return self.s["torque"].copy()
    def get_act_joint_positions(self): return list(self.s["pose"])
    def check_rfoot_floor_collision(self): return len(self.s["con_r"]) > 0
    def check_lfoot_floor_collision(self): return len(self.s["con_l"]) > 0
    def get_rfoot_floor_contacts(self): return [(i, Named(pos=p)) for i, p in enumerate(self.s["con_r"])]
    def get_lfoot_floor_contacts(self): return [(i, Named(pos=p)) for i, p in enumerate(self.s["con_l"])]
    def check_self_collisions(self): return self.s["selfcol"]


def main():
    from oracle.oracle import Oracle
    pkg = types.ModuleType("tasks")
    pkg.__path__ = [os.path.join(REF, "tasks")]
    sys.modules["tasks"] = pkg
    spec = importlib.util.spec_from_file_location("tasks.walking_task", os.path.join(REF, "tasks/walking_task.py"))
    wt = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(wt)
    feed = WordFeed()
    wt.np.random.choice, wt.np.random.uniform, wt.np.random.randint = feed.choice, feed.uniform, feed.randint
    MODES = {wt.WalkModes.STANDING: 0, wt.WalkModes.INPLACE: 1, wt.WalkModes.FORWARD: 2}
    o = Oracle("jvrc_walk")
    cfg = o.mj["cfg"]
    nominal = np.array(cfg["nominal_qpos"][7:])
    rng = np.random.RandomState(77)
    ref_words = lambda mode, w: ([w[0], w[1], w[2]] if mode == wt.WalkModes.STANDING else [w[0]])

    cases = []
    for case in range(10):
        seed, env_id = int(rng.randint(1 << 30)), int(rng.randint(1 << 20))
        ctr = int(rng.randint(1, 1 << 16))
        c = Client()
        task = wt.WalkingTask(client=c, dt=cfg["control_dt"], neutral_foot_orient=np.array([1, 0, 0, 0]), neutral_pose=nominal,
                              root_body="root", lfoot_body="lfoot", rfoot_body="rfoot", head_body="head", manip_hfield=True)
        t = cfg["task"]
        task._goal_height_ref, task._total_duration, task._swing_duration, task._stance_duration = \
            t["goal_height"], t["total_duration"], t["swing_duration"], t["stance_duration"]
        # ---- reset: choice(mode), sample_ref, randint(0, period)
        u3, u4 = o.philox(seed, env_id, ctr, 3), o.philox(seed, env_id, ctr, 4)
        cm = (u3[0] >> 8) / 16777216.0
        mode0 = wt.WalkModes.STANDING if cm < 0.6 else wt.WalkModes.INPLACE if cm < 0.8 else wt.WalkModes.FORWARD
        feed.words = [u3[0]] + ref_words(mode0, u4) + [u3[1]]
        task.reset(iter_count=0)
        assert not feed.words
        rec = dict(seed=seed, env_id=env_id, reset_ctr=ctr, mode=MODES[task.mode], mode_ref=[float(x) for x in task.mode_ref],
                   phase=int(task._phase), period=float(task._period), steps=[])
        # ---- control steps at event counters where the random switches fire (searched), plus ordinary ones
        period = int(task._period)
        dbl = [ph for ph in range(period) if task.right_clock[0](ph) == 1 and task.left_clock[0](ph) == 1]
        want = ["plain", "sw100", "sw200", "hook", "plain", "sw100", "sw200", "hook"]
        k = ctr
        for kind in want:
            while True:
                k += 1
                u0 = o.philox(seed, env_id, k, 0)
                f100, f200, fh = (u0[0] * 100) >> 32 == 0, (u0[1] * 200) >> 32 == 0, (u0[2] * 200) >> 32 == 0
                if kind == "plain" and not (f100 or f200 or fh): break
                if kind == "sw100" and f100 and not f200: break
                if kind == "sw200" and f200 and not f100: break
                if kind == "hook" and fh and not (f100 or f200): break
            if kind == "sw100":
                task._phase = dbl[int(rng.randint(len(dbl)))] - 1          # the switch needs double support
                if task.mode == wt.WalkModes.FORWARD:
                    task.mode = wt.WalkModes.INPLACE
            if kind in ("sw200", "hook") and task.mode == wt.WalkModes.STANDING:
                task.mode = wt.WalkModes.INPLACE
            pre = dict(mode=MODES[task.mode], mode_ref=[float(x) for x in task.mode_ref], phase=int(task._phase))
            # the reference's draw order inside step(): randint(100) [sample_ref], randint(200) [sample_ref], randint(200) [3 x uniform]
            words = [u0[0]]
            mode = task.mode
            ph = (task._phase + 1) % period
            in_dbl = task.right_clock[0](ph) == 1 and task.left_clock[0](ph) == 1
            if f100 and in_dbl:
                mode = {wt.WalkModes.INPLACE: wt.WalkModes.STANDING, wt.WalkModes.STANDING: wt.WalkModes.INPLACE}.get(mode, mode)
                words += ref_words(mode, o.philox(seed, env_id, k, 1))
            words.append(u0[1])
            if f200 and mode != wt.WalkModes.STANDING:
                mode = {wt.WalkModes.FORWARD: wt.WalkModes.INPLACE, wt.WalkModes.INPLACE: wt.WalkModes.FORWARD}[mode]
                words += ref_words(mode, o.philox(seed, env_id, k, 2))
            words.append(u0[2])
            if fh and mode != wt.WalkModes.STANDING:
                words += o.philox(seed, env_id, k, 5)[:3]
            feed.words = words
            task.step()
            assert not feed.words, (kind, feed.words)
            # ---- a random robot state for calc_reward / done
            q = rng.normal(size=4)
            s = dict(lvel=rng.normal(size=3) * 0.3, rvel=rng.normal(size=3) * 0.3, lgrf=float(abs(rng.normal()) * 300),
                     rgrf=float(abs(rng.normal()) * 300), root=np.array([rng.normal() * 0.1, rng.normal() * 0.1, rng.uniform(0.55, 0.9)]),
                     root_vloc=rng.normal(size=3) * 0.4, qvel=rng.normal(size=18), qacc=rng.normal(size=18) * 3, torque=rng.normal(size=12) * 30,
                     pose=nominal + rng.normal(size=12) * 0.2, selfcol=bool(rng.randint(8) == 0),
                     con_r=[rng.normal(size=3) * 0.02 for _ in range(rng.randint(0, 3))], con_l=[rng.normal(size=3) * 0.02 for _ in range(rng.randint(0, 3))])
            s["head"] = s["root"] + np.array([rng.normal() * 0.05, rng.normal() * 0.05, 0.6])
            s["qpos"] = np.concatenate((s["root"], q / np.linalg.norm(q), s["pose"]))
            c.s = s
            prev_torque, prev_action, action = rng.normal(size=12) * 30, nominal + rng.normal(size=12) * 0.1, nominal + rng.normal(size=12) * 0.1
            r = task.calc_reward(prev_torque, prev_action, action)
            cz = min(p[2] for p in s["con_r"] + s["con_l"]) if (s["con_r"] or s["con_l"]) else 0.0
            rec["steps"].append(dict(kind=kind, ctr=k, pre=pre, mode=MODES[task.mode], mode_ref=[float(x) for x in task.mode_ref],
                                     phase=int(task._phase), hfield_pos=c.hfield.pos.tolist(),
                                     state={kk: (vv.tolist() if isinstance(vv, np.ndarray) else vv) for kk, vv in s.items() if kk not in ("con_r", "con_l")},
                                     ncon_r=len(s["con_r"]), ncon_l=len(s["con_l"]), contact_z_min=float(cz),
                                     prev_torque=prev_torque.tolist(), prev_action=prev_action.tolist(), action=action.tolist(),
                                     terms=[float(x) for x in r.values()], done=bool(task.done())))
            rec["names"] = list(r.keys())
        cases.append(rec)
    json.dump(cases, open(os.path.join(OUT, "walk_task.json"), "w"))
    kinds = [s["kind"] for c_ in cases for s in c_["steps"]]
    switched = sum(s["mode"] != s["pre"]["mode"] for c_ in cases for s in c_["steps"])
    print("wrote walk_task.json:", len(cases), "cases,", len(kinds), "steps, mode switches", switched,
          "hook moves", sum(any(s["hfield_pos"]) for c_ in cases for s in c_["steps"]))


if __name__ == "__main__":
    main()
