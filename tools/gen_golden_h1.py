#!/usr/bin/env python3
"""Generate tests/golden/h1_*.json by RUNNING the reference's own code for the H1 standing path in this container.

/root/reference does not exist on the GPU box, so the vectors are committed; this script is the provenance.
mujoco / transforms3d are not installable here, so the reference functions run against small stand-ins:

  * tasks/standing_task.py:StandingTask.calc_reward        — real class, RobotInterface replaced by a recorder
  * envs/common/domain_randomization.py (both functions)    — real module, mjModel/mjData replaced by name-indexed
                                                              containers with the attributes the functions touch
  * BaseHumanoidEnv._apply_init_noise / _apply_observation_noise (envs/common/base_humanoid_env.py:281-338) —
    the two function bodies are lifted out of the file by AST (the module itself imports mujoco) and executed
    with `tf3.euler.euler2quat` supplied by scipy ('sxyz' static == scipy extrinsic 'xyz')

The reference draws from numpy's global Mersenne twister; the oracle and the kernel draw from counter-based Philox
streams (oracle/sim_oracle.c).  To pin the draw ORDER, ranges and arithmetic, `np.random.uniform/randint` are
replaced during generation by functions that consume the oracle's Philox words in the order documented in
sim_oracle.c (randomize_dynamics: streams 16.., apply_perturbation: 32.., observation noise: 40.., init noise: 50..),
so the reference code, fed the same uniforms, must produce exactly what the oracle produces for that key.
"""
import ast
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


def load_by_path(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class WordFeed:
    """np.random stand-in fed with 32-bit words; same word -> value maps as oracle/sim_oracle.c (u01, randint)."""

    def __init__(self, words):
        self.words, self.i = list(words), 0

    def _u(self):
        w = self.words[self.i]
        self.i += 1
        return w

    def uniform(self, lo, hi, size=None):
        vec = size is not None or np.ndim(lo) > 0
        n = int(size) if size is not None else (len(lo) if np.ndim(lo) > 0 else 1)
        u = np.array([(self._u() >> 8) * (1.0 / 16777216.0) for _ in range(n)])
        out = lo + (hi - lo) * u
        return out if vec else float(out[0])

    def randint(self, n):
        return (self._u() * n) >> 32

    def skip(self, k):
        self.i += k


def main():
    from oracle.oracle import Oracle, load_model_json
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.RandomState(4321)
    o = Oracle("h1")
    mj = load_model_json("h1")
    cfg = mj["cfg"]

    # ------------------------------------------------------------------ StandingTask.calc_reward
    # tasks/__init__ imports every task (-> transforms3d); register a bare package so only standing_task loads
    pkg = types.ModuleType("tasks")
    pkg.__path__ = [os.path.join(REF, "tasks")]
    sys.modules["tasks"] = pkg
    st = load_by_path("tasks.standing_task", "tasks/standing_task.py")

    class Client:
        pass

    cases = []
    for _ in range(32):
        # a random pelvis pose, welded torso (same frame: envs/h1/h1_env.py removes the waist joint), joint state
        from scipy.spatial.transform import Rotation
        R = Rotation.from_rotvec(rng.normal(size=3) * 0.2).as_matrix()
        pos = np.array([rng.normal() * 0.1, rng.normal() * 0.1, rng.uniform(0.85, 1.05)])
        T = np.eye(4)
        T[:3, :3], T[:3, 3] = R, pos
        q = np.asarray(cfg["half_sitting_pose"]) + rng.normal(size=10) * 0.2
        tau = rng.normal(size=10) * 40
        vloc = rng.normal(size=3) * 0.5        # get_body_vel(frame=1)[0]: linear velocity in the pelvis frame
        qvel = rng.normal(size=16)
        c = Client()
        c.get_object_affine_by_name = lambda name, typ, T=T: T.copy()
        c.get_act_joint_positions = lambda q=q: list(q)
        c.get_act_joint_torques = lambda tau=tau: tau.copy()
        c.get_body_vel = lambda name, frame=0, vloc=vloc: [vloc.copy(), np.zeros(3)]
        c.get_qvel = lambda qvel=qvel: qvel.copy()
        task = st.StandingTask(c, np.asarray(cfg["half_sitting_pose"]))
        r = task.calc_reward(None, None, None)
        cases.append(dict(root_xmat=R.reshape(-1).tolist(), root_xpos=pos.tolist(), root_vlin_world=(R @ vloc).tolist(),
                          act_len=q.tolist(), act_force=tau.tolist(), qvel=qvel.tolist(),
                          names=list(r.keys()), terms=[float(v) for v in r.values()]))
    json.dump(cases, open(os.path.join(OUT, "h1_standing_reward.json"), "w"))

    # ------------------------------------------------------------------ randomize_dynamics / apply_perturbation
    dr = load_by_path("ref_domain_randomization", "envs/common/domain_randomization.py")
    links = mj["links"]
    joint_names = [lk["joint"]["name"] for lk in links[1:]]
    rp = mj["root_parts"]

    class Body:
        def __init__(self, name, mass, ipos):
            self.name, self.mass, self._ipos = name, np.array([mass]), np.array(ipos, dtype=float)

        @property
        def ipos(self):
            return self._ipos

        @ipos.setter
        def ipos(self, v):
            self._ipos = np.array(v, dtype=float)

    class Model:
        def __init__(self):
            self.bodies = {"pelvis": Body("pelvis", rp["pelvis"]["mass"], rp["pelvis"]["com"])}
            for lk in links[1:]:
                self.bodies[lk["name"]] = Body(lk["name"], lk["mass"], lk["com"])
            self.dof_frictionloss = np.zeros(16)
            self.dof_damping = np.array([0.0] * 6 + [lk["joint"]["damping"] for lk in links[1:]])

        def body(self, key):
            return self.bodies[key]

        def joint(self, name):
            return types.SimpleNamespace(bodyid=links[1 + joint_names.index(name)]["name"])

    interface = types.SimpleNamespace(get_jnt_qveladr_by_name=lambda jn: 6 + joint_names.index(jn))
    real_uniform, real_randint = np.random.uniform, np.random.randint
    dyn, pert = [], []
    try:
        for case in range(8):
            seed, env_id, ctr = int(rng.randint(1 << 30)), int(rng.randint(1 << 20)), int(rng.randint(1, 1 << 20))
            words = []
            for s in range(16, 21):
                words += o.philox(seed, env_id, ctr, s)
            for s in range(21, 32):
                words += o.philox(seed, env_id, ctr, s)
            feed = WordFeed(words)
            np.random.uniform, np.random.randint = feed.uniform, feed.randint
            model, default = Model(), Model()
            dr.randomize_dynamics(model, default, interface, joint_names, None)
            assert feed.i == 20 + 44
            dyn.append(dict(seed=seed, env_id=env_id, ctr=ctr,
                            frictionloss=model.dof_frictionloss[6:].tolist(), damping=model.dof_damping[6:].tolist(),
                            pelvis_mass=float(model.body("pelvis").mass[0]), pelvis_ipos=model.body("pelvis").ipos.tolist(),
                            link_mass=[float(model.body(lk["name"]).mass[0]) for lk in links[1:]],
                            link_ipos=[model.body(lk["name"]).ipos.tolist() for lk in links[1:]]))

            # apply_perturbation: force(3), torque(3), coin per body; oracle streams 32+2b (force + coin in lane 3), 33+2b
            words = []
            for b in range(2):
                f, t = o.philox(seed, env_id, ctr, 32 + 2 * b), o.philox(seed, env_id, ctr, 33 + 2 * b)
                words += f[:3] + t[:3] + [f[3]]
            feed = WordFeed(words)
            np.random.uniform, np.random.randint = feed.uniform, feed.randint

            class Data:
                def __init__(self):
                    self.xfrc_applied = np.zeros((3, 6))   # world, pelvis, torso_link

                def body(self, name):
                    return types.SimpleNamespace(xfrc_applied=self.xfrc_applied[{"pelvis": 1, "torso_link": 2}[name]])

            data = Data()
            pcfg = types.SimpleNamespace(force_magnitude=cfg["perturbation"]["force_magnitude"],
                                         torque_magnitude=cfg["perturbation"]["torque_magnitude"],
                                         bodies=cfg["perturbation"]["bodies"])
            dr.apply_perturbation(data, pcfg)
            pert.append(dict(seed=seed, env_id=env_id, ctr=ctr, xfrc=data.xfrc_applied[1:].reshape(-1).tolist()))
    finally:
        np.random.uniform, np.random.randint = real_uniform, real_randint
    json.dump(dict(randomize_dynamics=dyn, apply_perturbation=pert), open(os.path.join(OUT, "h1_domain_randomization.json"), "w"))

    # ------------------------------------------------------------------ init noise / observation noise (AST-lifted)
    src = open(os.path.join(REF, "envs/common/base_humanoid_env.py")).read()
    tree = ast.parse(src)
    wanted = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in ("_apply_init_noise", "_apply_observation_noise"):
            wanted[node.name] = ast.Module(body=[node], type_ignores=[])
    from scipy.spatial.transform import Rotation

    def euler2quat(ai, aj, ak):
        x, y, z, w = Rotation.from_euler("xyz", [ai, aj, ak]).as_quat()
        return np.array([w, x, y, z])

    tf3 = types.SimpleNamespace(euler=types.SimpleNamespace(euler2quat=euler2quat))
    ns = {"np": np, "tf3": tf3}
    for name, mod in wanted.items():
        exec(compile(mod, "<reference:base_humanoid_env.py:%s>" % name, "exec"), ns)
    scales = cfg["observation_noise"]["scales"]
    self_ = types.SimpleNamespace(
        cfg=types.SimpleNamespace(init_noise=cfg["init_noise_deg"],
                                  observation_noise=types.SimpleNamespace(enabled=True, type="uniform", multiplier=1.0,
                                                                          scales=types.SimpleNamespace(**scales))),
        _get_joint_names=lambda: joint_names,
        interface=types.SimpleNamespace(get_jnt_qposadr_by_name=lambda n: [0]))
    init, obsn = [], []
    try:
        for case in range(8):
            seed, env_id, ctr = int(rng.randint(1 << 30)), int(rng.randint(1 << 20)), int(rng.randint(1, 1 << 20))
            w50 = o.philox(seed, env_id, ctr, 50)
            words = w50[:3]
            for s in range(51, 54):
                words += o.philox(seed, env_id, ctr, s)
            feed = WordFeed(words)
            np.random.uniform, np.random.randint = feed.uniform, feed.randint
            q = ns["_apply_init_noise"](self_, list(cfg["nominal_qpos"]))
            init.append(dict(seed=seed, env_id=env_id, ctr=ctr, qpos=[float(v) for v in q]))

            words = []
            for s in range(40, 49):
                words += o.philox(seed, env_id, ctr, s)
            feed = WordFeed(words)
            np.random.uniform, np.random.randint = feed.uniform, feed.randint
            clean = dict(root_orient=rng.normal(size=2) * 0.1, root_ang_vel=rng.normal(size=3),
                         motor_pos=rng.normal(size=10), motor_vel=rng.normal(size=10), motor_tau=rng.normal(size=10) * 30)
            noisy = ns["_apply_observation_noise"](self_, dict(clean))
            order = ("root_orient", "root_ang_vel", "motor_pos", "motor_vel", "motor_tau")
            obsn.append(dict(seed=seed, env_id=env_id, ctr=ctr, clean=np.concatenate([clean[k] for k in order]).tolist(),
                             noisy=np.concatenate([noisy[k] for k in order]).tolist()))
    finally:
        np.random.uniform, np.random.randint = real_uniform, real_randint
    json.dump(dict(init_noise=init, observation_noise=obsn), open(os.path.join(OUT, "h1_noise.json"), "w"))

    # ------------------------------------------------------------------ RobotBase PD-gain randomisation (pdrand_k)
    rb = load_by_path("ref_robot_base", "robots/robot_base.py")

    class PDClient:
        def __init__(self):
            self.gains = []

        def nu(self): return 10
        def sim_dt(self): return 0.001
        def set_pd_gains(self, kp, kd): self.gains.append((np.array(kp, dtype=float), np.array(kd, dtype=float)))
        def step_pd(self, p, v): return np.zeros(10)
        def get_act_joint_velocities(self): return np.zeros(10)
        def get_gear_ratios(self): return np.ones(10)
        def set_motor_torque(self, tau, motor_dyn=False): pass
        def step(self): pass

    pd = []
    try:
        for case in range(6):
            seed, env_id, ctr = int(rng.randint(1 << 30)), int(rng.randint(1 << 20)), int(rng.randint(1, 1 << 20))
            k = float(rng.choice([0.1, 0.2, 0.3]))
            words = []
            for s_ in (8, 9, 10):
                words += o.philox(seed, env_id, ctr, s_)
            words = words[:10]
            w2 = []
            for s_ in (11, 12, 13):
                w2 += o.philox(seed, env_id, ctr, s_)
            words += w2[:10]
            client = PDClient()
            robot = rb.RobotBase(np.array([cfg["kp"], cfg["kd"]]), cfg["control_dt"], client, None, pdrand_k=k)
            feed = WordFeed(words)
            np.random.uniform, np.random.randint = feed.uniform, feed.randint
            robot._do_simulation(np.zeros(10), 2)
            kp, kd = client.gains[-1]
            pd.append(dict(seed=seed, env_id=env_id, ctr=ctr, k=k, kp=kp.tolist(), kd=kd.tolist()))
    finally:
        np.random.uniform, np.random.randint = real_uniform, real_randint
    json.dump(pd, open(os.path.join(OUT, "h1_pd_gain_randomization.json"), "w"))
    print("H1 golden vectors written to", os.path.abspath(OUT))


if __name__ == "__main__":
    main()
