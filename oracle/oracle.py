"""ctypes wrapper around oracle/sim_oracle.c — TEST INFRASTRUCTURE, never imported by the product.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs use it.
The model constants come from the committed compiled model (learninghumanoidwalking_b200/model/*.json,
produced by tools/compile_model.py) and the gait clocks from tests/golden/gait_clocks.json (produced by
running the reference's tasks/rewards.py:create_phase_reward — i.e. pinned to the reference itself).
"""
from __future__ import annotations

import ctypes
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "_build", "liboracle.so")
NOBS, NREW, NU, NV, NQ = 37, 10, 12, 18, 19   # jvrc sizes == the C struct capacities (ORC_NU, ORC_NV, ORC_NQ)
MAXLINK = 16


def build(force: bool = False) -> str:
    src = [os.path.join(HERE, "sim_oracle.c"), os.path.join(HERE, "sim_oracle.h")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src):
        os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-std=gnu11", "-o", LIB_PATH, src[0], "-lm"])
    return LIB_PATH


def load_model_json(name: str = "jvrc_walk") -> dict:
    return json.load(open(os.path.join(ROOT, "learninghumanoidwalking_b200", "model", name + ".json")))


def load_clocks() -> dict:
    return json.load(open(os.path.join(ROOT, "tests", "golden", "gait_clocks.json")))


def pack_model(mj: dict, clocks: dict, tolerance: float | None = None, solver: int = 0, iterations: int | None = None,
               self_collision: bool = True, pdrand_k: float = 0.0, iteration_count: float = float("inf")):
    """Flat double layout consumed by orc_model_from_flat (keep in sync with sim_oracle.c)."""
    b: list[float] = []
    links = mj["links"]
    b.append(len(links))
    for i, lk in enumerate(links):
        b.append(lk["parent"])
        b += lk["pos"]
        b += list(np.array(lk["rot"]).reshape(-1))
        b += lk["joint"].get("axis", [0.0, 0.0, 0.0])
        b.append(lk["mass"])
        b += lk["com"]
        a = lk["inertia"]
        b += [a[0], a[3], a[4], a[3], a[1], a[5], a[4], a[5], a[2]]
        b += mj["link_invweight0"][i]
    nv = 6 + len(links) - 1
    for d in range(nv):
        if d < 6:
            b += [0.0, 0.0, 0.0, 0.0, 0.0, mj["dof_invweight0"][d]]
        else:
            j = links[d - 5]["joint"]
            b += [j["armature"], j["damping"], j["range"][0], j["range"][1], 1.0 if j.get("limited", True) else 0.0,
                  mj["dof_invweight0"][d]]
    b.append(len(mj["geoms"]))
    for g in mj["geoms"]:
        b.append(g["link"])
        b += g.get("pos", [0.0, 0.0, 0.0])
        b += g.get("size", [0.0, 0.0, 0.0])
    o = mj["opt"]
    b.append(o["timestep"])
    b += o["gravity"]
    b += o["solref"]
    b += o["solimp"]
    b += [o["friction"][0], o["impratio"], mj["meaninertia"], o["tolerance"] if tolerance is None else tolerance,
          o["iterations"] if iterations is None else iterations, solver]
    c = mj["cfg"]
    b += c["kp"]
    b += c["kd"]
    b += c["nominal_qpos"]
    b += [c["frame_skip"], c["action_smoothing"], mj["rfoot_link"], mj["lfoot_link"]]
    stand = mj["name"] == "h1"
    b += mj.get("head_in_root", [0.0, 0.0, 0.0])
    # the mass the task normalises ground reaction forces with: RobotInterface.get_robot_mass() = mj_getTotalmass, which in
    # jvrc_step also counts the 20 static boxes
    task_mass = mj.get("stepping", {}).get("task_mass", mj["total_mass"])
    b += [task_mass, c.get("task", {}).get("goal_height", 0.98), 0 if stand else clocks["period"]]
    if not stand:
        for k in ("r_frc", "r_vel", "l_frc", "l_vel"):
            b += clocks[k]
    sc = mj.get("self_collision") if self_collision else None
    caps = sc["capsules"] if sc else []
    b.append(len(caps))
    for cap in caps:
        b += [cap["link"]] + cap["p0"] + cap["p1"] + [cap["radius"]]
    pairs = sc["pairs"] if sc else []
    b.append(len(pairs))
    for a_, b_ in pairs:
        b += [a_, b_]
    # task / robot variant tail (envs/h1/configs/base.yaml, tasks/standing_task.py)
    if stand:
        ns = c["observation_noise"]
        sc = ns["scales"]
        lvl = ns["multiplier"] if ns["enabled"] else 0.0
        pert, dyn = c["perturbation"], c["dynamics_randomization"]
        b += [1, 35, 0.9, 1.4]
        b += [lvl * sc[k] for k in ("root_orient", "root_ang_vel", "motor_pos", "motor_vel", "motor_tau")]
        b += [int(dyn["interval"] / c["control_dt"]) if dyn["enable"] else 0,
              int(pert["interval"] / c["control_dt"]) if pert["enable"] else 0,
              pert["force_magnitude"], pert["torque_magnitude"], c["init_noise_deg"]]
    elif mj["name"] == "jvrc_step":
        b += [2, 39, 0.6, 1e30] + [0.0] * 5 + [0, 0, 0.0, 0.0, 0.0]     # tasks/stepping_task.py:256 (no upper height bound)
    else:
        b += [0, 37, 0.6, 1.4] + [0.0] * 5 + [0, 0, 0.0, 0.0, 0.0]
    for g in mj["geoms"]:
        pts = g.get("points", [])
        b += [1 if g.get("type") == "spheres" else 0, len(pts), g.get("radius", 0.0)]
        for pt in pts:
            b += pt
    rp = mj.get("root_parts")
    if rp:
        b += [rp["pelvis"]["mass"]] + rp["pelvis"]["com"] + list(np.array(rp["pelvis"]["Ic"]).reshape(-1))
        b += [rp["rest"]["mass"]] + rp["rest"]["mc"] + list(np.array(rp["rest"]["Io"]).reshape(-1))
        b += rp["torso_com"]
    else:
        b += [0.0] * 29
    b.append(float(pdrand_k))
    tr = mj.get("terrain")
    b.append(1 if tr else 0)
    if tr:
        b += tr["strip_half"] + [tr["side_tol"], tr["pitch"], tr["bump"], tr["z_lo"], tr["z_hi"], tr["xy"], tr["interval"]]
        b += tr["contact_solref"]
        b.append(1 if tr.get("side_faces", True) else 0)
    if mj["name"] == "jvrc_step":
        st = mj["stepping"]
        for site in mj["foot_sites"]:
            b += site
        b += st["slab_half"]
        b += [st["target_radius"], st["side_tol"], st["delay_frames"], curriculum_height(iteration_count),
              1 if st.get("slab_contacts_are_floor") else 0, 1 if st.get("side_faces", True) else 0]
        b.append(len(st["plans"]))
        for plan in st["plans"]:
            b.append(len(plan))
            for row in plan:
                b += row
    # assumption switches (model JSON "assumptions", SURVEY.md Appendix A warnings): bit 0 = explicit Euler
    b.append(float(assumption_flags(mj)))
    return np.array(b, dtype=np.float64)


def assumption_flags(mj: dict) -> int:
    a = mj.get("assumptions", {})
    return 0 if a.get("implicit_damping", True) else 1


def curriculum_height(iteration_count: float) -> float:
    """tasks/stepping_task.py:312: h = clip((iteration_count - 3000) / 8000, 0, 1) * 0.1 (iteration_count = inf by default,
    robots/robot_base.py:35)."""
    return float(np.clip((iteration_count - 3000) / 8000, 0, 1) * 0.1)


class Oracle:
    """One compiled model + helpers to own N environments."""

    def __init__(self, name: str = "jvrc_walk", tolerance: float | None = None, solver: int = 0,
                 iterations: int | None = None, pdrand_k: float = 0.0, iteration_count: float = float("inf"),
                 model_dict: dict | None = None):
        self.lib = ctypes.CDLL(build())
        L = self.lib
        L.orc_energy.restype = ctypes.c_double
        self.mj = model_dict if model_dict is not None else load_model_json(name)
        self.clocks = load_clocks()
        flat = pack_model(self.mj, self.clocks, tolerance, solver, iterations, pdrand_k=pdrand_k,
                          iteration_count=iteration_count)
        self._model = ctypes.create_string_buffer(L.orc_sizeof_model())
        rc = L.orc_model_from_flat(self._model, flat.ctypes.data_as(ctypes.c_void_p), len(flat))
        if rc != 0:
            raise RuntimeError(f"orc_model_from_flat failed: {rc}")
        self.env_size = L.orc_sizeof_env()
        assert self.env_size == _ENV_SIZE, (self.env_size, _ENV_SIZE)
        self.nu = len(self.mj["links"]) - 1
        self.nv, self.nq = 6 + self.nu, 7 + self.nu
        self.nobs = {"h1": 35, "jvrc_step": 39}.get(name, NOBS)

    def set_iteration_count(self, iteration_count: float):
        """env.robot.iteration_count = itr (rl/workers/rollout_worker.py:95) -> the curriculum's step height."""
        self.lib.orc_set_step_height(self._model, ctypes.c_double(curriculum_height(iteration_count)))

    def contacts(self, envs, i=0):
        """(pos [n,3], dist [n], foot [n], slab [n]) of the contacts at env i's current qpos; slab: 0 floor plane, 1 top face of a
        stepping stone, 3 side face of one (riser contact)."""
        pos, dist = np.zeros((128, 3)), np.zeros(128)
        foot, slab = np.zeros(128, dtype=np.int32), np.zeros(128, dtype=np.int32)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        n = self.lib.orc_test_contacts(self._model, self.env_ptr(envs, i), p(pos), p(dist), p(foot), p(slab))
        return pos[:n], dist[:n], foot[:n], slab[:n]

    def task_reset(self, envs, i=0):
        self.lib.orc_test_task_reset(self._model, self.env_ptr(envs, i))

    def task_step(self, envs, i=0):
        self.lib.orc_test_task_step(self._model, self.env_ptr(envs, i))

    # ---- single/batched env management
    def make_envs(self, n: int, seed: int = 0, first_id: int = 0):
        buf = ctypes.create_string_buffer(self.env_size * n)
        for i in range(n):
            self.lib.orc_env_init(self._model, ctypes.byref(buf, i * self.env_size), ctypes.c_uint32(seed),
                                  ctypes.c_uint32(first_id + i))
        return buf

    def env_ptr(self, envs, i=0):
        return ctypes.byref(envs, i * self.env_size)

    def field(self, envs, i, name):
        """Read a field of env i by name (offsets mirror orc_env)."""
        off, cnt, typ = _ENV_FIELDS[name]
        raw = np.frombuffer(envs, dtype=np.uint8, count=self.env_size, offset=i * self.env_size)
        return raw[off:off + cnt * np.dtype(typ).itemsize].view(typ).copy()

    def set_field(self, envs, i, name, value):
        off, cnt, typ = _ENV_FIELDS[name]
        raw = np.frombuffer(envs, dtype=np.uint8, count=self.env_size, offset=i * self.env_size)
        raw[off:off + cnt * np.dtype(typ).itemsize] = np.asarray(value, dtype=typ).reshape(cnt).view(np.uint8)

    def reset(self, envs, i=0):
        obs = np.zeros(self.nobs)
        self.lib.orc_reset(self._model, self.env_ptr(envs, i), obs.ctypes.data_as(ctypes.c_void_p))
        return obs

    def mj_step(self, envs, i, ctrl):
        ctrl = np.ascontiguousarray(ctrl, dtype=np.float64)
        self.lib.orc_mj_step(self._model, self.env_ptr(envs, i), ctrl.ctypes.data_as(ctypes.c_void_p))

    def step(self, envs, i, action):
        action = np.ascontiguousarray(action, dtype=np.float64)
        obs, terms = np.zeros(self.nobs), np.zeros(NREW)
        rew, done = ctypes.c_double(), ctypes.c_int()
        self.lib.orc_step(self._model, self.env_ptr(envs, i), action.ctypes.data_as(ctypes.c_void_p),
                          obs.ctypes.data_as(ctypes.c_void_p), terms.ctypes.data_as(ctypes.c_void_p),
                          ctypes.byref(rew), ctypes.byref(done))
        return obs, rew.value, bool(done.value), terms

    def batch_reset(self, envs, n, nthreads=0):
        obs = np.zeros((n, self.nobs))
        self.lib.orc_batch_reset(self._model, envs, n, obs.ctypes.data_as(ctypes.c_void_p), nthreads)
        return obs

    def batch_step(self, envs, n, actions, max_traj_len=400, nthreads=0):
        actions = np.ascontiguousarray(actions, dtype=np.float64)
        assert actions.shape == (n, self.nu)
        obs, tobs, terms = np.zeros((n, self.nobs)), np.zeros((n, self.nobs)), np.zeros((n, NREW))
        rew = np.zeros(n)
        done, ended = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        self.lib.orc_batch_step_autoreset(self._model, envs, n, p(actions), max_traj_len, p(obs), p(tobs), p(terms),
                                          p(rew), p(done), p(ended), nthreads)
        return obs, tobs, terms, rew, done, ended

    # ---- building blocks
    def mass_matrix(self, qpos):
        qpos = np.ascontiguousarray(qpos, dtype=np.float64)
        M = np.zeros((self.nv, self.nv))
        self.lib.orc_mass_matrix(self._model, qpos.ctypes.data_as(ctypes.c_void_p), M.ctypes.data_as(ctypes.c_void_p))
        return M

    def bias(self, qpos, qvel):
        qpos = np.ascontiguousarray(qpos, dtype=np.float64)
        qvel = np.ascontiguousarray(qvel, dtype=np.float64)
        c = np.zeros(self.nv)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        self.lib.orc_bias(self._model, p(qpos), p(qvel), p(c))
        return c

    def energy(self, qpos, qvel):
        qpos = np.ascontiguousarray(qpos, dtype=np.float64)
        qvel = np.ascontiguousarray(qvel, dtype=np.float64)
        ke, pe = ctypes.c_double(), ctypes.c_double()
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        self.lib.orc_energy(self._model, p(qpos), p(qvel), ctypes.byref(ke), ctypes.byref(pe))
        return ke.value, pe.value

    def philox(self, seed, env_id, ctr, stream):
        out = (ctypes.c_uint32 * 4)()
        self.lib.orc_philox(ctypes.c_uint32(seed), ctypes.c_uint32(env_id), ctypes.c_uint32(ctr),
                            ctypes.c_uint32(stream), out)
        return list(out)

    def quat2rp(self, quat):
        quat = np.ascontiguousarray(quat, dtype=np.float64)
        r, p = ctypes.c_double(), ctypes.c_double()
        self.lib.orc_quat2rp(quat.ctypes.data_as(ctypes.c_void_p), ctypes.byref(r), ctypes.byref(p))
        return r.value, p.value

    def calc_reward(self, envs, i, target):
        target = np.ascontiguousarray(target, dtype=np.float64)
        t = np.zeros(NREW)
        self.lib.orc_calc_reward(self._model, self.env_ptr(envs, i), target.ctypes.data_as(ctypes.c_void_p),
                                 t.ctypes.data_as(ctypes.c_void_p))
        return t

    def max_threads(self):
        return self.lib.orc_max_threads()


def _layout():
    """Offsets of orc_env fields (mirrors the struct in sim_oracle.h; doubles first then ints, natural alignment)."""
    fields = [
        ("qpos", NQ, "f8"), ("qvel", NV, "f8"), ("qacc_warm", NV, "f8"), ("qacc", NV, "f8"),
        ("act_len", NU, "f8"), ("act_vel", NU, "f8"), ("act_force", NU, "f8"),
        ("root_xpos", 3, "f8"), ("root_xmat", 9, "f8"), ("head_xpos", 3, "f8"), ("root_vlin", 3, "f8"),
        ("rfoot_vel", 3, "f8"), ("lfoot_vel", 3, "f8"), ("rfoot_grf", 1, "f8"), ("lfoot_grf", 1, "f8"),
        ("contact_z_min", 1, "f8"),
        ("ncon_r", 1, "i4"), ("ncon_l", 1, "i4"), ("ncon", 1, "i4"), ("self_collision", 1, "i4"),
        ("prev_prediction", NU, "f8"), ("prev_action", NU, "f8"), ("prev_torque", NU, "f8"),
        ("have_prev", 1, "i4"), ("phase", 1, "i4"), ("mode", 1, "i4"), ("_pad0", 1, "i4"),
        ("mode_ref", 3, "f8"),
        ("traj_len", 1, "i4"), ("ep_len", 1, "i4"), ("ep_rew", 1, "f8"),
        ("seed", 1, "u4"), ("env_id", 1, "u4"), ("rng_ctr", 1, "u4"), ("last_solver_iter", 1, "i4"),
        ("last_kkt_residual", 1, "f8"), ("status", 1, "i4"), ("nsubsteps", 1, "i4"),
        # orc_params P, then xfrc[2][6]
        ("P_mass", MAXLINK, "f8"), ("P_com", MAXLINK * 3, "f8"), ("P_inertia", MAXLINK * 9, "f8"),
        ("P_damping", NV, "f8"), ("P_frictionloss", NV, "f8"), ("P_pel_mass", 1, "f8"), ("P_pel_com", 3, "f8"),
        ("xfrc", 12, "f8"),
        # SteppingTask state
        ("seq", 80, "f8"), ("goal_steps", 8, "f8"), ("site_pos", 6, "f8"), ("foot_xpos", 6, "f8"), ("root_quat", 4, "f8"),
        ("seq_len", 1, "i4"), ("t1", 1, "i4"), ("t2", 1, "i4"), ("target_reached", 1, "i4"),
        ("target_reached_frames", 1, "i4"), ("con_overflow", 1, "i4"),
        ("iter_trace", 32, "i4"), ("nrow_trace", 32, "i4"),
    ]
    out, off = {}, 0
    for name, cnt, typ in fields:
        out[name] = (off, cnt, np.dtype(typ))
        off += cnt * np.dtype(typ).itemsize
    return out, off


_ENV_FIELDS, _ENV_SIZE = _layout()
