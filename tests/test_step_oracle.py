"""jvrc_step (BASELINE configs[2]): the oracle's SteppingTask restatement against vectors produced by
running the reference's tasks/stepping_task.py (tools/gen_golden_step.py), the stepping-stone contact model against
physical invariants, and the product's kernel source (CPU emulation, tests/emu) against the oracle."""
import json
import os

import numpy as np
import pytest

from emu import Emu
from learninghumanoidwalking_b200.model import load_model, pack_model
from oracle.oracle import Oracle

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["foot_frc_score", "foot_vel_score", "orient_cost", "height_error", "step_reward", "upper_body_reward"]


def quat2mat(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


@pytest.fixture(scope="module")
def orc():
    return Oracle("jvrc_step", tolerance=1e-14)


def test_stepping_task_matches_the_reference_class(orc):
    """SteppingTask.reset / step / calc_reward / done run from the reference's own file on recorded inputs."""
    o = orc
    cases = json.load(open(os.path.join(G, "step_task.json")))
    assert {c["mode"] for c in cases} == {0, 1, 2, 3, 4}
    n_adv = n_done = 0
    for c in cases:
        assert c["names"] == NAMES and c["period"] == 88 and c["delay_frames"] == 30
        assert c["box_size"] == [0.15, 1.0, 0.1]
        envs = o.make_envs(1, seed=c["seed"], first_id=c["env_id"])
        o.set_iteration_count(np.inf if c["iteration_count"] is None else c["iteration_count"])
        o.set_field(envs, 0, "rng_ctr", c["rng_ctr"])
        o.set_field(envs, 0, "root_xpos", c["root_xpos"])
        o.set_field(envs, 0, "root_xmat", c["root_xmat"])
        o.set_field(envs, 0, "foot_xpos", c["rfoot_xpos"] + c["lfoot_xpos"])
        o.task_reset(envs, 0)
        n = len(c["seq"])
        assert int(o.field(envs, 0, "mode")[0]) == c["mode"] and int(o.field(envs, 0, "phase")[0]) == c["phase"]
        assert int(o.field(envs, 0, "seq_len")[0]) == n
        assert (int(o.field(envs, 0, "t1")[0]), int(o.field(envs, 0, "t2")[0])) == (c["t1"], c["t2"])
        seq = o.field(envs, 0, "seq").reshape(20, 4)
        assert np.abs(seq[:n] - np.array(c["seq"])).max() < 1e-13
        # the boxes the reference re-poses (stepping_task.py:318-334): top face through the step, yawed with it; the rest parked
        bp, by = np.array(c["box_pos"]), np.array(c["box_yaw"])
        assert np.abs(seq[:, :2] - bp[:, :2]).max() < 1e-13 and np.abs(seq[:, 2] - (bp[:, 2] + 0.1)).max() < 1e-13
        assert np.abs(np.angle(np.exp(1j * (seq[:, 3] - by)))).max() < 1e-12
        assert (c["floor_z"] == -2.0) == (c["mode"] == 4)
        for s in c["steps"]:
            o.set_field(envs, 0, "root_xpos", s["root_xpos"])
            o.set_field(envs, 0, "root_xmat", quat2mat(s["root_quat"]).reshape(-1))
            o.set_field(envs, 0, "root_quat", s["root_quat"])
            o.set_field(envs, 0, "head_xpos", s["head_xpos"])
            o.set_field(envs, 0, "site_pos", s["rsite"] + s["lsite"])
            o.set_field(envs, 0, "rfoot_vel", s["rvel"]); o.set_field(envs, 0, "lfoot_vel", s["lvel"])
            o.set_field(envs, 0, "rfoot_grf", s["rgrf"]); o.set_field(envs, 0, "lfoot_grf", s["lgrf"])
            o.set_field(envs, 0, "ncon_r", s["ncon_r"]); o.set_field(envs, 0, "ncon_l", s["ncon_l"])
            o.set_field(envs, 0, "contact_z_min", s["contact_z_min"])
            o.task_step(envs, 0)
            assert int(o.field(envs, 0, "phase")[0]) == s["phase"]
            assert (int(o.field(envs, 0, "t1")[0]), int(o.field(envs, 0, "t2")[0])) == (s["t1"], s["t2"])
            assert bool(o.field(envs, 0, "target_reached")[0]) == s["target_reached"]
            assert int(o.field(envs, 0, "target_reached_frames")[0]) == s["frames"]
            assert np.abs(o.field(envs, 0, "goal_steps") - np.array(s["goal_steps"])).max() < 1e-12
            t = o.calc_reward(envs, 0, np.zeros(12))
            assert np.abs(t[:6] - np.array(s["terms"])).max() < 1e-13 and (t[6:] == 0).all()
            fz = min(s["rsite"][2], s["lsite"][2])
            assert (s["root_xpos"][2] - fz < 0.6) == s["done"]      # SteppingTask.done with no self collision
            n_done += s["done"]
        n_adv += c["steps"][-1]["t1"] > 0
    assert n_adv >= 6 and n_done >= 6


def _stand_on(o, envs, seq_rows, mode):
    """Nominal pose, zero velocity, given slab layout."""
    q = np.array(o.mj["cfg"]["nominal_qpos"])
    o.set_field(envs, 0, "qpos", q)
    o.set_field(envs, 0, "qvel", np.zeros(18))
    seq = np.tile([0.0, 0.0, -1.0, 0.0], (20, 1))
    if len(seq_rows):
        seq[:len(seq_rows)] = seq_rows
    o.set_field(envs, 0, "seq", seq.reshape(-1))
    o.set_field(envs, 0, "mode", mode)


def test_slab_contacts_equal_floor_contacts_and_stack_as_multiplicity(orc):
    """A slab whose top face is the plane z = 0 must act on a foot inside its footprint exactly like the floor; floor and
    k coplanar slabs together are k + 1 identical contacts per corner (MuJoCo would list them all), which stiffens the
    support but leaves the static force balance sum(GRF) = m g intact.  (Run with slab_contacts_are_floor=True so that the
    recorded GRF covers every contact; the reference's own floor-contact query is checked in the quirks test below.)"""
    import copy
    m2 = copy.deepcopy(load_model("jvrc_step"))
    m2["stepping"]["slab_contacts_are_floor"] = True
    o = Oracle("jvrc_step", tolerance=1e-14, model_dict=m2)
    mg = o.mj["total_mass"] * 9.81
    res = {}
    for name, rows, mode in (("floor", [], 1), ("slab", [[0.12, 0, 0, 0.05]], 4), ("floor+slab", [[0.12, 0, 0, 0.05]], 1),
                             ("floor+3", [[0.12, 0, 0, 0.05], [0.12, 0.2, 0, -0.05], [0.13, -0.1, 0, 0.0]], 1)):
        envs = o.make_envs(1)
        _stand_on(o, envs, rows, mode)
        kp, kd = np.array(o.mj["cfg"]["kp"]), np.array(o.mj["cfg"]["kd"])
        nom = np.array(o.mj["cfg"]["nominal_qpos"])[7:]
        for _ in range(600):   # PD-held nominal stance settles within ~0.5 s
            ctrl = kp * (nom - o.field(envs, 0, "qpos")[7:]) - kd * o.field(envs, 0, "qvel")[6:]
            o.mj_step(envs, 0, ctrl)
        res[name] = (o.field(envs, 0, "qpos").copy(), int(o.field(envs, 0, "ncon")[0]),
                     float(o.field(envs, 0, "rfoot_grf")[0] + o.field(envs, 0, "lfoot_grf")[0]))
    assert res["floor"][1] == res["slab"][1] and np.abs(res["floor"][0] - res["slab"][0]).max() < 1e-12
    assert res["floor+slab"][1] == 2 * res["floor"][1] and res["floor+3"][1] == 4 * res["floor"][1]
    z = [res[k][0][2] for k in ("floor", "floor+slab", "floor+3")]
    assert z[0] < z[1] < z[2] and z[2] - z[0] < 2e-3          # stiffer support, less static penetration
    for k in res:   # quasi-static (the PD-held stance sways slowly): sum of contact-force norms ~ m g in every layout
        assert 0.95 * mg < res[k][2] < 1.05 * mg, (k, res[k][2], mg)
    assert abs(res["floor+3"][2] - res["floor"][2]) < 0.01 * mg


def test_reference_quirks_of_the_stepping_env_are_reproduced_and_switchable(orc):
    """SURVEY Appendix C-2 / C-3.  (C-3) SteppingTask normalises foot forces with get_robot_mass() = mj_getTotalmass, which
    also sums the 20 static 800 kg boxes: 16062.4 kg, not 62.4.  (C-2) get_*_floor_contacts drops contacts whose geom1 is a
    robot geom, i.e. every foot-on-stone contact: the stones carry the robot but the task sees no GRF / contact height from
    them.  `slab_contacts_are_floor` switches the second one off; kernel source and oracle agree in both settings."""
    import copy
    mj = load_model("jvrc_step")
    assert mj["stepping"]["task_mass"] == pytest.approx(62.4 + 20 * 800.0) and mj["stepping"]["slab_contacts_are_floor"] is False
    mg = mj["total_mass"] * 9.81
    seen = {}
    for flag in (False, True):
        m2 = copy.deepcopy(mj)
        m2["stepping"]["slab_contacts_are_floor"] = flag
        o = Oracle("jvrc_step", tolerance=1e-14, model_dict=m2)
        envs = o.make_envs(1)
        kp, kd = np.array(mj["cfg"]["kp"]), np.array(mj["cfg"]["kd"])
        nom = np.array(mj["cfg"]["nominal_qpos"])[7:]
        for mode in (4, 1):   # FORWARD: stones only (floor at -2); STANDING: floor + one coplanar stone
            _stand_on(o, envs, [[0.12, 0, 0, 0.0]], mode)
            for _ in range(300):
                ctrl = kp * (nom - o.field(envs, 0, "qpos")[7:]) - kd * o.field(envs, 0, "qvel")[6:]
                o.mj_step(envs, 0, ctrl)
            seen[(flag, mode)] = (float(o.field(envs, 0, "rfoot_grf")[0] + o.field(envs, 0, "lfoot_grf")[0]),
                                  int(o.field(envs, 0, "ncon_r")[0] + o.field(envs, 0, "ncon_l")[0]), int(o.field(envs, 0, "ncon")[0]))
        # kernel source == oracle under this flag (closed loop, rewards include the GRF / contact-height terms)
        e = Emu(pack_model(m2, tolerance=1e-14), 64, 4, seed=3)
        envs = o.make_envs(4, seed=3)
        assert np.abs(o.batch_reset(envs, 4) - e.reset()).max() < 1e-12
        rng = np.random.RandomState(5)
        for _ in range(40):
            a = rng.normal(size=(4, 12)) * 0.2
            oo, _, tt, rr, dd, ee = o.batch_step(envs, 4, a, max_traj_len=30)
            eo, _, etm, er, ed, een, _, _ = e.step(a, max_traj_len=30)
            assert (dd == ed).all() and np.abs(tt - etm).max() < 1e-10 and np.abs(oo - eo).max() < 1e-9
    assert seen[(False, 4)][:2] == (0.0, 0) and seen[(False, 4)][2] == 8          # carried by 8 stone contacts, task sees none
    assert 0.9 * mg < seen[(True, 4)][0] < 1.1 * mg and seen[(True, 4)][1] == 8
    assert 0.45 * mg < seen[(False, 1)][0] < 0.55 * mg and seen[(False, 1)][1:] == (8, 16)   # the floor's half of the load
    assert 0.9 * mg < seen[(True, 1)][0] < 1.1 * mg and seen[(True, 1)][1] == 16


def test_foot_overhanging_a_slab_edge_is_held_by_edge_contacts(orc):
    """FORWARD mode (no floor): with the slab ending under the middle of the feet, the sole-edge x slab-boundary crossing
    vertices carry the load at the slab edge (without them the two inner corners alone would let the feet pitch over it)."""
    o = orc
    envs = o.make_envs(1)
    # feet span x in [0.022, 0.222] at the nominal pose; slab covers x <= 0.15
    _stand_on(o, envs, [[0.0, 0, 0, 0.0]], 4)
    o.mj_step(envs, 0, np.zeros(12))
    q0 = o.field(envs, 0, "qpos").copy()
    kp, kd = np.array(o.mj["cfg"]["kp"]), np.array(o.mj["cfg"]["kd"])
    nom = np.array(o.mj["cfg"]["nominal_qpos"])[7:]
    ncon = set()
    for _ in range(150):
        ctrl = kp * (nom - o.field(envs, 0, "qpos")[7:]) - kd * o.field(envs, 0, "qvel")[6:]
        o.mj_step(envs, 0, ctrl)
        ncon.add(int(o.field(envs, 0, "ncon")[0]))
    assert 8 in ncon                                  # 2 corners + 2 crossings per foot
    assert o.field(envs, 0, "qpos")[2] > 0.775         # rests at the floor-supported stance height (0.787), did not tip over the edge


def test_kernel_source_matches_oracle_on_stepping_stones():
    """csrc/sim_core.h (Cfg<6,1>) emulated on the CPU vs the oracle: closed loop, falls, truncations, auto-resets, all five
    walk modes, stairs (iteration_count = inf -> 0.1 m steps), contact counts up to 100+ in the oracle (one contact per
    surface) against the kernel's merged multiplicities."""
    o = Oracle("jvrc_step", tolerance=1e-14)
    N = 10
    e = Emu(pack_model(load_model("jvrc_step"), tolerance=1e-14), 64, N, seed=5, first_id=10)
    assert e.nobs == 39 and e.nr == 204
    envs = o.make_envs(N, seed=5, first_id=10)
    assert np.abs(o.batch_reset(envs, N) - e.reset()).max() < 1e-12
    rng = np.random.RandomState(2)
    n_end, modes, ncons = 0, set(), set()
    for _ in range(120):
        a = rng.normal(size=(N, 12)) * 0.25
        oo, to, tt, rr, dd, ee = o.batch_step(envs, N, a, max_traj_len=50)
        eo, et, etm, er, ed, een, eplen, eprew = e.step(a, max_traj_len=50)
        assert (dd == ed).all() and (ee == een).all()
        assert np.abs(oo - eo).max() < 1e-9 and np.abs(rr - er).max() < 1e-10 and np.abs(tt - etm).max() < 1e-10
        m = ee.astype(bool)
        if m.any():
            assert np.abs(to[m] - et[m]).max() < 1e-9
            n_end += int(m.sum())
        for i in range(N):
            modes.add(int(o.field(envs, i, "mode")[0]))
            ncons.add(int(o.field(envs, i, "ncon")[0]))
    assert n_end >= 15 and modes == {0, 1, 2, 3, 4} and max(ncons) > 40
    base = 19 + 18 + 18 + 5 * 12 + 3 + 1
    seq_o = np.stack([o.field(envs, i, "seq") for i in range(N)])
    assert np.abs(e.sr[:, base:base + 80] - seq_o).max() < 1e-12
    assert (e.si[:, 1] == [int(o.field(envs, i, "mode")[0]) for i in range(N)]).all()
    assert all(int(o.field(envs, i, "con_overflow")[0]) == 0 for i in range(N))


def test_height_curriculum_follows_iteration_count():
    """robot.iteration_count -> h = clip((it - 3000) / 8000, 0, 1) * 0.1 (stepping_task.py:312): flat below 3000."""
    mj = load_model("jvrc_step")
    for it, h in ((0, 0.0), (3000, 0.0), (7000, 0.05), (20000, 0.1), (np.inf, 0.1)):
        o = Oracle("jvrc_step", iteration_count=it)
        e = Emu(pack_model(mj, iteration_count=it), 64, 24, seed=11)
        envs = o.make_envs(24, seed=11)
        assert np.abs(o.batch_reset(envs, 24) - e.reset()).max() < 1e-12
        zs = np.stack([o.field(envs, i, "seq").reshape(20, 4)[:, 2] for i in range(24)])
        fwd = np.array([int(o.field(envs, i, "mode")[0]) == 4 for i in range(24)])
        assert fwd.any() and np.abs(np.abs(zs[fwd]).max(axis=1) - h * 15).max() < 1e-12 + h * 1.01   # 16 or 15 raised steps
        assert (zs[~fwd][zs[~fwd] > -1] == 0).all()


# ---------------------------------------------------------------- uneven / compliant terrain extension (BASELINE configs[4])
def test_terrain_extension_kernel_source_matches_oracle_and_reduces_to_flat_walk():
    """jvrc_walk_terrain = WalkingTask on 20 terraces re-posed with the ranges of the reference's (unused) manip_hfield hook
    (tasks/walking_task.py:172-179) + softer foot-ground contacts.  Not a parity target of the reference; checked:
    (i) kernel source == oracle in closed loop incl. re-poses, (ii) with the terraces sunk below the floor and the default
    contact solref the environment IS jvrc_walk (to roundoff: the slab code path sums the contact point differently), (iii) compliance: the softer contact sinks deeper at rest."""
    import copy
    from oracle import oracle as orc
    mj = load_model("jvrc_walk_terrain")
    assert mj["terrain"]["interval"] == 200 and mj["terrain"]["z_lo"] == -0.035 and mj["terrain"]["z_hi"] == -0.015
    o = Oracle("jvrc_walk_terrain", tolerance=1e-14)
    N = 8
    e = Emu(pack_model(mj, tolerance=1e-14), 64, N, seed=5, first_id=10)
    assert e.nobs == 37 and e.nr == 204
    envs = o.make_envs(N, seed=5, first_id=10)
    assert np.abs(o.batch_reset(envs, N) - e.reset()).max() < 1e-12
    rng = np.random.RandomState(2)
    n_end, reposed = 0, 0
    prev = np.stack([o.field(envs, i, "seq") for i in range(N)])
    assert (prev.reshape(N, 20, 4)[:, :, 2].max(axis=1) > 0).any()          # some terraces stick out of the floor
    for _ in range(220):
        a = rng.normal(size=(N, 12)) * 0.2
        oo, to, tt, rr, dd, ee = o.batch_step(envs, N, a, max_traj_len=80)
        eo, et, etm, er, ed, een, _, _ = e.step(a, max_traj_len=80)
        assert (dd == ed).all() and (ee == een).all()
        assert np.abs(oo - eo).max() < 1e-9 and np.abs(rr - er).max() < 1e-10 and np.abs(tt - etm).max() < 1e-10
        n_end += int(ee.sum())
        cur = np.stack([o.field(envs, i, "seq") for i in range(N)])
        reposed += int((np.abs(cur - prev).max(axis=1) > 0).sum())
        prev = cur
    assert n_end >= 10 and reposed >= 2
    assert np.abs(e.sr[:, 119:199] - prev).max() == 0

    # (ii) terraces out of reach + default solref == plain jvrc_walk
    flat = copy.deepcopy(mj)
    flat["terrain"].update(bump=0.0, z_lo=-0.3, z_hi=-0.3, contact_solref=[0.02, 1.0])
    ow = Oracle("jvrc_walk", tolerance=1e-14)
    ef = Emu(pack_model(flat, tolerance=1e-14), 64, 2, seed=9)
    ew = Emu(pack_model(load_model("jvrc_walk"), tolerance=1e-14), 64, 2, seed=9)
    assert np.array_equal(ef.reset(), ew.reset())
    for _ in range(30):
        a = rng.normal(size=(2, 12)) * 0.2
        rf, rw = ef.step(a), ew.step(a)
        assert np.abs(rf[0] - rw[0]).max() < 1e-10 and np.abs(rf[3] - rw[3]).max() < 1e-11 and (rf[5] == rw[5]).all()
    # (iii) compliance: PD-held stance on the floor, default vs soft contact
    z = {}
    for name, sr in (("default", [0.02, 1.0]), ("soft", [0.04, 1.0])):
        m2 = copy.deepcopy(flat)
        m2["terrain"]["contact_solref"] = sr
        em = Emu(pack_model(m2, tolerance=1e-14), 64, 1, seed=1)
        em.reset()
        for _ in range(12):
            em.step(np.zeros((1, 12)), autoreset=0)
        z[name] = em.qpos[0, 2]
    assert 1e-4 < z["default"] - z["soft"] < 5e-3


def test_slab_contact_set_against_an_independent_polygon_clip(orc):
    """The contact generation of the stepping stones (oracle: corners + Liang-Barsky edge crossings) against an independent
    formulation: Sutherland-Hodgman clipping of the sole quadrilateral by the slab's four half-planes (numpy, below).  The
    polygon's vertices with z below the top face must be exactly the oracle's slab contacts for that (foot, slab) pair."""
    o = orc
    mj = o.mj
    hx, hy, hz = mj["stepping"]["slab_half"]
    size, gpos = np.array(mj["geoms"][0]["size"]), np.array(mj["geoms"][0]["pos"])
    rng = np.random.RandomState(7)

    def clip_poly(poly, a, b, c):          # keep a*x + b*y <= c ; poly: list of (x, y, z)
        out = []
        for i in range(len(poly)):
            p, q = poly[i], poly[(i + 1) % len(poly)]
            fp, fq = a * p[0] + b * p[1] - c, a * q[0] + b * q[1] - c
            if fp <= 0:
                out.append(p)
            if (fp < 0 < fq) or (fq < 0 < fp):
                t = fp / (fp - fq)
                out.append(p + t * (q - p))
        return out

    from tools.compile_model import kinematics
    n_cross = n_corner = 0
    for trial in range(120):
        envs = o.make_envs(1)
        q = np.array(mj["cfg"]["nominal_qpos"])
        q[0:2] = rng.uniform(-0.1, 0.1, 2)
        q[2] = 0.806 + rng.uniform(-0.012, 0.003)      # sole between ~1.5 cm inside the slab and just above it
        ang = rng.normal(size=3) * np.array([0.03, 0.03, 0.5])
        q[3:7] = [np.cos(np.linalg.norm(ang) / 2), *(np.sin(np.linalg.norm(ang) / 2) * ang / np.linalg.norm(ang))]
        q[7:] += rng.uniform(-0.08, 0.08, 12)
        slab = np.array([rng.uniform(-0.1, 0.35), rng.uniform(-0.3, 0.3), rng.uniform(-0.004, 0.004), rng.uniform(-0.8, 0.8)])
        seq = np.tile([0.0, 0.0, -1.0, 0.0], (20, 1))
        seq[3] = slab
        o.set_field(envs, 0, "qpos", q); o.set_field(envs, 0, "seq", seq.reshape(-1)); o.set_field(envs, 0, "mode", 4)   # no floor
        pos, dist, foot, is_slab = o.contacts(envs, 0)
        assert is_slab.all()
        top = is_slab == 1          # this test is about the top-face manifold; riser contacts (3) have their own test below
        pos, dist, foot = pos[top], dist[top], foot[top]
        xpos, xmat = kinematics(mj, q)
        c, s = np.cos(slab[3]), np.sin(slab[3])
        for f, lk in enumerate((mj["rfoot_link"], mj["lfoot_link"])):
            R, p0 = np.array(xmat[lk]).reshape(3, 3), np.array(xpos[lk])
            # sole rectangle in the order the kernel walks it (corner ids 0, 1, 3, 2), world coordinates
            sole = [p0 + R @ (gpos + np.array([sx * size[0], sy * size[1], -size[2]])) for sx, sy in ((-1, -1), (1, -1), (1, 1), (-1, 1))]
            to_slab = lambda p: np.array([c * (p[0] - slab[0]) + s * (p[1] - slab[1]), -s * (p[0] - slab[0]) + c * (p[1] - slab[1]), p[2]])
            poly = [to_slab(p) for p in sole]
            for a_, b_, c_ in ((1, 0, hx), (-1, 0, hx), (0, 1, hy), (0, -1, hy)):
                poly = clip_poly(poly, a_, b_, c_)
            tol = mj["stepping"]["side_tol"]      # the model's side-face rule: deeper than tol only if as far from the sides
            keep = lambda v: -2 * hz < v[2] - slab[2] < 0 and (slab[2] - v[2] <= tol or slab[2] - v[2] <= min(hx - abs(v[0]), hy - abs(v[1])) + 1e-15)
            is_corner = lambda v: any(np.abs(v - to_slab(p)).max() < 1e-13 for p in sole)
            crossings = [v for v in poly if keep(v) and not is_corner(v)]
            n_riser = int(((is_slab == 3) & (o.contacts(envs, 0)[2] == f)).sum())
            if len(crossings) + n_riser > 4:
                continue                               # the per-foot cap of the extra slots is order dependent; not this test's subject
            # corners: mjc_PlaneBox's rule over all 8 box corners in index order (x sign = bit 0, y = bit 1, z = bit 2), 4 at
            # most: below the box centre, inside the footprint, supported by the top face (a tilted foot can offer a corner
            # of its TOP face as well, exactly as it would to the floor plane)
            ctr = p0 + R @ gpos
            corners = []
            for i in range(8):
                pt = p0 + R @ (gpos + np.array([(1 if i & 1 else -1) * size[0], (1 if i & 2 else -1) * size[1], (1 if i & 4 else -1) * size[2]]))
                v = to_slab(pt)
                if len(corners) < 4 and pt[2] - ctr[2] <= 0 and abs(v[0]) <= hx and abs(v[1]) <= hy and keep(v):
                    corners.append(v)
            exp = corners + crossings
            got = pos[foot == f]
            got_d = dist[foot == f]
            # the oracle caps crossings at 4 per foot and corners at 4: sizes here never exceed that (one slab)
            assert len(got) == len(exp), (trial, f, len(got), len(exp))
            for v in exp:
                w = np.array([slab[0] + c * v[0] - s * v[1], slab[1] + s * v[0] + c * v[1]])
                d = v[2] - slab[2]
                j = np.argmin(np.abs(got[:, 0] - w[0]) + np.abs(got[:, 1] - w[1]))
                assert np.abs(got[j, :2] - w).max() < 1e-12 and abs(got_d[j] - d) < 1e-12 and abs(got[j, 2] - (v[2] - 0.5 * d)) < 1e-12
            inside = [abs(to_slab(p)[0]) <= hx and abs(to_slab(p)[1]) <= hy for p in sole]
            n_corner += sum(inside)
            n_cross += len(crossings)
    assert n_cross > 40 and n_corner > 100      # both kinds of vertices were exercised


def test_newton_and_dual_pgs_agree_on_stepping_stone_contacts():
    """The stepping-stone rows (corner contacts on a slab, sole-edge crossings at its boundary, stacked coplanar supports) go
    through the oracle's two independent solvers — primal Newton with exact line search and dual projected Gauss-Seidel on
    J M^-1 J' + R — and give the same constrained acceleration: the rows are well-posed constraints, not solver artefacts."""
    newton = Oracle("jvrc_step", tolerance=1e-14, solver=0)
    pgs = Oracle("jvrc_step", tolerance=1e-14, solver=1, iterations=100)
    rng = np.random.RandomState(11)
    seen_rows = set()
    for layout, mode in (([[0.12, 0, 0, 0.05]], 4),                               # fully on one slab, no floor
                         ([[0.0, 0, 0, 0.0]], 4),                                 # slab ends under the feet: crossings
                         ([[0.12, 0, 0, 0.05], [0.12, 0.2, 0, -0.05]], 1)):       # floor + two coplanar slabs: multiplicity 3
        q = np.array(newton.mj["cfg"]["nominal_qpos"])
        q[2] = 0.797                         # every sole corner a few mm inside the supporting surface
        q[3:7] += rng.normal(size=4) * 0.003
        q[3:7] /= np.linalg.norm(q[3:7])
        q[7:] += rng.uniform(-0.01, 0.01, 12)
        v = rng.normal(size=18) * 0.2
        accs = []
        for o in (newton, pgs):
            envs = o.make_envs(1)
            o.set_field(envs, 0, "qpos", q); o.set_field(envs, 0, "qvel", v)
            seq = np.tile([0.0, 0.0, -1.0, 0.0], (20, 1)); seq[:len(layout)] = layout
            o.set_field(envs, 0, "seq", seq.reshape(-1)); o.set_field(envs, 0, "mode", mode)
            o.mj_step(envs, 0, np.zeros(12))
            n = int(o.field(envs, 0, "ncon")[0])
            assert 0 < n <= 34                       # <= 160 rows: inside the dual solver's table
            assert o.field(envs, 0, "last_kkt_residual")[0] < 1e-6
            accs.append(o.field(envs, 0, "qacc"))
            seen_rows.add(n)
        assert np.abs(accs[0] - accs[1]).max() < 1e-6 * max(1.0, np.abs(accs[0]).max())
    assert len(seen_rows) >= 3 and max(seen_rows) >= 20, seen_rows


def test_fp32_kernel_source_tracks_oracle_on_stepping_stones_for_a_short_horizon():
    o = Oracle("jvrc_step", tolerance=1e-14)
    N = 4
    e = Emu(pack_model(load_model("jvrc_step"), tolerance=1e-6), 32, N, seed=1)
    envs = o.make_envs(N, seed=1)
    assert np.abs(o.batch_reset(envs, N) - e.reset()).max() < 1e-5
    std = np.concatenate(([0.2, 0.2, 1, 1, 1], 0.5 * np.ones(12), 4 * np.ones(12), [1, 1], np.ones(8)))
    for _ in range(8):
        a = np.zeros((N, 12))
        oo, _, _, rr, dd, ee = o.batch_step(envs, N, a)
        eo, _, _, er, ed, een, _, _ = e.step(a)
        assert (ee == een).all()
        assert (np.abs(oo - eo) / std).max() < 5e-3 and np.abs(rr - er).max() < 5e-3


def test_a_foot_inside_a_riser_is_pushed_out():
    """Side faces of the stepping stones (tasks/stepping_task.py:318-334 poses real boxes; MuJoCo's box-box test gives riser
    contacts).  The robot stands on the floor with both toes 1 cm INSIDE the near face of a 0.1 m stair (10 cm below its top:
    the top face does not support them).  With `side_faces` the contact set holds riser contacts whose normal faces the robot
    and the toes are pushed back out to the soft-contact equilibrium; without it (round-1 behaviour) nothing touches the
    toes and they stay inside.  The push-out creates no energy beyond what the penetration stored."""
    import copy
    from oracle.oracle import Oracle, load_model_json
    from tools.compile_model import kinematics
    res = {}
    for side in (True, False):
        mj = copy.deepcopy(load_model_json("jvrc_step"))
        mj["stepping"]["side_faces"] = side
        mj["cfg"]["kp"] = [20 * k for k in mj["cfg"]["kp"]]      # quasi-rigid legs (see the static stance test)
        mj["cfg"]["kd"] = [5 * k for k in mj["cfg"]["kd"]]
        o = Oracle("jvrc_step", tolerance=1e-12, model_dict=mj)
        envs = o.make_envs(1)
        o.reset(envs)
        o.set_field(envs, 0, "mode", 1)                          # STANDING: the floor stays at z = 0
        size, gpos = np.array(mj["geoms"][0]["size"]), np.array(mj["geoms"][0]["pos"])

        def toe_x():
            xpos, xmat = kinematics(mj, o.field(envs, 0, "qpos"))
            return max((np.array(xpos[lk]) + np.array(xmat[lk]).reshape(3, 3) @ (gpos + np.array([size[0], sy * size[1], -size[2]])))[0]
                       for lk in (mj["rfoot_link"], mj["lfoot_link"]) for sy in (-1, 1))
        face = toe_x() - 0.01                                    # near (-x) face of the stair: 1 cm behind the toe corners
        seq = np.tile([0.0, 0.0, -1.0, 0.0], (20, 1))
        seq[0] = [face + mj["stepping"]["slab_half"][0], 0.0, 0.1, 0.0]
        o.set_field(envs, 0, "seq", seq.reshape(-1))
        o.set_field(envs, 0, "seq_len", 1)
        _, dist, foot, kind = o.contacts(envs, 0)
        n_riser0 = int((kind == 3).sum())
        ke_max = 0.0
        for _ in range(12):                                      # 0.3 s of PD-held stance (a = 0: target = nominal pose)
            o.step(envs, 0, np.zeros(12))
            ke_max = max(ke_max, o.energy(o.field(envs, 0, "qpos"), o.field(envs, 0, "qvel"))[0])
        res[side] = dict(inside=toe_x() - face, riser0=n_riser0, ke=ke_max)
    assert res[True]["riser0"] == 4 and res[False]["riser0"] == 0, res       # two toe corners per foot
    assert res[True]["inside"] < 0.002 < 0.006 < res[False]["inside"], res
    assert res[True]["ke"] < 5.0, res


def test_riser_contacts_kernel_source_matches_oracle():
    """The kernel source (CPU emulation) against the oracle with riser contacts ACTIVE: every env gets a 0.1 m stair whose
    near face sits 0.5 .. 2 cm behind its toe corners (yawed differently per env), then 12 closed-loop control steps."""
    from emu import Emu
    from learninghumanoidwalking_b200.model import load_model, pack_model
    from oracle.oracle import Oracle
    from tools.compile_model import kinematics
    mj = load_model("jvrc_step")
    o = Oracle("jvrc_step", tolerance=1e-14)
    n = 4
    e = Emu(pack_model(mj, tolerance=1e-14), 64, n, seed=11, first_id=2)
    envs = o.make_envs(n, seed=11, first_id=2)
    assert np.abs(e.reset() - o.batch_reset(envs, n)).max() < 1e-9
    size, gpos = np.array(mj["geoms"][0]["size"]), np.array(mj["geoms"][0]["pos"])
    hx = mj["stepping"]["slab_half"][0]
    for i in range(n):
        q = o.field(envs, i, "qpos")
        xpos, xmat = kinematics(mj, q)
        toe = max((np.array(xpos[lk]) + np.array(xmat[lk]).reshape(3, 3) @ (gpos + np.array([size[0], sy * size[1], -size[2]])))[0]
                  for lk in (mj["rfoot_link"], mj["lfoot_link"]) for sy in (-1, 1))
        yaw = 0.15 * (i - 1.5)
        face = toe - 0.005 * (i + 1)
        seq = np.tile([0.0, 0.0, -1.0, 0.0], (20, 1))
        seq[0] = [face + hx * np.cos(yaw), q[1] + hx * np.sin(yaw), 0.1, yaw]
        o.set_field(envs, i, "seq", seq.reshape(-1)); o.set_field(envs, i, "seq_len", 1)
        o.set_field(envs, i, "t1", 0); o.set_field(envs, i, "t2", 0); o.set_field(envs, i, "mode", 1)      # STANDING: floor at z = 0
        e.sr[i, 119:199] = seq.reshape(-1)
        e.sr[i, 199:204] = [1, 0, 0, 0, 0]
        e.si[i, 1] = 1
    rng = np.random.RandomState(4)
    n_riser = 0
    for k in range(12):
        for i in range(n):
            n_riser += int((o.contacts(envs, i)[3] == 3).sum())
        a = rng.normal(size=(n, 12)) * 0.1
        oo, _, _, orew, odone, oend = o.batch_step(envs, n, a, 400)
        r = e.step(a, max_traj_len=400)
        assert (r[4] == odone).all() and (r[5] == oend).all(), k
        assert np.abs(r[0] - oo).max() < 1e-8 and np.abs(r[3] - orew).max() < 1e-9, (k, np.abs(r[0] - oo).max())
        oq = np.stack([o.field(envs, i, "qpos") for i in range(n)])
        assert np.abs(e.sr[:, :19] - oq).max() < 1e-9
    assert n_riser >= 8, n_riser
