"""H1 standing environment (BASELINE config "h1 standing task, domain-randomised"): pin the oracle's env-level pieces to
vectors produced by the reference's own code (tools/gen_golden_h1.py ran tasks/standing_task.py,
envs/common/domain_randomization.py and the two noise functions of envs/common/base_humanoid_env.py), anchor the new
physics (friction-loss rows, sphere-foot contacts, external wrenches, per-env inertial parameters) on invariants, and
check the product's kernel source (CPU lane emulation, tests/emu) against the oracle."""
import ctypes
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gold(name):
    return json.load(open(os.path.join(GOLD, name)))


@pytest.fixture(scope="module")
def h1():
    return O.Oracle("h1", tolerance=1e-14)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# ---------------------------------------------------------------------------------- golden vectors (reference code)
def test_standing_reward_matches_reference(h1):
    o = h1
    envs = o.make_envs(1)
    for c in gold("h1_standing_reward.json"):
        assert c["names"] == ["com_vel_error", "yaw_vel_error", "height", "upperbody", "joint_torque_reward", "posture"]
        o.set_field(envs, 0, "root_xmat", c["root_xmat"])
        o.set_field(envs, 0, "root_xpos", c["root_xpos"])
        o.set_field(envs, 0, "root_vlin", c["root_vlin_world"])
        o.set_field(envs, 0, "qvel", c["qvel"] + [0.0, 0.0])
        o.set_field(envs, 0, "act_len", c["act_len"] + [0.0, 0.0])
        o.set_field(envs, 0, "act_force", c["act_force"] + [0.0, 0.0])
        t = o.calc_reward(envs, 0, np.zeros(12))
        assert np.abs(t[:6] - np.array(c["terms"])).max() < 1e-13
        assert (t[6:] == 0).all()


def test_randomize_dynamics_matches_reference_draw_order_and_ranges(h1):
    o = h1
    for c in gold("h1_domain_randomization.json")["randomize_dynamics"]:
        envs = o.make_envs(1, seed=c["seed"], first_id=c["env_id"])
        o.set_field(envs, 0, "rng_ctr", [c["ctr"]])
        o.lib.orc_test_randomize_dynamics(o._model, o.env_ptr(envs, 0))
        assert np.abs(o.field(envs, 0, "P_frictionloss")[6:16] - c["frictionloss"]).max() < 1e-15
        assert np.abs(o.field(envs, 0, "P_damping")[6:16] - c["damping"]).max() < 1e-15
        assert abs(o.field(envs, 0, "P_pel_mass")[0] - c["pelvis_mass"]) < 1e-14
        assert np.abs(o.field(envs, 0, "P_pel_com") - c["pelvis_ipos"]).max() < 1e-16
        assert np.abs(o.field(envs, 0, "P_mass")[1:11] - c["link_mass"]).max() < 1e-14
        assert np.abs(o.field(envs, 0, "P_com").reshape(-1, 3)[1:11] - np.array(c["link_ipos"])).max() < 1e-16
        # the welded root link = randomised pelvis body + constant rest: mass and first moment add up
        rp = o.mj["root_parts"]
        M = o.field(envs, 0, "P_mass")[0]
        assert abs(M - (c["pelvis_mass"] + rp["rest"]["mass"])) < 1e-12
        mc = c["pelvis_mass"] * np.array(c["pelvis_ipos"]) + np.array(rp["rest"]["mc"])
        assert np.abs(M * o.field(envs, 0, "P_com")[:3] - mc).max() < 1e-12
        # composite inertia about the com stays symmetric positive definite
        Ic = o.field(envs, 0, "P_inertia")[:9].reshape(3, 3)
        assert np.abs(Ic - Ic.T).max() < 1e-14 and np.linalg.eigvalsh(Ic).min() > 0.1


def test_apply_perturbation_matches_reference(h1):
    o = h1
    seen_zero = seen_set = False
    for c in gold("h1_domain_randomization.json")["apply_perturbation"]:
        envs = o.make_envs(1, seed=c["seed"], first_id=c["env_id"])
        o.set_field(envs, 0, "rng_ctr", [c["ctr"]])
        o.lib.orc_test_apply_perturbation(o._model, o.env_ptr(envs, 0))
        assert np.abs(o.field(envs, 0, "xfrc") - c["xfrc"]).max() < 1e-14
        seen_zero |= not np.any(c["xfrc"])
        seen_set |= bool(np.any(c["xfrc"]))
    assert seen_zero and seen_set


def test_init_and_observation_noise_match_reference(h1):
    o = h1
    g = gold("h1_noise.json")
    for c in g["init_noise"]:
        envs = o.make_envs(1, seed=c["seed"], first_id=c["env_id"])
        o.set_field(envs, 0, "rng_ctr", [c["ctr"]])
        o.lib.orc_test_init_pose(o._model, o.env_ptr(envs, 0))
        assert np.abs(o.field(envs, 0, "qpos")[:17] - c["qpos"]).max() < 1e-15
    for c in g["observation_noise"]:
        envs = o.make_envs(1, seed=c["seed"], first_id=c["env_id"])
        o.set_field(envs, 0, "rng_ctr", [c["ctr"]])
        clean = np.array(c["clean"])
        # drive get_obs: roll/pitch through a quaternion with those euler angles, the rest through the lagged fields
        r, p = clean[0], clean[1]
        cr, sr, cp, sp = np.cos(r / 2), np.sin(r / 2), np.cos(p / 2), np.sin(p / 2)
        q = np.zeros(19)
        q[3:7] = [cp * cr, cp * sr, sp * cr, -sp * sr]
        o.set_field(envs, 0, "qpos", q)
        v = np.zeros(18)
        v[3:6] = clean[2:5]
        o.set_field(envs, 0, "qvel", v)
        for name, lo in (("act_len", 5), ("act_vel", 15), ("act_force", 25)):
            o.set_field(envs, 0, name, list(clean[lo:lo + 10]) + [0.0, 0.0])
        obs = np.zeros(35)
        o.lib.orc_test_get_obs(o._model, o.env_ptr(envs, 0), _p(obs))
        assert np.abs(obs - np.array(c["noisy"])).max() < 1e-13


# ---------------------------------------------------------------------------------- physics invariants
def _quiet(name="h1", **kw):
    """Oracle with noise / randomisation switched off (deterministic nominal model)."""
    o = O.Oracle(name, **kw)
    mj = json.loads(json.dumps(o.mj))
    mj["cfg"]["observation_noise"]["enabled"] = False
    mj["cfg"]["perturbation"]["enable"] = False
    mj["cfg"]["dynamics_randomization"]["enable"] = False
    mj["cfg"]["init_noise_deg"] = 0
    flat = O.pack_model(mj, o.clocks, kw.get("tolerance"), kw.get("solver", 0), kw.get("iterations"))
    assert o.lib.orc_model_from_flat(o._model, _p(flat), len(flat)) == 0
    return o


def test_static_stance_contact_force_is_weight():
    o = _quiet(tolerance=1e-12)
    envs = o.make_envs(1)
    o.reset(envs, 0)
    # stiff PD towards the reset pose so the robot stands still, then the ground reaction must carry m g
    q0 = o.field(envs, 0, "qpos")[7:17].copy()
    for _ in range(1500):
        al, av = o.field(envs, 0, "act_len")[:10], o.field(envs, 0, "act_vel")[:10]
        ctrl = 4000 * (q0 - al) - 50 * av
        o.mj_step(envs, 0, np.concatenate([ctrl, [0, 0]]))
    assert np.abs(o.field(envs, 0, "qvel")[:16]).max() < 5e-2
    grf = o.field(envs, 0, "rfoot_grf")[0] + o.field(envs, 0, "lfoot_grf")[0]
    assert abs(grf - o.mj["total_mass"] * 9.81) / (o.mj["total_mass"] * 9.81) < 0.03


def test_external_wrench_changes_momentum_at_the_commanded_rate():
    o = _quiet(tolerance=1e-12)
    envs = o.make_envs(1)
    q = np.zeros(19)
    q[:17] = o.mj["cfg"]["nominal_qpos"]
    q[2] = 5.0                                      # free flight: no contacts
    o.set_field(envs, 0, "qpos", q)
    f = np.array([3.0, -7.0, 5.0])
    xf = np.zeros(12)
    xf[0:3] = f                                     # force on the pelvis body
    xf[6 + 3:6 + 6] = [0.5, -1.0, 2.0]              # torque on the torso body
    o.set_field(envs, 0, "xfrc", xf)
    n = 200
    for _ in range(n):
        o.mj_step(envs, 0, np.zeros(12))
    # total linear momentum = M_total * v_com; with zero joint torques the com accelerates by g + f / M
    # (check through the root translation of the first step being dominated by f: use the exact com instead)
    M = o.mj["total_mass"]
    t = n * 1e-3
    # centre-of-mass velocity from the generalised momentum p = (M qvel)[0:3] (world-frame translation dofs)
    Mq = o.mass_matrix(o.field(envs, 0, "qpos")[:17])
    p = (Mq @ o.field(envs, 0, "qvel")[:16])[:3]
    expect = (f + np.array([0, 0, -9.81 * M])) * t
    assert np.abs(p - expect).max() < 2e-2 * np.abs(expect).max()


def test_friction_loss_newton_equals_pgs_and_saturates():
    """Friction-loss rows make the constraint cost a Huber function; the primal Newton solution must agree with the
    independent dual PGS (box-projected) solution, and the joint must not move under a torque below the loss level."""
    oN = _quiet(tolerance=1e-14)
    oP = _quiet(tolerance=1e-14, solver=1, iterations=100)
    rng = np.random.RandomState(0)
    for trial in range(4):
        q = np.zeros(19)
        q[:17] = oN.mj["cfg"]["nominal_qpos"]
        q[2] = 0.97 if trial % 2 == 0 else 5.0      # with / without ground contact
        q[7:17] += rng.uniform(-0.1, 0.1, 10)
        v = np.zeros(18)
        v[:16] = rng.normal(size=16) * 0.2
        fl = np.zeros(18)
        fl[6:16] = rng.uniform(0.2, 2.0, 10)
        ctrl = np.concatenate([rng.uniform(-10, 10, 10), [0, 0]])
        out = []
        for o in (oN, oP):
            envs = o.make_envs(1)
            o.set_field(envs, 0, "qpos", q)
            o.set_field(envs, 0, "qvel", v)
            o.set_field(envs, 0, "P_frictionloss", fl)
            o.mj_step(envs, 0, ctrl)
            out.append(o.field(envs, 0, "qacc")[:16].copy())
            assert o.field(envs, 0, "last_kkt_residual")[0] < 1e-6
        assert np.abs(out[0] - out[1]).max() < 1e-6 * max(1.0, np.abs(out[0]).max())
    # stiction: floating robot at rest, one joint driven with 1 N m against 1.5 N m of friction loss -> stays at rest
    envs = oN.make_envs(1)
    q = np.zeros(19)
    q[:17] = oN.mj["cfg"]["nominal_qpos"]
    q[2] = 5.0
    oN.set_field(envs, 0, "qpos", q)
    fl = np.zeros(18)
    fl[6:16] = 1.5
    oN.set_field(envs, 0, "P_frictionloss", fl)
    g0 = oN.bias(q[:17], np.zeros(16))               # gravity torque on the joints is zero in free fall (all links fall)
    ctrl = np.zeros(12)
    ctrl[3] = 1.0
    for _ in range(50):
        oN.mj_step(envs, 0, ctrl)
    # (soft constraint: the regulariser R lets the joint creep, two orders below the ~0.25 rad/s of a free joint)
    assert abs(oN.field(envs, 0, "qvel")[6 + 3]) < 2e-2, (oN.field(envs, 0, "qvel")[6:16], g0[6:])
    ctrl[3] = 6.0                                    # above the loss level it breaks away
    for _ in range(50):
        oN.mj_step(envs, 0, ctrl)
    assert abs(oN.field(envs, 0, "qvel")[6 + 3]) > 0.05


# ---------------------------------------------------------------------------------- kernel source (lane emulation) vs oracle
def test_kernel_source_matches_oracle_h1_fp64(h1):
    from emu import Emu
    from learninghumanoidwalking_b200.model import load_model, pack_model
    o = h1
    N = 6
    e = Emu(pack_model(load_model("h1"), tolerance=1e-14), 64, N, seed=9, first_id=100)
    assert e.nobs == 35 and e.nu == 10
    envs = o.make_envs(N, seed=9, first_id=100)
    assert np.abs(o.batch_reset(envs, N) - e.reset()).max() < 1e-12
    rng = np.random.RandomState(3)
    n_end, pushed = 0, False
    for t in range(450):
        a = rng.normal(size=(N, 10)) * 0.3
        oo, to, tt, rr, dd, ee = o.batch_step(envs, N, a, max_traj_len=60)
        eo, et, etm, er, ed, een, eplen, eprew = e.step(a, max_traj_len=60)
        assert (dd == ed).all() and (ee == een).all()
        assert np.abs(oo - eo).max() < 1e-8 and np.abs(rr - er).max() < 1e-10 and np.abs(tt - etm).max() < 1e-10
        m = ee.astype(bool)
        if m.any():
            assert np.abs(to[m] - et[m]).max() < 1e-8
            n_end += int(m.sum())
        x = np.array([o.field(envs, i, "xfrc") for i in range(N)])
        pushed |= bool(np.any(x))
        assert np.abs(x - e.sr[:, -12:]).max() < 1e-12      # same pushes, same coin flips
    assert n_end >= 20 and pushed
    # per-env randomised parameters agree (record tail: mass, com, inertia0, damping, floss, pelvis com, xfrc)
    for i in range(N):
        assert np.abs(o.field(envs, i, "P_mass")[:11] - e.sr[i, 103:114]).max() < 1e-13
        assert np.abs(o.field(envs, i, "P_frictionloss")[6:16] - e.sr[i, 163:173]).max() < 1e-15


def test_kernel_source_h1_self_collision_flag(h1):
    """Large actions make the legs cross: the leg-vs-leg capsule test must fire in the kernel source exactly when it
    does in the oracle (stepping without auto-reset so the terminal state can be inspected)."""
    from emu import Emu
    from learninghumanoidwalking_b200.model import load_model, pack_model
    o, N = h1, 8
    e = Emu(pack_model(load_model("h1"), tolerance=1e-14), 64, N, seed=2, first_id=0)
    envs = o.make_envs(N, seed=2, first_id=0)
    o.batch_reset(envs, N)
    e.reset()
    rng = np.random.RandomState(5)
    n_self = n_done = 0
    for t in range(140):
        a = rng.normal(size=(N, 10))
        eo, _, _, er, ed, *_ = e.step(a, autoreset=0)
        for i in range(N):
            ob, r, d, _ = o.step(envs, i, a[i])
            assert bool(ed[i]) == d
            assert np.abs(ob - eo[i]).max() < 1e-8
            if d:
                n_done += 1
                z = o.field(envs, i, "qpos")[2]
                sc = int(o.field(envs, i, "self_collision")[0])
                n_self += sc
                assert sc or z < 0.9 or z > 1.4
                assert np.abs(o.reset(envs, i) - e.reset_one(i)).max() < 1e-12
    assert n_done > 30 and n_self >= 2, (n_done, n_self)


def test_kernel_source_h1_thigh_against_the_torso_hip_capsule():
    """The torso's "hip" capsule (h1.xml:154, welded to the pelvis) against thighs / shins: the self-contact large actions reach
    first (tests/golden/h1_self_collision_eval.json).  Spreading both hips ends the episode on that pair while every leg-vs-leg
    pair is still open; kernel source and oracle terminate on the same step."""
    import json
    from emu import Emu
    from learninghumanoidwalking_b200.model import load_model, pack_model
    from tools.eval_h1_self_collision import seg_seg
    from tools.shin_clearance import link_poses
    mj = load_model("h1")
    sc = mj["self_collision"]
    assert len(sc["capsules"]) == 15 and len(sc["pairs"]) == 57 and sc["capsules"][14]["link"] == 0
    ev = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "h1_self_collision_eval.json")))
    assert ev["rollout_sigma_1.0"]["false_negative_rate"] == 0.0 and ev["uniform_joint_ranges"]["false_negative_rate"] == 0.0
    assert ev["before_the_hip_capsule_pairs_were_added"]["rollout_sigma_1.0"]["false_negative_rate"] == 1.0
    o = O.Oracle("h1", tolerance=1e-14)
    e = Emu(pack_model(mj, tolerance=1e-14), 64, 1, seed=2, first_id=0)
    envs = o.make_envs(1, seed=2, first_id=0)
    o.reset(envs)
    e.reset()
    a = np.zeros(10)
    a[1], a[6] = 2.0, -2.0      # left / right hip roll targets outwards
    hit = None
    for k in range(20):
        _, _, d, _ = o.step(envs, 0, a)
        out = e.step(a[None], autoreset=0)
        assert bool(out[4][0]) == d, k
        if d:
            hit = k
            break
    q = np.asarray(o.field(envs, 0, "qpos"))
    assert hit is not None and int(o.field(envs, 0, "self_collision")[0]) == 1 and 0.9 < q[2] < 1.4
    R, p = link_poses(mj["links"], q)
    E = [(p[c["link"]] + R[c["link"]] @ np.array(c["p0"]), p[c["link"]] + R[c["link"]] @ np.array(c["p1"]), c["radius"])
         for c in sc["capsules"]]
    gap = lambda a_, b_: seg_seg(E[a_][0], E[a_][1], E[b_][0], E[b_][1]) - E[a_][2] - E[b_][2]
    assert min(gap(a_, b_) for a_, b_ in sc["pairs"] if b_ == 14) < 0 < min(gap(a_, b_) for a_, b_ in sc["pairs"] if b_ != 14)


def test_kernel_source_h1_fp32_stays_close(h1):
    from emu import Emu
    from learninghumanoidwalking_b200.model import load_model, pack_model
    o = O.Oracle("h1", tolerance=1e-10)
    e = Emu(pack_model(load_model("h1"), tolerance=1e-6), 32, 2, seed=4, first_id=0)
    envs = o.make_envs(2, seed=4, first_id=0)
    assert np.abs(o.batch_reset(envs, 2) - e.reset()).max() < 2e-3
    rng = np.random.RandomState(0)
    for t in range(8):
        a = rng.normal(size=(2, 10)) * 0.2
        oo, _, _, rr, dd, _ = o.batch_step(envs, 2, a)
        eo, _, _, er, ed, *_ = e.step(a)
        if dd.any() or ed.any():
            break
        # torque observations (scale 100) dominate the absolute error; compare in normalised units
        std = np.concatenate(([0.2, 0.2, 1, 1, 1], 0.5 * np.ones(10), 4 * np.ones(10), 100 * np.ones(10)))
        assert (np.abs(oo - eo) / std).max() < 5e-2 and np.abs(rr - er).max() < 5e-3


# ---------------------------------------------------------------------------------- PD-gain randomisation (RobotBase pdrand_k)
def test_pd_gain_randomisation_matches_reference_robot_base():
    """robots/robot_base.py:41-47 fed with the oracle's Philox words must give the oracle's gains."""
    for c in gold("h1_pd_gain_randomization.json"):
        o = O.Oracle("h1", pdrand_k=c["k"])
        envs = o.make_envs(1, seed=c["seed"], first_id=c["env_id"])
        o.set_field(envs, 0, "rng_ctr", [c["ctr"]])
        kp, kd = np.zeros(12), np.zeros(12)
        o.lib.orc_test_pd_gains(o._model, o.env_ptr(envs, 0), _p(kp), _p(kd))
        assert np.abs(kp[:10] - c["kp"]).max() < 1e-12 and np.abs(kd[:10] - c["kd"]).max() < 1e-13


@pytest.mark.parametrize("name,nu", [("h1", 10), ("jvrc_walk", 12)])
def test_kernel_source_matches_oracle_with_pd_gain_randomisation(name, nu):
    from emu import Emu
    from learninghumanoidwalking_b200.model import load_model, pack_model
    o = O.Oracle(name, tolerance=1e-14, pdrand_k=0.3)
    N = 3
    e = Emu(pack_model(load_model(name), tolerance=1e-14, pd_gain_randomization=0.3), 64, N, seed=6, first_id=2)
    envs = o.make_envs(N, seed=6, first_id=2)
    assert np.abs(o.batch_reset(envs, N) - e.reset()).max() < 1e-12
    rng = np.random.RandomState(1)
    o0 = O.Oracle(name, tolerance=1e-14)            # same seeds, gains not randomised: must diverge from the above
    envs0 = o0.make_envs(N, seed=6, first_id=2)
    o0.batch_reset(envs0, N)
    differs = False
    for t in range(40):
        a = rng.normal(size=(N, nu)) * 0.3
        oo, _, tt, rr, dd, ee = o.batch_step(envs, N, a, max_traj_len=30)
        eo, _, etm, er, ed, een, *_ = e.step(a, max_traj_len=30)
        assert (dd == ed).all() and (ee == een).all()
        assert np.abs(oo - eo).max() < 1e-8 and np.abs(rr - er).max() < 1e-10
        if t < 5:
            differs |= np.abs(o0.batch_step(envs0, N, a, max_traj_len=30)[0] - oo).max() > 1e-6
    assert differs
