"""The oracle has no MuJoCo golden vectors to lean on (parity unpinned), so it is anchored on physics:
energy consistency of (M, bias, gravity), M symmetric positive definite and equal to an independent numpy
assembly, static contact force = m g, unique constraint solution (Newton == PGS, KKT residual ~ 0)."""
import copy
import ctypes
import os
import sys

import numpy as np

from oracle import oracle as O

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(__file__)), "tools"))


def _oracle_with(mutate, **kw):
    o = O.Oracle(**kw)
    mj = copy.deepcopy(o.mj)
    mutate(mj)
    flat = O.pack_model(mj, o.clocks, kw.get("tolerance"), kw.get("solver", 0), kw.get("iterations"))
    assert o.lib.orc_model_from_flat(o._model, flat.ctypes.data_as(ctypes.c_void_p), len(flat)) == 0
    o.mj = mj
    return o


def _random_state(mj, rng, height=5.0):
    q = np.array(mj["cfg"]["nominal_qpos"])
    q[2] = height
    q[3:7] = rng.normal(size=4)
    q[3:7] /= np.linalg.norm(q[3:7])
    q[7:] += rng.uniform(-0.2, 0.2, 12)
    return q, rng.normal(size=18)


def test_mass_matrix_spd_and_matches_independent_numpy_assembly(oracle_tight):
    import compile_model as cm
    rng = np.random.RandomState(0)
    for _ in range(5):
        q, _ = _random_state(oracle_tight.mj, rng)
        M = oracle_tight.mass_matrix(q)
        assert np.abs(M - M.T).max() < 1e-12
        assert np.linalg.eigvalsh(M).min() > 1e-3
        assert np.abs(M - cm.mass_matrix(oracle_tight.mj, q)).max() < 1e-11
        assert np.abs(M[6:12, 12:18]).max() == 0.0  # the two legs only couple through the root


def test_energy_drift_is_first_order_in_timestep():
    """Free flight, no damping: semi-implicit Euler drifts O(h); halving h must halve the drift.  This ties the
    bias forces (Coriolis + gravity) to the mass matrix: an inconsistent pair shows an h-independent drift."""
    drifts = []
    for h in (1e-3, 5e-4, 2.5e-4):
        def mut(mj, h=h):
            mj["opt"]["timestep"] = h
            for lk in mj["links"][1:]:
                lk["joint"]["damping"] = 0.0
        o = _oracle_with(mut)
        envs = o.make_envs(1)
        q, v = _random_state(o.mj, np.random.RandomState(3))
        o.set_field(envs, 0, "qpos", q)
        o.set_field(envs, 0, "qvel", v)
        e0 = sum(o.energy(q, v))
        for _ in range(int(round(0.2 / h))):
            o.mj_step(envs, 0, np.zeros(12))
        drifts.append(sum(o.energy(o.field(envs, 0, "qpos"), o.field(envs, 0, "qvel"))) - e0)
    assert abs(drifts[0]) < 1.0 and abs(drifts[0] / drifts[1] - 2) < 0.05 and abs(drifts[1] / drifts[2] - 2) < 0.05, drifts


def test_free_fall_acceleration_is_g(oracle_tight):
    o = oracle_tight
    envs = o.make_envs(1)
    q = np.array(o.mj["cfg"]["nominal_qpos"])
    q[2] = 3.0
    o.set_field(envs, 0, "qpos", q)
    o.set_field(envs, 0, "qvel", np.zeros(18))
    o.mj_step(envs, 0, np.zeros(12))
    qacc = o.field(envs, 0, "qacc")
    assert abs(qacc[2] + 9.81) < 1e-9 and np.abs(qacc[:2]).max() < 1e-9 and np.abs(qacc[3:6]).max() < 1e-9


def test_static_stance_contact_force_equals_weight():
    """Stiff PD (kp x40, kd x5: explicit PD damping is unstable beyond kd h / I = 2) makes the robot quasi-rigid; after settling, sum of normal forces = m g and GRF >= normal."""
    def mut(mj):
        mj["cfg"]["kp"] = [40 * k for k in mj["cfg"]["kp"]]
        mj["cfg"]["kd"] = [5 * k for k in mj["cfg"]["kd"]]
    o = _oracle_with(mut, tolerance=1e-14)
    envs = o.make_envs(1)
    o.reset(envs)
    for _ in range(200):
        _, _, done, _ = o.step(envs, 0, np.zeros(12))
        assert not done
    grf = o.field(envs, 0, "rfoot_grf")[0] + o.field(envs, 0, "lfoot_grf")[0]
    weight = o.mj["total_mass"] * 9.81
    assert o.field(envs, 0, "ncon")[0] == 8
    assert np.abs(o.field(envs, 0, "qvel")).max() < 5e-2
    assert weight * 0.995 < grf < weight * 1.10, (grf, weight)   # norm includes friction => slightly above m g
    assert o.field(envs, 0, "last_kkt_residual")[0] < 1e-9


def test_newton_and_pgs_reach_the_same_constrained_acceleration():
    newton = O.Oracle(tolerance=1e-14, solver=O.ctypes.c_int(0).value)
    pgs = O.Oracle(tolerance=1e-14, solver=1)
    rng = np.random.RandomState(5)
    for _ in range(3):
        q = np.array(newton.mj["cfg"]["nominal_qpos"])
        q[2] = 0.805
        q[3:7] += rng.normal(size=4) * 0.02
        q[3:7] /= np.linalg.norm(q[3:7])
        q[7:] += rng.uniform(-0.1, 0.1, 12)
        v = rng.normal(size=18) * 0.3
        accs = []
        for o in (newton, pgs):
            envs = o.make_envs(1)
            o.set_field(envs, 0, "qpos", q)
            o.set_field(envs, 0, "qvel", v)
            o.mj_step(envs, 0, rng.uniform(-5, 5, 12) * 0)
            assert o.field(envs, 0, "ncon")[0] > 0
            assert o.field(envs, 0, "last_kkt_residual")[0] < 1e-7
            accs.append(o.field(envs, 0, "qacc"))
        assert np.abs(accs[0] - accs[1]).max() < 1e-6 * max(1.0, np.abs(accs[0]).max())


def test_joint_limit_pushes_back(oracle_tight):
    o = oracle_tight
    envs = o.make_envs(1)
    q = np.array(o.mj["cfg"]["nominal_qpos"])
    q[2] = 3.0
    q[7 + 3] = -0.05  # right knee below its lower limit 0
    o.set_field(envs, 0, "qpos", q)
    o.set_field(envs, 0, "qvel", np.zeros(18))
    o.mj_step(envs, 0, np.zeros(12))
    free = O.Oracle(tolerance=1e-14)
    e2 = free.make_envs(1)
    q2 = q.copy()
    q2[7 + 3] = 0.05
    free.set_field(e2, 0, "qpos", q2)
    free.set_field(e2, 0, "qvel", np.zeros(18))
    free.mj_step(e2, 0, np.zeros(12))
    assert o.field(envs, 0, "qacc")[6 + 3] > free.field(e2, 0, "qacc")[6 + 3] + 10.0


def test_philox_stream_is_counter_based(oracle_tight):
    a = oracle_tight.philox(1, 2, 3, 4)
    assert a == oracle_tight.philox(1, 2, 3, 4) and a != oracle_tight.philox(1, 2, 4, 4) and a != oracle_tight.philox(2, 2, 3, 4)


def _momenta(o, q, v):
    """Total linear momentum P (world) and angular momentum about the centre of mass L_C (world) from the generalised momentum
    p = M(q) v alone: p[0:3] = P; R p[3:6] = angular momentum about the root origin O (the free joint's angular velocity is
    body-frame); M[0:3,3:6] = -m R [c]x gives the centre of mass c relative to O; L_C = L_O - (R c) x P."""
    M = o.mass_matrix(q)
    p = M @ v
    w_, x, y, z = q[3:7]
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w_ * z), 2 * (x * z + w_ * y)],
                  [2 * (x * y + w_ * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w_ * x)],
                  [2 * (x * z - w_ * y), 2 * (y * z + w_ * x), 1 - 2 * (x * x + y * y)]])
    m = M[0, 0]
    S = -R.T @ M[0:3, 3:6] / m
    assert np.abs(S + S.T).max() < 1e-12                     # a cross-product matrix, as the derivation says
    c = np.array([S[2, 1], S[0, 2], S[1, 0]])
    return m, p[0:3], R @ p[3:6] - np.cross(R @ c, p[0:3]), q[0:3] + R @ c


def test_momentum_balance_in_free_flight_under_internal_torques_and_damping():
    """Joint torques and joint damping are internal: in free flight the total linear momentum changes at exactly m g and the
    angular momentum about the centre of mass not at all, whatever the legs do.  Semi-implicit Euler keeps both to O(h): the
    error must halve with the time step (an inconsistent Coriolis / mass-matrix pair leaves an h-independent residue), and the
    centre of mass recovered from the mass matrix must be the one the potential energy sees."""
    errs = []
    for h in (1e-3, 5e-4, 2.5e-4):
        o = _oracle_with(lambda mj, h=h: mj["opt"].__setitem__("timestep", h))
        assert any(lk["joint"]["damping"] > 0 for lk in o.mj["links"][1:])
        envs = o.make_envs(1)
        rng = np.random.RandomState(11)
        q, v = _random_state(o.mj, rng)
        v[6:] *= 3.0
        o.set_field(envs, 0, "qpos", q)
        o.set_field(envs, 0, "qvel", v)
        m, P0, L0, C0 = _momenta(o, q, v)
        assert abs(m - o.mj["total_mass"]) < 1e-9
        ke, pe = o.energy(q, v)
        assert abs(pe - m * 9.81 * C0[2]) < 1e-8 * abs(pe)     # same centre of mass as the gravity term
        T = 0.1
        for k in range(int(round(T / h))):
            o.mj_step(envs, 0, 40.0 * np.sin(0.05 * k * h / 1e-3 + np.arange(12)))   # strong, time-varying internal torques
        q1, v1 = o.field(envs, 0, "qpos"), o.field(envs, 0, "qvel")
        _, P1, L1, _ = _momenta(o, q1, v1)
        errs.append((np.linalg.norm(P1 - P0 - m * T * np.array([0, 0, -9.81])), np.linalg.norm(L1 - L0), np.linalg.norm(L0)))
    (p0, l0, scale), (p1, l1, _), (p2, l2, _) = errs
    assert scale > 5.0 and l0 < 0.02 * scale and p0 < 0.02 * m, errs          # small at the reference's time step already
    assert abs(l0 / l1 - 2) < 0.15 and abs(l1 / l2 - 2) < 0.15, errs           # and first order in h
    assert abs(p0 / p1 - 2) < 0.15 and abs(p1 / p2 - 2) < 0.15, errs
