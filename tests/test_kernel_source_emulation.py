"""The product's kernel source (csrc/sim_core.h) compiled for the host with the 32 lanes of a warp run
sequentially (tests/emu/, test harness only) against the oracle: the same arithmetic the GPU executes,
checked on the CPU tier.  The GPU tier (test_gpu_parity.py) repeats this through the C-ABI on the device."""
import numpy as np
import pytest

from emu import Emu
from learninghumanoidwalking_b200.model import load_model, pack_model


def test_substep_parity_through_contact_switches(oracle_tight):
    o = oracle_tight
    mj = load_model()
    e = Emu(pack_model(mj, tolerance=1e-14), 64, 1)
    envs = o.make_envs(1)
    rng = np.random.RandomState(1)
    q = np.array(mj["cfg"]["nominal_qpos"])
    q[2] = 0.805
    q[3:7] += rng.normal(size=4) * 0.02
    q[3:7] /= np.linalg.norm(q[3:7])
    q[7:] += rng.uniform(-0.1, 0.1, 12)
    v = rng.normal(size=18) * 0.3
    o.set_field(envs, 0, "qpos", q)
    o.set_field(envs, 0, "qvel", v)
    e.sr[0, 0:19], e.sr[0, 19:37] = q, v
    ctrl = rng.uniform(-20, 20, 12)
    seen = set()
    for _ in range(400):
        o.mj_step(envs, 0, ctrl)
        e.substep(0, ctrl, 1)
        seen.add(int(o.field(envs, 0, "ncon")[0]))
        assert np.abs(o.field(envs, 0, "qpos") - e.qpos[0]).max() < 1e-11
        assert np.abs(o.field(envs, 0, "qvel") - e.qvel[0]).max() < 1e-10
    assert len(seen) >= 3, seen  # went through several contact configurations


def test_env_level_parity_with_autoreset_fp64(oracle_tight):
    o = oracle_tight
    N = 3
    e = Emu(pack_model(load_model(), tolerance=1e-14), 64, N, seed=5, first_id=10)
    envs = o.make_envs(N, seed=5, first_id=10)
    assert np.abs(o.batch_reset(envs, N) - e.reset()).max() < 1e-12
    rng = np.random.RandomState(2)
    n_end = 0
    for _ in range(150):
        a = rng.normal(size=(N, 12)) * 0.3
        oo, to, tt, rr, dd, ee = o.batch_step(envs, N, a, max_traj_len=40)
        eo, et, etm, er, ed, een, eplen, eprew = e.step(a, max_traj_len=40)
        assert (dd == ed).all() and (ee == een).all()
        assert np.abs(oo - eo).max() < 1e-9 and np.abs(rr - er).max() < 1e-10 and np.abs(tt - etm).max() < 1e-10
        m = ee.astype(bool)
        if m.any():
            assert np.abs(to[m] - et[m]).max() < 1e-9
            n_end += int(m.sum())
    assert n_end >= 6
    assert (e.si[:, 1] == [int(o.field(envs, i, "mode")[0]) for i in range(N)]).all()   # same RNG stream
    assert (e.si[:, 0] == [int(o.field(envs, i, "phase")[0]) for i in range(N)]).all()


def test_fp32_kernel_source_stays_close_for_a_short_horizon(oracle_tight):
    o = oracle_tight
    N = 2
    e = Emu(pack_model(load_model(), tolerance=1e-6), 32, N, seed=1)
    envs = o.make_envs(N, seed=1)
    assert np.abs(o.batch_reset(envs, N) - e.reset()).max() < 1e-5
    for _ in range(10):
        a = np.zeros((N, 12))
        oo, _, _, rr, dd, ee = o.batch_step(envs, N, a)
        eo, _, _, er, ed, een, _, _ = e.step(a)
        assert (ee == een).all()
        assert np.abs(oo - eo).max() < 2e-3 and np.abs(rr - er).max() < 2e-3


def test_self_collision_proxies_terminate_when_legs_cross(oracle_tight):
    """reference row S7 (check_self_collisions): capsule proxies flag leg-leg contact; oracle and kernel agree."""
    o = oracle_tight
    mj = load_model()
    assert len(mj["self_collision"]["capsules"]) == 12 and len(mj["self_collision"]["pairs"]) == 36     # two capsules per thigh / shin hull
    e = Emu(pack_model(mj, tolerance=1e-14), 64, 1, seed=2)
    envs = o.make_envs(1, seed=2)
    o.reset(envs)
    e.reset()
    # nominal stance: no self collision, episode continues
    _, _, d0, _ = o.step(envs, 0, np.zeros(12))
    out = e.step(np.zeros((1, 12)), autoreset=0)
    assert not d0 and not out[4][0]
    # swing the right hip inwards (roll) hard: thighs / shanks cross -> both must terminate on the same step
    a = np.zeros(12)
    a[1] = 1.2    # R_HIP_R target +1.2 rad on top of nominal (towards the left leg)
    a[7] = -1.2   # L_HIP_R mirrored
    hit = None
    for k in range(30):
        _, _, d_o, _ = o.step(envs, 0, a)
        out = e.step(a[None], autoreset=0)
        assert bool(out[4][0]) == d_o, k
        if d_o:
            hit = k
            sc = int(o.field(envs, 0, "self_collision")[0])
            z = o.field(envs, 0, "qpos")[2]
            break
    assert hit is not None and sc == 1 and 0.6 < z < 1.4, (hit, sc, z)


@pytest.mark.parametrize("poison", ["nan", "nan_pose", "runaway"])
def test_diverged_environment_is_flagged_terminated_and_reset_without_touching_its_neighbours(oracle_tight, poison):
    """The build's counterpart of MuJoCo's mj_checkAcc auto-reset (SURVEY §8b, C-ABI row): an environment whose state has gone
    non-finite (a NaN velocity -> the factorisation sees a non-positive pivot) or runaway (|qacc| > 1e10) gets a status word,
    terminates on that control step and comes back from the auto-reset with a finite observation; oracle and kernel source
    agree on which step that is, and the other environments of the batch are bit-identical to an unpoisoned run."""
    o = oracle_tight
    N, BAD = 3, 1
    flat = pack_model(load_model(), tolerance=1e-14)
    e, clean = Emu(flat, 64, N, seed=9, first_id=4), Emu(flat, 64, N, seed=9, first_id=4)
    envs = o.make_envs(N, seed=9, first_id=4)
    o.batch_reset(envs, N); e.reset(); clean.reset()
    rng = np.random.RandomState(3)
    for _ in range(3):
        a = rng.normal(size=(N, 12)) * 0.2
        o.batch_step(envs, N, a); e.step(a); clean.step(a)
    qp, qv = o.field(envs, BAD, "qpos"), o.field(envs, BAD, "qvel")
    if poison == "nan":
        qv[7] = np.nan
    elif poison == "nan_pose":
        qp[9] = np.nan                                           # a NaN joint angle: the mass matrix itself is NaN (pivot check)
    else:
        qv[6:] = 1e13
    o.set_field(envs, BAD, "qpos", qp); o.set_field(envs, BAD, "qvel", qv)
    e.sr[BAD, :19], e.sr[BAD, 19:19 + 18] = qp, qv               # record layout: [qpos 19 | qvel 18 | ...]
    a = rng.normal(size=(N, 12)) * 0.2
    oo, to, tt, rr, dd, ee = o.batch_step(envs, N, a)
    eo, et, etm, er, ed, een, eplen, eprew = e.step(a)
    co = clean.step(a)
    assert dd[BAD] == 1 and ee[BAD] == 1 and ed[BAD] == 1 and een[BAD] == 1
    assert np.isfinite(eo).all() and np.isfinite(oo).all()        # the observation returned is the fresh episode's
    assert np.abs(eo[BAD] - oo[BAD]).max() < 1e-12
    assert e.si[BAD, 6] == 0 and e.si[BAD, 2] == 0                # status cleared, traj_len restarted
    keep = [i for i in range(N) if i != BAD]
    assert (eo[keep] == co[0][keep]).all() and (er[keep] == co[3][keep]).all() and (e.sr[keep] == clean.sr[keep]).all()
    for _ in range(3):                                           # and the batch carries on in step with the oracle
        a = rng.normal(size=(N, 12)) * 0.2
        oo, _, _, rr, dd, ee = o.batch_step(envs, N, a)
        eo, _, _, er, ed, een, _, _ = e.step(a)
        assert (dd == ed).all() and (ee == een).all() and np.abs(oo - eo).max() < 1e-9


def test_candidate_defines_reach_the_emulation_build():
    """Kernel candidates are compile-time flags of sim_core.h (-DLHW_X_<name>=1; tools/ab_variants.py builds and times one
    library per flag) and are kept honest on the CPU tier through LHW_EMU_DEFINES: the emulation-vs-oracle check in a fresh
    process whose emulation is compiled with the flags on.  No candidate is pending at the moment (round 2 promoted CF / RSQ /
    GMODEL and deleted SPLITBAR, profiles/r02_ab_variants.md); this keeps the mechanism itself under test with an inert define."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, LHW_EMU_DEFINES="LHW_X_NONE=1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", os.path.join(here, "test_kernel_source_emulation.py"), "-k",
                        "diverged or substep_parity"], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and " passed" in r.stdout, r.stdout[-1500:] + r.stderr[-500:]


_LANE_ORDER_SCRIPT = r"""
import sys
import numpy as np
sys.path[:0] = [sys.argv[2], sys.argv[2] + "/tests/emu"]
from emu import Emu
from learninghumanoidwalking_b200.model import load_model, pack_model
out = {}
for name in ("jvrc_walk", "h1", "jvrc_step", "jvrc_walk_terrain"):
    for prec in (64, 32):
        e = Emu(pack_model(load_model(name), tolerance=1e-10 if prec == 64 else 1e-6), prec, 3, seed=3, first_id=7)
        e.reset()
        rng, acc = np.random.RandomState(5), []
        for k in range(40):
            r = e.step(rng.normal(size=(3, e.nu)) * 0.4, max_traj_len=20)
            acc.append(np.concatenate([np.asarray(x, dtype=np.float64).ravel() for x in (r[0], r[2], r[3], r[4], r[5])] + [e.sr.astype(np.float64).ravel()]))
        out[f"{name}_{prec}"] = np.stack(acc)
np.savez(sys.argv[1], **out)
"""


def test_no_phase_depends_on_the_order_its_lanes_run_in(tmp_path):
    """Race check of the kernel source on the CPU tier.  The emulation runs the 32 lanes of a phase one after the other; a phase in
    which a lane reads what another lane of the SAME phase writes (a missing LHW_SYNC) gives different results when the lanes run
    in descending instead of ascending order.  Every variant and both precisions, closed loop with contacts, falls, truncations,
    resets and (H1) randomisation: observations, reward terms, flags and the whole state record must be bit-identical.  (The one
    phase that emulates a warp-wide prefix count keeps ascending lanes in both builds: LHW_LANES_ORDERED.)"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    runs = {}
    for tag, defs in (("fwd", ""), ("rev", "LHW_EMU_REVERSE=1")):
        out = str(tmp_path / f"{tag}.npz")
        r = subprocess.run([sys.executable, "-c", _LANE_ORDER_SCRIPT, out, root], env=dict(os.environ, LHW_EMU_DEFINES=defs),
                           capture_output=True, text=True, timeout=1200)
        assert r.returncode == 0, r.stderr[-1500:]
        runs[tag] = np.load(out)
    assert len(runs["fwd"].files) == 8
    for k in runs["fwd"].files:
        a, b = runs["fwd"][k], runs["rev"][k]
        assert a.shape == b.shape and np.isfinite(a).all()
        assert (a == b).all(), (k, np.argwhere(a != b)[:3])
