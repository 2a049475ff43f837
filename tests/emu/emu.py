"""ctypes wrapper of the CPU emulation of the kernel source (tests only; see sim_emu.cpp)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
# LHW_EMU_DEFINES="LHW_X_CF=1 ..." builds (and loads) the emulation of a candidate kernel variant instead (see sim_core.h)
DEFINES = os.environ.get("LHW_EMU_DEFINES", "").split()
LIB = os.path.join(HERE, "_build", "libsim_emu" + "".join("." + d.replace("=", "_") for d in DEFINES) + ".so")
NREW, NSTATE_I = 10, 8


def build():
    srcs = [os.path.join(HERE, "sim_emu.cpp"), os.path.join(ROOT, "learninghumanoidwalking_b200", "csrc", "sim_core.h"),
            os.path.join(ROOT, "learninghumanoidwalking_b200", "csrc", "model_pack.h")]
    if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17"] + ["-D" + d for d in DEFINES] + ["-o", LIB, srcs[0]])
    return LIB


class Emu:
    def __init__(self, flat, precision=64, n=1, seed=0, first_id=0):
        self.lib = ctypes.CDLL(build())
        self.lib.emu_create.restype = ctypes.c_void_p
        flat = np.ascontiguousarray(flat, dtype=np.float64)
        self.h = ctypes.c_void_p(self.lib.emu_create(flat.ctypes.data_as(ctypes.c_void_p), len(flat), precision))
        assert self.h.value, "emu_create failed"
        self.prec, self.n, self.seed, self.first_id = precision, n, seed, first_id
        self.dt = np.float64 if precision == 64 else np.float32
        self.nr = self.lib.emu_state_words(self.h)
        self.nobs = self.lib.emu_obs_dim(self.h)
        self.nu = 2 * (int(flat[0]) % 100)
        self.nq, self.nv = 7 + self.nu, 6 + self.nu
        self.sr = np.zeros((n, self.nr), dtype=self.dt)
        self.si = np.zeros((n, NSTATE_I), dtype=np.int32)

    def _p(self, a):
        return a.ctypes.data_as(ctypes.c_void_p)

    def reset(self):
        obs = np.zeros((self.n, self.nobs), dtype=self.dt)
        self.lib.emu_reset(self.h, self.prec, self._p(self.sr), self._p(self.si), self.n, ctypes.c_uint32(self.seed),
                           ctypes.c_uint32(self.first_id), self._p(obs), 1)
        return obs

    def reset_one(self, i):
        """Reset env i only (what the RolloutWorker does after a terminal step when autoreset is off)."""
        obs = np.zeros((1, self.nobs), dtype=self.dt)
        self.lib.emu_reset(self.h, self.prec, self._p(self.sr[i]), self._p(self.si[i]), 1, ctypes.c_uint32(self.seed),
                           ctypes.c_uint32(self.first_id + i), self._p(obs), 0)
        return obs[0]

    def step(self, actions, max_traj_len=400, autoreset=1):
        a = np.ascontiguousarray(actions, dtype=self.dt).reshape(self.n, self.nu)
        n = self.n
        obs, tobs = np.zeros((n, self.nobs), self.dt), np.zeros((n, self.nobs), self.dt)
        rew, terms, eprew = np.zeros(n, self.dt), np.zeros((n, NREW), self.dt), np.zeros(n, self.dt)
        done, ended, eplen = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32)
        self.lib.emu_step(self.h, self.prec, self._p(self.sr), self._p(self.si), n, ctypes.c_uint32(self.seed),
                          ctypes.c_uint32(self.first_id), self._p(a), max_traj_len, autoreset, self._p(obs),
                          self._p(tobs), self._p(rew), self._p(terms), self._p(done), self._p(ended), self._p(eplen),
                          self._p(eprew))
        return obs, tobs, terms, rew, done, ended, eplen, eprew

    def substep(self, i, ctrl, nsteps=1):
        assert self.prec == 64
        c = np.ascontiguousarray(ctrl, dtype=np.float64)
        self.lib.emu_substep64(self.h, self._p(self.sr[i]), self._p(self.si[i]), self._p(c), nsteps)

    @property
    def qpos(self):
        return self.sr[:, 0:self.nq]

    @property
    def qvel(self):
        return self.sr[:, self.nq:self.nq + self.nv]
