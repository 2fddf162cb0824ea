"""ctypes wrapper of oracle/_build/libbmi_oracle.so (TEST ORACLE — see oracle/__init__.py)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libbmi_oracle.so")
MODEL = os.path.join(os.path.dirname(HERE), "rl_arm_under_sparse_reward_b200", "assets", "bmirobot_model.bin")
SELFCOL = os.path.join(os.path.dirname(HERE), "rl_arm_under_sparse_reward_b200", "assets", "bmirobot_selfcol.bin")
HULLS = os.path.join(os.path.dirname(HERE), "rl_arm_under_sparse_reward_b200", "assets", "bmirobot_hulls.bin")
_dp = ctypes.POINTER(ctypes.c_double)


def build():
    subprocess.run(["make", "-s", "-C", HERE], check=True)
    return LIB


def _lib():
    if not os.path.exists(LIB):
        build()
    lib = ctypes.CDLL(LIB)
    lib.bmo_create.restype = ctypes.c_void_p
    lib.bmo_create.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]
    lib.bmo_get_param.restype = ctypes.c_double
    lib.bmo_get_param.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.bmo_set_param.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double]
    lib.bmo_set_kernel_schedule.restype = None
    lib.bmo_set_kernel_schedule.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.bmo_set_caps.restype = None
    lib.bmo_set_caps.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    lib.bmo_set_selfcol_table.restype = None
    lib.bmo_set_selfcol_table.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
    lib.bmo_contacts.restype = ctypes.c_int
    lib.bmo_contacts.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    lib.bmo_set_hulls.restype = ctypes.c_int
    lib.bmo_set_hulls.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
    for name, n in (("bmo_destroy", 1), ("bmo_reset", 5), ("bmo_step", 6), ("bmo_get_state", 2), ("bmo_set_state", 2),
                    ("bmo_stats", 2), ("bmo_ik", 4), ("bmo_fk_ee", 4), ("bmo_mass_matrix", 3)):
        getattr(lib, name).argtypes = [ctypes.c_void_p] * n
        getattr(lib, name).restype = None
    return lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class OracleEnv:
    """One bmirobot env instance in double precision."""

    def __init__(self, task=0, model_path=MODEL):
        self.lib = _lib()
        self.blob = np.fromfile(model_path, dtype="<f4")
        self.h = self.lib.bmo_create(_p(self.blob), self.blob.shape[0], task)
        if not self.h:
            raise RuntimeError("oracle: bad model blob " + model_path)
        self.task = task
        if os.path.exists(HULLS):
            self.hulls = np.fromfile(HULLS, dtype="<f4")
            if self.lib.bmo_set_hulls(self.h, _p(self.hulls), self.hulls.shape[0]) != 0:
                raise RuntimeError("oracle: bad hull file " + HULLS)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.bmo_destroy(self.h)
            self.h = None

    def use_pair_tables(self, path=SELFCOL):
        """kernel geometry: self-collision pairs from the baked tables instead of GJK / EPA (MP_SELF_TABLE)"""
        self.sc_table = np.fromfile(path, dtype="<f4")
        self.lib.bmo_set_selfcol_table(self.h, _p(self.sc_table), self.sc_table.shape[0])
        self.set_param(55, 1.0)

    def set_caps(self, contacts=9, arm=6):
        """apply the CUDA kernel's contact lane budget (csrc/physics.cu MAXC / MAXA); (0, 0) = keep everything"""
        self.lib.bmo_set_caps(self.h, int(contacts), int(arm))

    def kernel_mode(self, schedule=True):
        """everything the CUDA kernel approximates, so that kernel-vs-oracle tests compare like with like: baked pair
        tables, the 9 / 6 contact lane budget and (schedule=True) the compressed solver iteration schedule.  The default
        OracleEnv is the faithful restatement: GJK + EPA, every contact, Bullet's plain 150-iteration loop."""
        self.use_pair_tables()
        self.set_caps(9, 6)
        self.lib.bmo_set_kernel_schedule(self.h, int(bool(schedule)))
        return self

    def set_param(self, idx, v):
        self.lib.bmo_set_param(self.h, int(idx), float(v))

    def get_param(self, idx):
        return self.lib.bmo_get_param(self.h, int(idx))

    def reset(self, init8):
        init8 = np.ascontiguousarray(init8, dtype=np.float64)
        obs, ag, g = np.zeros(27), np.zeros(3), np.zeros(3)
        self.lib.bmo_reset(self.h, _p(init8), _p(obs), _p(ag), _p(g))
        return obs, ag, g

    def step(self, action):
        a = np.ascontiguousarray(action, dtype=np.float64)
        obs, ag, r, s = np.zeros(27), np.zeros(3), np.zeros(1), np.zeros(1)
        self.lib.bmo_step(self.h, _p(a), _p(obs), _p(ag), _p(r), _p(s))
        return obs, ag, float(r[0]), float(s[0])

    def get_state(self):
        st = np.zeros(48)
        self.lib.bmo_get_state(self.h, _p(st))
        return st

    def set_state(self, st):
        st = np.ascontiguousarray(st, dtype=np.float64)
        self.lib.bmo_set_state(self.h, _p(st))

    def stats(self):
        o = np.zeros(3, dtype=np.int32)
        self.lib.bmo_stats(self.h, _p(o))
        return tuple(int(x) for x in o)

    def contacts(self):
        """contact list of the current state: rows of [link1, link, has_block, dist, n(3), x(3), mu, id]"""
        out = np.zeros(48 * 12)
        n = self.lib.bmo_contacts(self.h, _p(out), 48)
        return out[:12 * n].reshape(n, 12).copy()

    def ik(self, q0, target):
        q0 = np.ascontiguousarray(q0, dtype=np.float64)
        t = np.ascontiguousarray(target, dtype=np.float64)
        out = np.zeros(9)
        self.lib.bmo_ik(self.h, _p(q0), _p(t), _p(out))
        return out

    def fk_ee(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        p, e = np.zeros(3), np.zeros(3)
        self.lib.bmo_fk_ee(self.h, _p(q), _p(p), _p(e))
        return p, e

    def mass_matrix(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        M = np.zeros((9, 9))
        self.lib.bmo_mass_matrix(self.h, _p(q), _p(M))
        return M
