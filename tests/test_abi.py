"""CPU-only: the C-ABI library loads and exports every symbol include/bmi.h declares; the ctypes
binding table covers exactly that set (no compute calls here)."""
import ctypes
import os
import re

from rl_arm_under_sparse_reward_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "bmi.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bmi_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    syms = _header_symbols()
    assert len(syms) >= 35
    for must in ("bmi_her_sample", "bmi_env_step", "bmi_ddpg_backward", "bmi_comm_allreduce_sum_f32", "bmi_norm_update"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in _header_symbols():
        assert hasattr(lib, s), "libbmi_b200.so does not export " + s
    assert lib.bmi_abi_version() == 1


def test_binding_table_matches_header():
    assert sorted(_lib.SIGNATURES.keys()) == _header_symbols()


def test_errors_are_reported_not_swallowed():
    lib = _lib.load()
    rc = lib.bmi_compute_reward(None, None, 4, 3, 7, 0.05, None, None)   # bad dtype code -> argument error, no launch
    assert rc == -1
    assert b"dtype" in lib.bmi_last_error()
    rc = lib.bmi_her_sample(None, 0, None, None, None, None, 1, 0.8, 0.05, None, None)
    assert rc == -1


def test_env_create_validates_the_model_blob_before_touching_the_gpu():
    """bmi_env_create checks magic / size / topology / staging capacity on the host first: these calls fail with an
    argument error and a message, without a CUDA call (so they run on a CPU-only box)."""
    import numpy as np
    lib = _lib.load()
    blob = np.fromfile(os.path.join(ROOT, "rl_arm_under_sparse_reward_b200", "assets", "bmirobot_model.bin"), dtype="<f4")
    h = ctypes.c_void_p()

    def create(b, n_envs=4, task=0, nbytes=None):
        return lib.bmi_env_create(ctypes.byref(h), n_envs, task, b.ctypes.data_as(ctypes.c_void_p),
                                  int(b.nbytes if nbytes is None else nbytes))

    bad = blob.copy(); bad[0] = 1.0
    assert create(bad) == -1 and b"magic" in lib.bmi_last_error()
    assert create(blob, nbytes=blob.nbytes - 4) == -1                       # MP_TOTAL disagrees with the byte count
    assert create(blob, n_envs=0) == -1 and b"n_envs" in lib.bmi_last_error()
    assert create(blob, task=7) == -1 and b"task" in lib.bmi_last_error()
    bad = blob.copy(); bad[64 + 3 * 32 + 0] = 0.0                           # link 3 claims parent 0: not the compiled arm topology
    assert create(bad) == -1 and b"parent" in lib.bmi_last_error()
    assert lib.bmi_env_create(None, 4, 0, blob.ctypes.data_as(ctypes.c_void_p), int(blob.nbytes)) == -1
    assert lib.bmi_env_num_envs(None) == -1
