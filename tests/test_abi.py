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
