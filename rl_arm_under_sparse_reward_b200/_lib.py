"""ctypes binding of libbmi_b200.so (the C-ABI declared in include/bmi.h).

The product path has no CPU fallback: if the shared library is missing or a call fails the
caller gets an exception (``BmiError``).  torch is imported first so that the CUDA runtime
libraries it bundles (cuBLASLt, NCCL) are already mapped when the library is loaded.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, byref, c_char_p, c_double, c_float, c_int32, c_int64,
                    c_uint64, c_void_p)

HERE = os.path.dirname(os.path.abspath(__file__))
# BMI_B200_LIB points the binding at an instrumented build of the same sources (tools/prof_rollout_phases.py)
LIB_PATH = os.environ.get("BMI_B200_LIB") or os.path.join(HERE, "_C", "libbmi_b200.so")

BMI_F32, BMI_F64 = 0, 1
TASK_PUSH, TASK_PICK = 0, 1
ENV_STATE_DIM = 48


class BmiError(RuntimeError):
    pass


class Episodes(Structure):
    _fields_ = [("obs", c_void_p), ("ag", c_void_p), ("g", c_void_p), ("actions", c_void_p),
                ("n_episodes", c_int64), ("T", c_int32), ("obs_dim", c_int32),
                ("goal_dim", c_int32), ("act_dim", c_int32), ("dtype", c_int32), ("_pad", c_int32)]


class Transitions(Structure):
    _fields_ = [("obs", c_void_p), ("ag", c_void_p), ("g", c_void_p), ("actions", c_void_p),
                ("obs_next", c_void_p), ("ag_next", c_void_p), ("r", c_void_p)]


class DdpgConfig(Structure):
    _fields_ = [("obs_dim", c_int32), ("goal_dim", c_int32), ("act_dim", c_int32),
                ("hidden", c_int32), ("batch", c_int32), ("max_act_rows", c_int32),
                ("action_max", c_float), ("gamma", c_float), ("action_l2", c_float),
                ("lr_actor", c_float), ("lr_critic", c_float), ("polyak", c_float),
                ("adam_beta1", c_float), ("adam_beta2", c_float), ("adam_eps", c_float),
                ("clip_return", c_float), ("one_minus_polyak", c_float), ("_pad", c_float)]


class RolloutArgs(Structure):
    _fields_ = [("T", c_int32), ("explore", c_int32), ("actor_t", c_void_p), ("o_mean", c_void_p), ("o_std", c_void_p),
                ("g_mean", c_void_p), ("g_std", c_void_p), ("clip_range", c_float), ("action_max", c_float),
                ("noise_eps", c_float), ("random_eps", c_float), ("late_clip", c_float), ("seed", c_uint64),
                ("counter", c_void_p), ("episodes", POINTER(Episodes)), ("init", c_void_p), ("obs", c_void_p),
                ("ag", c_void_p), ("g", c_void_p), ("success", c_void_p)]


# name -> (restype, argtypes); every symbol here must be declared in include/bmi.h
SIGNATURES = {
    "bmi_abi_version": (c_int32, []),
    "bmi_last_error": (c_char_p, []),
    "bmi_launch_count": (c_int64, []),
    "bmi_set_l2_fetch_granularity": (c_int32, [c_int32]),
    "bmi_buffer_store": (c_int32, [POINTER(Episodes), POINTER(Episodes), c_void_p, c_void_p]),
    "bmi_compute_reward": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_double,
                                     c_void_p, c_void_p]),
    "bmi_her_sample": (c_int32, [POINTER(Episodes), c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_int64, c_double, c_double, POINTER(Transitions), c_void_p]),
    "bmi_her_sample_inputs": (c_int32, [POINTER(Episodes), c_int64, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_int64, c_double, c_double, c_double, c_double,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p]),
    "bmi_her_draw": (c_int32, [c_uint64, c_void_p, c_int64, c_void_p, c_int32, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_void_p]),
    "bmi_norm_update": (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_double, c_void_p, c_void_p,
                                  c_void_p, c_void_p]),
    "bmi_norm_recompute": (c_int32, [c_void_p] * 8 + [c_int32, c_float, c_float, c_void_p]),
    "bmi_norm_normalize": (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p,
                                     c_double, c_void_p, c_int32, c_void_p]),
    "bmi_preproc_inputs": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_void_p,
                                     c_void_p]),
    "bmi_ddpg_actor_param_count": (c_int64, [POINTER(DdpgConfig)]),
    "bmi_ddpg_critic_param_count": (c_int64, [POINTER(DdpgConfig)]),
    "bmi_ddpg_create": (c_int32, [POINTER(c_void_p), POINTER(DdpgConfig), c_void_p, c_void_p,
                                  c_void_p, c_void_p]),
    "bmi_ddpg_destroy": (c_int32, [c_void_p]),
    "bmi_ddpg_act": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p]),
    "bmi_ddpg_backward": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p]),
    "bmi_ddpg_grad_buffer": (c_int32, [c_void_p, POINTER(c_void_p), POINTER(c_int64)]),
    "bmi_ddpg_adam_step": (c_int32, [c_void_p, c_void_p]),
    "bmi_ddpg_p2p_export": (c_int32, [c_void_p, c_void_p]),
    "bmi_ddpg_p2p_attach": (c_int32, [c_void_p, c_int32, c_int32, c_void_p]),
    "bmi_ddpg_adam_step_p2p": (c_int32, [c_void_p, c_void_p]),
    "bmi_ddpg_p2p_status": (c_int32, [c_void_p, POINTER(c_int32)]),
    "bmi_ddpg_soft_update": (c_int32, [c_void_p, c_void_p]),
    "bmi_select_actions": (c_int32, [c_void_p, c_int64, c_int32, c_float, c_float, c_float, c_float,
                                     c_uint64, c_void_p, c_void_p, c_void_p]),
    "bmi_env_create": (c_int32, [POINTER(c_void_p), c_int32, c_int32, c_void_p, c_int64]),
    "bmi_env_destroy": (c_int32, [c_void_p]),
    "bmi_env_num_envs": (c_int32, [c_void_p]),
    "bmi_env_set_selfcol": (c_int32, [c_void_p, c_void_p, c_int64]),
    "bmi_env_contact_drops": (c_int32, [c_void_p, c_void_p, c_int32]),
    "bmi_env_reset": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p]),
    "bmi_env_sample_init": (c_int32, [c_void_p, c_uint64, c_void_p, c_void_p, c_void_p]),
    "bmi_env_step": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p]),
    "bmi_env_rollout": (c_int32, [c_void_p, POINTER(RolloutArgs), c_void_p]),
    "bmi_actor_transpose": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "bmi_env_get_state": (c_int32, [c_void_p, c_void_p, c_void_p]),
    "bmi_env_set_state": (c_int32, [c_void_p, c_void_p, c_void_p]),
    "bmi_rollout_record": (c_int32, [POINTER(Episodes), c_int32, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p]),
    "bmi_comm_unique_id": (c_int32, [c_void_p]),
    "bmi_comm_init": (c_int32, [POINTER(c_void_p), c_int32, c_int32, c_void_p]),
    "bmi_comm_destroy": (c_int32, [c_void_p]),
    "bmi_comm_allreduce_sum_f32": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p]),
    "bmi_comm_bcast_f32": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_void_p]),
}

_lib = None


def load():
    """Load the shared library once; raises BmiError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    import torch  # noqa: F401  (maps libcublasLt / libnccl / libcudart before our .so)
    if not os.path.exists(LIB_PATH):
        raise BmiError(
            "libbmi_b200.so is not built (%s missing). Run `python -c 'import __graft_entry__ as g; "
            "g.build()'` at the repo root; there is no CPU fallback for the product path." % LIB_PATH)
    try:
        lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    except OSError as e:  # pragma: no cover - depends on the box
        raise BmiError("cannot load %s: %s" % (LIB_PATH, e))
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            raise BmiError("libbmi_b200.so does not export %s (stale build?)" % name)
        fn.restype = res
        fn.argtypes = args
    if lib.bmi_abi_version() != 1:
        raise BmiError("ABI version mismatch: library %d, binding 1" % lib.bmi_abi_version())
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().bmi_last_error()
        raise BmiError("%s failed (%d): %s" % (what or "bmi call", rc, (msg or b"").decode()))


def call(name, *args):
    lib = load()
    check(getattr(lib, name)(*args), name)


def stream_ptr():
    """cudaStream_t of torch's current stream (so torch.cuda.graph capture sees our launches)."""
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def dtype_code(torch_dtype):
    import torch
    if torch_dtype == torch.float32:
        return BMI_F32
    if torch_dtype == torch.float64:
        return BMI_F64
    raise BmiError("unsupported storage dtype %s (float32/float64 only)" % torch_dtype)


def launch_count():
    return int(load().bmi_launch_count())
