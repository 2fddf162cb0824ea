"""Device-resident vectorised bmirobot env: N instances stepped by one kernel launch.

The per-instance semantics are the reference's ``bmirobotGymEnv`` (bmirobot_env_push_F.py:25-245,
pick deltas bmirobot_env_pickandplace_v2.py:92-95,116-131); all arrays stay on the GPU.
"""
import ctypes
import os

import numpy as np
import torch

from .. import _lib

MODEL_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets", "bmirobot_model.bin")
SELFCOL_PATH = os.path.join(os.path.dirname(MODEL_PATH), "bmirobot_selfcol.bin")
_SELFCOL_CACHE = {}


def _selfcol_table(path):
    """The baked self-collision pair tables (48 MB, git-ignored, written by __graft_entry__.build() /
    tools/bake_selfcol.py); one host copy per process."""
    if path not in _SELFCOL_CACHE:
        if not os.path.exists(path):
            raise _lib.BmiError("self-collision pair tables missing: %s (run `python tools/bake_selfcol.py` or "
                                "__graft_entry__.build())" % path)
        _SELFCOL_CACHE[path] = np.fromfile(path, dtype="<f4")
    return _SELFCOL_CACHE[path]


class BmiVecEnv:
    obs_dim, goal_dim, act_dim = 27, 3, 4
    action_max = 0.5
    distance_threshold = 0.05
    n_substeps = 20

    def __init__(self, n_envs, task="push", seed=125, device=None, model_path=MODEL_PATH, selfcol_path=SELFCOL_PATH):
        self.n_envs = int(n_envs)
        self.task = {"push": _lib.TASK_PUSH, "pick": _lib.TASK_PICK}[task]
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        blob = np.fromfile(model_path, dtype="<f4")
        self._blob = blob
        h = ctypes.c_void_p()
        _lib.call("bmi_env_create", ctypes.byref(h), self.n_envs, self.task, blob.ctypes.data_as(ctypes.c_void_p),
                  int(blob.nbytes))
        self._h = h
        if blob[47] > 0.5:   # MP_SELF_COLLISION (bmirobot.py:58 flags=9)
            t = _selfcol_table(selfcol_path)
            _lib.call("bmi_env_set_selfcol", h, t.ctypes.data_as(ctypes.c_void_p), int(t.nbytes))
        f = lambda *s: torch.zeros(s, dtype=torch.float32, device=self.device)
        self.obs, self.ag, self.g = f(n_envs, 27), f(n_envs, 3), f(n_envs, 3)
        self.reward, self.success = f(n_envs), f(n_envs)
        self.init = f(n_envs, 8)
        self.seed_value = int(seed)
        self.counter = torch.zeros(1, dtype=torch.int64, device=self.device)  # Philox counter (uint64 bits)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.call("bmi_env_destroy", h)
            except Exception:
                pass
            self._h = None

    def seed(self, s):
        self.seed_value = int(s)
        self.counter.zero_()

    def reset(self, init=None, mask=None):
        """init: optional (n,8) float32 tensor [block x,y,z,yaw, goal x,y,z, 0]; drawn on the device with
        the reference's ranges and rejection rule when omitted.  Returns (obs, ag, g) views."""
        if init is None:
            _lib.call("bmi_env_sample_init", self._h, ctypes.c_uint64(self.seed_value), _lib.ptr(self.counter),
                      _lib.ptr(self.init), _lib.stream_ptr())
        else:
            self.init.copy_(torch.as_tensor(init, dtype=torch.float32).reshape(self.n_envs, 8))
        m = None if mask is None else torch.as_tensor(mask, dtype=torch.uint8, device=self.device).contiguous()
        _lib.call("bmi_env_reset", self._h, _lib.ptr(m), _lib.ptr(self.init), _lib.ptr(self.obs), _lib.ptr(self.ag),
                  _lib.ptr(self.g), _lib.stream_ptr())
        return self.obs, self.ag, self.g

    def step(self, actions):
        """actions: (n,4) float32 CUDA tensor.  Returns (obs, ag, reward, success) views (overwritten by
        the next call)."""
        if actions.dtype != torch.float32 or not actions.is_cuda or not actions.is_contiguous():
            actions = actions.to(self.device, torch.float32).contiguous()
        _lib.call("bmi_env_step", self._h, _lib.ptr(actions), _lib.ptr(self.obs), _lib.ptr(self.ag),
                  _lib.ptr(self.reward), _lib.ptr(self.success), _lib.stream_ptr())
        return self.obs, self.ag, self.reward, self.success

    def rollout(self, T, actor_t, o_norm, g_norm, clip_range, explore, noise_eps=0.0, random_eps=0.0, late_clip=0.0,
                seed=0, counter=None, episodes=None, reset=True):
        """Fused rollout: ONE kernel launch runs T policy + env steps for every env (bmi_env_rollout).
        actor_t: transposed flat actor parameters (bmi_actor_transpose); o_norm/g_norm: normalizer objects;
        episodes: dict of float32 staging tensors obs/ag/g/actions or None.  Returns (obs, ag, g, success)."""
        if reset:
            _lib.call("bmi_env_sample_init", self._h, ctypes.c_uint64(self.seed_value), _lib.ptr(self.counter),
                      _lib.ptr(self.init), _lib.stream_ptr())
        eps = None
        if episodes is not None:
            eps = _lib.Episodes(_lib.ptr(episodes['obs']), _lib.ptr(episodes['ag']), _lib.ptr(episodes['g']),
                                _lib.ptr(episodes['actions']), self.n_envs, int(T), 27, 3, 4, _lib.BMI_F32, 0)
        ra = _lib.RolloutArgs(int(T), int(bool(explore)), _lib.ptr(actor_t), _lib.ptr(o_norm.mean_dev), _lib.ptr(o_norm.std_dev),
                              _lib.ptr(g_norm.mean_dev), _lib.ptr(g_norm.std_dev), float(clip_range), float(self.action_max),
                              float(noise_eps), float(random_eps), float(late_clip), int(seed), _lib.ptr(counter),
                              ctypes.pointer(eps) if eps is not None else None, _lib.ptr(self.init) if reset else None,
                              _lib.ptr(self.obs), _lib.ptr(self.ag), _lib.ptr(self.g), _lib.ptr(self.success))
        _lib.call("bmi_env_rollout", self._h, ctypes.byref(ra), _lib.stream_ptr())
        return self.obs, self.ag, self.g, self.success

    def contact_drops(self, reset=False):
        """contacts dropped by the kernel's lane budget (9 contacts, 6 on arm links) since the counter was last reset"""
        out = ctypes.c_uint64(0)
        _lib.call("bmi_env_contact_drops", self._h, ctypes.byref(out), int(bool(reset)))
        return int(out.value)

    def get_state(self):
        st = torch.empty((self.n_envs, _lib.ENV_STATE_DIM), dtype=torch.float32, device=self.device)
        _lib.call("bmi_env_get_state", self._h, _lib.ptr(st), _lib.stream_ptr())
        return st

    def set_state(self, st):
        st = torch.as_tensor(st, dtype=torch.float32).to(self.device).contiguous()
        _lib.call("bmi_env_set_state", self._h, _lib.ptr(st), _lib.stream_ptr())

    def compute_reward(self, achieved_goal, goal, info=None):
        """bmirobot_env_push_F.py:84-90 for CUDA tensors or numpy arrays (same type out)."""
        host = isinstance(achieved_goal, np.ndarray)
        a = torch.as_tensor(np.ascontiguousarray(achieved_goal)).to(self.device) if host else achieved_goal.contiguous()
        b = torch.as_tensor(np.ascontiguousarray(goal)).to(self.device) if host else goal.contiguous()
        if a.shape != b.shape:
            raise AssertionError("goal shapes differ")  # goal_distance's assert (push_F.py:21)
        if a.dtype != b.dtype or a.dtype not in (torch.float32, torch.float64):
            a, b = a.to(torch.float64), b.to(torch.float64)
        out = torch.empty(a.shape[:-1], dtype=torch.float32, device=self.device)
        n = out.numel()
        _lib.call("bmi_compute_reward", _lib.ptr(a), _lib.ptr(b), n, int(a.shape[-1]), _lib.dtype_code(a.dtype),
                  float(self.distance_threshold), _lib.ptr(out), _lib.stream_ptr())
        return out.cpu().numpy() if host else out
