"""HER 'future' sampler backed by the sm_100a gather/relabel/reward kernel.

Mirror of the reference ``her.py:3-41`` (class name, constructor arguments, ``future_p`` and
``sample_her_transitions(episode_batch, batch_size_in_transitions)``).  The four random arrays
are drawn on the host from numpy's global legacy stream in the reference order
(``randint, randint, uniform, uniform`` — her.py:24-31) so a run seeded like the reference
consumes the stream identically; the gather, the relabelling, the index arithmetic and the
sparse reward run on the GPU (``bmi_her_sample``), bit-exact in float64.
"""
import ctypes

import numpy as np
import torch

from . import _lib

_KEYS = ("obs", "ag", "g", "actions")


def _is_shifted_alias(base, nxt):
    """True when ``nxt`` is ``base[:, 1:, :]`` (the only form the reference ever passes)."""
    if isinstance(base, np.ndarray):
        return (isinstance(nxt, np.ndarray) and nxt.shape == (base.shape[0], base.shape[1] - 1, base.shape[2])
                and nxt.strides == base.strides
                and nxt.__array_interface__["data"][0] == base.__array_interface__["data"][0] + base.strides[1])
    return (torch.is_tensor(nxt) and tuple(nxt.shape) == (base.shape[0], base.shape[1] - 1, base.shape[2])
            and nxt.stride() == base.stride()
            and nxt.data_ptr() == base.data_ptr() + base.stride(1) * base.element_size())


class her_sampler:
    def __init__(self, replay_strategy, replay_k, reward_func=None, distance_threshold=None):
        self.replay_strategy = replay_strategy
        self.replay_k = replay_k
        if self.replay_strategy == 'future':
            self.future_p = 1 - (1. / (1 + replay_k))
        else:
            self.future_p = 0
        self.reward_func = reward_func
        # the kernel evaluates the sparse reward of bmirobot_env_push_F.py:84-90 itself; the
        # threshold comes from the env that owns reward_func (env.compute_reward) when bound
        if distance_threshold is None:
            owner = getattr(reward_func, "__self__", None)
            distance_threshold = getattr(owner, "distance_threshold", 0.05)
        self.distance_threshold = float(distance_threshold)

    # ---- host-side draws (reference order on numpy's global stream) ----------------------
    @staticmethod
    def draw(rollout_batch_size, T, batch_size):
        episode_idxs = np.random.randint(0, rollout_batch_size, batch_size)
        t_samples = np.random.randint(T, size=batch_size)
        u_her = np.random.uniform(size=batch_size)
        u_off = np.random.uniform(size=batch_size)
        return (episode_idxs.astype(np.int64), t_samples.astype(np.int64), u_her, u_off)

    # ---- device path -----------------------------------------------------------------------
    def sample_device(self, obs, ag, g, actions, n_valid, draws, out=None):
        """obs/ag/g/actions: contiguous CUDA tensors [E,T+1,Do],[E,T+1,Dg],[E,T,Dg],[E,T,Da]
        (float32 or float64).  draws: 4 host numpy arrays or 4 CUDA tensors.  Returns a dict of
        CUDA tensors with the reference keys plus 'r' of shape (B,1) float32."""
        dev = obs.device
        E, Tp1, Do = obs.shape
        T, Dg, Da = Tp1 - 1, ag.shape[2], actions.shape[2]
        for name, t_ in (("obs", obs), ("ag", ag), ("g", g), ("actions", actions)):
            if not t_.is_contiguous() or t_.dtype != obs.dtype or not t_.is_cuda:
                raise _lib.BmiError("her_sampler.sample_device: %s must be a contiguous CUDA tensor of dtype %s" % (name, obs.dtype))
        if torch.is_tensor(draws[0]):
            ep_d, t_d, uh_d, uo_d = draws
        else:
            ep_d = torch.as_tensor(np.ascontiguousarray(draws[0], dtype=np.int64)).to(dev, non_blocking=True)
            t_d = torch.as_tensor(np.ascontiguousarray(draws[1], dtype=np.int64)).to(dev, non_blocking=True)
            uh_d = torch.as_tensor(np.ascontiguousarray(draws[2], dtype=np.float64)).to(dev, non_blocking=True)
            uo_d = torch.as_tensor(np.ascontiguousarray(draws[3], dtype=np.float64)).to(dev, non_blocking=True)
        B = int(ep_d.shape[0])
        if out is None:
            dt = obs.dtype
            out = {"obs": torch.empty((B, Do), dtype=dt, device=dev), "ag": torch.empty((B, Dg), dtype=dt, device=dev),
                   "g": torch.empty((B, Dg), dtype=dt, device=dev), "actions": torch.empty((B, Da), dtype=dt, device=dev),
                   "obs_next": torch.empty((B, Do), dtype=dt, device=dev),
                   "ag_next": torch.empty((B, Dg), dtype=dt, device=dev),
                   "r": torch.empty((B, 1), dtype=torch.float32, device=dev)}
        eps = _lib.Episodes(_lib.ptr(obs), _lib.ptr(ag), _lib.ptr(g), _lib.ptr(actions), E, T, Do, Dg, Da,
                            _lib.dtype_code(obs.dtype), 0)
        tr = _lib.Transitions(_lib.ptr(out.get("obs")), _lib.ptr(out.get("ag")), _lib.ptr(out.get("g")),
                              _lib.ptr(out.get("actions")), _lib.ptr(out.get("obs_next")),
                              _lib.ptr(out.get("ag_next")), _lib.ptr(out.get("r")))
        _lib.call("bmi_her_sample", ctypes.byref(eps), int(n_valid), _lib.ptr(ep_d), _lib.ptr(t_d), _lib.ptr(uh_d),
                  _lib.ptr(uo_d), B, float(self.future_p), self.distance_threshold, ctypes.byref(tr), _lib.stream_ptr())
        return out

    # ---- reference entry point ---------------------------------------------------------------
    def sample_her_transitions(self, episode_batch, batch_size_in_transitions):
        """her.py:13-41.  ``episode_batch`` is a dict of numpy arrays (host, any float dtype —
        uploaded, sampled on the GPU, returned as numpy like the reference) or of CUDA tensors
        (returned as CUDA tensors).  'obs_next'/'ag_next', when present, must be the shifted
        views of 'obs'/'ag' that every reference call site passes (replay_buffer.py:51-52,
        ddpg_agent.py:191-192)."""
        for k in _KEYS:
            if k not in episode_batch:
                raise KeyError("episode_batch is missing key %r" % k)
        for k, base in (("obs_next", "obs"), ("ag_next", "ag")):
            if k in episode_batch and not _is_shifted_alias(episode_batch[base], episode_batch[k]):
                raise _lib.BmiError("%s must be the [:, 1:, :] view of %s" % (k, base))
        extra = set(episode_batch.keys()) - set(_KEYS) - {"obs_next", "ag_next"}
        if extra:
            raise _lib.BmiError("unsupported episode_batch keys: %s" % sorted(extra))
        T = episode_batch['actions'].shape[1]
        rollout_batch_size = episode_batch['actions'].shape[0]
        batch_size = int(batch_size_in_transitions)
        if rollout_batch_size == 0:
            raise ValueError("cannot sample from an empty episode batch")  # numpy randint raises too
        draws = self.draw(rollout_batch_size, T, batch_size)
        on_host = isinstance(episode_batch['obs'], np.ndarray)
        if on_host:
            dev = torch.device("cuda", torch.cuda.current_device())
            arrs = [torch.as_tensor(np.ascontiguousarray(episode_batch[k], dtype=np.float64)).to(dev) for k in _KEYS]
        else:
            arrs = [episode_batch[k].contiguous() for k in _KEYS]
        out = self.sample_device(arrs[0], arrs[1], arrs[2], arrs[3], rollout_batch_size, draws)
        if on_host:
            return {k: v.cpu().numpy() for k, v in out.items()}
        return out
