"""HBM-resident episode replay buffer.

Mirror of the reference ``replay_buffer.py:10-71``: same constructor, ``store_episode``,
``sample``, ``_get_storage_idx`` and attributes (``T size current_size n_transitions_stored
buffers lock``).  ``buffers`` holds CUDA tensors in the reference's struct-of-arrays,
episode-major layout (``obs[size,T+1,obs] ag[size,T+1,goal] g[size,T,goal]
actions[size,T,action]``); slot selection (append, then RANDOM overwrite once full —
replay_buffer.py:57-71) is the reference's host logic on numpy's global stream, the copies
and the sampling are CUDA kernels (``bmi_buffer_store`` / ``bmi_her_sample``).
"""
import ctypes
import threading

import numpy as np
import torch

from . import _lib
from .her import her_sampler


class replay_buffer:
    def __init__(self, env_params, buffer_size, sample_func, dtype=torch.float64, device=None, verbose=True):
        self.env_params = env_params
        self.T = env_params['max_timesteps']
        self.size = int(buffer_size // self.T)
        self.current_size = 0
        self.n_transitions_stored = 0
        self.sample_func = sample_func
        self.dtype = dtype
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if verbose:
            print("Buffer_size:", self.size, "max_timesteps:", self.T, "env_params:", self.env_params)
        Do, Dg, Da = env_params['obs'], env_params['goal'], env_params['action']
        mk = lambda *shape: torch.empty(shape, dtype=dtype, device=self.device)
        self.buffers = {'obs': mk(self.size, self.T + 1, Do),
                        'ag': mk(self.size, self.T + 1, Dg),
                        'g': mk(self.size, self.T, Dg),
                        'actions': mk(self.size, self.T, Da)}
        # device copy of current_size for graph-captured samplers (bmi_her_draw reads it)
        self.current_size_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.lock = threading.Lock()

    # ---- helpers -----------------------------------------------------------------------------
    def _episodes_struct(self, tensors, n):
        obs, ag, g, act = tensors
        return _lib.Episodes(_lib.ptr(obs), _lib.ptr(ag), _lib.ptr(g), _lib.ptr(act), int(n), int(self.T),
                             int(obs.shape[2]), int(ag.shape[2]), int(act.shape[2]), _lib.dtype_code(obs.dtype), 0)

    def _to_device(self, a):
        if torch.is_tensor(a):
            t = a.to(self.device)
            if t.dtype not in (torch.float32, torch.float64):
                t = t.to(torch.float64)
            return t.contiguous()
        a = np.asarray(a)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        return torch.as_tensor(np.ascontiguousarray(a)).to(self.device)

    # ---- reference API -----------------------------------------------------------------------
    def store_episode(self, episode_batch):
        """replay_buffer.py:32-43.  Accepts numpy arrays (reference) or torch tensors (host or
        device) shaped (R,T+1,obs) (R,T+1,goal) (R,T,goal) (R,T,action)."""
        mb_obs, mb_ag, mb_g, mb_actions = episode_batch
        batch_size = mb_obs.shape[0]
        exp = [(batch_size, self.T + 1, self.env_params['obs']), (batch_size, self.T + 1, self.env_params['goal']),
               (batch_size, self.T, self.env_params['goal']), (batch_size, self.T, self.env_params['action'])]
        for a, e, name in zip(episode_batch, exp, ("obs", "ag", "g", "actions")):
            if tuple(a.shape) != e:
                raise ValueError("store_episode: %s has shape %s, expected %s" % (name, tuple(a.shape), e))
        if batch_size == 0:
            return
        with self.lock:
            idxs = self._get_storage_idx(inc=batch_size)
            idxs = np.atleast_1d(np.asarray(idxs, dtype=np.int64)).copy()
            # numpy fancy assignment with repeated indices keeps the LAST write; mark the
            # earlier duplicates as skipped so the parallel copy is deterministic
            _, last_pos = np.unique(idxs[::-1], return_index=True)
            keep = np.zeros(idxs.shape[0], dtype=bool)
            keep[idxs.shape[0] - 1 - last_pos] = True
            idxs[~keep] = -1
            src = [self._to_device(a) for a in episode_batch]
            if len({t.dtype for t in src}) != 1:
                src = [t.to(torch.float64) for t in src]
            slots = torch.as_tensor(idxs).to(self.device)
            s = self._episodes_struct(src, batch_size)
            d = self._episodes_struct([self.buffers[k] for k in ('obs', 'ag', 'g', 'actions')], self.size)
            _lib.call("bmi_buffer_store", ctypes.byref(d), ctypes.byref(s), _lib.ptr(slots), _lib.stream_ptr())
            self.current_size_dev.fill_(self.current_size)
            self.n_transitions_stored += self.T * batch_size

    def sample_device(self, batch_size):
        """sample() without the device->host copy: dict of CUDA tensors."""
        owner = getattr(self.sample_func, "__self__", None)
        with self.lock:
            n_valid = self.current_size
        if isinstance(owner, her_sampler):
            if n_valid == 0:
                raise ValueError("cannot sample from an empty replay buffer")
            draws = owner.draw(n_valid, self.T, int(batch_size))
            b = self.buffers
            return owner.sample_device(b['obs'], b['ag'], b['g'], b['actions'], n_valid, draws)
        # arbitrary sample_func: hand it the reference-shaped dict of (device) views
        temp = {k: v[:n_valid] for k, v in self.buffers.items()}
        temp['obs_next'] = temp['obs'][:, 1:, :]
        temp['ag_next'] = temp['ag'][:, 1:, :]
        return self.sample_func(temp, batch_size)

    def sample(self, batch_size):
        """replay_buffer.py:46-55: returns a dict of numpy arrays like the reference."""
        out = self.sample_device(batch_size)
        return {k: (v.cpu().numpy() if torch.is_tensor(v) else v) for k, v in out.items()}

    def _get_storage_idx(self, inc=None):
        """Slot choice of replay_buffer.py:57-71 (host logic on numpy's global stream): fill
        the free tail first, then overwrite uniformly-random occupied slots."""
        inc = inc or 1
        free = self.size - self.current_size
        if inc <= free:
            idx = np.arange(self.current_size, self.current_size + inc)
        elif free > 0:
            tail = np.arange(self.current_size, self.size)
            idx = np.concatenate([tail, np.random.randint(0, self.current_size, inc - free)])
        else:
            idx = np.random.randint(0, self.size, inc)
        self.current_size = min(self.size, self.current_size + inc)
        return idx[0] if inc == 1 else idx
