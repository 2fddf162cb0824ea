"""Running mean/std normaliser with device-resident statistics.

Mirror of the reference ``normalizer.py:5-70``: ``normalizer(size, eps, default_clip_range)``,
``update``, ``recompute_stats``, ``normalize``, ``sync`` and the attributes the checkpoint
writer reads (``mean``, ``std`` — ddpg_agent.py:158).  The float32 accumulators live on the GPU;
``update``/``recompute_stats``/``normalize`` are CUDA kernels (csrc/normalizer.cu) that reproduce
numpy's float64-sum / float32-accumulate arithmetic bit for bit; the cross-rank average of
``_mpi_average`` (normalizer.py:60-64) is an NCCL sum followed by the divide.
"""
import threading

import numpy as np
import torch

from . import _lib
from . import utils as _utils


class normalizer:
    def __init__(self, size, eps=1e-2, default_clip_range=np.inf, device=None):
        self.size = size
        self.eps = eps
        self.default_clip_range = default_clip_range
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        z = lambda n: torch.zeros(n, dtype=torch.float32, device=self.device)
        # one flat allocation so the cross-rank sum of (sum, sumsq, count) is ONE collective
        self._local = z(2 * size + 1)
        self.local_sum_dev = self._local[:size]
        self.local_sumsq_dev = self._local[size:2 * size]
        self.local_count_dev = self._local[2 * size:]
        self.total_sum_dev = z(size)
        self.total_sumsq_dev = z(size)
        self.total_count_dev = torch.ones(1, dtype=torch.float32, device=self.device)
        self.mean_dev = z(size)
        self.std_dev = torch.ones(size, dtype=torch.float32, device=self.device)
        self.lock = threading.Lock()

    # numpy views of the device state, for reference-compatible attribute access
    def _np(self, t):
        return t.detach().cpu().numpy().copy()

    mean = property(lambda self: self._np(self.mean_dev))
    std = property(lambda self: self._np(self.std_dev))
    local_sum = property(lambda self: self._np(self.local_sum_dev))
    local_sumsq = property(lambda self: self._np(self.local_sumsq_dev))
    local_count = property(lambda self: self._np(self.local_count_dev))
    total_sum = property(lambda self: self._np(self.total_sum_dev))
    total_sumsq = property(lambda self: self._np(self.total_sumsq_dev))
    total_count = property(lambda self: self._np(self.total_count_dev))

    @mean.setter
    def mean(self, v):  # ddpg_agent.py:55-62 (resume) assigns these
        self.mean_dev.copy_(torch.as_tensor(np.asarray(v, dtype=np.float32)))

    @std.setter
    def std(self, v):
        self.std_dev.copy_(torch.as_tensor(np.asarray(v, dtype=np.float32)))

    def _as_dev(self, v):
        if torch.is_tensor(v):
            t = v.to(self.device)
            if t.dtype not in (torch.float32, torch.float64):
                t = t.to(torch.float64)
            return t.contiguous(), False
        a = np.asarray(v)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        return torch.as_tensor(np.ascontiguousarray(a)).to(self.device), True

    def update(self, v, pre_clip=0.0):
        """normalizer.py:25-31.  pre_clip > 0 fuses ddpg_agent._preproc_og's clip into the kernel."""
        t, _ = self._as_dev(v)
        t = t.reshape(-1, self.size)
        with self.lock:
            _lib.call("bmi_norm_update", _lib.ptr(t), int(t.shape[0]), int(self.size), _lib.dtype_code(t.dtype),
                      float(pre_clip), _lib.ptr(self.local_sum_dev), _lib.ptr(self.local_sumsq_dev), _lib.ptr(self.local_count_dev),
                      _lib.stream_ptr())

    def sync(self, local_sum, local_sumsq, local_count):
        """normalizer.py:34-38 for host arrays (kept for API parity)."""
        local_sum[...] = self._mpi_average(local_sum)
        local_sumsq[...] = self._mpi_average(local_sumsq)
        local_count[...] = self._mpi_average(local_count)
        return local_sum, local_sumsq, local_count

    def _mpi_average(self, x):
        t = torch.as_tensor(np.asarray(x, dtype=np.float32)).to(self.device)
        _utils.allreduce_sum_(t)
        t /= _utils.world_size()
        return t.cpu().numpy()

    def recompute_stats(self):
        """normalizer.py:40-57: one fused collective instead of three, then one kernel."""
        with self.lock:
            _utils.allreduce_sum_(self._local)
            _lib.call("bmi_norm_recompute", _lib.ptr(self.local_sum_dev), _lib.ptr(self.local_sumsq_dev),
                      _lib.ptr(self.local_count_dev), _lib.ptr(self.total_sum_dev), _lib.ptr(self.total_sumsq_dev),
                      _lib.ptr(self.total_count_dev), _lib.ptr(self.mean_dev), _lib.ptr(self.std_dev), int(self.size),
                      float(self.eps), float(_utils.world_size()), _lib.stream_ptr())

    def normalize(self, v, clip_range=None):
        """normalizer.py:67-70.  numpy in -> float64 numpy out (reference); CUDA tensor in ->
        CUDA tensor out (float64)."""
        if clip_range is None:
            clip_range = self.default_clip_range
        t, was_host = self._as_dev(v)
        shape = t.shape
        t2 = t.reshape(-1, self.size)
        out = torch.empty(t2.shape, dtype=torch.float64, device=self.device)
        _lib.call("bmi_norm_normalize", _lib.ptr(t2), int(t2.shape[0]), int(self.size), _lib.dtype_code(t2.dtype),
                  _lib.ptr(self.mean_dev), _lib.ptr(self.std_dev), float(clip_range), _lib.ptr(out), _lib.BMI_F64,
                  _lib.stream_ptr())
        out = out.reshape(shape)
        return out.cpu().numpy() if was_host else out
