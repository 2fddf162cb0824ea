"""Cross-rank plumbing of the data-parallel path: one process per GPU.

Mirror of the reference ``utils.py:6-69`` (``sync_networks``, ``sync_grads``) with mpi4py
replaced by NCCL over NVLink (``bmi_comm_*`` in csrc/comm.cu).  Semantics kept from the
reference: parameters are broadcast from rank 0 (utils.py:13), gradients are SUMMED, not
averaged (utils.py:47).  With world size 1 every collective is the identity.  On CPU (unit
tests, gloo) the same entry points go through ``torch.distributed``.
"""
import ctypes
import os

import torch

from . import _lib

_state = {"comm": None, "rank": 0, "world": 1}


def rank():
    return _state["rank"]


def world_size():
    return _state["world"]


def init_comm(backend=None):
    """Join the job described by RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun).  Creates
    the torch.distributed group (bootstrap + CPU tensors) and, on GPUs, the NCCL communicator
    used by the CUDA-graph-captured collectives."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rk = int(os.environ.get("RANK", "0"))
    if world <= 1:
        _state.update(rank=0, world=1, comm=None)
        return 0, 1
    import torch.distributed as dist
    use_cuda = torch.cuda.is_available()
    if use_cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rk)) % torch.cuda.device_count())
    if not dist.is_initialized():
        dist.init_process_group(backend=backend or ("nccl" if use_cuda else "gloo"), rank=rk, world_size=world)
    _state.update(rank=rk, world=world)
    if use_cuda:
        buf = (ctypes.c_uint8 * 128)()
        if rk == 0:
            _lib.call("bmi_comm_unique_id", ctypes.cast(buf, ctypes.c_void_p))
        obj = [bytes(buf)]
        dist.broadcast_object_list(obj, src=0)
        idbuf = (ctypes.c_uint8 * 128).from_buffer_copy(obj[0])
        h = ctypes.c_void_p()
        _lib.call("bmi_comm_init", ctypes.byref(h), rk, world, ctypes.cast(idbuf, ctypes.c_void_p))
        _state["comm"] = h
    return rk, world


def shutdown_comm():
    """Tear the communicators down.  CUDA graphs that captured NCCL operations keep the communicator busy:
    release them first (``ddpg_agent.release_graphs()``), otherwise ncclCommDestroy waits for ever."""
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    if _state["comm"] is not None:
        _lib.call("bmi_comm_destroy", _state["comm"])
        _state["comm"] = None
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()
    _state.update(rank=0, world=1)


def allreduce_sum_(t):
    """In-place SUM over ranks of a contiguous float32 tensor."""
    if _state["world"] == 1:
        return t
    if t.is_cuda and _state["comm"] is not None:
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise _lib.BmiError("allreduce_sum_: contiguous float32 CUDA tensor required")
        _lib.call("bmi_comm_allreduce_sum_f32", _state["comm"], _lib.ptr(t), t.numel(), _lib.stream_ptr())
        return t
    import torch.distributed as dist
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def bcast_(t, root=0):
    if _state["world"] == 1:
        return t
    if t.is_cuda and _state["comm"] is not None:
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise _lib.BmiError("bcast_: contiguous float32 CUDA tensor required")
        _lib.call("bmi_comm_bcast_f32", _state["comm"], _lib.ptr(t), t.numel(), int(root), _lib.stream_ptr())
        return t
    import torch.distributed as dist
    dist.broadcast(t, src=root)
    return t


def _flat_of(network, attr):
    flat = getattr(network, attr, None)
    if flat is None:
        raise _lib.BmiError("network has no %s buffer; build it with rl_arm_under_sparse_reward_b200.models" % attr)
    return flat


def sync_networks(network):
    """utils.py:6-15: broadcast rank 0's parameters (one collective on the flat buffer)."""
    bcast_(_flat_of(network, "flat"), root=0)


def sync_grads(network):
    """utils.py:43-48: SUM the flat gradients over ranks (no divide)."""
    allreduce_sum_(_flat_of(network, "flat_grad"))
