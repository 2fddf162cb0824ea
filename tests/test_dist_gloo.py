"""CPU-only, world_size 2 over gloo: the cross-rank semantics of the data-parallel path
(gradients SUMMED not averaged — utils.py:43-48; normaliser sums AVERAGED — normalizer.py:60-64;
parameters broadcast from rank 0 — utils.py:6-15)."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from rl_arm_under_sparse_reward_b200 import utils
    r, w = utils.init_comm(backend="gloo")
    assert (r, w) == (rank, world) and utils.world_size() == world and utils.rank() == rank
    g = torch.full((10,), float(rank + 1))
    utils.allreduce_sum_(g)                      # SUM, no divide
    p = torch.full((5,), float(rank + 7))
    utils.bcast_(p, root=0)

    class Net:
        flat = torch.full((4,), float(rank))
        flat_grad = torch.full((4,), 0.5 * (rank + 1))
    utils.sync_networks(Net)
    utils.sync_grads(Net)
    # normaliser averaging: local sums differ per rank
    local = torch.tensor([1.0 + rank, 10.0 * (rank + 1), 100.0])
    utils.allreduce_sum_(local)
    local /= utils.world_size()
    q.put((rank, g.numpy().copy(), p.numpy().copy(), Net.flat.numpy().copy(), Net.flat_grad.numpy().copy(), local.numpy().copy()))
    utils.shutdown_comm()


def test_two_rank_collectives():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, g, p, flat, fgrad, local in res:
        assert np.array_equal(g, np.full(10, 3.0))          # 1 + 2
        assert np.array_equal(p, np.full(5, 7.0))           # rank 0's value
        assert np.array_equal(flat, np.zeros(4))            # rank 0's parameters
        assert np.array_equal(fgrad, np.full(4, 1.5))       # 0.5 + 1.0, summed
        assert np.allclose(local, [1.5, 15.0, 100.0])       # averaged
