"""GPU parity of the DDPG learner (cuBLASLt GEMMs + fused kernels) against the torch-CPU fp32 oracle and
against the reference agent's own _update_network (golden).  Floating-point path: tolerances are stated
per assertion (losses rtol 1e-4, parameters after updates rtol 1e-3 as in BASELINE.md section 4)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import ddpg_oracle as do
from oracle import learner_oracle as lo

pytestmark = pytest.mark.gpu
PARAMS = {'obs': 27, 'goal': 3, 'action': 4, 'action_max': 0.5, 'max_timesteps': 100}


def _seeded(seed):
    L = do.Learner()
    wr = np.random.RandomState(seed)
    for net in (L.actor, L.critic):
        for _, p in net.named_parameters():
            bound = 1.0 / np.sqrt(p.shape[-1] if p.dim() > 1 else 256)
            p.data.copy_(torch.tensor(wr.uniform(-bound, bound, tuple(p.shape)).astype(np.float32)))
    L.actor_t.load_state_dict(L.actor.state_dict())
    L.critic_t.load_state_dict(L.critic.state_dict())
    return L


class Trainer:
    """thin test harness around the C-ABI learner"""

    def __init__(self, L, batch=256, max_rows=512, dims=(27, 3, 4), amax=0.5, gamma=0.98, l2=1.0):
        from rl_arm_under_sparse_reward_b200 import _lib
        self._lib = _lib
        dev = torch.device("cuda")
        self.pa = do.flat_params(L.actor).to(dev).contiguous()
        self.pc = do.flat_params(L.critic).to(dev).contiguous()
        self.ta = do.flat_params(L.actor_t).to(dev).contiguous()
        self.tc = do.flat_params(L.critic_t).to(dev).contiguous()
        self.cfg = _lib.DdpgConfig(dims[0], dims[1], dims[2], 256, batch, max_rows, amax, gamma, l2, 1e-3, 1e-3, 0.95, 0.9, 0.999,
                                   1e-8, float(1.0 / (1.0 - gamma)), float(1.0 - 0.95), 0.0)
        self.h = ctypes.c_void_p()
        _lib.call("bmi_ddpg_create", ctypes.byref(self.h), ctypes.byref(self.cfg), _lib.ptr(self.pa), _lib.ptr(self.pc),
                  _lib.ptr(self.ta), _lib.ptr(self.tc))
        self.na = int(_lib.load().bmi_ddpg_actor_param_count(ctypes.byref(self.cfg)))
        self.ncr = int(_lib.load().bmi_ddpg_critic_param_count(ctypes.byref(self.cfg)))
        self.losses = torch.zeros(2, dtype=torch.float32, device=dev)

    def backward(self, x, xn, a, r):
        _lib = self._lib
        t = [torch.as_tensor(v).cuda().contiguous() for v in (x, xn, a, r.reshape(-1))]
        _lib.call("bmi_ddpg_backward", self.h, _lib.ptr(t[0]), _lib.ptr(t[1]), _lib.ptr(t[2]), _lib.ptr(t[3]),
                  _lib.ptr(self.losses), _lib.stream_ptr())
        torch.cuda.synchronize()

    def grads(self):
        gp, gn = ctypes.c_void_p(), ctypes.c_int64()
        self._lib.call("bmi_ddpg_grad_buffer", self.h, ctypes.byref(gp), ctypes.byref(gn))
        n = int(gn.value)
        out = torch.empty(n, dtype=torch.float32, device="cuda")
        ctypes.cdll.LoadLibrary("libcudart.so.12") if False else None
        # copy through torch: wrap the raw pointer with the CUDA array interface
        class _W:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (gp.value, False), "version": 2}
        g = torch.as_tensor(_W(), device="cuda").clone()
        na_pad = (self.na + 63) // 64 * 64
        return g[:self.na].cpu().numpy(), g[na_pad:na_pad + self.ncr].cpu().numpy(), g[self.na:na_pad].cpu().numpy()

    def adam(self):
        self._lib.call("bmi_ddpg_adam_step", self.h, self._lib.stream_ptr())

    def soft(self):
        self._lib.call("bmi_ddpg_soft_update", self.h, self._lib.stream_ptr())

    def close(self):
        self._lib.call("bmi_ddpg_destroy", self.h)


def _batch(seed, B=256):
    rng = np.random.RandomState(seed)
    x = np.clip(rng.standard_normal((B, 30)), -5, 5).astype(np.float32)
    xn = np.clip(x + 0.1 * rng.standard_normal((B, 30)), -5, 5).astype(np.float32)
    a = rng.uniform(-0.5, 0.5, (B, 4)).astype(np.float32)
    r = -(rng.uniform(size=(B, 1)) > 0.2).astype(np.float32)
    return x, xn, a, r


def test_parameter_counts_match_reference_nets():
    L = do.Learner()
    T = Trainer(L)
    assert T.na == 140548 and T.ncr == 140801
    assert T.na == do.flat_params(L.actor).numel() and T.ncr == do.flat_params(L.critic).numel()
    T.close()


def test_losses_and_gradients_vs_torch_oracle():
    torch.set_num_threads(1)
    L = _seeded(21)
    T = Trainer(L)
    x, xn, a, r = _batch(0)
    la, lc, ga, gc = L.losses_and_grads(torch.tensor(x), torch.tensor(xn), torch.tensor(a), torch.tensor(r))
    T.backward(x, xn, a, r)
    los = T.losses.cpu().numpy()
    assert np.allclose(los, [la, lc], rtol=1e-4, atol=1e-6), (los, la, lc)
    mya, myc, pad = T.grads()
    assert np.all(pad == 0)
    for mine, ref in ((mya, ga.numpy()), (myc, gc.numpy())):
        scale = np.abs(ref).max()
        assert np.abs(mine - ref).max() <= 2e-4 * scale + 1e-8, np.abs(mine - ref).max() / scale
        # relative error of the whole vector
        assert np.linalg.norm(mine - ref) <= 1e-4 * np.linalg.norm(ref)
    T.close()


def test_fused_update_matches_cublaslt_chain(monkeypatch):
    """The hand-written two-launch update (ddpg_fused.cuh, the default at hidden = 256) against the cuBLASLt chain
    (BMI_DDPG_CUBLAS=1) on the same batch: same gradients and losses up to fp32 summation order, and 2 launches instead of ~54."""
    L = _seeded(21)
    x, xn, a, r = _batch(3)
    out = {}
    for name, flag in (("fused", "0"), ("lt", "1")):
        monkeypatch.setenv("BMI_DDPG_CUBLAS", flag)
        T = Trainer(L)
        T.backward(x, xn, a, r)                    # first call of the cuBLASLt path builds its plans
        n0 = T._lib.launch_count()
        T.backward(x, xn, a, r)
        out[name] = (T.losses.cpu().numpy().copy(), T.grads(), T._lib.launch_count() - n0)
        T.close()
    assert out["fused"][2] == 2 and out["lt"][2] > 40, (out["fused"][2], out["lt"][2])
    assert np.allclose(out["fused"][0], out["lt"][0], rtol=1e-5, atol=1e-7)
    for mine, ref in zip(out["fused"][1][:2], out["lt"][1][:2]):
        scale = np.abs(ref).max()
        assert np.abs(mine - ref).max() <= 1e-4 * scale + 1e-9, np.abs(mine - ref).max() / scale
        assert np.linalg.norm(mine - ref) <= 2e-5 * np.linalg.norm(ref)
    assert np.all(out["fused"][1][2] == 0)


@pytest.mark.parametrize("dims,batch,amax,gamma,l2", [((40, 5, 7), 96, 1.0, 0.9, 0.5), ((10, 2, 1), 32, 0.25, 0.95, 0.0),
                                                     ((27, 3, 4), 1024, 0.5, 0.98, 1.0)])
def test_fused_update_other_shapes_vs_torch_oracle(dims, batch, amax, gamma, l2):
    """The two-launch update is not specialised to 27 + 3 + 4 / batch 256: other input widths (up to 64 with the action), action
    counts (up to 8) and batch sizes (multiples of 32) against the torch oracle, same tolerances as the reference shape."""
    torch.set_num_threads(1)
    torch.manual_seed(3)
    L = do.Learner(n_obs=dims[0], n_goal=dims[1], n_act=dims[2], max_action=amax, gamma=gamma, action_l2=l2)
    T = Trainer(L, batch=batch, dims=dims, amax=amax, gamma=gamma, l2=l2)
    rng = np.random.RandomState(5)
    Dx = dims[0] + dims[1]
    x = np.clip(rng.standard_normal((batch, Dx)), -5, 5).astype(np.float32)
    xn = np.clip(x + 0.1 * rng.standard_normal((batch, Dx)), -5, 5).astype(np.float32)
    a = rng.uniform(-amax, amax, (batch, dims[2])).astype(np.float32)
    r = -(rng.uniform(size=(batch, 1)) > 0.2).astype(np.float32)
    la, lc, ga, gc = L.losses_and_grads(torch.tensor(x), torch.tensor(xn), torch.tensor(a), torch.tensor(r))
    n0 = T._lib.launch_count()
    T.backward(x, xn, a, r)
    assert T._lib.launch_count() - n0 == 2                   # the fused path was taken
    assert np.allclose(T.losses.cpu().numpy(), [la, lc], rtol=1e-4, atol=1e-6), (T.losses.cpu().numpy(), la, lc)
    mya, myc, pad = T.grads()
    assert np.all(pad == 0)
    for mine, ref in ((mya, ga.numpy()), (myc, gc.numpy())):
        scale = np.abs(ref).max()
        assert np.abs(mine - ref).max() <= 2e-4 * scale + 1e-8, np.abs(mine - ref).max() / scale
        assert np.linalg.norm(mine - ref) <= 1e-4 * np.linalg.norm(ref)
    T.close()


def test_actor_forward_matches_torch():
    L = _seeded(5)
    T = Trainer(L, max_rows=4096)
    _lib = T._lib
    for n in (1, 7, 256, 4096):
        x = np.random.RandomState(n).standard_normal((n, 30)).astype(np.float32)
        out = torch.empty((n, 4), dtype=torch.float32, device="cuda")
        xt = torch.as_tensor(x).cuda()
        _lib.call("bmi_ddpg_act", T.h, _lib.ptr(xt), n, 0, _lib.ptr(out), _lib.stream_ptr())
        with torch.no_grad():
            ref = L.actor(torch.tensor(x)).numpy()
        assert np.allclose(out.cpu().numpy(), ref, rtol=1e-4, atol=2e-6), n
    with pytest.raises(_lib.BmiError):
        _lib.call("bmi_ddpg_act", T.h, _lib.ptr(xt), 5000, 0, _lib.ptr(out), _lib.stream_ptr())
    T.close()


def test_forty_updates_and_polyak_vs_torch_oracle():
    """parameters after 40 updates within rtol 1e-3 of torch CPU fp32."""
    torch.set_num_threads(4)
    L = _seeded(33)
    T = Trainer(L)
    for i in range(40):
        x, xn, a, r = _batch(100 + i)
        L.update(torch.tensor(x), torch.tensor(xn), torch.tensor(a), torch.tensor(r))
        T.backward(x, xn, a, r)
        T.adam()
    L.soft_update()
    T.soft()
    torch.cuda.synchronize()
    for mine, ref in ((T.pa, do.flat_params(L.actor)), (T.pc, do.flat_params(L.critic)), (T.ta, do.flat_params(L.actor_t)),
                      (T.tc, do.flat_params(L.critic_t))):
        m, r_ = mine.cpu().numpy(), ref.numpy()
        assert np.linalg.norm(m - r_) <= 1e-3 * np.linalg.norm(r_)
        # Adam moves every element by ~lr per step whatever the gradient's size, so an element whose
        # gradient is rounding noise can drift by a few steps of lr = 1e-3 between two fp32 back-ends
        assert np.abs(m - r_).max() <= 5e-3
    T.close()


def test_polyak_bit_exact():
    L = _seeded(2)
    for p in L.actor.parameters():
        p.data.add_(0.01)
    for p in L.critic.parameters():
        p.data.mul_(1.01)
    T = Trainer(L)
    L.soft_update()
    T.soft()
    torch.cuda.synchronize()
    assert np.array_equal(T.ta.cpu().numpy(), do.flat_params(L.actor_t).numpy())
    assert np.array_equal(T.tc.cpu().numpy(), do.flat_params(L.critic_t).numpy())
    T.close()


def test_update_chain_vs_reference_agent_golden(golden_dir):
    """HER draws (numpy stream) -> fused inputs kernel -> 3 updates -> Polyak, against the parameters the
    UNMODIFIED reference ddpg_agent produced (tests/golden/learner_update.npz)."""
    from rl_arm_under_sparse_reward_b200 import _lib
    g = np.load(os.path.join(golden_dir, "learner_update.npz"))
    L = _seeded(int(g["weight_seed"]))
    T = Trainer(L)
    dev = torch.device("cuda")
    buf = {k: torch.as_tensor(g["buf_" + k]).to(dev).contiguous() for k in ("obs", "ag", "g", "actions")}
    eps = _lib.Episodes(_lib.ptr(buf["obs"]), _lib.ptr(buf["ag"]), _lib.ptr(buf["g"]), _lib.ptr(buf["actions"]), 8, 100, 27, 3, 4,
                        _lib.BMI_F64, 0)
    st = [torch.as_tensor(v).to(dev) for v in (g["o_mean"], g["o_std"].astype(np.float32), g["g_mean"], g["g_std"].astype(np.float32))]
    np.random.seed(int(g["np_seed"]))
    lo.her_draw_numpy(2, 100, 100)      # the draws _update_normalizer consumed first
    mk = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
    X, XN, A, R = mk(256, 30), mk(256, 30), mk(256, 4), mk(256)
    for i in range(3):
        d = [torch.as_tensor(v).to(dev) for v in lo.her_draw_numpy(8, 100, 256)]
        _lib.call("bmi_her_sample_inputs", ctypes.byref(eps), 8, _lib.ptr(d[0]), _lib.ptr(d[1]), _lib.ptr(d[2]), _lib.ptr(d[3]),
                  256, 0.8, 0.05, 200.0, 5.0, _lib.ptr(st[0]), _lib.ptr(st[1]), _lib.ptr(st[2]), _lib.ptr(st[3]), _lib.ptr(X),
                  _lib.ptr(XN), _lib.ptr(A), _lib.ptr(R), _lib.stream_ptr())
        _lib.call("bmi_ddpg_backward", T.h, _lib.ptr(X), _lib.ptr(XN), _lib.ptr(A), _lib.ptr(R), _lib.ptr(T.losses), _lib.stream_ptr())
        T.adam()
        torch.cuda.synchronize()
        flat = torch.cat([T.pa, T.pc]).cpu().numpy()
        assert np.allclose(flat[g["pick"]], g["params_after"][i], rtol=1e-3, atol=2e-6), i
        assert abs(flat.astype(np.float64).sum() - g["param_sums"][i]) < 5e-2
    T.soft()
    torch.cuda.synchronize()
    tgt = torch.cat([T.ta, T.tc]).cpu().numpy()
    assert np.allclose(tgt[g["pick"]], g["target_after"], rtol=1e-3, atol=2e-6)
    T.close()


def test_select_actions_vs_oracle():
    from rl_arm_under_sparse_reward_b200 import _lib
    dev = torch.device("cuda")
    n = 4096
    pi = torch.as_tensor(np.random.RandomState(0).uniform(-0.5, 0.5, (n, 4)).astype(np.float32)).to(dev)
    ctr = torch.tensor([77], dtype=torch.int64, device=dev)
    out = torch.empty_like(pi)
    for late in (0.0, 0.15):
        ctr.fill_(77)
        _lib.call("bmi_select_actions", _lib.ptr(pi), n, 4, 0.5, 0.01, 0.3, late, ctypes.c_uint64(125), _lib.ptr(ctr), _lib.ptr(out),
                  _lib.stream_ptr())
        want = lo.select_actions_philox(pi.cpu().numpy(), 125, 77, late_clip=late)
        got = out.cpu().numpy()
        assert np.abs(got - want).max() < 2e-6          # float32 sin/cos/log differ by an ulp between libm and CUDA
        assert np.abs(got).max() <= (0.15 if late else 0.5)
    assert int(ctr.item()) == 77 + n
    rand_frac = (np.abs(got - np.clip(pi.cpu().numpy(), -0.15, 0.15)).max(axis=1) > 0.03).mean()
    assert abs(rand_frac - 0.3) < 0.05


def test_fused_p2p_adam_world1_equals_plain_adam():
    """The peer-memory sum + Adam kernel with a single rank (flags and peer pointers all local) must reproduce the
    plain Adam kernel bit for bit; the 2-GPU equivalence with the NCCL path is tests/test_gpu_multi.py."""
    from rl_arm_under_sparse_reward_b200 import _lib
    L = _seeded(8)
    T1, T2 = Trainer(L), Trainer(L)
    buf = (ctypes.c_uint8 * 128)()
    _lib.call("bmi_ddpg_p2p_export", T2.h, ctypes.cast(buf, ctypes.c_void_p))
    _lib.call("bmi_ddpg_p2p_attach", T2.h, 0, 1, ctypes.cast(buf, ctypes.c_void_p))
    for i in range(5):
        x, xn, a, r = _batch(300 + i)
        T1.backward(x, xn, a, r)
        T1.adam()
        T2.backward(x, xn, a, r)
        _lib.call("bmi_ddpg_adam_step_p2p", T2.h, _lib.stream_ptr())
    torch.cuda.synchronize()
    t = ctypes.c_int32(7)
    _lib.call("bmi_ddpg_p2p_status", T2.h, ctypes.byref(t))
    assert t.value == 0
    assert torch.equal(T1.pa, T2.pa) and torch.equal(T1.pc, T2.pc)
    T1.close()
    T2.close()
