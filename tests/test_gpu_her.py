"""GPU parity of the episode store / HER relabel / sparse reward / fused network-input kernels:
bit-exact against the reference's own outputs (golden) and against the numpy oracle."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import learner_oracle as lo

pytestmark = pytest.mark.gpu
PARAMS = {'obs': 27, 'goal': 3, 'action': 4, 'action_max': 0.5, 'max_timesteps': 100}


def _mods():
    from rl_arm_under_sparse_reward_b200 import _lib
    from rl_arm_under_sparse_reward_b200.her import her_sampler
    from rl_arm_under_sparse_reward_b200.replay_buffer import replay_buffer
    return _lib, her_sampler, replay_buffer


def _golden_buffer(golden_dir, dtype):
    _lib, her_sampler, replay_buffer = _mods()
    g = np.load(os.path.join(golden_dir, "learner_her.npz"))
    hs = her_sampler('future', 4, None)
    rb = replay_buffer(PARAMS, 12 * 100, hs.sample_her_transitions, dtype=dtype, verbose=False)
    n = int(g["current_size"])
    rb.store_episode([g["buf_obs"][:n], g["buf_ag"][:n], g["buf_g"][:n], g["buf_actions"][:n]])
    return g, hs, rb


def test_sample_bit_exact_vs_reference_output(golden_dir):
    """replay_buffer.sample through the C-ABI == reference her.py output on the same numpy stream."""
    g, hs, rb = _golden_buffer(golden_dir, torch.float64)
    assert rb.current_size == int(g["current_size"]) and rb.n_transitions_stored == 1200
    np.random.seed(int(g["seed"]))
    tr = rb.sample(int(g["batch"]))
    for k in ("obs", "ag", "g", "actions", "obs_next", "ag_next", "r"):
        assert tr[k].dtype == g["tr_" + k].dtype and tr[k].shape == g["tr_" + k].shape, k
        assert np.array_equal(tr[k], g["tr_" + k]), k


def test_f32_storage_equals_oracle_on_rounded_values(golden_dir):
    g, hs, rb = _golden_buffer(golden_dir, torch.float32)
    np.random.seed(7)
    tr = rb.sample(512)
    np.random.seed(7)
    buf = {k: g["buf_" + k][: int(g["current_size"])].astype(np.float32).astype(np.float64) for k in ("obs", "ag", "g", "actions")}
    want = lo.her_sample_with_draws(buf, lo.her_draw_numpy(int(g["current_size"]), 100, 512), hs.future_p)
    for k in ("obs", "ag", "g", "actions", "obs_next", "ag_next"):
        assert tr[k].dtype == np.float32 and np.array_equal(tr[k].astype(np.float64), want[k]), k
    assert np.array_equal(tr["r"], want["r"])


@pytest.mark.parametrize("B", [1, 3, 255, 4096])
def test_batch_sizes_and_dict_entry_point(golden_dir, B):
    """her_sampler.sample_her_transitions on a host dict (the _update_normalizer call shape)."""
    _lib, her_sampler, _ = _mods()
    g = np.load(os.path.join(golden_dir, "learner_her.npz"))
    hs = her_sampler('future', 4, None)
    eb = {k: g["buf_" + k][:2] for k in ("obs", "ag", "g", "actions")}
    eb["obs_next"], eb["ag_next"] = eb["obs"][:, 1:, :], eb["ag"][:, 1:, :]
    np.random.seed(B)
    tr = hs.sample_her_transitions(eb, B)
    np.random.seed(B)
    want = lo.her_sample_with_draws({k: eb[k] for k in ("obs", "ag", "g", "actions")}, lo.her_draw_numpy(2, 100, B), 0.8)
    for k in want:
        assert np.array_equal(tr[k], want[k]), k


def test_edge_indices_last_step_and_forced_relabel(golden_dir):
    """t = T-1 with u_off -> future_t = T exactly; u_her just below / at future_p."""
    _lib, her_sampler, _ = _mods()
    g = np.load(os.path.join(golden_dir, "learner_her.npz"))
    hs = her_sampler('future', 4, None)
    dev = torch.device("cuda")
    arrs = [torch.as_tensor(g["buf_" + k][:4]).to(dev) for k in ("obs", "ag", "g", "actions")]
    ep = np.array([0, 3, 1, 2, 0], dtype=np.int64)
    t = np.array([99, 99, 0, 0, 50], dtype=np.int64)
    u_her = np.array([0.0, np.nextafter(0.8, 0), 0.8, 0.9999, 0.5])
    u_off = np.array([0.999999, 0.0, np.nextafter(1.0, 0), 0.0, 0.5])
    out = hs.sample_device(*arrs, 4, (ep, t, u_her, u_off))
    buf = {k: g["buf_" + k][:4] for k in ("obs", "ag", "g", "actions")}
    want = lo.her_sample_with_draws(buf, (ep, t, u_her, u_off), hs.future_p)
    for k in want:
        assert np.array_equal(out[k].cpu().numpy(), want[k]), k
    assert np.array_equal(out["g"][0].cpu().numpy(), buf["ag"][0, 100])      # future_t == T
    assert np.array_equal(out["g"][2].cpu().numpy(), buf["g"][1, 0])         # u_her == future_p -> not relabelled


def test_empty_buffer_and_bad_inputs_raise():
    _lib, her_sampler, replay_buffer = _mods()
    hs = her_sampler('future', 4, None)
    rb = replay_buffer(PARAMS, 400, hs.sample_her_transitions, verbose=False)
    with pytest.raises(ValueError):
        rb.sample(8)
    with pytest.raises(_lib.BmiError):
        hs.sample_her_transitions({"obs": np.zeros((1, 101, 27)), "ag": np.zeros((1, 101, 3)), "g": np.zeros((1, 100, 3)),
                                   "actions": np.zeros((1, 100, 4)), "obs_next": np.zeros((1, 100, 27))}, 4)


def test_overwrite_when_full_matches_numpy_semantics():
    """random overwrite with duplicate slots: last write wins, like numpy fancy assignment."""
    _lib, her_sampler, replay_buffer = _mods()
    hs = her_sampler('future', 4, None)
    rb = replay_buffer(PARAMS, 300, hs.sample_her_transitions, verbose=False)     # 3 slots
    ref = {k: np.zeros(tuple(v.shape)) for k, v in rb.buffers.items()}
    rng = np.random.RandomState(0)
    np.random.seed(11)
    cs = 0
    state = None
    for n in (2, 4, 5):
        ep = [rng.standard_normal((n, 101, 27)), rng.standard_normal((n, 101, 3)), rng.standard_normal((n, 100, 3)),
              rng.standard_normal((n, 100, 4))]
        st = np.random.get_state()
        rb.store_episode(ep)
        np.random.set_state(st)
        idx, cs = lo.storage_idx(cs, 3, n)
        for k, a in zip(("obs", "ag", "g", "actions"), ep):
            ref[k][idx] = a
    for k in ref:
        assert np.array_equal(rb.buffers[k].cpu().numpy(), ref[k]), k


def test_reward_kernel_matches_numpy(golden_dir):
    from rl_arm_under_sparse_reward_b200.bmirobot_env.vec_env import BmiVecEnv
    env = BmiVecEnv(1)
    rng = np.random.RandomState(3)
    ag = 0.3 + rng.standard_normal((2000, 3)) * 0.03
    g = 0.3 + rng.standard_normal((2000, 3)) * 0.03
    g[:50] = ag[:50] + np.array([0.05, 0, 0])            # exactly at the threshold (d > 0.05 is False or rounding)
    r = env.compute_reward(ag, g, None)
    assert r.dtype == np.float32 and np.array_equal(r, lo.compute_reward(ag, g))
    r2 = env.compute_reward(ag.reshape(20, 100, 3), g.reshape(20, 100, 3), None)
    assert r2.shape == (20, 100)
    with pytest.raises(AssertionError):
        env.compute_reward(ag, g[:10], None)


def test_fused_network_inputs_bit_exact(golden_dir):
    """bmi_her_sample_inputs == reference chain clip -> normalize -> concat -> float32 cast."""
    _lib, her_sampler, _ = _mods()
    g = np.load(os.path.join(golden_dir, "learner_update.npz"))
    dev = torch.device("cuda")
    buf = {k: g["buf_" + k] for k in ("obs", "ag", "g", "actions")}
    on, gn = lo.Normalizer(27, clip=5), lo.Normalizer(3, clip=5)
    on.mean, on.std = g["o_mean"], g["o_std"].astype(np.float32)
    gn.mean, gn.std = g["g_mean"], g["g_std"].astype(np.float32)
    np.random.seed(5)
    draws = lo.her_draw_numpy(8, 100, 256)
    x, xn, a, r = lo.network_inputs(lo.her_sample_with_draws(buf, draws, 0.8), on, gn)
    for dt in (torch.float64, torch.float32):
        t = {k: torch.as_tensor(v).to(dev, dt).contiguous() for k, v in buf.items()}
        if dt == torch.float32:
            b32 = {k: v.astype(np.float32).astype(np.float64) for k, v in buf.items()}
            x, xn, a, r = lo.network_inputs(lo.her_sample_with_draws(b32, draws, 0.8), on, gn)
        eps = _lib.Episodes(_lib.ptr(t["obs"]), _lib.ptr(t["ag"]), _lib.ptr(t["g"]), _lib.ptr(t["actions"]), 8, 100, 27, 3, 4,
                            _lib.dtype_code(dt), 0)
        d = [torch.as_tensor(v).to(dev) for v in draws]
        mk = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        X, XN, A, Rr = mk(256, 30), mk(256, 30), mk(256, 4), mk(256)
        st = [torch.as_tensor(v).to(dev) for v in (on.mean, on.std, gn.mean, gn.std)]
        _lib.call("bmi_her_sample_inputs", ctypes.byref(eps), 8, _lib.ptr(d[0]), _lib.ptr(d[1]), _lib.ptr(d[2]), _lib.ptr(d[3]),
                  256, 0.8, 0.05, 200.0, 5.0, _lib.ptr(st[0]), _lib.ptr(st[1]), _lib.ptr(st[2]), _lib.ptr(st[3]), _lib.ptr(X),
                  _lib.ptr(XN), _lib.ptr(A), _lib.ptr(Rr), _lib.stream_ptr())
        assert np.array_equal(X.cpu().numpy(), x) and np.array_equal(XN.cpu().numpy(), xn)
        assert np.array_equal(A.cpu().numpy(), a) and np.array_equal(Rr.cpu().numpy(), r[:, 0])


def test_fused_inputs_large_ragged_batch_bit_exact():
    """The lane-per-element sampler (multiplication by 1 / std with an exact-division slow path) == numpy float64 division,
    on 100 003 samples (ragged tail) of a random buffer, both storage dtypes; a few values sit exactly on float32 rounding
    boundaries so that the slow path runs."""
    _lib, _, _ = _mods()
    dev = torch.device("cuda")
    rng = np.random.RandomState(11)
    E, T, B = 300, 100, 100003
    buf = {"obs": rng.standard_normal((E, T + 1, 27)) * 0.7, "ag": 0.3 + rng.standard_normal((E, T + 1, 3)) * 0.03,
           "g": 0.3 + rng.standard_normal((E, T, 3)) * 0.03, "actions": rng.uniform(-0.5, 0.5, (E, T, 4))}
    on, gn = lo.Normalizer(27, clip=5), lo.Normalizer(3, clip=5)
    on.mean, on.std = (rng.standard_normal(27) * 0.2).astype(np.float32), rng.uniform(0.05, 1.5, 27).astype(np.float32)
    gn.mean, gn.std = (0.3 + rng.standard_normal(3) * 0.01).astype(np.float32), rng.uniform(0.01, 0.1, 3).astype(np.float32)
    on.mean[5], on.std[5] = 0.0, 1.0
    buf["obs"][:, :, 5] = 1.0 + 2.0 ** -24 * rng.randint(-3, 4, (E, T + 1))        # float32 midpoints / representable values
    buf["obs"][:, ::7, 6] = on.mean[6]                                              # exact zeros after the subtraction
    np.random.seed(9)
    draws = lo.her_draw_numpy(E, T, B)
    for dt in (torch.float64, torch.float32):
        src = buf if dt == torch.float64 else {k: v.astype(np.float32).astype(np.float64) for k, v in buf.items()}
        x, xn, a, r = lo.network_inputs(lo.her_sample_with_draws(src, draws, 0.8), on, gn)
        t = {k: torch.as_tensor(v).to(dev, dt).contiguous() for k, v in buf.items()}
        eps = _lib.Episodes(_lib.ptr(t["obs"]), _lib.ptr(t["ag"]), _lib.ptr(t["g"]), _lib.ptr(t["actions"]), E, T, 27, 3, 4,
                            _lib.dtype_code(dt), 0)
        d = [torch.as_tensor(v).to(dev) for v in draws]
        mk = lambda *s: torch.full(s, 7.0, dtype=torch.float32, device=dev)
        X, XN, A, Rr = mk(B + 4, 30), mk(B + 4, 30), mk(B + 4, 4), mk(B + 4)
        st = [torch.as_tensor(v).to(dev) for v in (on.mean, on.std, gn.mean, gn.std)]
        _lib.call("bmi_her_sample_inputs", ctypes.byref(eps), E, _lib.ptr(d[0]), _lib.ptr(d[1]), _lib.ptr(d[2]), _lib.ptr(d[3]),
                  B, 0.8, 0.05, 200.0, 5.0, _lib.ptr(st[0]), _lib.ptr(st[1]), _lib.ptr(st[2]), _lib.ptr(st[3]), _lib.ptr(X),
                  _lib.ptr(XN), _lib.ptr(A), _lib.ptr(Rr), _lib.stream_ptr())
        assert np.array_equal(X[:B].cpu().numpy(), x) and np.array_equal(XN[:B].cpu().numpy(), xn)
        assert np.array_equal(A[:B].cpu().numpy(), a) and np.array_equal(Rr[:B].cpu().numpy(), r[:, 0])
        for tail in (X, XN, A, Rr):                                                  # nothing written past the batch
            assert bool((tail[B:] == 7.0).all())


def test_fused_inputs_wide_observation_fallback_bit_exact():
    """obs_dim + goal_dim > 32 takes the tiled sampler (her_inputs_kernel) instead of the lane-per-element one: same bit-exact
    contract, both tile sizes (batch <= 8192 and above)."""
    _lib, _, _ = _mods()
    dev = torch.device("cuda")
    rng = np.random.RandomState(4)
    E, T, Do, Dg, Da = 40, 50, 40, 5, 6
    buf = {"obs": rng.standard_normal((E, T + 1, Do)), "ag": rng.standard_normal((E, T + 1, Dg)) * 0.05,
           "g": rng.standard_normal((E, T, Dg)) * 0.05, "actions": rng.uniform(-1, 1, (E, T, Da))}
    on, gn = lo.Normalizer(Do, clip=5), lo.Normalizer(Dg, clip=5)
    on.mean, on.std = (rng.standard_normal(Do) * 0.2).astype(np.float32), rng.uniform(0.05, 1.5, Do).astype(np.float32)
    gn.mean, gn.std = (rng.standard_normal(Dg) * 0.01).astype(np.float32), rng.uniform(0.01, 0.1, Dg).astype(np.float32)
    for B in (777, 9001):
        np.random.seed(B)
        draws = lo.her_draw_numpy(E, T, B)
        x, xn, a, r = lo.network_inputs(lo.her_sample_with_draws(buf, draws, 0.8), on, gn)
        t = {k: torch.as_tensor(v).to(dev, torch.float64).contiguous() for k, v in buf.items()}
        eps = _lib.Episodes(_lib.ptr(t["obs"]), _lib.ptr(t["ag"]), _lib.ptr(t["g"]), _lib.ptr(t["actions"]), E, T, Do, Dg, Da,
                            _lib.dtype_code(torch.float64), 0)
        d = [torch.as_tensor(v).to(dev) for v in draws]
        mk = lambda *sh: torch.empty(sh, dtype=torch.float32, device=dev)
        X, XN, A, Rr = mk(B, Do + Dg), mk(B, Do + Dg), mk(B, Da), mk(B)
        st = [torch.as_tensor(v).to(dev) for v in (on.mean, on.std, gn.mean, gn.std)]
        _lib.call("bmi_her_sample_inputs", ctypes.byref(eps), E, _lib.ptr(d[0]), _lib.ptr(d[1]), _lib.ptr(d[2]), _lib.ptr(d[3]),
                  B, 0.8, 0.05, 200.0, 5.0, _lib.ptr(st[0]), _lib.ptr(st[1]), _lib.ptr(st[2]), _lib.ptr(st[3]), _lib.ptr(X),
                  _lib.ptr(XN), _lib.ptr(A), _lib.ptr(Rr), _lib.stream_ptr())
        assert np.array_equal(X.cpu().numpy(), x) and np.array_equal(XN.cpu().numpy(), xn), B
        assert np.array_equal(A.cpu().numpy(), a) and np.array_equal(Rr.cpu().numpy(), r[:, 0]), B


def test_device_philox_draws_match_oracle():
    _lib, _, _ = _mods()
    dev = torch.device("cuda")
    B, T, nv = 1000, 100, 37
    ctr = torch.tensor([12345], dtype=torch.int64, device=dev)
    n_valid = torch.tensor([nv], dtype=torch.int64, device=dev)
    ep, t = torch.empty(B, dtype=torch.int64, device=dev), torch.empty(B, dtype=torch.int64, device=dev)
    uh, uo = torch.empty(B, dtype=torch.float64, device=dev), torch.empty(B, dtype=torch.float64, device=dev)
    _lib.call("bmi_her_draw", ctypes.c_uint64(125), _lib.ptr(ctr), B, _lib.ptr(n_valid), T, _lib.ptr(ep), _lib.ptr(t),
              _lib.ptr(uh), _lib.ptr(uo), _lib.stream_ptr())
    want = lo.her_draw_philox(125, 12345, B, nv, T)
    for got, w in zip((ep, t, uh, uo), want):
        assert np.array_equal(got.cpu().numpy(), w)
    assert int(ctr.item()) == 12345 + B          # counter advanced for the next (graph-replayed) draw


def test_full_size_buffer_properties():
    """BASELINE size (5000 episodes = 5e5 transitions): size-independent properties of a 65536 batch."""
    _lib, her_sampler, replay_buffer = _mods()
    hs = her_sampler('future', 4, None)
    rb = replay_buffer(PARAMS, 5e5, hs.sample_her_transitions, dtype=torch.float32, verbose=False)
    dev = rb.device
    gen = torch.Generator(device=dev).manual_seed(0)
    E = 5000
    ag = 0.3 + torch.cumsum(torch.randn(E, 101, 3, device=dev, generator=gen) * 0.01, dim=1)
    obs = torch.randn(E, 101, 27, device=dev, generator=gen)
    obs[:, :, 12:15] = ag
    g = (0.3 + torch.randn(E, 1, 3, device=dev, generator=gen) * 0.05).expand(E, 100, 3).contiguous()
    act = torch.rand(E, 100, 4, device=dev, generator=gen) - 0.5
    rb.store_episode([obs, ag, g, act])
    assert rb.current_size == 5000
    np.random.seed(0)
    tr = rb.sample_device(65536)
    assert torch.equal(tr["ag"], tr["obs"][:, 12:15]) and torch.equal(tr["ag_next"], tr["obs_next"][:, 12:15])
    d = (tr["ag_next"].double() - tr["g"].double()).pow(2).sum(-1).sqrt()
    assert torch.equal(tr["r"][:, 0], -(d > 0.05).float())
    # every sampled row exists in the buffer at (ep, t): check via the draws replayed on the host
    np.random.seed(0)
    ep, t, uh, uo = lo.her_draw_numpy(5000, 100, 65536)
    ep_t, t_t = torch.as_tensor(ep, device=dev), torch.as_tensor(t, device=dev)
    assert torch.equal(tr["obs"], rb.buffers["obs"][ep_t, t_t]) and torch.equal(tr["obs_next"], rb.buffers["obs"][ep_t, t_t + 1])
    relabel = torch.as_tensor(uh < 0.8, device=dev)
    assert torch.equal(tr["g"][~relabel], rb.buffers["g"][ep_t, t_t][~relabel])
    assert abs(relabel.float().mean().item() - 0.8) < 0.01
