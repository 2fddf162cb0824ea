"""Pin the learner-side oracle (oracle/learner_oracle.py, oracle/ddpg_oracle.py) against the
fixtures produced by the UNMODIFIED reference (tests/golden/make_learner_goldens.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import learner_oracle as lo
from oracle import ddpg_oracle as do
from oracle import philox


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_her_bit_exact_vs_reference(golden_dir):
    g = _load(golden_dir, "learner_her.npz")
    buf = {k: g["buf_" + k][: int(g["current_size"])] for k in ("obs", "ag", "g", "actions")}
    np.random.seed(int(g["seed"]))
    draws = lo.her_draw_numpy(int(g["current_size"]), 100, int(g["batch"]))
    tr = lo.her_sample_with_draws(buf, draws, float(g["future_p"]))
    for k in ("obs", "ag", "g", "actions", "obs_next", "ag_next", "r"):
        assert tr[k].dtype == g["tr_" + k].dtype
        assert np.array_equal(tr[k], g["tr_" + k]), k
    # both reward classes and both relabel branches are exercised
    assert 0 < (g["tr_r"] == 0).sum() < g["tr_r"].size


def test_storage_idx_vs_reference(golden_dir):
    g = _load(golden_dir, "learner_storage_idx.npz")
    np.random.seed(int(g["seed"]))
    cs, out = 0, []
    for inc in g["incs"]:
        idx, cs = lo.storage_idx(cs, int(g["size"]), int(inc))
        out.append(np.atleast_1d(idx))
    assert np.array_equal(np.concatenate(out), g["idx"])
    assert cs == int(g["final_size"])


def test_normalizer_vs_reference(golden_dir):
    g = _load(golden_dir, "learner_norm.npz")
    on, gn = lo.Normalizer(27, clip=5), lo.Normalizer(3, clip=5)
    for v, w in zip(g["feeds_o"], g["feeds_g"]):
        on.update(v)
        gn.update(w)
        on.recompute_stats()
        gn.recompute_stats()
    assert np.array_equal(on.total_sum, g["o_total_sum"])
    assert np.array_equal(on.total_sumsq, g["o_total_sumsq"])
    assert np.array_equal(on.total_count, g["o_total_count"])
    assert np.array_equal(on.mean, g["o_mean"]) and np.array_equal(gn.mean, g["g_mean"])
    # the reference ran under numpy 2 (std computed in float64); numpy 1.19 — the version the
    # reference pins — keeps float32: the float32 rounding of the golden must equal ours
    assert np.array_equal(on.std, g["o_std"].astype(np.float32))
    assert np.array_equal(gn.std, g["g_std"].astype(np.float32))
    ref = np.clip((g["probe"] - g["o_mean"]) / g["o_std"].astype(np.float32), -5, 5)
    assert np.array_equal(on.normalize(g["probe"]), ref)
    assert np.allclose(on.normalize(g["probe"]), g["probe_norm"], rtol=1e-6, atol=1e-7)


def _seeded_learner(seed):
    L = do.Learner()
    wr = np.random.RandomState(seed)
    for net in (L.actor, L.critic):
        for _, p in net.named_parameters():
            bound = 1.0 / np.sqrt(p.shape[-1] if p.dim() > 1 else 256)
            p.data.copy_(torch.tensor(wr.uniform(-bound, bound, tuple(p.shape)).astype(np.float32)))
    L.actor_t.load_state_dict(L.actor.state_dict())
    L.critic_t.load_state_dict(L.critic.state_dict())
    return L


def test_update_chain_vs_reference_agent(golden_dir):
    """buffer -> HER draws -> clip/normalise -> 3 DDPG updates -> Polyak, against the reference
    ddpg_agent._update_network run on the same numpy stream."""
    torch.set_num_threads(1)
    g = _load(golden_dir, "learner_update.npz")
    buf = {k: g["buf_" + k] for k in ("obs", "ag", "g", "actions")}
    np.random.seed(int(g["np_seed"]))
    on, gn = lo.Normalizer(27, clip=5), lo.Normalizer(3, clip=5)
    # _update_normalizer (ddpg_agent.py:187-212): HER-sample T transitions from the 2 new episodes
    two = {k: v[:2] for k, v in buf.items()}
    tr = lo.her_sample_with_draws(two, lo.her_draw_numpy(2, 100, 100), 0.8)
    on.update(np.clip(tr["obs"], -200, 200))
    gn.update(np.clip(tr["g"], -200, 200))
    on.recompute_stats()
    gn.recompute_stats()
    assert np.array_equal(on.mean, g["o_mean"]) and np.array_equal(on.std, g["o_std"].astype(np.float32))
    L = _seeded_learner(int(g["weight_seed"]))
    for i in range(3):
        tr = lo.her_sample_with_draws(buf, lo.her_draw_numpy(8, 100, 256), 0.8)
        x, xn, a, r = lo.network_inputs(tr, on, gn)
        L.update(torch.tensor(x), torch.tensor(xn), torch.tensor(a), torch.tensor(r))
        flat = torch.cat([do.flat_params(L.actor), do.flat_params(L.critic)]).numpy()
        assert np.allclose(flat[g["pick"]], g["params_after"][i], rtol=1e-5, atol=1e-7), i
        assert abs(flat.astype(np.float64).sum() - g["param_sums"][i]) < 1e-3
    L.soft_update()
    tgt = torch.cat([do.flat_params(L.actor_t), do.flat_params(L.critic_t)]).numpy()
    assert np.allclose(tgt[g["pick"]], g["target_after"], rtol=1e-6, atol=1e-8)


def test_philox_known_answer():
    # Random123 known-answer vectors for philox4x32-10
    out = philox.philox4x32_10(0, np.array([0], dtype=np.uint64), 0)[0]
    assert [hex(int(v)) for v in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    # counter = ff..f, key = ff..f
    full = (1 << 64) - 1
    out = philox.philox4x32_10(full, np.array([full], dtype=np.uint64), full)[0]
    assert [hex(int(v)) for v in out] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]


def test_philox_draw_ranges():
    ep, t, uh, uo = lo.her_draw_philox(125, 0, 4096, 37, 100)
    assert ep.min() >= 0 and ep.max() < 37 and t.min() >= 0 and t.max() < 100
    assert 0 <= uh.min() and uh.max() < 1 and 0 <= uo.min() and uo.max() < 1
    assert abs((uh < 0.8).mean() - 0.8) < 0.03
    a = lo.select_actions_philox(np.zeros((4096, 4), np.float32), 125, 0)
    assert np.abs(a).max() <= 0.5
    frac_random = (np.abs(a).max(axis=1) > 0.05).mean()
    assert abs(frac_random - 0.3) < 0.05
