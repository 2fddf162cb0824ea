"""CPU-only tests of the host-side logic of the Python mirrors (no kernels are launched)."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import learner_oracle as lo
from rl_arm_under_sparse_reward_b200.arguments import Args
from rl_arm_under_sparse_reward_b200.her import her_sampler, _is_shifted_alias
from rl_arm_under_sparse_reward_b200.replay_buffer import replay_buffer

PARAMS = {'obs': 27, 'goal': 3, 'action': 4, 'action_max': 0.5, 'max_timesteps': 100}


def test_storage_idx_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "learner_storage_idx.npz"))
    rb = replay_buffer(PARAMS, int(g["size"]) * 100, None, device="cpu", verbose=False)
    np.random.seed(int(g["seed"]))
    out = [np.atleast_1d(rb._get_storage_idx(inc=int(i))) for i in g["incs"]]
    assert np.array_equal(np.concatenate(out), g["idx"])
    assert rb.current_size == int(g["final_size"])
    assert rb.size == 12 and rb.T == 100


def test_buffer_shapes_and_reference_size_rule():
    a = Args()
    rb = replay_buffer(PARAMS, a.buffer_size, None, device="cpu", verbose=False, dtype=torch.float32)
    assert rb.size == 5000            # int(5e5 // 100), replay_buffer.py:16
    assert tuple(rb.buffers['obs'].shape) == (5000, 101, 27) and tuple(rb.buffers['g'].shape) == (5000, 100, 3)
    with pytest.raises(ValueError):
        rb.store_episode([np.zeros((2, 100, 27)), np.zeros((2, 101, 3)), np.zeros((2, 100, 3)), np.zeros((2, 100, 4))])


def test_her_sampler_parameters_and_draw_order():
    hs = her_sampler('future', 4, None)
    assert hs.future_p == 1 - 1. / 5 and hs.distance_threshold == 0.05
    assert her_sampler('none', 4, None).future_p == 0
    np.random.seed(3)
    mine = hs.draw(37, 100, 64)
    np.random.seed(3)
    ref = lo.her_draw_numpy(37, 100, 64)
    for a, b in zip(mine, ref):
        assert np.array_equal(a, b)


def test_shifted_alias_detection():
    obs = np.zeros((4, 11, 5))
    assert _is_shifted_alias(obs, obs[:, 1:, :])
    assert not _is_shifted_alias(obs, obs[:, 1:, :].copy())
    t = torch.zeros(4, 11, 5)
    assert _is_shifted_alias(t, t[:, 1:, :]) and not _is_shifted_alias(t, t[:, :-1, :])


def test_args_defaults_match_reference():
    a = Args()
    assert (a.n_epochs, a.n_cycles, a.n_batches, a.batch_size, a.replay_k) == (200, 50, 40, 256, 4)
    assert (a.seed, a.gamma, a.polyak, a.noise_eps, a.random_eps, a.clip_obs, a.clip_range) == (125, 0.98, 0.95, 0.01, 0.3, 200, 5)
    assert a.buffer_size == 5e5 and a.num_rollouts_per_mpi == 2 and a.n_test_rollouts == 25


def test_env_placement_draw_order_matches_reference_semantics():
    """6 python-random draws per push attempt, 7 per pick attempt, rejection below 15 cm."""
    from rl_arm_under_sparse_reward_b200.bmirobot_env import bmirobot_push_F as push, bmirobot_pickandplace_v2 as pick
    for mod, n_draws in ((push, 6), (pick, 7)):
        env = mod.bmirobotGymEnv.__new__(mod.bmirobotGymEnv)
        random.seed(5)
        p = env._sample_placement()
        state_after = random.getstate()
        random.seed(5)
        attempts = 0
        while True:
            d = [random.random() for _ in range(n_draws)]
            attempts += 1
            x, y = 0.15 + 0.2 * d[0], d[1] * 0.3 + 0.2
            xt = 0.35 * d[3]
            yt, zt = (d[4] * 0.3 + 0.2, 0.2) if n_draws == 6 else (d[4] * 0.25 + 0.3, 0.3 + 0.2 * d[5])
            if ((x - xt) ** 2 + (y - yt) ** 2 + (0.2 - zt) ** 2) ** 0.5 >= 0.15:
                break
        assert random.getstate() == state_after
        assert np.allclose(p[:3], [x, y, 0.2]) and np.allclose(p[4:7], [xt, yt, zt])


def test_kernel_lane_budget_and_model_constants():
    """The kernel's solver gives one lane to every contact (lane budget below); the oracle keeps EVERY contact it finds
    (storage bound only) and takes the kernel's caps as a run-time option of the kernel-vs-oracle tests (bmo_set_caps)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cu = open(os.path.join(root, "rl_arm_under_sparse_reward_b200", "csrc", "physics.cu")).read()
    oc = open(os.path.join(root, "oracle", "bmi_physics_oracle.c")).read()
    maxc = int(re.search(r"constexpr int MAXC = (\d+);", cu).group(1))
    assert int(re.search(r"#define MAX_CONTACTS (\d+)", oc).group(1)) >= 4 * maxc
    # one solver lane per contact next to 9 joint lanes and 6 block-velocity lanes
    assert 16 + maxc <= 32 and 3 * maxc <= 32
    blob = np.fromfile(os.path.join(root, "rl_arm_under_sparse_reward_b200", "assets", "bmirobot_model.bin"), "<f4")
    staged = int(re.search(r"constexpr int STAGED_FULL = (\d+);", cu).group(1))
    assert blob.shape[0] == int(blob[7]) <= staged          # MP_TOTAL; the env kernels stage the whole blob
    assert int(blob[2]) == 9 and int(blob[3]) <= 4           # links, shapes
