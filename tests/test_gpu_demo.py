"""GPU: the vectorised scripted-controller demo generators (reference get_demo_data_push.py / get_demo_data_pick.py)
produce files in the reference's format that the agent's demo loader accepts; the closed loop is also a regression of
the physics kernel (the push script must actually bring blocks to their goals)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_push_demo_generator_format_and_success(tmp_path, golden_dir):
    from rl_arm_under_sparse_reward_b200.get_demo_data import get_demo, save_demo
    out, rate = get_demo("push", demo_num=48, n_envs=256, seed=125, max_batches=2, verbose=False)
    ref = np.load(os.path.join(golden_dir, "demo_small.npz"), allow_pickle=True)       # slice of the reference's own demo file
    n = out["acs"].shape[0]
    assert n > 0 and rate > 0.05, rate          # the scripted push succeeds for a fair share of placements
    for k in ("acs", "obs", "g", "ag"):
        assert out[k].dtype == ref[k].dtype == np.float64 and out[k].shape[1:] == ref[k].shape[1:], k
    assert out["info"].shape == (n, 100) and out["info"].dtype == object
    assert all(out["info"][i, -1]['is_success'] == np.float32(1.0) for i in range(n))   # only successful episodes are kept
    assert isinstance(out["info"][0, 0]['is_success'], np.float32)
    assert np.array_equal(out["ag"], out["obs"][:, :, 12:15])                           # achieved goal = block position
    assert np.array_equal(out["g"], np.repeat(out["g"][:, :1], 100, axis=1))            # goal constant per episode
    assert np.abs(out["acs"][:, :, 3]).max() == 0.0                                     # push: the script never grips
    d = np.linalg.norm(out["ag"][:, -1] - out["g"][:, -1], axis=1)
    assert (d < 0.05).all()
    # the file round-trips through the agent's demo loader (ddpg_agent.py:82-90)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        name = save_demo(out, "push")
        from rl_arm_under_sparse_reward_b200.arguments import Args
        from rl_arm_under_sparse_reward_b200.bmirobot_env.vec_env import BmiVecEnv
        from rl_arm_under_sparse_reward_b200.ddpg_agent import ddpg_agent
        from rl_arm_under_sparse_reward_b200.train import get_env_params
        a = Args()
        a.add_demo, a.demo_name, a.n_envs, a.verbose, a.buffer_size, a.save_dir = True, os.path.join(str(tmp_path), name), 8, False, 256 * 100, str(tmp_path) + "/"
        env = BmiVecEnv(8, task="push", seed=1)
        agent = ddpg_agent(a, env, get_env_params(env))
        assert agent.buffer.current_size == n
        assert np.array_equal(agent.buffer.buffers['obs'][:n].double().cpu().numpy().astype(np.float32), out["obs"].astype(np.float32))
    finally:
        os.chdir(cwd)


def test_pick_demo_controller_runs_and_grips():
    from rl_arm_under_sparse_reward_b200.get_demo_data import pick_controller, run_scripted_batch
    from rl_arm_under_sparse_reward_b200.bmirobot_env.vec_env import BmiVecEnv
    env = BmiVecEnv(64, task="pick", seed=9)
    obs_b, ag_b, g_b, act_b, suc_b = run_scripted_batch(env, pick_controller)
    assert torch.isfinite(obs_b).all() and (ag_b[:, :, 2] > 0.15).all()       # nothing fell through the table
    assert (act_b[:, 30:50, 3] == 0.1).all() and (act_b[:, 70:90, 3] == -0.1).all()   # open / close phases of the script
    # the hand reaches the block's neighbourhood in the approach phases
    assert ((obs_b[:, 70, :3] - obs_b[:, 70, 12:15]).norm(dim=1) < 0.15).float().mean() > 0.5
