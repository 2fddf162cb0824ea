"""GPU end-to-end: one vectorised training cycle through the reference-shaped agent API."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _agent(tmp_path, n_envs=32, **kw):
    from rl_arm_under_sparse_reward_b200.arguments import Args
    from rl_arm_under_sparse_reward_b200.bmirobot_env.vec_env import BmiVecEnv
    from rl_arm_under_sparse_reward_b200.ddpg_agent import ddpg_agent
    from rl_arm_under_sparse_reward_b200.train import get_env_params
    a = Args()
    a.add_demo, a.n_envs, a.verbose, a.buffer_size, a.save_dir = False, n_envs, False, 256 * 100, str(tmp_path) + "/"
    for k, v in kw.items():
        setattr(a, k, v)
    env = BmiVecEnv(n_envs, task=a.train_type, seed=a.seed)
    torch.manual_seed(a.seed)
    return ddpg_agent(a, env, get_env_params(env)), a


def test_one_cycle_graphed_equals_eager(tmp_path):
    ag1, _ = _agent(tmp_path, use_cuda_graphs=True)
    ag2, _ = _agent(tmp_path, use_cuda_graphs=False)
    ag2.actor_network.flat.copy_(ag1.actor_network.flat)
    ag2.critic_network.flat.copy_(ag1.critic_network.flat)
    ag2.actor_target_network.flat.copy_(ag1.actor_target_network.flat)
    ag2.critic_target_network.flat.copy_(ag1.critic_target_network.flat)
    for ag in (ag1, ag2):
        for _ in range(2):      # second call replays the captured graph
            ag.rollout(0)
            ag.buffer.store_episode([ag.ep['obs'], ag.ep['ag'], ag.ep['g'], ag.ep['actions']])
            ag._update_normalizer()
            ag.update_many(4)
            ag._soft_update_target_network()
    torch.cuda.synchronize()
    assert torch.equal(ag1.ep['obs'], ag2.ep['obs']) and torch.equal(ag1.ep['actions'], ag2.ep['actions'])
    assert torch.equal(ag1.actor_network.flat, ag2.actor_network.flat)
    assert torch.equal(ag1.critic_target_network.flat, ag2.critic_target_network.flat)
    assert ag1.buffer.current_size == 64 and ag1.env_steps == 2 * 32 * 100 and ag1.updates == 8
    # episode arrays are self-consistent: ag == obs[12:15], g constant per episode
    ep = ag1.ep
    assert torch.equal(ep['ag'], ep['obs'][:, :, 12:15])
    assert torch.equal(ep['g'], ep['g'][:, :1].expand_as(ep['g']))
    assert ep['actions'].abs().max() <= 0.5


def test_numpy_stream_update_is_bit_reproducible_and_matches_oracle_sampling(tmp_path):
    from oracle import learner_oracle as lo
    ag, a = _agent(tmp_path, device_rng=False, buffer_dtype="float64")
    ag.rollout(0)
    ag.buffer.store_episode([ag.ep['obs'], ag.ep['ag'], ag.ep['g'], ag.ep['actions']])
    np.random.seed(4)
    ag._update_normalizer()
    ag._update_network()
    torch.cuda.synchronize()
    x_gpu = ag._x.cpu().numpy().copy()
    # replay the same numpy stream through the oracle
    np.random.seed(4)
    bufs = {k: v[:32].cpu().numpy() for k, v in ag.buffer.buffers.items()}
    new = {k: ag.ep[k].double().cpu().numpy() for k in ('obs', 'ag', 'g', 'actions')}
    on, gn = lo.Normalizer(27, clip=5), lo.Normalizer(3, clip=5)
    tr = lo.her_sample_with_draws(new, lo.her_draw_numpy(32, 100, 100 * 16), 0.8)
    on.update(np.clip(tr['obs'], -200, 200))
    gn.update(np.clip(tr['g'], -200, 200))
    on.recompute_stats()
    gn.recompute_stats()
    assert np.array_equal(ag.o_norm.mean, on.mean) and np.array_equal(ag.o_norm.std, on.std)
    x, xn, act, r = lo.network_inputs(lo.her_sample_with_draws(bufs, lo.her_draw_numpy(32, 100, 256), 0.8), on, gn)
    assert np.array_equal(x_gpu, x) and np.array_equal(ag._r.cpu().numpy(), r[:, 0])


def test_eval_and_checkpoint_format(tmp_path):
    ag, a = _agent(tmp_path, n_envs=8, n_test_rollouts=8)
    rate = ag._eval_agent()
    assert 0.0 <= rate <= 1.0
    path = ag.save_checkpoint()
    o_mean, o_std, g_mean, g_std, sd = torch.load(path, weights_only=False)
    assert o_mean.shape == (27,) and g_std.shape == (3,) and o_mean.dtype == np.float32
    assert list(sd.keys()) == ['fc1.weight', 'fc1.bias', 'fc2.weight', 'fc2.bias', 'fc3.weight', 'fc3.bias',
                               'action_out.weight', 'action_out.bias']
    assert sd['fc1.weight'].shape == (256, 30) and not sd['fc1.weight'].is_cuda
    # the reference's own actor class shape (models.py:11-26) loads it
    from oracle.ddpg_oracle import Actor
    Actor().load_state_dict(sd)


def test_demo_buffer_preload(tmp_path, golden_dir):
    demo = os.path.join(golden_dir, "demo_small.npz")
    ag, a = _agent(tmp_path, n_envs=4, add_demo=True, demo_name=demo)
    d = np.load(demo)
    assert ag.buffer.current_size == d["obs"].shape[0]
    assert np.allclose(ag.buffer.buffers['obs'][:2].cpu().numpy(), d["obs"][:2].astype(np.float32))
    assert float(ag.o_norm.total_count[0]) == 1.0     # normalisers are NOT updated from demos (ddpg_agent.py:49-53)


def test_fused_rollout_matches_stepwise_pipeline(tmp_path):
    """The fused per-env rollout kernel (policy MLP in-kernel) against the step-wise pipeline (cuBLASLt actor +
    one env kernel per step): same Philox noise, same physics; only the MLP's fp32 summation order differs."""
    ag1, _ = _agent(tmp_path, n_envs=64, fused_rollout=True)
    ag2, _ = _agent(tmp_path, n_envs=64, fused_rollout=False, use_cuda_graphs=False)
    ag2.actor_network.flat.copy_(ag1.actor_network.flat)
    # non-trivial normaliser statistics
    for ag in (ag1, ag2):
        g = torch.Generator(device="cuda").manual_seed(0)
        ag.o_norm.update(torch.randn(200, 27, device="cuda", generator=g) * 0.3)
        ag.g_norm.update(0.3 + torch.randn(200, 3, device="cuda", generator=g) * 0.1)
        ag.o_norm.recompute_stats()
        ag.g_norm.recompute_stats()
    ag1.rollout(0)
    ag2.rollout(0)
    torch.cuda.synchronize()
    e1, e2 = ag1.ep, ag2.ep
    assert torch.equal(e1['obs'][:, 0], e2['obs'][:, 0]) and torch.equal(e1['g'], e2['g'])     # same reset
    assert (e1['actions'][:, 0] - e2['actions'][:, 0]).abs().max() < 1e-5                      # same policy + noise
    # random (epsilon) actions are bit-identical whatever the MLP rounding
    big = e2['actions'].abs().amax(dim=2) > 0.05
    assert big.float().mean() > 0.2
    d1 = (e1['obs'][:, 1, :3] - e2['obs'][:, 1, :3]).abs().max()
    assert d1 < 1e-4, d1
    # trajectories stay statistically aligned early on (contacts make them diverge later)
    med = (e1['obs'][:, 10, :3] - e2['obs'][:, 10, :3]).abs().amax(dim=1).median()
    assert med < 5e-3, med
    assert torch.equal(e1['ag'], e1['obs'][:, :, 12:15]) and e1['actions'].abs().max() <= 0.5
    assert torch.isfinite(e1['obs']).all()
    # evaluation path runs and reports a rate
    assert 0.0 <= ag1._eval_agent() <= 1.0


def test_pick_task_cycle_with_demo_preload(tmp_path, golden_dir):
    """BASELINE config 3 in miniature: pick-and-place env, add_demo=True with (a slice of) the reference's pick demo file."""
    demo = os.path.join(golden_dir, "demo_small_pick.npz")
    ag, a = _agent(tmp_path, n_envs=16, add_demo=True, demo_name=demo, train_type="pick")
    assert ag.vec.task == 1 and ag.buffer.current_size == 8
    d = np.load(demo)
    # recorded pick episodes end in success: the stored goals/achieved goals keep that property through the f32 store
    last = np.linalg.norm(d["ag"][:, -1] - d["g"][:, -1], axis=1)
    assert (last < 0.05).all()
    ag.rollout(0)
    ag.buffer.store_episode([ag.ep['obs'], ag.ep['ag'], ag.ep['g'], ag.ep['actions']])
    ag._update_normalizer()
    ag.update_many(4)
    ag._soft_update_target_network()
    torch.cuda.synchronize()
    assert ag.buffer.current_size == 24 and np.isfinite(ag.losses()).all()
    g = ag.ep['g'][:, 0]
    assert (g[:, 1] >= 0.3 - 1e-6).all() and (g[:, 1] <= 0.55 + 1e-6).all() and (g[:, 2] >= 0.3 - 1e-6).all() and (g[:, 2] <= 0.5 + 1e-6).all()
    # the 4x4x8 cm block settles on the table (upright z = 0.175 + 0.04; toppled by the arm z = 0.175 + 0.02) unless it is lifted
    z = ag.ep['ag'][:, 5, 2]
    assert ((z - 0.215).abs() < 5e-3).float().mean() > 0.6
    assert (z > 0.19).all() and (z < 0.30).all()


def test_launch_learn_one_epoch_with_reference_demo_and_checkpoint(tmp_path, golden_dir, monkeypatch):
    """train.launch() -> ddpg_agent.learn() end to end (train.py:26-45, ddpg_agent.py:92-161): add_demo with a slice of the
    reference's own demo file, 1 epoch x 2 cycles at 64 envs, per-epoch eval, checkpoint in the reference's tuple layout
    that the reference's demo_push.py loader accepts (models.py state_dict keys), success_rates saved by
    plot_success_rate()."""
    from rl_arm_under_sparse_reward_b200 import train
    from rl_arm_under_sparse_reward_b200.arguments import Args
    monkeypatch.chdir(tmp_path)
    a = Args()
    a.n_envs, a.n_epochs, a.n_cycles, a.n_batches, a.verbose = 64, 1, 2, 3, False
    a.buffer_size, a.save_dir = 256 * 100, str(tmp_path) + "/saved_models/"
    a.add_demo, a.demo_name = True, os.path.join(golden_dir, "demo_small.npz")
    tr = train.launch(a)
    assert tr.buffer.current_size == 16 + 2 * 64 and tr.env_steps == 2 * 64 * 100 and tr.updates == 6
    assert len(tr.success_rates) == 1 and 0.0 <= tr.success_rates[0] <= 1.0
    assert os.path.exists(tmp_path / "test_rates" / ("%d_True_success_rates.npy" % a.seed))
    ckpt = [f for f in os.listdir(tr.model_path) if f.endswith("_model.pt")]
    assert len(ckpt) == 1
    o_mean, o_std, g_mean, g_std, sd = torch.load(os.path.join(tr.model_path, ckpt[0]), weights_only=False)
    assert np.asarray(o_mean).shape == (27,) and np.asarray(g_std).shape == (3,)
    ref = torch.nn.ModuleDict({"fc1": torch.nn.Linear(30, 256), "fc2": torch.nn.Linear(256, 256), "fc3": torch.nn.Linear(256, 256),
                               "action_out": torch.nn.Linear(256, 4)})   # the reference actor's layers (models.py:15-18)
    ref.load_state_dict(sd)
    # a missing demo file fails with a message that says what to do (Args() default: add_demo=True, cwd-relative name)
    b = Args()
    b.n_envs, b.verbose, b.save_dir, b.demo_name = 8, False, str(tmp_path) + "/m2/", "does_not_exist.npz"
    from rl_arm_under_sparse_reward_b200.bmirobot_env.vec_env import BmiVecEnv
    from rl_arm_under_sparse_reward_b200.ddpg_agent import ddpg_agent
    with pytest.raises(FileNotFoundError, match="get_demo_data"):
        ddpg_agent(b, BmiVecEnv(8), train.get_env_params(None) if False else {'obs': 27, 'goal': 3, 'action': 4, 'action_max': 0.5, 'max_timesteps': 100})


def test_soft_update_honours_target_and_source(tmp_path):
    """ddpg_agent.py:220-222 with the reference's (target, source) arguments updates exactly that net"""
    ag, a = _agent(tmp_path, n_envs=8)
    ag.actor_network.flat.add_(0.25)
    ag.critic_network.flat.add_(0.5)
    at, ct = ag.actor_target_network.flat.clone(), ag.critic_target_network.flat.clone()
    ag._soft_update_target_network(ag.critic_target_network, ag.critic_network)
    assert torch.equal(ag.actor_target_network.flat, at)
    want = (1 - a.polyak) * ag.critic_network.flat + a.polyak * ct
    assert torch.equal(ag.critic_target_network.flat, want)
    ag._soft_update_target_network(ag.actor_target_network, ag.actor_network)
    assert torch.equal(ag.actor_target_network.flat, (1 - a.polyak) * ag.actor_network.flat + a.polyak * at)
    with pytest.raises(ValueError):
        ag._soft_update_target_network(ag.actor_target_network, None)
