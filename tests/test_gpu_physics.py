"""GPU parity of the fp32 articulated-body + IK kernel against the double-precision C oracle on identical
states/actions, and against what the reference's recorded trajectories pin (tests/golden/physics_golden.npz).

Stated tolerances (fp32 kernel vs fp64 oracle, same algorithm):
  one env-step from an identical state:  |obs error| <= 2e-3 (positions/angles/velocities as mixed units)
  for >= 95 % of the envs (contact-set flips between fp32/fp64 near a threshold are discrete events);
  10-step open-loop rollouts: median position error <= 1e-3 m.
"""
import os

import numpy as np
import pytest
import torch

from oracle.physics_oracle import OracleEnv

pytestmark = pytest.mark.gpu


def _env(n, task="push", seed=125):
    from rl_arm_under_sparse_reward_b200.bmirobot_env.vec_env import BmiVecEnv
    return BmiVecEnv(n, task=task, seed=seed)


def test_reset_pose_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "physics_golden.npz"))
    env = _env(4)
    init = np.tile(g["push_init"], (4, 1)).astype(np.float32)
    obs, ag, goal = env.reset(init=torch.as_tensor(init))
    o = obs.cpu().numpy()
    # EE (0.241, 0.3265, 0.294), euler (0, 0, pi/2), zero velocities: exact in the reference (std 0 over 1000 episodes)
    assert np.abs(o[:, :12] - g["push_obs"][0, :12]).max() < 2e-6
    assert np.abs(o[:, 12:15] - g["push_obs"][0, 12:15]).max() < 1e-6
    assert np.abs(o[:, 18:21] - g["push_obs"][0, 18:21]).max() < 2e-6
    assert np.array_equal(ag.cpu().numpy(), o[:, 12:15]) and np.allclose(goal.cpu().numpy(), g["push_init"][4:7])


@pytest.mark.parametrize("task", ["push", "pick"])
def test_block_settle_transient_matches_reference_golden(golden_dir, task):
    """block drop / depenetration transient recorded by the reference env (pins dt, g, ERP 0.08, slop)."""
    g = np.load(os.path.join(golden_dir, "physics_golden.npz"))
    env = _env(1, task=task)
    env.reset(init=torch.as_tensor(g[task + "_init"][None].astype(np.float32)))
    for t in range(6):
        obs, _, _, _ = env.step(torch.as_tensor(g[task + "_acs"][t][None].astype(np.float32)).cuda())
        o = obs.cpu().numpy()[0]
        assert abs(o[14] - g[task + "_obs"][t + 1, 14]) < 5e-6, (t, o[14], g[task + "_obs"][t + 1, 14])
        assert abs(o[23] - g[task + "_obs"][t + 1, 23]) < 5e-5, (t, o[23], g[task + "_obs"][t + 1, 23])


def _kernel_oracle(task=0):
    """the oracle with the kernel's geometry: self-collision pairs from the same baked tables the kernel reads (the
    table-vs-GJK/EPA deviation is measured on the CPU, tests/test_selfcol_tables.py), the kernel's lane budget and its
    compressed solver schedule (deviation from Bullet's plain loop: tests/test_oracle_physics.py)"""
    return OracleEnv(task).kernel_mode()


def _oracle_rollout(task, init, acts, state=None):
    o = _kernel_oracle({"push": 0, "pick": 1}[task])
    o.reset(init)
    if state is not None:
        o.set_state(state)
    out = []
    for a in acts:
        obs, _, r, s = o.step(a)
        out.append((obs, r, s))
    return out, o.get_state()


def _oracle_sensitivity(task, init, st, act, ref, n_pert=4, seed=1):
    """max |oracle(step from st + 1e-6 perturbation) - oracle(step from st)|: large where st sits on a discontinuity"""
    rs = np.random.RandomState(seed)
    sens = 0.0
    for _ in range(n_pert):
        o = _kernel_oracle(0 if task == "push" else 1)
        o.reset(init)
        s2 = st.copy()
        s2[:9] += rs.uniform(-1e-6, 1e-6, 9)             # joint angles: a few float32 ulps
        s2[27:30] += rs.uniform(-1e-6, 1e-6, 3)          # block position: a micrometre
        o.set_state(s2)
        w2, _, _, _ = o.step(act)
        sens = max(sens, float(np.abs(w2 - ref).max()))
    return sens


def _unexplained_outliers(task, init, st, act, want, errs, tol):
    """envs whose kernel-vs-oracle error exceeds tol although the oracle itself is smooth there (error > 10 x sensitivity)"""
    bad = []
    for e in np.nonzero(errs > tol)[0]:
        sens = _oracle_sensitivity(task, init[e], st[e], act[e], want[e])
        if errs[e] > 10.0 * sens:
            bad.append((int(e), float(errs[e]), sens))
    return bad


@pytest.mark.parametrize("task", ["push", "pick"])
def test_single_step_vs_oracle_from_random_states(task):
    n = 64
    env = _env(n, task=task, seed=7)
    env.reset()
    rng = np.random.RandomState(0)
    # move away from the reset pose first (5 steps), then compare ONE step from the exact fp32 state
    for t in range(5):
        env.step(torch.as_tensor(rng.uniform(-0.3, 0.3, (n, 4)).astype(np.float32)).cuda())
    st = env.get_state().cpu().numpy().astype(np.float64)
    init = env.init.cpu().numpy().astype(np.float64)
    act = rng.uniform(-0.5, 0.5, (n, 4)).astype(np.float32)
    obs, ag, r, s = env.step(torch.as_tensor(act).cuda())
    got = obs.cpu().numpy()
    errs, wants = [], []
    for e in range(n):
        (res,), _ = _oracle_rollout(task, init[e], [act[e]], state=st[e])
        errs.append(np.abs(got[e] - res[0]).max())
        wants.append(res[0])
        assert r[e].item() == res[1] and s[e].item() == res[2] or errs[-1] > 1e-4
    errs = np.array(errs)
    # measured: median 1e-4, p90 2e-4 .. 3e-4, 95 % of the envs within 2e-3; the rest sit on discontinuities -- the wrist
    # pair's kink (the two hull features of the link6 x link8 penetration depth swap within one 5 mrad table cell) or a
    # contact that opens / closes -- where the ORACLE itself moves by as much when its joint angles are perturbed by
    # 1e-6 rad (tests/diag/diag_worst_env.py prints both columns): discrete events, not drift
    print("one step vs oracle (%s): median %.2e, p90 %.2e, max %.2e" % (task, np.median(errs), np.percentile(errs, 90), errs.max()))
    assert np.mean(errs <= 2e-3) >= 0.85, np.sort(errs)[-8:]
    assert np.mean(errs <= 1e-3) >= 0.90, np.sort(errs)[-8:]     # north_star's 1e-3 (observations are O(0.1 .. 1)): measured 95 %
    assert np.median(errs) <= 2e-4, np.median(errs)
    # ... and the exempt envs are justified one by one: an error above 2e-3 is only accepted where the oracle's own answer
    # moves by at least a tenth of it under a 1e-6 perturbation of the state (at most two unexplained envs of 64)
    bad = _unexplained_outliers(task, init, st, act, wants, errs, 2e-3)
    assert len(bad) <= 2, bad


def test_single_step_vs_oracle_from_contact_rich_states():
    """Same one-step comparison, from states where the hand is pressed onto the table next to / against the block
    (arm-table and arm-block contacts, friction cones on arm links, joint-limit rows): the rows that couple the arm
    and the block in the solver's table.  Contact-set flips between fp32 and fp64 are more frequent here, so the
    bound is looser: <= 5e-3 for >= 85 % of the envs, median <= 5e-4."""
    n = 64
    env = _env(n, seed=21)
    obs, ag, g = env.reset()
    dev = obs.device
    rng = np.random.RandomState(3)
    for t in range(30):   # drive the hand to the block at table height, then push down and sideways
        grip, blk = obs[:, :3], obs[:, 12:15]
        tgt = blk + torch.tensor([0.0, 0.0, 0.01], device=dev)
        a = torch.cat([(tgt - grip).clamp(-0.2, 0.2), torch.zeros(n, 1, device=dev)], 1)
        a[:, 2] -= 0.05 * (t > 15)
        a[:, :2] += torch.as_tensor(rng.uniform(-0.05, 0.05, (n, 2)).astype(np.float32), device=dev)
        obs, ag, _, _ = env.step(a.float().contiguous())
    st = env.get_state().cpu().numpy().astype(np.float64)
    init = env.init.cpu().numpy().astype(np.float64)
    act = rng.uniform(-0.3, 0.3, (n, 4)).astype(np.float32)
    act[:, 2] = -np.abs(act[:, 2])
    obs, _, r, s = env.step(torch.as_tensor(act).cuda())
    got = obs.cpu().numpy()
    assert np.isfinite(got).all()
    errs, ncs, wants = [], [], []
    for e in range(n):
        o = _kernel_oracle(0)
        o.reset(init[e])
        o.set_state(st[e])
        want, _, _, _ = o.step(act[e])
        ncs.append(o.stats()[2])
        errs.append(np.abs(got[e] - want).max())
        wants.append(want)
    errs, ncs = np.array(errs), np.array(ncs)
    assert (ncs > 4).mean() > 0.3, ncs          # the scenario really is contact rich (more than the 4 block-table contacts)
    print("one step vs oracle (contact rich): median %.2e, p90 %.2e, max %.2e" % (np.median(errs), np.percentile(errs, 90), errs.max()))
    assert np.mean(errs <= 5e-3) >= 0.85, np.sort(errs)[-12:]
    assert np.median(errs) <= 5e-4, np.median(errs)
    bad = _unexplained_outliers("push", init, st, act, wants, errs, 5e-3)    # see the random-state test
    assert len(bad) <= 2, bad


def test_arm_trajectory_of_reference_episode0(golden_dir):
    """The CUDA kernel itself against the reference's recorded fresh-process trajectory (episode 0, 10 steps x 12 arm
    dims, identical in the push and pick demo files): same per-horizon bound as the oracle's CPU test + 1.5 mm for
    fp32 and the 5 mrad pair tables; the wrist hold angle and the elbow stall are reproduced."""
    from test_oracle_physics import ARM_TOL, GOLD_Q146
    g = np.load(os.path.join(golden_dir, "physics_golden.npz"))
    for task in ("push", "pick"):
        env = _env(2, task=task)
        env.reset(init=torch.as_tensor(np.tile(g[task + "_init"], (2, 1)).astype(np.float32)))
        qs = []
        for t in range(10):
            obs, _, _, _ = env.step(torch.as_tensor(np.tile(g[task + "_acs"][t], (2, 1)).astype(np.float32)).cuda())
            o = obs.cpu().numpy()
            assert np.array_equal(o[0], o[1])
            assert np.abs(o[0, :3] - g[task + "_obs"][t + 1, :3]).max() < ARM_TOL[t] + 0.0015, (task, t, o[0, :3])
            qs.append(env.get_state().cpu().numpy()[0, [0, 3, 5]])
        qs = np.array(qs)
        assert np.abs(qs[1:, 2] - 0.1975).max() < 0.008, qs[:, 2]
        assert np.abs(qs[6:, 1] - GOLD_Q146[6:, 1]).max() < 0.006, qs[:, 1]


def test_contact_lane_budget_drop_rate():
    """contacts dropped by the solver's lane budget (9 contacts, 6 on arm links) during a scripted push episode of 256
    envs: counted by the kernel (bmi_env_contact_drops), reported per env sub-step"""
    n = 256
    env = _env(n, seed=9)
    obs, ag, g = env.reset()
    env.contact_drops(reset=True)
    for t in range(1, 61):
        grip, blk = obs[:, :3], obs[:, 12:15]
        if t <= 10:
            a = torch.tensor([0, -0.1, 0.1, 0.0], device=obs.device).repeat(n, 1)
        elif t <= 20:
            a = torch.cat([(g - blk) * (-0.5) + blk - grip, torch.zeros(n, 1, device=obs.device)], 1)
        else:
            a = torch.cat([g - blk, torch.zeros(n, 1, device=obs.device)], 1)
        obs, ag, r, s = env.step(a.float().contiguous())
    drops = env.contact_drops()
    rate = drops / (n * 60 * 20)
    print("contacts dropped per env sub-step: %.4f" % rate)
    assert rate < 0.2, rate      # measured 0.097 (what the budget costs against the unbounded oracle)


def test_ten_step_rollout_vs_oracle():
    n = 32
    env = _env(n, seed=11)
    env.reset()
    init = env.init.cpu().numpy().astype(np.float64)
    rng = np.random.RandomState(1)
    acts = rng.uniform(-0.2, 0.2, (10, n, 4)).astype(np.float32)
    for t in range(10):
        obs, _, _, _ = env.step(torch.as_tensor(acts[t]).cuda())
    got = obs.cpu().numpy()
    err = []
    for e in range(n):
        res, _ = _oracle_rollout("push", init[e], acts[:, e])
        err.append(np.abs(got[e, :3] - res[-1][0][:3]).max())
    assert np.median(err) <= 1e-3, np.sort(err)


def test_state_roundtrip_and_determinism():
    env = _env(16, seed=3)
    env.reset()
    a = torch.as_tensor(np.random.RandomState(2).uniform(-0.3, 0.3, (16, 4)).astype(np.float32)).cuda()
    env.step(a)
    st = env.get_state().clone()
    o1 = env.step(a)[0].clone()
    env.set_state(st)
    o2 = env.step(a)[0].clone()
    assert torch.equal(o1, o2)              # bit-reproducible
    assert torch.isfinite(o1).all()


def test_scripted_push_moves_block_and_reward_is_consistent():
    """closed-loop regression: the reference's scripted controller (get_demo_data_push.py:40-58) pushes the block."""
    n = 64
    env = _env(n, seed=5)
    obs, ag, g = env.reset()
    start = ag.clone()
    for t in range(1, 101):
        grip, blk = obs[:, :3], obs[:, 12:15]
        if t <= 10:
            a = torch.tensor([0, -0.1, 0.1, 0.0], device=obs.device).repeat(n, 1)
        elif t <= 20 or 60 < t <= 80:
            a = torch.cat([(g - blk) * (-0.5) + blk - grip, torch.zeros(n, 1, device=obs.device)], 1)
        elif t <= 40 or t > 80:
            a = torch.cat([g - blk, torch.zeros(n, 1, device=obs.device)], 1)
        else:
            a = torch.cat([torch.tensor([0.241, 0.3265, 0.294], device=obs.device) - grip, torch.zeros(n, 1, device=obs.device)], 1)
        a = torch.where(((blk - g).norm(dim=1) < 0.05)[:, None], torch.zeros_like(a), a)
        obs, ag, r, s = env.step(a.float().contiguous())
    d = (ag - g).norm(dim=1)
    assert torch.equal(s, (d < 0.05).float()) and torch.equal(r, -(d > 0.05).float())
    moved = (ag - start).norm(dim=1)
    assert (moved > 0.02).float().mean() > 0.5          # the hand reaches and displaces most blocks
    assert torch.isfinite(obs).all() and (ag[:, 2] > 0.15).all()   # nothing fell through the table


def test_gym_style_wrapper_surface():
    import random
    from rl_arm_under_sparse_reward_b200.bmirobot_env.bmirobot_push_F import bmirobotGympushEnv
    random.seed(125)
    env = bmirobotGympushEnv()
    assert env.action_space.shape == (4,) and env.action_space.high[0] == 0.5
    o = env.reset()
    assert set(o) == {'observation', 'achieved_goal', 'desired_goal'} and o['observation'].shape == (27,)
    assert o['observation'].dtype == np.float64 and np.allclose(o['observation'][:3], [0.241, 0.3265, 0.294], atol=1e-5)
    assert np.array_equal(o['desired_goal'], env.goal)
    o2, r, done, info = env.step(np.array([0.9, -0.1, 0.1, 1.0]))      # clipped to +-0.5, action[3] forced to 0
    assert done is False and r in (-1.0, 0.0) and info['is_success'] in (0.0, 1.0)
    assert env.compute_reward(o2['achieved_goal'], o2['desired_goal'], info) == r
    assert env._is_success(o2['achieved_goal'], o2['desired_goal']) == info['is_success']


def test_single_step_vs_oracle_from_gripping_states():
    """Pick-and-place (bmirobot_env_pickandplace_v2.py:92-95: auto-grip when the arm touches the block; finger friction 10 / 1,
    robotarm_description.urdf:370,398): the reference's scripted pick controller is run for 80 steps, then ONE step is
    compared with the oracle from the states in which a finger presses on the block (finger-block rows with mu = 5 / 0.5
    next to the fingers' own self-contact row).  Same bound as the contact-rich push test."""
    from rl_arm_under_sparse_reward_b200.get_demo_data import pick_controller
    n = 128
    env = _env(n, task="pick", seed=31)
    obs, ag, g = env.reset()
    for t in range(80):     # the fingers reach the block around step 55-65 and close on it from step 71 (get_demo_data_pick.py:53-68)
        obs, ag, _, _ = env.step(pick_controller(t + 1, obs, g))
    st = env.get_state().cpu().numpy().astype(np.float64)
    init = env.init.cpu().numpy().astype(np.float64)
    act = pick_controller(81, obs, g)
    got = env.step(act)[0].cpu().numpy()
    act = act.cpu().numpy()
    errs, grip, wants = [], [], []
    for e in range(n):
        o = _kernel_oracle(1)
        o.reset(init[e])
        o.set_state(st[e])
        c = o.contacts()
        fingers = [r for r in c if r[2] == 1 and r[1] >= 7 and r[3] < 1e-3]      # block x hand1 / hand2, touching
        if not fingers:
            continue
        want, _, _, _ = o.step(act[e])
        grip.append(e)
        errs.append(np.abs(got[e] - want).max())
        wants.append(want)
    errs = np.array(errs)
    print("gripping states: %d of %d envs; one-step error median %.2e p90 %.2e" % (len(grip), n, np.median(errs), np.percentile(errs, 90)))
    assert len(grip) >= 10, len(grip)
    assert np.mean(errs <= 5e-3) >= 0.8, np.sort(errs)[-12:]
    assert np.median(errs) <= 1e-3, np.median(errs)
    # the exempt envs sit on discontinuities of the oracle itself (see test_single_step_vs_oracle_from_random_states)
    bad = _unexplained_outliers("pick", init[grip], st[grip], act[grip], wants, errs, 5e-3)
    print("gripping states: %d envs beyond 5e-3, unexplained by the oracle's own sensitivity: %s" % (int((errs > 5e-3).sum()), bad))
    assert len(bad) <= 3, bad


def test_scripted_pick_success_rate():
    """closed loop: the reference's scripted pick controller (get_demo_data_pick.py:53-68) lifts the block to the goal in a
    fraction of the episodes (the reference kept 1000 of an unrecorded number of attempts); recorded in DESIGN.md"""
    from rl_arm_under_sparse_reward_b200.get_demo_data import pick_controller, run_scripted_batch
    env = _env(512, task="pick", seed=41)
    obs_b, ag_b, g_b, act_b, suc_b = run_scripted_batch(env, pick_controller)
    rate = float((suc_b[:, -1] == 1.0).float().mean())
    lifted = float((ag_b[:, -1, 2] > 0.25).float().mean())
    print("scripted pick: success %.3f, block above z = 0.25 at the end in %.3f of 512 episodes" % (rate, lifted))
    assert torch.isfinite(obs_b).all()
    assert lifted > 0.02 or rate > 0.01


def _replay_recordings(task, golden_dir):
    """Open-loop replay of the reference's recorded episodes by the CUDA kernel itself: recorded actions from the recorded
    reset, one env per (episode, block yaw) -- the reference does not record the block's initial yaw, so 8 yaws are tried
    and the one with the best final block position counts (same protocol as tests/diag/replay_reference.py for the oracle)."""
    d = np.load(os.path.join(golden_dir, "demo_small.npz" if task == "push" else "demo_small_pick.npz"))
    obs, acs, g = d["obs"], d["acs"], d["g"][:, 0]
    E = obs.shape[0]
    yaws = np.linspace(1.57, 4.71, 9)[:-1]
    n = E * len(yaws)
    init = np.zeros((n, 8), np.float32)
    for e in range(E):
        for k, y in enumerate(yaws):
            init[e * len(yaws) + k] = [obs[e, 0, 12], obs[e, 0, 13], 0.2, y, g[e, 0], g[e, 1], g[e, 2], 0]
    env = _env(n, task=task, seed=1)
    env.reset(init=torch.as_tensor(init).cuda())
    ee = np.zeros((n, 100))
    for t in range(100):
        a = np.repeat(acs[:, t], len(yaws), axis=0).astype(np.float32)
        o, ag, r, s = env.step(torch.as_tensor(a).cuda())
        o = o.cpu().numpy()
        ee[:, t] = np.abs(o[:, :3] - np.repeat(obs[:, t + 1, :3], len(yaws), axis=0)).max(axis=1)
    blk = np.linalg.norm(o[:, 12:15] - np.repeat(obs[:, 100, 12:15], len(yaws), axis=0), axis=1).reshape(E, len(yaws))
    best = blk.argmin(axis=1)
    ee = ee.reshape(E, len(yaws), 100)
    ee_best = np.array([ee[e, best[e]] for e in range(E)])
    return ee_best, blk.min(axis=1)


@pytest.mark.parametrize("task", ["push", "pick"])
def test_kernel_replays_reference_recordings(task, golden_dir):
    """The CUDA env against the REFERENCE'S OWN recordings (not against the oracle): 16 push / 8 pick episodes of
    bmirobot_1000_*_demo.npz, 100 env-steps open loop.  Measured (final round-2 kernel): push -- EE within 14.0 mm over the
    first 10 env-steps in every episode, within 35 mm over all 100 steps in 11 of 16 episodes, median final block error
    22.9 mm; pick -- 16.3 mm, 5 of 8, 34.3 mm (the oracle's faithful mode: 12 of 16 / 24.0 mm and 5 of 8 / 37.7 mm,
    profiles/r02_reference_replay.md; contact-rich episodes diverge chaotically, all but episode 0 start from a leaked
    solver state in the reference).  The bounds below leave room for one or two episodes to flip."""
    ee, blk = _replay_recordings(task, golden_dir)
    whole = (ee.max(axis=1) < 0.035)
    print("%s: EE max error steps 1-10: %.1f mm (worst episode); EE within 35 mm over the whole episode in %d of %d; "
          "median final block error %.1f mm" % (task, 1e3 * ee[:, :10].max(), whole.sum(), len(whole), 1e3 * np.median(blk)))
    first10, n_whole, med_blk = {"push": (0.018, 8, 0.035), "pick": (0.020, 3, 0.060)}[task]
    assert ee[:, :10].max() < first10, ee[:, :10].max(axis=1)
    assert whole.sum() >= n_whole, ee.max(axis=1)
    assert np.median(blk) <= med_blk, np.sort(blk)
