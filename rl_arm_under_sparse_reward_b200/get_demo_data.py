"""Scripted-controller demonstrations on the vectorised CUDA env.

Mirror of the reference ``get_demo_data_push.py:24-94`` and ``get_demo_data_pick.py:34-100``: the same open-loop /
proportional controllers (phase boundaries at the reference's step counts), episodes of ``max_timesteps`` = 100, only
episodes whose LAST step reports ``is_success`` are kept, and the result is written in the reference's on-disk format
(``np.savez_compressed`` with ``acs (N,100,4) obs (N,101,27) info (N,100) g (N,100,3) ag (N,101,3)``, float64, ``info``
an object array of ``{'is_success': float32}`` dicts) so that ``ddpg_agent._init_demo_buffer`` (ddpg_agent.py:82-90) of
either implementation loads it.  What differs: ``n_envs`` episodes run simultaneously, one ``bmi_env_step`` launch per
step, and the controller is evaluated on the device for all of them at once.

    python -m rl_arm_under_sparse_reward_b200.get_demo_data --task push --demo-num 1000 --n-envs 1024
"""
import argparse

import numpy as np
import torch

from .bmirobot_env.vec_env import BmiVecEnv

HOME = (0.241, 0.3265, 0.294)   # the reset pose of the gripper the push script returns to (get_demo_data_push.py:52)


def push_controller(step_time, obs, g):
    """get_demo_data_push.py:40-62 for a batch: obs (n,27), g (n,3) CUDA tensors -> actions (n,4)."""
    n, dev = obs.shape[0], obs.device
    grip, blk = obs[:, :3], obs[:, 12:15]
    zero = torch.zeros(n, 1, device=dev)
    if step_time <= 10:
        a = torch.tensor([0.0, -0.1, 0.1, 0.0], device=dev).repeat(n, 1)
    elif step_time <= 20 or 60 < step_time <= 80:      # behind the block, on the far side from the goal
        a = torch.cat([(g - blk) * (-0.5) + blk - grip, zero], 1)
    elif step_time <= 40 or step_time > 80:            # push towards the goal
        a = torch.cat([g - blk, zero], 1)
    else:                                              # back to the reset pose
        a = torch.cat([torch.tensor(HOME, device=dev) - grip, zero], 1)
    done = (blk - g).norm(dim=1) < 0.05                # get_demo_data_push.py:60-62
    return torch.where(done[:, None], torch.zeros_like(a), a).float().contiguous()


def pick_controller(step_time, obs, g):
    """get_demo_data_pick.py:53-68 for a batch."""
    n, dev = obs.shape[0], obs.device
    grip, blk = obs[:, :3], obs[:, 12:15]
    const = lambda v: torch.tensor(v, device=dev).repeat(n, 1)
    if step_time <= 10:
        a = const([0.0, -0.1, 0.1, 0.0])
    elif step_time <= 30:
        a = torch.cat([blk - grip + torch.tensor([0.0, -0.2, 0.1], device=dev), torch.zeros(n, 1, device=dev)], 1)
    elif step_time <= 50:
        a = const([0.0, 0.0, 0.0, 0.1])
    elif step_time <= 70:
        a = torch.cat([blk - grip + torch.tensor([0.0, -0.05, 0.05], device=dev), torch.zeros(n, 1, device=dev)], 1)
    elif step_time <= 90:
        a = const([0.0, 0.0, 0.0, -0.1])
    else:
        a = torch.cat([g - blk, torch.zeros(n, 1, device=dev)], 1)
    return a.float().contiguous()


def run_scripted_batch(env, controller, T=100):
    """One batch of env.n_envs scripted episodes.  Returns device tensors obs (n,T+1,27) ag (n,T+1,3) g (n,T,3)
    acs (n,T,4) success (n,T) — success[:, t] is info['is_success'] of step t."""
    n, dev = env.n_envs, env.device
    obs_b = torch.empty((n, T + 1, 27), device=dev)
    ag_b = torch.empty((n, T + 1, 3), device=dev)
    g_b = torch.empty((n, T, 3), device=dev)
    act_b = torch.empty((n, T, 4), device=dev)
    suc_b = torch.empty((n, T), device=dev)
    obs, ag, g = env.reset()
    for t in range(T):
        a = controller(t + 1, obs, g)
        obs_b[:, t], ag_b[:, t], g_b[:, t], act_b[:, t] = obs, ag, g, a
        obs, ag, _, s = env.step(a)
        suc_b[:, t] = s
    obs_b[:, T], ag_b[:, T] = obs, ag
    return obs_b, ag_b, g_b, act_b, suc_b


def get_demo(task="push", demo_num=1000, n_envs=1024, seed=125, max_batches=64, verbose=True):
    """Collect `demo_num` successful episodes (get_push_demo / get_demo_data of the reference).  Returns the dict that
    is written to disk plus the fraction of attempted episodes that were kept."""
    env = BmiVecEnv(n_envs, task=task, seed=seed)
    controller = push_controller if task == "push" else pick_controller
    keep = {k: [] for k in ("acs", "obs", "info", "g", "ag")}
    kept = tried = 0
    for _ in range(max_batches):
        if kept >= demo_num:
            break
        obs_b, ag_b, g_b, act_b, suc_b = run_scripted_batch(env, controller)
        ok = suc_b[:, -1] == 1.0                       # `if info['is_success'] == 1.0` after the last step
        tried += n_envs
        idx = torch.nonzero(ok).flatten()[: demo_num - kept]
        kept += int(idx.numel())
        if idx.numel():
            keep["acs"].append(act_b[idx].double().cpu().numpy())
            keep["obs"].append(obs_b[idx].double().cpu().numpy())
            keep["g"].append(g_b[idx].double().cpu().numpy())
            keep["ag"].append(ag_b[idx].double().cpu().numpy())
            s = suc_b[idx].cpu().numpy()
            info = np.empty(s.shape, dtype=object)
            for i in range(s.shape[0]):
                for t in range(s.shape[1]):
                    info[i, t] = {'is_success': np.float32(s[i, t])}
            keep["info"].append(info)
        if verbose:
            print("This is %d savetime (%d episodes tried)" % (kept, tried))
    shapes = {"acs": (0, 100, 4), "obs": (0, 101, 27), "g": (0, 100, 3), "ag": (0, 101, 3)}
    out = {k: (np.concatenate(v) if v else (np.empty((0, 100), dtype=object) if k == "info" else np.zeros(shapes[k])))
           for k, v in keep.items()}
    return out, (kept / tried if tried else 0.0)


def save_demo(out, task):
    """File name and container of the reference (get_demo_data_push.py:90-93)."""
    name = "bmirobot_%d_%s_demo.npz" % (out["acs"].shape[0], task)
    np.savez_compressed(name, acs=out["acs"], obs=out["obs"], info=out["info"], g=out["g"], ag=out["ag"])
    return name


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", default="push", choices=["push", "pick"])
    ap.add_argument("--demo-num", type=int, default=1000)
    ap.add_argument("--n-envs", type=int, default=1024)
    ap.add_argument("--seed", type=int, default=125)
    a = ap.parse_args()
    demo, rate = get_demo(a.task, a.demo_num, a.n_envs, a.seed)
    print("kept %.1f %% of the scripted episodes; wrote %s" % (100 * rate, save_demo(demo, a.task)))
