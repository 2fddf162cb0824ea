"""Small driver for ncu: N env-steps of 4096 envs with exploration-like actions (eager launches)."""
import sys
import os
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_arm_under_sparse_reward_b200.bmirobot_env.vec_env import BmiVecEnv

n_envs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
env = BmiVecEnv(n_envs, seed=125)
env.reset()
g = torch.Generator(device="cuda").manual_seed(0)
evs = []
for t in range(n_steps):
    u = torch.rand(n_envs, 1, device="cuda", generator=g)
    a = torch.where(u < 0.3, torch.rand(n_envs, 4, device="cuda", generator=g) - 0.5,
                    0.005 * torch.randn(n_envs, 4, device="cuda", generator=g)).contiguous()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    env.step(a)
    e.record()
    evs.append((s, e))
torch.cuda.synchronize()
ms = np.array([s.elapsed_time(e) for s, e in evs])
print("env_step ms: first %.3f  median %.3f  last %.3f  -> %.0f env-steps/s" % (ms[0], np.median(ms), ms[-1], n_envs / (np.median(ms) * 1e-3)))
