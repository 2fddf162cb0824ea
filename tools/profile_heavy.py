"""Tail-regime driver for ncu: few envs (about one warp per SM sub-partition) with the hand pressed on the table."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_arm_under_sparse_reward_b200.bmirobot_env.vec_env import BmiVecEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 592
env = BmiVecEnv(n, seed=1)
env.reset()
down = torch.tensor([[0.0, 0.05, -0.5, 0.0]], device="cuda").repeat(n, 1).contiguous()
for t in range(12):
    env.step(down)
torch.cuda.synchronize()
ev = []
for t in range(6):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); env.step(down); e.record(); ev.append((s, e))
torch.cuda.synchronize()
print("heavy env_step ms:", [round(s.elapsed_time(e), 3) for s, e in ev], "EE z", env.obs[:4, 2].tolist())
