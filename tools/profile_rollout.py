"""Driver for ncu: a few fused rollouts of 4096 envs (random-init policy + exploration)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_arm_under_sparse_reward_b200.arguments import Args
from rl_arm_under_sparse_reward_b200.bmirobot_env.vec_env import BmiVecEnv
from rl_arm_under_sparse_reward_b200.ddpg_agent import ddpg_agent
from rl_arm_under_sparse_reward_b200.train import get_env_params
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
T = int(sys.argv[3]) if len(sys.argv) > 3 else 100
a = Args()
a.add_demo, a.verbose, a.n_envs, a.buffer_size, a.save_dir = False, False, n, 8192 * 100, "/tmp/bmi_prof/"
torch.manual_seed(125)
env = BmiVecEnv(n, seed=125)
p = get_env_params(env); p['max_timesteps'] = T
ag = ddpg_agent(a, env, p)
for i in range(reps):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); ag.rollout(0); e.record(); torch.cuda.synchronize()
    print("rollout %d: %.1f ms -> %.0f env-steps/s" % (i, s.elapsed_time(e), n * T / (s.elapsed_time(e) * 1e-3)))
