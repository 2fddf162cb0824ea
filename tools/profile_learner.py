"""Driver for ncu captures of the learner-side kernels: a few eager (un-graphed) DDPG updates at batch 256 and the fused HER
sampler at batch 65 536 on a 65 536-episode buffer (larger than L2).
    ncu --set full -k regex:'her_inputs_kernel|drelu_bgrad_kernel|adam_kernel' -c 9 -o gpurun_out/r2_learner python tools/profile_learner.py"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_arm_under_sparse_reward_b200 import _lib
from rl_arm_under_sparse_reward_b200.arguments import Args
from rl_arm_under_sparse_reward_b200.bmirobot_env.vec_env import BmiVecEnv
from rl_arm_under_sparse_reward_b200.ddpg_agent import ddpg_agent
from rl_arm_under_sparse_reward_b200.train import get_env_params
a = Args()
a.add_demo, a.verbose, a.n_envs, a.buffer_size, a.save_dir, a.use_cuda_graphs = False, False, 256, 65536 * 100, "/tmp/bmi_prof_l/", False
torch.manual_seed(125)
env = BmiVecEnv(a.n_envs, seed=125)
ag = ddpg_agent(a, env, get_env_params(env))
ag.rollout(0)
ag.buffer.store_episode([ag.ep['obs'], ag.ep['ag'], ag.ep['g'], ag.ep['actions']])
ag._update_normalizer()
ag.update_many(3)                       # her_draw + her_inputs (3 x 256) + 3 x (backward + adam)
torch.cuda.synchronize()
# the fused sampler at a bandwidth-relevant batch: fill the whole buffer with random episodes first
b = ag.buffer
for k in b.buffers:
    b.buffers[k].normal_()
b.current_size = b.size
b.current_size_dev.fill_(b.size)
n = 256
ag.args.batch_size = 256
ag._XA = None
ag._sample_batches(n)                   # ONE her_inputs launch of 65 536 samples
torch.cuda.synchronize()
print("learner profile driver done; losses", ag.losses())
