"""Time per DDPG update (HER draw + fused sampler + backward + Adam, graph-replayed like learn()) on one GPU.
    python tools/time_update.py [n_updates]            # BMI_DDPG_CUBLAS=1 selects the cuBLASLt chain"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_arm_under_sparse_reward_b200 import _lib
from rl_arm_under_sparse_reward_b200.arguments import Args
from rl_arm_under_sparse_reward_b200.bmirobot_env.vec_env import BmiVecEnv
from rl_arm_under_sparse_reward_b200.ddpg_agent import ddpg_agent
from rl_arm_under_sparse_reward_b200.train import get_env_params

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
a = Args()
a.add_demo, a.verbose, a.n_envs, a.buffer_size, a.save_dir = False, False, 256, 4096 * 100, "/tmp/bmi_time_u/"
torch.manual_seed(125)
env = BmiVecEnv(a.n_envs, seed=125)
ag = ddpg_agent(a, env, get_env_params(env))
ag.rollout(0)
ag.buffer.store_episode([ag.ep['obs'], ag.ep['ag'], ag.ep['g'], ag.ep['actions']])
ag._update_normalizer()
for _ in range(3):
    ag.update_many(n)
torch.cuda.synchronize()
n0 = _lib.launch_count()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
s.record()
for _ in range(reps):
    ag.update_many(n)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / reps
print("path=%s  %d updates: %.3f ms = %.1f us per update (launch count of the eager region since warm-up: %d)" % (
    "cublaslt" if os.environ.get("BMI_DDPG_CUBLAS") == "1" else "fused", n, ms, 1e3 * ms / n, _lib.launch_count() - n0), "losses", ag.losses())
