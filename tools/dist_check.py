"""Staged 2+-rank check of the NCCL path (run under torchrun); prints progress so a hang can be located."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_arm_under_sparse_reward_b200 import utils


def log(*a):
    # ONE write per line: two ranks share the pipe and print() emits its arguments piecewise
    sys.stdout.write("[rank %s %.1fs] %s\n" % (os.environ.get("RANK"), time.time() - T0, " ".join(str(x) for x in a)))
    sys.stdout.flush()


T0 = time.time()
rank, world = utils.init_comm()
log("init_comm ok", rank, world, torch.cuda.current_device())
t = torch.full((1000,), float(rank + 1), device="cuda")
utils.allreduce_sum_(t)
torch.cuda.synchronize()
log("allreduce ok", t[0].item())
p = torch.full((10,), float(rank), device="cuda")
utils.bcast_(p, 0)
torch.cuda.synchronize()
log("bcast ok", p[0].item())
from rl_arm_under_sparse_reward_b200.arguments import Args
from rl_arm_under_sparse_reward_b200.bmirobot_env.vec_env import BmiVecEnv
from rl_arm_under_sparse_reward_b200.ddpg_agent import ddpg_agent
from rl_arm_under_sparse_reward_b200.train import get_env_params
a = Args()
a.add_demo, a.verbose, a.n_envs, a.buffer_size, a.save_dir = False, False, 64, 1024 * 100, "/tmp/bmi_dc_%d/" % rank
torch.manual_seed(a.seed + rank)
env = BmiVecEnv(a.n_envs, seed=a.seed + rank)
a.p2p_adam = "--nccl" not in sys.argv
agent = ddpg_agent(a, env, get_env_params(env))
torch.cuda.synchronize()
log("p2p adam attached:", agent._p2p)
log("agent ok; params identical across ranks:", agent.actor_network.flat.sum().item())
agent.rollout(0)
agent.buffer.store_episode([agent.ep['obs'], agent.ep['ag'], agent.ep['g'], agent.ep['actions']])
torch.cuda.synchronize()
log("rollout+store ok")
agent._update_normalizer()
torch.cuda.synchronize()
log("normalizer ok", agent.o_norm.total_count)
a.use_cuda_graphs = False
agent.update_many(2)
torch.cuda.synchronize()
log("eager updates ok", agent.losses())
a.use_cuda_graphs = True
agent.update_many(3)
torch.cuda.synchronize()
log("graph capture ok")
agent.update_many(3)
torch.cuda.synchronize()
log("graph replay ok; params", agent.actor_network.flat.sum().item(), agent.critic_network.flat.sum().item(),
    "p2p timed out:", agent.p2p_timed_out())
import hashlib
log("param digest", hashlib.md5(agent.actor_network.flat.cpu().numpy().tobytes() + agent.critic_network.flat.cpu().numpy().tobytes()).hexdigest())
r = agent._eval_agent()
log("eval ok", r)
agent.release_graphs()
log("graphs released")
utils.shutdown_comm()
log("done")
