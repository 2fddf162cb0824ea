"""Entry point of the training path — mirror of the reference ``train.py:15-60``.

``python -m rl_arm_under_sparse_reward_b200.train`` on one GPU, or one rank per GPU with
``python -m torch.distributed.run --nproc-per-node N -m rl_arm_under_sparse_reward_b200.train``
(the reference uses ``mpirun -np N python train.py``).  Seeding keeps the reference contract
(``seed + rank`` for env / random / numpy / torch — train.py:34-39).
"""
import random

import numpy as np
import torch

from . import utils
from .arguments import Args
from .bmirobot_env.vec_env import BmiVecEnv
from .ddpg_agent import ddpg_agent


def get_env_params(env):
    """train.py:15-23"""
    return {'obs': env.obs_dim, 'goal': env.goal_dim, 'action': env.act_dim, 'action_max': env.action_max,
            'max_timesteps': 100}


def launch(args):
    rank, world = utils.init_comm()
    n_envs = args.n_envs or args.num_rollouts_per_mpi
    env = BmiVecEnv(n_envs, task=args.train_type, seed=args.seed + rank)
    random.seed(args.seed + rank)
    np.random.seed(args.seed + rank)
    torch.manual_seed(args.seed + rank)
    torch.cuda.manual_seed(args.seed + rank)
    env_params = get_env_params(env)
    trainer = ddpg_agent(args, env, env_params)
    trainer.learn()
    trainer.plot_success_rate()
    trainer.release_graphs()
    utils.shutdown_comm()
    return trainer


if __name__ == '__main__':
    launch(Args())
