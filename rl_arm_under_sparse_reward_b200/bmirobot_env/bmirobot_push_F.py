"""Drop-in for the reference's ``bmirobot_env/bmirobot_push_F.py:8-21``: ``bmirobotGympushEnv`` with the
gym-style ``reset() / step(a) / compute_reward(ag, g, info) / seed(s) / action_space`` surface
(bmirobot_env_push_F.py:84-165), one env instance, numpy dicts at the boundary.  The PyBullet calls are
replaced by the CUDA env (``BmiVecEnv`` with n_envs=1); block/goal placement consumes Python's ``random``
stream in the reference's order (6 draws per attempt, push_F.py:117-132) so ``random.seed`` reproduces
the reference's episode sequence.
"""
import math
import random

import numpy as np
import torch

from .vec_env import BmiVecEnv


class _Box:
    """Minimal stand-in for gym.spaces.Box (gym is not a dependency): shape, low, high, sample()."""

    def __init__(self, low, high):
        self.low = np.asarray(low, dtype=np.float32)
        self.high = np.asarray(high, dtype=np.float32)
        self.shape = self.low.shape
        self.dtype = np.float32

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(np.float32)


class bmirobotGymEnv:
    _task = "push"

    def __init__(self, model_path=None, n_substeps=20, gripper_extra_height=0.0, block_gripper=False, has_object=True,
                 target_in_the_air=False, target_offset=0.0, obj_range=0.15, target_range=0.15,
                 distance_threshold=0.05, initial_qpos=None, reward_type='sparse'):
        if reward_type != 'sparse':
            raise NotImplementedError("only the sparse reward of the reference's configs is implemented")
        self.n_substeps, self.distance_threshold, self.reward_type = n_substeps, distance_threshold, reward_type
        self.has_object, self.n_actions = has_object, 4
        self._action_bound = 0.5
        self.action_space = _Box([-self._action_bound] * 4, [self._action_bound] * 4)
        self._vec = BmiVecEnv(1, task=self._task)
        self._vec.distance_threshold = distance_threshold
        self.goal = np.zeros(3)
        self.seed()
        self.reset()

    # ---- placement (python `random`, reference draw order) --------------------------------------
    def _sample_placement(self):
        for _ in range(100):
            xpos = 0.15 + 0.2 * random.random()
            ypos = (random.random() * 0.3) + 0.2
            zpos = 0.2
            ang = 3.14 * 0.5 + 3.1415925438 * random.random()
            xt = 0.35 * random.random()
            yt, zt = self._sample_target_yz()
            random.random()  # ang_target: drawn by the reference, only used for the marker's orientation
            if math.sqrt((xpos - xt) ** 2 + (ypos - yt) ** 2 + (zpos - zt) ** 2) >= 0.15:
                break
        return [xpos, ypos, zpos, ang, xt, yt, zt, 0.0]

    def _sample_target_yz(self):
        return (random.random() * 0.3) + 0.2, 0.2

    def _dict(self, obs, ag, g):
        return {'observation': obs[0].double().cpu().numpy().copy(), 'achieved_goal': ag[0].double().cpu().numpy().copy(),
                'desired_goal': g[0].double().cpu().numpy().copy()}

    def reset(self):
        init = self._sample_placement()
        obs, ag, g = self._vec.reset(init=torch.tensor([init], dtype=torch.float32))
        out = self._dict(obs, ag, g)
        self.goal = out['desired_goal'].copy()
        return out

    def step(self, action):
        a = np.clip(np.asarray(action, dtype=np.float64), -0.5, 0.5)
        obs, ag, r, s = self._vec.step(torch.as_tensor(a.reshape(1, 4), dtype=torch.float32, device=self._vec.device))
        out = self._dict(obs, ag, self._vec.g)
        info = {'is_success': np.float32(s[0].item())}
        return out, np.float32(r[0].item()), False, info

    def compute_reward(self, achieved_goal, goal, info):
        return self._vec.compute_reward(np.asarray(achieved_goal), np.asarray(goal), info)

    def _is_success(self, achieved_goal, desired_goal):
        d = np.linalg.norm(np.asarray(achieved_goal) - np.asarray(desired_goal), axis=-1)
        return (d < self.distance_threshold).astype(np.float32)

    def seed(self, seed=None):
        self.np_random = np.random.RandomState(seed)  # like the reference, nothing consumes it
        return [seed]


class bmirobotGympushEnv(bmirobotGymEnv):
    def __init__(self, reward_type='sparse'):
        self.maxtimesteps = 150
        super().__init__(n_substeps=20, distance_threshold=0.05, reward_type=reward_type)
