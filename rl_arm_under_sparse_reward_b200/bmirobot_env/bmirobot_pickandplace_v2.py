"""Drop-in for the reference's ``bmirobot_env/bmirobot_pickandplace_v2.py:7-20`` (same class name as the
push env, as in the reference).  Differences from push (bmirobot_env_pickandplace_v2.py:92-95,116-131):
4x4x8 cm 2 kg block, goal in the air (y in [.3,.55], z in [.3,.5], 7 draws per placement attempt) and the
auto-grip rule (action[3] = -1 while the arm touches the block), all inside the CUDA env.
"""
import random

from .bmirobot_push_F import bmirobotGymEnv as _Base


class bmirobotGymEnv(_Base):
    _task = "pick"

    def _sample_target_yz(self):
        yt = (random.random() * 0.25) + 0.3
        zt = 0.3 + 0.2 * random.random()
        return yt, zt


class bmirobotGympushEnv(bmirobotGymEnv):
    def __init__(self, reward_type='sparse'):
        super().__init__(n_substeps=20, distance_threshold=0.05, reward_type=reward_type)
