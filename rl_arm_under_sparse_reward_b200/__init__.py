"""B200-native (sm_100a) rollout -> HER-relabel -> DDPG-update path of
PiggyCh/RL_arm_under_sparse_reward behind the reference's own Python API.

Public names mirror the reference modules: ``her.her_sampler``,
``replay_buffer.replay_buffer``, ``normalizer.normalizer``, ``models.actor/critic``,
``utils.sync_networks/sync_grads``, ``ddpg_agent.ddpg_agent``, ``arguments.Args`` and
``bmirobot_env.bmirobot_push_F.bmirobotGympushEnv``.  All arithmetic runs in
``_C/libbmi_b200.so`` (C-ABI in include/bmi.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
