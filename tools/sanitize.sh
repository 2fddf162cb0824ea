#!/bin/bash
# compute-sanitizer memcheck + racecheck of the env kernels and the fused HER sampler at 64 envs.
# usage (GPU box): bash tools/sanitize.sh gpurun_out/r2_sanitizer
out=${1:-gpurun_out/sanitizer}
cat > /tmp/bmi_san.py <<'PY'
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from rl_arm_under_sparse_reward_b200.arguments import Args
from rl_arm_under_sparse_reward_b200.bmirobot_env.vec_env import BmiVecEnv
from rl_arm_under_sparse_reward_b200.ddpg_agent import ddpg_agent
from rl_arm_under_sparse_reward_b200.train import get_env_params
a = Args(); a.add_demo, a.verbose, a.n_envs, a.buffer_size, a.save_dir, a.use_cuda_graphs = False, False, 64, 256 * 100, "/tmp/bmi_san/", False
for task in ("push", "pick"):
    env = BmiVecEnv(64, task=task, seed=3)
    p = get_env_params(env); p['max_timesteps'] = 2
    ag = ddpg_agent(a, env, p)
    ag.rollout(0)                                   # rollout_kernel (policy + IK + physics + record)
    env.reset(); env.step(torch.zeros(64, 4, device="cuda"))   # env_reset_kernel, env_step_kernel
    ag.buffer.store_episode([ag.ep['obs'], ag.ep['ag'], ag.ep['g'], ag.ep['actions']])
    ag._update_normalizer()                         # her_draw, her gather, norm_update
    ag.update_many(2)                               # her_inputs_lane_kernel (4-sample chunks), ddpg_rows / ddpg_wgrad, adam
    ag._sample_batches(80)                          # her_inputs_lane_kernel, 16-sample chunks (20 480 samples)
    ag._soft_update_target_network()
    torch.cuda.synchronize()
    print(task, "ok", ag.losses())
PY
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python /tmp/bmi_san.py > ${out}_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' ${out}_$tool.log | tail -1)"
done
