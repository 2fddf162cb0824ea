"""Run ddpg_agent.learn() on the GPU and record the eval success rate per epoch (the reference's only published result
is this curve: README.md:99, push task, ~0.5 at epoch 10 and ~0.9 at epoch 45, one reference epoch = 10 000 env-steps and
2 000 updates per MPI worker).

    python tools/learning_curve.py --task push --n-batches 40 --cycles-per-epoch 10 --epochs 15 --out gpurun_out/curve.jsonl

Every line of the output: {"epoch", "env_steps", "updates", "success_rate", "wall_s", "updates_per_env_step"}.
"""
import argparse, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_arm_under_sparse_reward_b200 import get_demo_data
from rl_arm_under_sparse_reward_b200.arguments import Args
from rl_arm_under_sparse_reward_b200.bmirobot_env.vec_env import BmiVecEnv
from rl_arm_under_sparse_reward_b200.ddpg_agent import ddpg_agent
from rl_arm_under_sparse_reward_b200.train import get_env_params

ap = argparse.ArgumentParser()
ap.add_argument("--task", default="push")
ap.add_argument("--n-envs", type=int, default=4096)
ap.add_argument("--n-batches", type=int, default=40)
ap.add_argument("--batch-size", type=int, default=256)
ap.add_argument("--cycles-per-epoch", type=int, default=10)
ap.add_argument("--epochs", type=int, default=15)
ap.add_argument("--demos", type=int, default=1000)
ap.add_argument("--seed", type=int, default=125)
ap.add_argument("--out", default="gpurun_out/curve.jsonl")
a_ = ap.parse_args()

a = Args()
a.verbose, a.n_envs, a.n_batches, a.batch_size = False, a_.n_envs, a_.n_batches, a_.batch_size
a.n_epochs, a.n_cycles, a.seed = 1, a_.cycles_per_epoch, a_.seed
a.buffer_size = 65536 * 100
a.eval_all_envs = True
a.save_dir = "/tmp/bmi_curve/"
a.train_type = a_.task
a.add_demo = a_.demos > 0
if a.add_demo:
    demo, rate = get_demo_data.get_demo(a_.task, a_.demos, n_envs=2048, seed=1000, max_batches=24, verbose=False)
    a.demo_name = "/tmp/bmi_curve_demo.npz"
    np.savez_compressed(a.demo_name, acs=demo["acs"], obs=demo["obs"], info=demo["info"], g=demo["g"], ag=demo["ag"])
    print("demos: %d kept (%.1f %% of the scripted episodes)" % (demo["acs"].shape[0], 100 * rate), flush=True)
    a.add_demo = demo["acs"].shape[0] > 0
torch.manual_seed(a.seed)
np.random.seed(a.seed)
env = BmiVecEnv(a.n_envs, task=a_.task, seed=a.seed)
agent = ddpg_agent(a, env, get_env_params(env))
os.makedirs(os.path.dirname(a_.out) or ".", exist_ok=True)
t0 = time.time()
with open(a_.out, "w") as f:
    for epoch in range(a_.epochs):
        agent.learn()                      # one epoch: cycles_per_epoch x (rollout, store, normaliser, n_batches updates, Polyak) + eval + checkpoint
        rec = {"epoch": epoch, "env_steps": agent.env_steps, "updates": agent.updates, "success_rate": agent.success_rates[-1],
               "wall_s": time.time() - t0, "updates_per_env_step": a.n_batches / (a.n_envs * 100.0), "losses": agent.losses().tolist()}
        f.write(json.dumps(rec) + "\n"); f.flush()
        print(json.dumps(rec), flush=True)
