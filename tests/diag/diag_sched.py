"""Diagnostic (GPU): one env-step from identical random states, CUDA kernel vs the oracle in kernel mode, for the plain
(K = 0) and the compressed (K = 2) solver schedule on either side.  Prints error quantiles of the 27-float observation."""
import os, sys, tempfile
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.physics_oracle import OracleEnv, MODEL
from rl_arm_under_sparse_reward_b200.bmirobot_env.vec_env import BmiVecEnv

n = 64
blob = np.fromfile(MODEL, dtype="<f4")
paths = {}
for K in (0.0, 2.0):
    b = blob.copy(); b[56] = K
    f = tempfile.NamedTemporaryFile(suffix=".bin", delete=False); b.tofile(f.name); paths[K] = f.name
rng = np.random.RandomState(0)
env = BmiVecEnv(n, seed=7, model_path=paths[0.0])
env.reset()
for t in range(5):
    env.step(torch.as_tensor(rng.uniform(-0.3, 0.3, (n, 4)).astype(np.float32)).cuda())
st = env.get_state().clone()
init = env.init.cpu().numpy().astype(np.float64)
act = rng.uniform(-0.5, 0.5, (n, 4)).astype(np.float32)
res = {}
for K in (0.0, 2.0):
    e2 = BmiVecEnv(n, seed=7, model_path=paths[K])
    e2.reset(init=torch.as_tensor(init.astype(np.float32)))
    e2.set_state(st)
    res[("gpu", K)] = e2.step(torch.as_tensor(act).cuda())[0].cpu().numpy().astype(np.float64)
for sched in (False, True):
    out = []
    for e in range(n):
        o = OracleEnv(0).kernel_mode(schedule=sched)
        o.reset(init[e]); o.set_state(st[e].cpu().numpy().astype(np.float64))
        out.append(o.step(act[e])[0])
    res[("cpu", 2.0 if sched else 0.0)] = np.array(out)
def q(a, b):
    e = np.abs(a - b).max(1)
    return "median %.2e p90 %.2e p95 %.2e max %.2e" % (np.median(e), np.percentile(e, 90), np.percentile(e, 95), e.max())
for a in (("gpu", 0.0), ("gpu", 2.0)):
    for b in (("cpu", 0.0), ("cpu", 2.0)):
        print(a, "vs", b, q(res[a], res[b]))
print("cpu K0 vs cpu K2", q(res[("cpu", 0.0)], res[("cpu", 2.0)]))
print("gpu K0 vs gpu K2", q(res[("gpu", 0.0)], res[("gpu", 2.0)]))
