"""Which observation components carry the worst kernel-vs-oracle one-step errors (tests/test_gpu_physics.py scenario)?
    python tests/diag/diag_worst_env.py [push|pick]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from rl_arm_under_sparse_reward_b200.bmirobot_env.vec_env import BmiVecEnv
from oracle.physics_oracle import OracleEnv

task = sys.argv[1] if len(sys.argv) > 1 else "push"
n = 64
env = BmiVecEnv(n, task=task, seed=7)
env.reset()
rng = np.random.RandomState(0)
for t in range(5):
    env.step(torch.as_tensor(rng.uniform(-0.3, 0.3, (n, 4)).astype(np.float32)).cuda())
st = env.get_state().cpu().numpy().astype(np.float64)
init = env.init.cpu().numpy().astype(np.float64)
act = rng.uniform(-0.5, 0.5, (n, 4)).astype(np.float32)
obs, ag, r, s = env.step(torch.as_tensor(act).cuda())
got = obs.cpu().numpy()
st2 = env.get_state().cpu().numpy().astype(np.float64)
errs, wants, states = [], [], []
for e in range(n):
    o = OracleEnv(0 if task == "push" else 1)
    o.kernel_mode()
    o.reset(init[e]); o.set_state(st[e])
    want, _, _, _ = o.step(act[e])
    wants.append(want); states.append(o.get_state())
    errs.append(np.abs(got[e] - want).max())
errs = np.array(errs)
np.set_printoptions(precision=4, suppress=True, linewidth=200)
for e in np.argsort(-errs)[:4]:
    d = got[e] - wants[e]
    k = np.argsort(-np.abs(d))[:6]
    print("env %d: max err %.3e; worst obs dims %s diffs %s" % (e, errs[e], k, d[k]))
    print("   got ", got[e][k], "\n   want", wants[e][k])
    ds = st2[e] - states[e]
    ks = np.argsort(-np.abs(ds))[:8]
    print("   state dims %s diffs %s" % (ks, ds[ks]))
# sensitivity of the ORACLE itself at those states: the same step from states whose joint angles differ by +-1e-6 rad
# (a few float32 ulps): where the kernel's outliers sit on a discontinuity (contact feature switch), the oracle moves as much
print("env  kernel-vs-oracle  oracle-vs-perturbed-oracle (max over 4 perturbations)")
rs = np.random.RandomState(1)
for e in np.argsort(-errs)[:10].tolist() + np.argsort(errs)[:3].tolist():
    sens = 0.0
    for k in range(4):
        o = OracleEnv(0 if task == "push" else 1)
        o.kernel_mode()
        o.reset(init[e])
        s2 = st[e].copy()
        s2[:9] += rs.uniform(-1e-6, 1e-6, 9)
        o.set_state(s2)
        w2, _, _, _ = o.step(act[e])
        sens = max(sens, np.abs(w2 - wants[e]).max())
    print("%3d  %.3e  %.3e" % (e, errs[e], sens))
print("median %.2e" % np.median(errs))
