"""Which observation components carry the worst kernel-vs-oracle one-step errors (tests/test_gpu_physics.py scenarios)?
    python tests/diag/diag_worst_env.py [push|pick] [random|contact]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from rl_arm_under_sparse_reward_b200.bmirobot_env.vec_env import BmiVecEnv
from oracle.physics_oracle import OracleEnv

task = sys.argv[1] if len(sys.argv) > 1 else "push"
scenario = sys.argv[2] if len(sys.argv) > 2 else "random"      # random | contact (hand pressed onto the table next to the block)
n = 64
if scenario == "contact":
    env = BmiVecEnv(n, task=task, seed=21)
    obs, ag, g = env.reset()
    dev = obs.device
    rng = np.random.RandomState(3)
    for t in range(30):
        grip, blk = obs[:, :3], obs[:, 12:15]
        tgt = blk + torch.tensor([0.0, 0.0, 0.01], device=dev)
        a = torch.cat([(tgt - grip).clamp(-0.2, 0.2), torch.zeros(n, 1, device=dev)], 1)
        a[:, 2] -= 0.05 * (t > 15)
        a[:, :2] += torch.as_tensor(rng.uniform(-0.05, 0.05, (n, 2)).astype(np.float32), device=dev)
        obs, ag, _, _ = env.step(a.float().contiguous())
    act = rng.uniform(-0.3, 0.3, (n, 4)).astype(np.float32)
    act[:, 2] = -np.abs(act[:, 2])
else:
    env = BmiVecEnv(n, task=task, seed=7)
    env.reset()
    rng = np.random.RandomState(0)
    for t in range(5):
        env.step(torch.as_tensor(rng.uniform(-0.3, 0.3, (n, 4)).astype(np.float32)).cuda())
    act = rng.uniform(-0.5, 0.5, (n, 4)).astype(np.float32)
st = env.get_state().cpu().numpy().astype(np.float64)
init = env.init.cpu().numpy().astype(np.float64)
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
        s2[27:30] += rs.uniform(-1e-6, 1e-6, 3)          # and the block position by a micrometre
        o.set_state(s2)
        w2, _, _, _ = o.step(act[e])
        sens = max(sens, np.abs(w2 - wants[e]).max())
    print("%3d  %.3e  %.3e" % (e, errs[e], sens))
print("median %.2e" % np.median(errs))
