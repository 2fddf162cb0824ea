"""Where does a fused rollout's time go, per env?  Builds an instrumented copy of the library (-DBMI_PROF: clock64
counters around policy / IK / fk+dynamics / Cholesky+M^-1 / contacts / rows+Delassus table / PGS loop+integrate, plus
solver iteration counts) and prints the distribution over envs.  Debug tool: the numbers are cycles of the env's own
warp, not a benchmark.

    python tools/prof_rollout_phases.py [n_envs] [T] [--no-build]
    BMI_PROF_LIB=<lib built with -DBMI_PROF -DBMI_ENVS_PER_BLOCK=1> ... 148 100 --no-build    (one env per SM: lone-warp latencies)
"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rl_arm_under_sparse_reward_b200 import _build
lib_path = os.environ.get("BMI_PROF_LIB", os.path.join(_build.OUT_DIR, "libbmi_b200_prof.so"))
if "--no-build" not in sys.argv:
    _build.build(force=True, verbose=False, extra_flags=["-DBMI_PROF"], lib_path=lib_path, obj_suffix="_prof")
os.environ["BMI_B200_LIB"] = lib_path
import numpy as np, torch
from rl_arm_under_sparse_reward_b200 import _lib
from rl_arm_under_sparse_reward_b200.arguments import Args
from rl_arm_under_sparse_reward_b200.bmirobot_env.vec_env import BmiVecEnv
from rl_arm_under_sparse_reward_b200.ddpg_agent import ddpg_agent
from rl_arm_under_sparse_reward_b200.train import get_env_params
argv = [a for a in sys.argv[1:] if not a.startswith("--")]
n = int(argv[0]) if len(argv) > 0 else 4096
T = int(argv[1]) if len(argv) > 1 else 100
a = Args(); a.add_demo, a.verbose, a.n_envs, a.buffer_size, a.save_dir = False, False, n, 8192 * 100, "/tmp/bmi_prof/"
torch.manual_seed(125)
env = BmiVecEnv(n, seed=125)
p = get_env_params(env); p['max_timesteps'] = T
ag = ddpg_agent(a, env, p)
dbg = ctypes.CDLL(lib_path).bmi_debug_prof
dbg.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
for rep in range(2):
    dbg(None, 0, 1)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); ag.rollout(0); e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e)
    out = np.zeros((8192, 8), dtype=np.uint64)
    dbg(out.ctypes.data_as(ctypes.c_void_p), out.size, 0)
    out = out[:n].astype(np.float64)
    names = ["policy", "ik+begin", "fk+dynamics", "chol+Minv+u", "contacts", "rows+Delassus", "PGS loop+integrate", "iters"]
    tot = out[:, :7].sum(1)
    print("rollout %d: %.1f ms (%.0f env-steps/s); per-env busy cycles: median %.3g max %.3g (=%.1f ms at 1.965 GHz)" % (
        rep, ms, n * T / ms * 1e3, np.median(tot), tot.max(), tot.max() / 1.965e6))
    nsub = T * 20
    for k, nm in enumerate(names):
        c = out[:, k]
        unit = "cycles/substep" if k < 7 else "per substep"
        print("  %-18s mean %10.1f  median %10.1f  p90 %10.1f  max %10.1f  %s" % (
            nm, c.mean() / nsub, np.median(c) / nsub, np.percentile(c, 90) / nsub, c.max() / nsub, unit))
    slow = np.argsort(-tot)[:3]
    for e_ in slow:
        print("  slow env %5d: total %.3g  " % (e_, tot[e_]) + " ".join("%s=%.0f" % (nm, out[e_, k] / nsub) for k, nm in enumerate(names)))
    per_it = out[:, 6] / np.maximum(out[:, 7], 1)
    print("  PGS cycles per iteration: median %.0f  p90 %.0f  max %.0f" % (np.median(per_it), np.percentile(per_it, 90), per_it.max()))
