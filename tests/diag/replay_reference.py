"""Open-loop replay of the reference's recorded episodes (tests/golden/demo_small*.npz = the first episodes of
bmirobot_1000_{push,pick}_demo.npz) through the C oracle: the recorded ACTIONS are applied from the reset and every state
is compared with the recording.  The block's initial yaw is not recorded by the reference, so each
episode is replayed for 8 yaws and the best final block position counts.  CPU only.

    python tests/diag/replay_reference.py [--mode faithful|kernel] [--procs N]     -> markdown table on stdout
"""
import argparse, os, sys
from multiprocessing import Pool
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.physics_oracle import OracleEnv

G = os.path.join(ROOT, "tests", "golden")
DATA = {}
for task, f in (("push", "demo_small.npz"), ("pick", "demo_small_pick.npz")):
    d = np.load(os.path.join(G, f))
    DATA[task] = {k: d[k] for k in d.files}
MODE = "faithful"


def one(args):
    task, ep, yaw = args
    d = DATA[task]
    obs, acs, g = d["obs"][ep], d["acs"][ep], d["g"][ep, 0]
    e = OracleEnv(0 if task == "push" else 1)
    if MODE == "kernel":
        e.kernel_mode()
    e.reset([obs[0, 12], obs[0, 13], 0.2, yaw, g[0], g[1], g[2], 0])
    ee, s = [], 0.0
    for t in range(100):
        o, _, _, s = e.step(acs[t])
        ee.append(np.abs(o[:3] - obs[t + 1, :3]).max())
    ref_s = float(np.linalg.norm(obs[100, 12:15] - g) < 0.05)
    return task, ep, yaw, max(ee[:10]), max(ee), float(np.linalg.norm(o[12:15] - obs[100, 12:15])), float(np.linalg.norm(obs[100, 12:15] - obs[0, 12:15])), s, ref_s


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="faithful")
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    a = ap.parse_args()
    MODE = a.mode
    yaws = np.linspace(1.57, 4.71, 9)[:-1]
    jobs = [(task, ep, y) for task in ("push", "pick") for ep in range(DATA[task]["obs"].shape[0]) for y in yaws]
    with Pool(a.procs) as p:
        res = p.map(one, jobs)
    print("| task | episode | EE max error, steps 1-10 (mm) | EE max error, steps 1-100, best yaw (mm) | block travel in the recording (mm) | final block error, best of 8 yaws (mm) | success agrees with the recording (best yaw) |")
    print("|---|---|---|---|---|---|---|")
    agg = {}
    for task in ("push", "pick"):
        for ep in range(DATA[task]["obs"].shape[0]):
            r = [x for x in res if x[0] == task and x[1] == ep]
            best = min(r, key=lambda x: x[5])
            agg.setdefault(task, []).append((best[4], best[5], best[7] == best[8]))
            print("| %s | %d | %.1f | %.0f | %.0f | %.1f | %s |" % (task, ep, 1e3 * best[3], 1e3 * best[4], 1e3 * best[6], 1e3 * best[5], "yes" if best[7] == best[8] else "no"))
    for task, v in agg.items():
        v = np.array(v, dtype=float)
        print("\n%s (%s oracle): %d episodes, EE within 35 mm over the whole episode in %d, median final block error %.1f mm, success flag agrees in %d"
              % (task, a.mode, len(v), int((v[:, 0] < 0.035).sum()), 1e3 * np.median(v[:, 1]), int(v[:, 2].sum())))
