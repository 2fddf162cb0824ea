"""Extract the small physics fixtures from the reference's recorded demo files (build container only).

    python tests/golden/make_physics_goldens.py
writes tests/golden/physics_golden.npz (episode 0 of the push and pick demo files: the fresh-process
trajectories) and tests/golden/demo_small.npz (first 16 push demo episodes, same keys as
the reference's bmirobot_1000_push_demo.npz, for the add_demo path)."""
import os

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
out = {}
for task in ("push", "pick"):
    d = np.load(os.path.join(REF, "bmirobot_1000_%s_demo.npz" % task), allow_pickle=True)
    obs, acs, g = d["obs"][0], d["acs"][0], d["g"][0, 0]
    out[task + "_obs"] = obs[:21]
    out[task + "_acs"] = acs[:20]
    # block yaw is not recorded by the reference; 1.57 is a placeholder that does not affect a flat drop
    out[task + "_init"] = np.array([obs[0, 12], obs[0, 13], 0.2, 1.57, g[0], g[1], g[2], 0.0])
np.savez_compressed(os.path.join(OUT, "physics_golden.npz"), **out)
d = np.load(os.path.join(REF, "bmirobot_1000_push_demo.npz"), allow_pickle=True)
np.savez_compressed(os.path.join(OUT, "demo_small.npz"), obs=d["obs"][:16], ag=d["ag"][:16], g=d["g"][:16], acs=d["acs"][:16])
d = np.load(os.path.join(REF, "bmirobot_1000_pick_demo.npz"), allow_pickle=True)
np.savez_compressed(os.path.join(OUT, "demo_small_pick.npz"), obs=d["obs"][:8], ag=d["ag"][:8], g=d["g"][:8], acs=d["acs"][:8])
print("wrote physics_golden.npz, demo_small.npz, demo_small_pick.npz")
