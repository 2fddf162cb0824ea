"""Bake the self-collision pair tables of the physics kernel (assets/bmirobot_selfcol.bin).

Every pair of right-arm links that can touch under the reference's self-collision rule (bmirobot.py:58 flags=9: all
pairs except child/parent) is separated by exactly TWO joints -- the grandparent pairs of the chain and the two
fingers -- so its narrow phase (signed distance, normal, witness point of the two convex hulls) is a function of two
joint angles.  The kernel looks that function up instead of running GJK / EPA on 150..750-vertex hulls every sub-step;
this tool samples it with the oracle's exact GJK / EPA (oracle/convex_epa.h through bmo_pair_query) on a regular grid
over the two joints' limit box.  Format: include/bmi_model.h (SC_*).

Needs only files inside the repo (assets/bmirobot_model.bin, assets/bmirobot_hulls.bin, the built oracle library), so it
runs from __graft_entry__.build(); the output (tens of MB) is git-ignored and travels to the GPU box with the snapshot.

    python tools/bake_selfcol.py [--h 0.005] [--procs N]

Pairs (hull indices = 1 + link index; hull 0 = right_link1, rigid with the base):
    right_link1 x right_link3   joints 0, 1     right_link4 x right_link6   joints 3, 4
    right_link6 x right_link8   joints 5, 6     right_hand1 x right_hand2   joints 7, 8
These are the only pairs that ever produced a row in 2000 env-steps of random exploration with the exact narrow phase
(the two last ones interpenetrate permanently); the other grandparent pairs stay more than 2 cm apart.
"""
import argparse
import ctypes
import os
import sys
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "rl_arm_under_sparse_reward_b200", "assets", "bmirobot_selfcol.bin")
MAGIC = 20261017.0
HDR, DESC = 8, 16
FAR = 0.012          # nodes whose cores are farther apart than this hold the "far" sentinel
# (hull a, hull b, joint a, joint b)
# ... and the grid spacing relative to --h (the base-link pair is rarely closer than 1 mm: coarser grid)
PAIRS = [(0, 2, 0, 1, 2.0), (3, 5, 3, 4, 1.0), (5, 7, 5, 6, 1.0), (8, 9, 7, 8, 1.0)]


def _env():
    from oracle.physics_oracle import OracleEnv, _p
    e = OracleEnv(0)
    e.lib.bmo_pair_query.restype = None
    e.lib.bmo_pair_query.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p]
    return e, _p


def _rows(args):
    (a, b, ja, jb), qa_vals, qb_vals = args
    e, _p = _env()
    out = np.zeros((len(qa_vals), len(qb_vals), 8), np.float32)
    q = np.zeros(9)
    tmp = np.zeros(8, np.float32)
    for i, qa in enumerate(qa_vals):
        for j, qb in enumerate(qb_vals):
            q[:] = 0
            q[ja], q[jb] = qa, qb
            e.lib.bmo_pair_query(e.h, a, b, _p(q), FAR, _p(tmp))
            out[i, j] = tmp
    return out


def bake(h0=0.005, procs=None, out_path=OUT, quiet=False):
    e, _ = _env()
    blob = e.blob
    links_off, stride = int(blob[4]), 32
    lo = [float(blob[links_off + stride * i + 16]) for i in range(9)]
    hi = [float(blob[links_off + stride * i + 17]) for i in range(9)]
    hulls = e.hulls
    mu, o = [], 1
    link_of = []
    for _i in range(int(hulls[0])):
        link_of.append(int(hulls[o]))
        mu.append(float(hulls[o + 1]))
        o += 3 + 3 * int(hulls[o + 2])
    procs = procs or os.cpu_count() or 1
    descs, tables, off = [], [], HDR + DESC * len(PAIRS)
    with Pool(procs) as pool:
        for (a, b, ja, jb, hs) in PAIRS:
            h = h0 * hs
            # q = 0 is a grid node: every episode starts from the all-zero pose, where the wrist pair is in a degenerate
            # face-face configuration whose witness point must be the exact narrow-phase answer, not an interpolation;
            # one node of slack beyond the limits (ERP lets joints overshoot)
            qa = h * np.arange(np.floor(lo[ja] / h) - 1, np.ceil(hi[ja] / h) + 2)
            qb = h * np.arange(np.floor(lo[jb] / h) - 1, np.ceil(hi[jb] / h) + 2)
            chunks = np.array_split(qa, max(1, min(len(qa), procs * 4)))
            parts = pool.map(_rows, [((a, b, ja, jb), c, qb) for c in chunks])
            t = np.concatenate(parts, axis=0)
            d = np.zeros(DESC, np.float32)
            d[:11] = [link_of[a], link_of[b], ja, jb, qa[0], qb[0], h, len(qa), len(qb), off, min(10.0, mu[a] * mu[b])]
            descs.append(d)
            tables.append(t.reshape(-1))
            off += t.size
            if not quiet:
                near = (t[..., 0] < 100).mean()
                print("pair hulls (%d, %d) joints (%d, %d): %d x %d nodes, %.1f %% within %.0f mm" % (a, b, ja, jb, len(qa), len(qb), 100 * near, 1e3 * FAR))
    hdr = np.zeros(HDR, np.float32)
    hdr[:3] = [MAGIC, len(PAIRS), off]
    assert off < 2 ** 24, "float offsets must stay exact in float32"
    np.concatenate([hdr] + descs + tables).astype("<f4").tofile(out_path)
    return out_path


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--h", type=float, default=0.005)
    ap.add_argument("--procs", type=int, default=None)
    ap.add_argument("--out", default=OUT)
    a = ap.parse_args()
    print(bake(a.h, a.procs, a.out))
