"""Bake the self-collision pair tables of the physics kernel (assets/bmirobot_selfcol.bin).

Every pair of right-arm links that can touch under the reference's self-collision rule (bmirobot.py:58 flags=9: all
pairs except child/parent) is separated by exactly TWO joints -- the grandparent pairs of the chain and the two
fingers -- so its narrow phase (signed distance, normal, witness point of the two convex hulls) is a function of two
joint angles.  The kernel looks that function up instead of running GJK / EPA on 150..750-vertex hulls every sub-step;
this tool samples it on a regular grid over the two joints' limit box with tools/geom/pair_table.c (forward kinematics
of the baked joint tree + exact GJK / EPA, tools/geom/convex_epa.h).  Format: include/bmi_model.h (SC_*).

Product build tooling: needs only assets/bmirobot_model.bin, assets/bmirobot_hulls.bin and gcc -- nothing under
oracle/ -- so it runs from __graft_entry__.build(); the output (tens of MB) is git-ignored and travels to the GPU box
with the snapshot.  (tests/test_selfcol_tables.py checks the baked nodes against the test oracle's own narrow phase.)

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
import subprocess
import sys
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ASSETS = os.path.join(ROOT, "rl_arm_under_sparse_reward_b200", "assets")
MODEL = os.path.join(ASSETS, "bmirobot_model.bin")
HULLS = os.path.join(ASSETS, "bmirobot_hulls.bin")
OUT = os.path.join(ASSETS, "bmirobot_selfcol.bin")
GEOM = os.path.join(ROOT, "tools", "geom")
LIB = os.path.join(GEOM, "_build", "libpairtable.so")
SOURCES = [os.path.join(GEOM, "pair_table.c"), os.path.join(GEOM, "convex_epa.h"), os.path.join(ROOT, "include", "bmi_model.h")]
MAGIC = 20261017.0
HDR, DESC = 8, 16
FAR = 0.012          # nodes whose cores are farther apart than this hold the "far" sentinel
# (hull a, hull b, joint a, joint b)
# ... and the grid spacing relative to --h (the base-link pair is rarely closer than 1 mm: coarser grid)
PAIRS = [(0, 2, 0, 1, 2.0), (3, 5, 3, 4, 1.0), (5, 7, 5, 6, 1.0), (8, 9, 7, 8, 1.0)]


def build_lib():
    """gcc the sampler (same floating-point flags as every other CPU build of the repo: no contraction)"""
    if os.path.exists(LIB) and all(os.path.getmtime(f) <= os.path.getmtime(LIB) for f in SOURCES):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.run([os.environ.get("CC", "gcc"), "-O2", "-fPIC", "-std=c11", "-ffp-contract=off", "-D_GNU_SOURCE", "-shared",
                    "-o", LIB, SOURCES[0], "-lm"], check=True)
    return LIB


def load_hulls(path=HULLS):
    """[(link index or -1, friction, vertices float64 (nv, 3))] of assets/bmirobot_hulls.bin"""
    h = np.fromfile(path, dtype="<f4")
    out, o = [], 1
    for _ in range(int(h[0])):
        nv = int(h[o + 2])
        out.append((int(h[o]), float(h[o + 1]), np.ascontiguousarray(h[o + 3:o + 3 + 3 * nv].astype(np.float64).reshape(nv, 3))))
        o += 3 + 3 * nv
    return out


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _rows(args):
    (a, b, ja, jb), qa_vals, qb_vals = args
    lib = ctypes.CDLL(build_lib())
    lib.pt_bake_rows.restype = None
    lib.pt_bake_rows.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                 ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                 ctypes.c_int, ctypes.c_double, ctypes.c_void_p]
    blob = np.fromfile(MODEL, dtype="<f4")
    hulls = load_hulls()
    (la, _, va), (lb, _, vb) = hulls[a], hulls[b]
    qa_vals, qb_vals = np.ascontiguousarray(qa_vals, np.float64), np.ascontiguousarray(qb_vals, np.float64)
    out = np.zeros((len(qa_vals), len(qb_vals), 8), np.float32)
    lib.pt_bake_rows(_p(blob), _p(va), len(va), la, _p(vb), len(vb), lb, ja, jb, _p(qa_vals), len(qa_vals), _p(qb_vals),
                     len(qb_vals), FAR, _p(out))
    return out


def bake(h0=0.005, procs=None, out_path=OUT, quiet=False):
    build_lib()
    blob = np.fromfile(MODEL, dtype="<f4")
    links_off, stride = int(blob[4]), 32
    lo = [float(blob[links_off + stride * i + 16]) for i in range(9)]
    hi = [float(blob[links_off + stride * i + 17]) for i in range(9)]
    hulls = load_hulls()
    link_of, mu = [h[0] for h in hulls], [h[1] for h in hulls]
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
