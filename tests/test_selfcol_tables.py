"""CPU: the baked self-collision pair tables (tools/bake_selfcol.py -> assets/bmirobot_selfcol.bin, the kernel's narrow
phase) against the oracle's exact GJK + EPA on the full hulls."""
import os

import numpy as np
import pytest

from oracle.physics_oracle import SELFCOL, OracleEnv


@pytest.fixture(scope="module")
def table():
    if not os.path.exists(SELFCOL):
        import tools.bake_selfcol as b
        b.bake(quiet=True)
    return SELFCOL


def test_table_header_and_descriptors(table):
    t = np.fromfile(table, dtype="<f4")
    assert t[0] == 20261017.0 and int(t[1]) == 4 and int(t[2]) == t.shape[0] and t.nbytes % 16 == 0
    pairs = [tuple(int(x) for x in t[8 + 16 * p:8 + 16 * p + 4]) for p in range(4)]
    # (link A, link B, joint a, joint b): right_link1 x right_link3, link4 x link6, link6 x link8, the two fingers
    assert pairs == [(-1, 1, 0, 1), (2, 4, 3, 4), (4, 6, 5, 6), (7, 8, 7, 8)]
    for p in range(4):
        d = t[8 + 16 * p:8 + 16 * (p + 1)]
        assert d[9] % 4 == 0 and d[9] + 8 * d[7] * d[8] <= t.shape[0]
        # q = 0 is a grid node (the reset pose is a degenerate face-face configuration of the wrist pair)
        assert abs(d[4] / d[6] - round(d[4] / d[6])) < 1e-3 and abs(d[5] / d[6] - round(d[5] / d[6])) < 1e-3


def test_table_contacts_match_gjk_epa(table):
    """contact list of random arm poses: table look-up vs exact narrow phase.  Away from feature switches the two agree to
    the grid's interpolation error; in the cells that contain a switch the table answers with its nearest node."""
    a, b = OracleEnv(0), OracleEnv(0)
    b.use_pair_tables(table)
    rng = np.random.RandomState(0)
    dd, dn, dx, n_pairs = [], [], [], 0
    blob = a.blob
    lo = np.array([blob[64 + 32 * i + 16] for i in range(9)]), 
    hi = np.array([blob[64 + 32 * i + 17] for i in range(9)])
    for _ in range(300):
        q = rng.uniform(-0.5, 0.5, 9)
        q[3] = rng.uniform(-0.8, 0.3)
        q = np.clip(q, lo[0], hi)                            # the tables cover the joint-limit box
        for e in (a, b):
            e.reset([0.3, 0.3, 0.2, 1.57, 0.0, 0.5, 0.2, 0.0])
            st = e.get_state()
            st[:9] = q
            e.set_state(st)
        ca = {(int(r[0]), int(r[1])): r for r in a.contacts() if r[11] >= 1000}
        cb = {(int(r[0]), int(r[1])): r for r in b.contacts() if r[11] >= 1000}
        for k in ca:
            if k not in cb:
                assert ca[k][3] > 0.0005, (k, ca[k])     # only pairs at the edge of the 1 mm reporting range may differ
                continue
            n_pairs += 1
            dd.append(abs(ca[k][3] - cb[k][3]))
            dn.append(1.0 - float(np.dot(ca[k][4:7], cb[k][4:7])))
            dx.append(np.linalg.norm(ca[k][7:10] - cb[k][7:10]))
    dd, dn, dx = np.array(dd), np.array(dn), np.array(dx)
    assert n_pairs >= 600                                    # the wrist pair and the fingers always touch
    assert np.median(dd) < 2e-6 and np.percentile(dd, 95) < 1e-4, (np.median(dd), np.percentile(dd, 95), dd.max())
    assert np.median(dn) < 1e-8 and np.percentile(dn, 90) < 1e-4, (np.median(dn), np.percentile(dn, 90))   # 1 - cos
    assert np.median(dx) < 2e-5, np.median(dx)


def test_table_mode_reproduces_reference_episode0(table, golden_dir):
    from test_oracle_physics import ARM_TOL
    g = np.load(os.path.join(golden_dir, "physics_golden.npz"))
    a, b = OracleEnv(0), OracleEnv(0)
    b.use_pair_tables(table)
    a.reset(g["push_init"])
    b.reset(g["push_init"])
    for t in range(10):
        oa, ob = a.step(g["push_acs"][t])[0], b.step(g["push_acs"][t])[0]
        assert np.abs(oa[:3] - ob[:3]).max() < 0.0015, (t, oa[:3], ob[:3])          # measured <= 1.2 mm
        assert np.abs(ob[:3] - g["push_obs"][t + 1, :3]).max() < ARM_TOL[t] + 0.001
