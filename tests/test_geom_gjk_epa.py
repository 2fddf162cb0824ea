"""CPU: the table baker's narrow phase (tools/geom/convex_epa.h through tools/geom/pair_table.c) against scipy's exact convex
hull of the Minkowski difference, on random polytopes -- the geometry behind the kernel's self-collision tables is checked
against an independent implementation, not only against the oracle that shares the header."""
import ctypes

import numpy as np
import pytest
from scipy.spatial import ConvexHull

from tools import bake_selfcol


@pytest.fixture(scope="module")
def lib():
    l = ctypes.CDLL(bake_selfcol.build_lib())
    l.pt_pair_world.restype = ctypes.c_int
    l.pt_pair_world.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double] + [ctypes.c_void_p] * 4
    return l


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _rot(rng):
    q = rng.standard_normal(4)
    x, y, z, w = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _query(lib, va, Ra, pa, vb, Rb, pb, far=10.0):
    d, n, wa, wb = np.zeros(1), np.zeros(3), np.zeros(3), np.zeros(3)
    ok = lib.pt_pair_world(_p(va), len(va), _p(np.ascontiguousarray(Ra)), _p(pa), _p(vb), len(vb), _p(np.ascontiguousarray(Rb)),
                           _p(pb), far, _p(d), _p(n), _p(wa), _p(wb))
    return ok, d[0], n, wa, wb


def test_penetration_depth_equals_minkowski_hull(lib):
    rng = np.random.RandomState(0)
    n_pen = 0
    for _ in range(200):
        va = np.ascontiguousarray(rng.uniform(-1, 1, (rng.randint(8, 40), 3)) * rng.uniform(0.3, 1.0, 3))
        vb = np.ascontiguousarray(rng.uniform(-1, 1, (rng.randint(8, 40), 3)) * rng.uniform(0.3, 1.0, 3))
        Ra, Rb = _rot(rng), _rot(rng)
        pa, pb = rng.uniform(-0.4, 0.4, 3), rng.uniform(-0.4, 0.4, 3)
        A, B = va @ Ra.T + pa, vb @ Rb.T + pb
        md = (A[:, None, :] - B[None, :, :]).reshape(-1, 3)          # vertices of A - B
        hull = ConvexHull(md)
        off = -hull.equations[:, 3]                                   # facet planes n.x = off, n outward
        ok, dist, n, wa, wb = _query(lib, va, Ra, pa, vb, Rb, pb)
        if off.min() > 1e-9:                                          # origin strictly inside A - B: the hulls overlap
            n_pen += 1
            k = int(np.argmin(off))
            assert ok and abs(-dist - off[k]) <= 1e-9 * max(1.0, off[k]), (dist, off[k])
            # n points from B towards A; separating A along n by the depth removes the overlap: n = -(outward facet normal)
            assert np.dot(-n, hull.equations[k, :3]) > 1 - 1e-6 or np.sum(np.abs(off - off[k]) < 1e-7) > 1
            assert abs(np.dot(wa - wb, n) - dist) < 1e-9                  # witness points realise the depth
        elif off.min() < -1e-9:
            assert ok and dist > 0
            # lower bound from the hull's facets (exact when the closest feature of A - B is a facet)
            assert dist >= -off.min() - 1e-9
            assert abs(np.linalg.norm(wa - wb) - dist) < 1e-9
            # the witness points lie on the hulls: moving B by the gap along n makes them touch (depth ~ 0 afterwards)
            ok2, d2, *_ = _query(lib, va, Ra, pa, vb, Rb, pb + n * (dist + 1e-6))
            assert ok2 and -2e-6 < d2 < 0
    assert n_pen > 50


def test_far_pairs_are_culled(lib):
    rng = np.random.RandomState(1)
    va = np.ascontiguousarray(rng.uniform(-0.1, 0.1, (12, 3)))
    ok, *_ = _query(lib, va, np.eye(3), np.zeros(3), va, np.eye(3), np.array([1.0, 0, 0]), far=0.012)
    assert not ok
