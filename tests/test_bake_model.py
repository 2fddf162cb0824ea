"""CPU: the model loader (tools/bake_model.py: URDF joint tree + STL meshes -> model blob + full hulls) reproduces the
committed assets from the reference's files.  Runs only where the reference tree is mounted (the build container)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
ASSETS = os.path.join(ROOT, "rl_arm_under_sparse_reward_b200", "assets")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "URDF_model")), reason="reference tree not mounted")
def test_rebake_from_reference_urdf_equals_committed_assets(tmp_path, monkeypatch):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bake_model
    monkeypatch.setattr(bake_model, "OUT_DIR", str(tmp_path))
    bake_model.main()
    for name in ("bmirobot_model.bin", "bmirobot_hulls.bin"):
        a = np.fromfile(os.path.join(tmp_path, name), "<f4")
        b = np.fromfile(os.path.join(ASSETS, name), "<f4")
        assert a.shape == b.shape and np.array_equal(a, b), name


def test_blob_describes_the_urdf_chain():
    """joint tree facts of robotarm_description.urdf:423-501 as stored in the blob: 9 revolute joints, parents, axes,
    limits, damping 0.7, unit masses with the COM one metre up the link z axis, PyBullet's solver constants."""
    b = np.fromfile(os.path.join(ASSETS, "bmirobot_model.bin"), "<f4")
    assert int(b[2]) == 9 and int(b[4]) == 64
    links = b[64:64 + 9 * 32].reshape(9, 32)
    assert [int(x) for x in links[:, 0]] == [-1, 0, 1, 2, 3, 4, 5, 6, 6]
    assert np.allclose(links[:, 18], 0.7) and np.allclose(links[:, 19], 1.0) and np.allclose(links[:, 20:23], [0, 0, 1])
    assert np.allclose(links[0, 13:16], [1, 0, 0]) and np.allclose(links[6, 13:16], [0, 0, 1])
    assert np.isclose(links[3, 16], -0.872664625997) and np.isclose(links[5, 16], -1.2217304764)
    assert np.isclose(b[8], 1 / 240) and b[9] == -10 and b[10] == 20 and b[11] == 150
    assert np.isclose(b[21], 0.5) and b[47] == 1 and np.isclose(b[49], 0.001)      # IK damping, self-collision, hull margin
    h = np.fromfile(os.path.join(ASSETS, "bmirobot_hulls.bin"), "<f4")
    assert int(h[0]) == 10
    o, links_h = 1, []
    for _ in range(10):
        links_h.append(int(h[o]))
        o += 3 + 3 * int(h[o + 2])
    assert links_h == [-1, 0, 1, 2, 3, 4, 5, 6, 7, 8] and o == h.shape[0]
