"""CPU checks of the two re-formulations the CUDA env kernel uses (float64 numpy restatements of the kernel's
formulas, no GPU): they are mathematically the oracle's algorithm, just organised for a warp.

* mass matrix by composite rigid bodies (csrc/physics.cu joint_space_dynamics) == the oracle's ten RNEA sweeps;
* projected Gauss-Seidel in constraint space with a precomputed coupling table (substep_solve) == the oracle's
  velocity-space PGS (pgs_solve in oracle/bmi_physics_oracle.c): same iterates, same iteration count;
* the IK step through the 3x3 push-through system == the oracle's 9x9 damped least squares step.
"""
import os

import numpy as np

from oracle.physics_oracle import MODEL, OracleEnv

HDR, LINK_STRIDE = 64, 32
ML_PARENT, ML_JPOS, ML_JROT, ML_AXIS, ML_LO, ML_HI, ML_MASS, ML_COM, ML_INERTIA = 0, 1, 4, 13, 16, 17, 19, 20, 23
MP_BASE_PX = 28


def _links():
    blob = np.fromfile(MODEL, dtype="<f4").astype(np.float64)
    return blob, [blob[HDR + i * LINK_STRIDE: HDR + (i + 1) * LINK_STRIDE] for i in range(9)]


def _rodrigues(u, q):
    K = np.array([[0, -u[2], u[1]], [u[2], 0, -u[0]], [-u[1], u[0], 0]])
    return np.eye(3) + np.sin(q) * K + (1 - np.cos(q)) * (K @ K)


def _fk(blob, links, q):
    R, p, z, c = [None] * 9, [None] * 9, [None] * 9, [None] * 9
    for i, lk in enumerate(links):
        pa = int(lk[ML_PARENT])
        Rl = lk[ML_JROT:ML_JROT + 9].reshape(3, 3) @ _rodrigues(lk[ML_AXIS:ML_AXIS + 3], q[i])
        if pa < 0:
            R[i], p[i] = Rl, blob[MP_BASE_PX:MP_BASE_PX + 3] + lk[ML_JPOS:ML_JPOS + 3]
        else:
            R[i], p[i] = R[pa] @ Rl, p[pa] + R[pa] @ lk[ML_JPOS:ML_JPOS + 3]
        z[i] = R[i] @ lk[ML_AXIS:ML_AXIS + 3]
        c[i] = p[i] + R[i] @ lk[ML_COM:ML_COM + 3]
    return R, p, z, c


def _in_subtree(j, l):      # physics.cu is_ancestor_or_self: chain 0..6, fingers 7 and 8 on link 6
    return j == l or (j <= 6 and l >= j)


def _crba(links, R, p, z, c):
    mass = np.array([lk[ML_MASS] for lk in links])
    Iw = [R[i] @ np.diag(links[i][ML_INERTIA:ML_INERTIA + 3]) @ R[i].T for i in range(9)]
    M = np.zeros((9, 9))
    comp = []
    for j in range(9):      # subtree mass, COM and inertia about the COM (parallel-axis sums)
        sub = [l for l in range(9) if _in_subtree(j, l)]
        m = mass[sub].sum()
        cc = sum(mass[l] * c[l] for l in sub) / m
        Ic = sum(Iw[l] + mass[l] * ((c[l] - cc) @ (c[l] - cc) * np.eye(3) - np.outer(c[l] - cc, c[l] - cc)) for l in sub)
        comp.append((m, cc, Ic))
    for k in range(9):
        m, cc, Ic = comp[k]
        N, F = Ic @ z[k], m * np.cross(z[k], cc - p[k])       # wrench of subtree k under a unit acceleration of joint k
        for j in range(k + 1):
            if _in_subtree(j, k):
                M[k, j] = M[j, k] = z[j] @ (N + np.cross(cc - p[j], F))
    return M


def test_composite_body_mass_matrix_equals_oracle_rnea_columns():
    blob, links = _links()
    o = OracleEnv(0)
    rng = np.random.RandomState(0)
    lo = np.array([lk[ML_LO] for lk in links]); hi = np.array([lk[ML_HI] for lk in links])
    for _ in range(20):
        q = lo + (hi - lo) * rng.uniform(0.05, 0.95, 9)
        M = _crba(links, *_fk(blob, links, q))
        assert np.abs(M - o.mass_matrix(q)).max() < 1e-9
        assert M[7, 8] == 0.0 and np.all(np.linalg.eigvalsh(M) > 0)


def _velocity_space_pgs(Minv, Ibinv6, rows, max_it=150, thresh=1e-7):
    """pgs_solve of the oracle: rows = list of dict(J (15,), lo, hi, rhs_scaled, friction_of / mu)."""
    W = [np.concatenate([Minv @ r["J"][:9], Ibinv6 @ r["J"][9:]]) for r in rows]
    invd = [1.0 / (r["J"] @ w) for r, w in zip(rows, W)]
    lam, dv, hist = np.zeros(len(rows)), np.zeros(15), []
    n_plain = sum(1 for r in rows if "friction_of" not in r)
    for it in range(max_it):
        resid = 0.0
        for i in range(n_plain):
            r = rows[i]
            d = (r["b"] - r["J"] @ dv) * invd[i]
            s = min(max(lam[i] + d, r["lo"]), r["hi"])
            d, lam[i] = s - lam[i], s
            dv += W[i] * d
            resid = max(resid, (d / invd[i]) ** 2)
        for i in range(n_plain, len(rows), 2):
            ra, rb = rows[i], rows[i + 1]
            lim = ra["mu"] * lam[ra["friction_of"]]
            sa = lam[i] + (ra["b"] - ra["J"] @ dv) * invd[i]
            sb = lam[i + 1] + (rb["b"] - rb["J"] @ dv) * invd[i + 1]
            n = np.hypot(sa, sb)
            if n > lim:
                sa, sb = (sa * lim / n, sb * lim / n) if n > 0 else (0.0, 0.0)
            da, db = sa - lam[i], sb - lam[i + 1]
            lam[i], lam[i + 1] = sa, sb
            dv += W[i] * da + W[i + 1] * db
            resid = max(resid, (da / invd[i]) ** 2, (db / invd[i + 1]) ** 2)
        hist.append(lam.copy())
        if resid <= thresh:
            break
    return dv, lam, len(hist), hist


def _constraint_space_pgs(Minv, Ibinv6, rows, max_it=150, thresh=1e-7):
    """substep_solve of the kernel: joint variables dv_j, block variables dvb, one variable v = J . dv per contact row,
    updated through the coupling table (M^-1, M^-1 Ja^T, block response, Delassus entries)."""
    nm = 9                                               # rows 0..8 are the motors (J = e_j)
    crow = [i for i in range(len(rows)) if i >= nm]      # contact rows (normals first, then friction pairs)
    Wa = {i: Minv @ rows[i]["J"][:9] for i in crow}
    Wb = {i: Ibinv6 @ rows[i]["J"][9:] for i in crow}
    A = {(y, x): rows[y]["J"][:9] @ Wa[x] + rows[y]["J"][9:] @ Wb[x] for y in crow for x in crow}
    diag = {i: (Minv[i, i] if i < nm else A[(i, i)]) for i in range(len(rows))}
    lam = np.zeros(len(rows)); dvj = np.zeros(9); dvb = np.zeros(6); v = {i: 0.0 for i in crow}; hist = []
    n_plain = sum(1 for r in rows if "friction_of" not in r)

    def fire_joint(j, d):
        nonlocal dvj
        dvj = dvj + Minv[:, j] * d
        for y in crow:
            v[y] += Wa[y][j] * d

    def fire_contact(x, d):
        nonlocal dvj, dvb
        dvj = dvj + Wa[x] * d
        dvb = dvb + Wb[x] * d
        for y in crow:
            v[y] += A[(y, x)] * d

    for it in range(max_it):
        resid = 0.0
        for i in range(n_plain):
            r = rows[i]
            vi = dvj[i] if i < nm else v[i]
            d = (r["b"] - vi) / diag[i]
            s = min(max(lam[i] + d, r["lo"]), r["hi"])
            d, lam[i] = s - lam[i], s
            fire_joint(i, d) if i < nm else fire_contact(i, d)
            resid = max(resid, (d * diag[i]) ** 2)
        for i in range(n_plain, len(rows), 2):
            ra, rb = rows[i], rows[i + 1]
            lim = ra["mu"] * lam[ra["friction_of"]]
            sa = lam[i] + (ra["b"] - v[i]) / diag[i]
            sb = lam[i + 1] + (rb["b"] - v[i + 1]) / diag[i + 1]
            n2 = sa * sa + sb * sb
            sc = lim / np.sqrt(n2) if n2 > lim * lim else 1.0
            sa, sb = sa * sc, sb * sc
            da, db = sa - lam[i], sb - lam[i + 1]
            lam[i], lam[i + 1] = sa, sb
            fire_contact(i, da); fire_contact(i + 1, db)
            resid = max(resid, (da * diag[i]) ** 2, (db * diag[i + 1]) ** 2)
        hist.append(lam.copy())
        if resid <= thresh:
            break
    return np.concatenate([dvj, dvb]), lam, len(hist), hist


def test_constraint_space_pgs_reproduces_the_velocity_space_iterates():
    rng = np.random.RandomState(1)
    for trial in range(8):
        Lm = rng.normal(size=(9, 9)) * 0.3 + np.eye(9)
        Minv = np.linalg.inv(Lm @ Lm.T + 0.1 * np.eye(9))
        Ibinv6 = np.diag([1.0, 1.0, 1.0, 3750.0, 3750.0, 3750.0])          # 1 kg, 4 cm cube
        rows = []
        for j in range(9):                                                   # motors: J = e_j, rhs in velocity units
            J = np.zeros(15); J[j] = 1.0
            rows.append(dict(J=J, lo=-2.0, hi=2.0, b=rng.normal() * 0.1))
        nc = 3 + trial % 4
        normals, frictions = [], []
        for c in range(nc):
            kind = c % 3                                                     # block-table / arm-table / arm-block contact
            Jn, Ja, Jb = (np.zeros(15) for _ in range(3))
            for J in (Jn, Ja, Jb):
                if kind != 0: J[:9] = rng.normal(size=9) * 0.2
                if kind != 1: J[9:] = rng.normal(size=6) * np.array([1, 1, 1, .02, .02, .02])
            normals.append(dict(J=Jn, lo=0.0, hi=1e10, b=abs(rng.normal()) * 0.05))
            frictions += [dict(J=Ja, b=rng.normal() * 0.02, friction_of=9 + c, mu=0.5),
                          dict(J=Jb, b=rng.normal() * 0.02, friction_of=9 + c, mu=0.5)]
        rows += normals + frictions
        dv1, lam1, it1, h1 = _velocity_space_pgs(Minv, Ibinv6, rows)
        dv2, lam2, it2, h2 = _constraint_space_pgs(Minv, Ibinv6, rows)
        assert it1 == it2
        assert max(np.abs(a - b).max() for a, b in zip(h1, h2)) < 1e-10     # every iterate, not only the last
        assert np.abs(dv1 - dv2).max() < 1e-10


def test_push_through_ik_step_equals_damped_least_squares():
    rng = np.random.RandomState(2)
    for _ in range(10):
        J = rng.normal(size=(3, 9)) * 0.3
        J[:, 7] = 0.0                                   # the first finger is not on the path to the end effector
        e, damp = rng.normal(size=3) * 0.1, 0.1
        a = np.linalg.solve(J.T @ J + damp * np.eye(9), J.T @ e)             # oracle solve_ik
        b = J.T @ np.linalg.solve(J @ J.T + damp * np.eye(3), e)             # kernel solve_ik (BussIK's form)
        assert np.abs(a - b).max() < 1e-12 and b[7] == 0.0


def test_demo_controllers_follow_the_reference_scripts():
    """get_demo_data.push_controller / pick_controller against a literal per-env restatement of
    get_demo_data_push.py:40-62 and get_demo_data_pick.py:53-68 (CPU tensors, no kernels)."""
    import math
    import torch
    from rl_arm_under_sparse_reward_b200.get_demo_data import pick_controller, push_controller
    rng = np.random.RandomState(3)
    obs = torch.as_tensor(rng.uniform(-0.3, 0.6, (16, 27)).astype(np.float32))
    g = torch.as_tensor(rng.uniform(0.0, 0.5, (16, 3)).astype(np.float32))
    obs[0, 12:15] = g[0] + 0.01                         # block already at its goal: the push script stops
    for step_time in (1, 10, 11, 20, 21, 40, 41, 60, 61, 80, 81, 100):
        got_push, got_pick = push_controller(step_time, obs, g).numpy(), pick_controller(step_time, obs, g).numpy()
        for e in range(16):
            grip, blk, gg = obs[e, :3].numpy(), obs[e, 12:15].numpy(), g[e].numpy()
            if step_time <= 10: a = [0, -0.1, 0.1, 0]
            elif step_time <= 20 or 60 < step_time <= 80: a = list((gg - blk) * (-0.5) + blk - grip) + [0]
            elif step_time <= 40 or step_time > 80: a = list(gg - blk) + [0]
            else: a = [0.241 - grip[0], 0.3265 - grip[1], 0.294 - grip[2], 0]
            if math.sqrt(float(((blk - gg) ** 2).sum())) < 0.05: a = [0, 0, 0, 0]
            assert np.allclose(got_push[e], a, atol=1e-6), (step_time, e)
            if step_time <= 10: p = [0, -0.1, 0.1, 0]
            elif step_time <= 30: p = [blk[0] - grip[0], blk[1] - grip[1] - 0.2, blk[2] - grip[2] + 0.1, 0]
            elif step_time <= 50: p = [0, 0, 0, 0.1]
            elif step_time <= 70: p = [blk[0] - grip[0], blk[1] - grip[1] - 0.05, blk[2] - grip[2] + 0.05, 0]
            elif step_time <= 90: p = [0, 0, 0, -0.1]
            else: p = list(gg - blk) + [0]
            assert np.allclose(got_pick[e], p, atol=1e-6), (step_time, e)
