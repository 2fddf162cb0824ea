"""numpy restatement of the learner-side data path (TEST ORACLE — see oracle/__init__.py).

Each function cites the reference lines it follows (paths relative to the reference tree).
Pinned against the reference's own her.py / replay_buffer.py / normalizer.py, imported
unmodified, by tests/golden/make_learner_goldens.py -> tests/golden/learner_*.npz.
"""
import numpy as np

from . import philox


def goal_distance(a, b):
    """bmirobot_env_push_F.py:20-23"""
    assert a.shape == b.shape
    return np.linalg.norm(a - b, axis=-1)


def compute_reward(ag, g, threshold=0.05):
    """bmirobot_env_push_F.py:84-90 (sparse)"""
    return -(goal_distance(ag, g) > threshold).astype(np.float32)


def her_sample_with_draws(buf, draws, future_p, threshold=0.05):
    """her.py:24-39 with the four random arrays given.  buf: dict obs[E,T+1,Do] ag g actions."""
    ep, t, u_her, u_off = draws
    T = buf['actions'].shape[1]
    full = dict(buf)
    full['obs_next'] = buf['obs'][:, 1:, :]
    full['ag_next'] = buf['ag'][:, 1:, :]
    tr = {k: full[k][ep, t].copy() for k in ('obs', 'ag', 'g', 'actions', 'obs_next', 'ag_next')}
    her = np.where(u_her < future_p)
    off = (u_off * (T - t)).astype(int)
    ft = (t + 1 + off)[her]
    tr['g'][her] = buf['ag'][ep[her], ft]
    tr['r'] = np.expand_dims(compute_reward(tr['ag_next'], tr['g'], threshold), 1)
    return tr


def her_draw_numpy(E, T, B):
    """her.py:24-25,28,30: the order in which the reference consumes numpy's global stream."""
    ep = np.random.randint(0, E, B)
    t = np.random.randint(T, size=B)
    u_her = np.random.uniform(size=B)
    u_off = np.random.uniform(size=B)
    return ep, t, u_her, u_off


def her_draw_philox(seed, counter, B, n_valid, T):
    """restates her_draw_kernel (csrc/her.cu)."""
    c = np.uint64(counter) + np.arange(B, dtype=np.uint64)
    p0 = philox.philox4x32_10(seed, np.uint64(2) * c, philox.STREAM_HER)
    p1 = philox.philox4x32_10(seed, np.uint64(2) * c + np.uint64(1), philox.STREAM_HER)
    ep = (philox.u53(p0[:, 0], p0[:, 1]) * float(n_valid)).astype(np.int64)
    ep = np.minimum(ep, n_valid - 1)
    t = ((p0[:, 2].astype(np.uint64) * np.uint64(T)) >> np.uint64(32)).astype(np.int64)
    return ep, t, philox.u53(p1[:, 0], p1[:, 1]), philox.u53(p1[:, 2], p1[:, 3])


def storage_idx(current_size, size, inc, rng=np.random):
    """replay_buffer.py:57-71; returns (idx, new_current_size)."""
    inc = inc or 1
    if current_size + inc <= size:
        idx = np.arange(current_size, current_size + inc)
    elif current_size < size:
        overflow = inc - (size - current_size)
        idx = np.concatenate([np.arange(current_size, size), rng.randint(0, current_size, overflow)])
    else:
        idx = rng.randint(0, size, inc)
    return idx, min(size, current_size + inc)


class Normalizer:
    """normalizer.py:5-70 for one rank; `world_sums` lets a test inject the other ranks' sums.
    numpy-1.19 float32 semantics for std (see csrc/normalizer.cu header)."""

    def __init__(self, size, eps=1e-2, clip=np.inf):
        self.size, self.eps, self.clip = size, eps, clip
        self.local_sum = np.zeros(size, np.float32)
        self.local_sumsq = np.zeros(size, np.float32)
        self.local_count = np.zeros(1, np.float32)
        self.total_sum = np.zeros(size, np.float32)
        self.total_sumsq = np.zeros(size, np.float32)
        self.total_count = np.ones(1, np.float32)
        self.mean = np.zeros(size, np.float32)
        self.std = np.ones(size, np.float32)

    def update(self, v):
        v = np.asarray(v, dtype=np.float64).reshape(-1, self.size)
        if v.shape[0] <= 1024:      # the reference's case (100 rows): numpy's own sequential row order
            s, q = v.sum(axis=0), np.square(v).sum(axis=0)
        else:                       # csrc/normalizer.cu norm_chunk_kernel: 1024-row chunks, chunk sums added in order
            chunks = [v[i:i + 1024] for i in range(0, v.shape[0], 1024)]
            s, q = chunks[0].sum(axis=0), np.square(chunks[0]).sum(axis=0)
            for c in chunks[1:]:
                s = s + c.sum(axis=0)
                q = q + np.square(c).sum(axis=0)
        self.local_sum += s
        self.local_sumsq += q
        self.local_count[0] += v.shape[0]

    def recompute_stats(self, others=(), world=1):
        ls, lq, lc = self.local_sum.copy(), self.local_sumsq.copy(), self.local_count.copy()
        for o in others:  # cross-rank SUM then /world (normalizer.py:60-64)
            ls, lq, lc = ls + o[0], lq + o[1], lc + o[2]
        ls, lq, lc = ls / np.float32(world), lq / np.float32(world), lc / np.float32(world)
        self.local_sum[...] = 0
        self.local_sumsq[...] = 0
        self.local_count[...] = 0
        self.total_sum += ls
        self.total_sumsq += lq
        self.total_count += lc
        self.mean = self.total_sum / self.total_count
        var = self.total_sumsq / self.total_count - np.square(self.total_sum / self.total_count)
        self.std = np.sqrt(np.maximum(np.float32(np.square(self.eps)), var)).astype(np.float32)

    def normalize(self, v, clip=None):
        clip = self.clip if clip is None else clip
        return np.clip((np.asarray(v, dtype=np.float64) - self.mean) / self.std, -clip, clip)


def network_inputs(tr, o_norm, g_norm, clip_obs=200.0):
    """ddpg_agent.py:229-248: clip, normalise, concat, cast -> x, x_next, actions, r (float32)."""
    o = np.clip(tr['obs'], -clip_obs, clip_obs)
    g = np.clip(tr['g'], -clip_obs, clip_obs)
    on = np.clip(tr['obs_next'], -clip_obs, clip_obs)
    x = np.concatenate([o_norm.normalize(o), g_norm.normalize(g)], axis=1).astype(np.float32)
    xn = np.concatenate([o_norm.normalize(on), g_norm.normalize(g)], axis=1).astype(np.float32)
    return x, xn, tr['actions'].astype(np.float32), tr['r'].astype(np.float32)


def select_actions_philox(pi, seed, counter, action_max=0.5, noise_eps=0.01, random_eps=0.3, late_clip=0.0):
    """restates select_actions_kernel (csrc/ddpg.cu), i.e. ddpg_agent.py:174-184 on a Philox stream."""
    pi = np.asarray(pi, dtype=np.float32)
    n, Da = pi.shape
    c = np.uint64(counter) + np.arange(n, dtype=np.uint64)
    pg = philox.philox4x32_10(seed, np.uint64(3) * c, philox.STREAM_EXPLORE)
    pu = philox.philox4x32_10(seed, np.uint64(3) * c + np.uint64(1), philox.STREAM_EXPLORE)
    pb = philox.philox4x32_10(seed, np.uint64(3) * c + np.uint64(2), philox.STREAM_EXPLORE)
    take = philox.u24(pb[:, 0]) < np.float32(random_eps)
    out = np.empty_like(pi)
    f = np.float32
    for j in range(Da):
        pair = (j >> 1) & 1
        u1 = f(1.0) - philox.u24(pg[:, 2 * pair])
        u2 = philox.u24(pg[:, 2 * pair + 1])
        rad = np.sqrt(f(-2.0) * np.log(u1)).astype(f)
        ang = (f(6.28318530717958647692) * u2).astype(f)
        gz = rad * (np.sin(ang) if (j & 1) else np.cos(ang)).astype(f)
        a = pi[:, j] + f(noise_eps) * f(action_max) * gz
        a = np.clip(a, -f(action_max), f(action_max))
        ra = -f(action_max) + f(2.0) * f(action_max) * philox.u24(pu[:, j & 3])
        a = np.where(take, ra, a)
        if late_clip > 0:
            a = np.clip(a, -f(late_clip), f(late_clip))
        out[:, j] = a
    return out
