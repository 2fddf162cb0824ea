"""Generate the learner-side golden fixtures by running the UNMODIFIED reference files.

Run in the build container only (needs /root/reference):
    python tests/golden/make_learner_goldens.py
Writes tests/golden/learner_*.npz.  mpi4py and matplotlib are not installed, so single-rank
stand-ins are placed on sys.path (Get_rank 0, Get_size 1, Allreduce = copy, Bcast = no-op);
the env is a stand-in that only supplies compute_reward (bmirobot_env_push_F.py:84-90 copied
semantics via the oracle) — no physics is involved in these fixtures.
"""
import os
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(OUT))
sys.path.insert(0, ROOT)


def install_stubs():
    d = tempfile.mkdtemp(prefix="bmi_stubs_")
    os.makedirs(os.path.join(d, "mpi4py"))
    with open(os.path.join(d, "mpi4py", "__init__.py"), "w") as f:
        f.write(
            "import numpy as np\n"
            "class _Comm:\n"
            "    def Get_rank(self): return 0\n"
            "    def Get_size(self): return 1\n"
            "    def Bcast(self, buf, root=0): pass\n"
            "    def Allreduce(self, a, b, op=None): b[...] = a\n"
            "    def allreduce(self, x, op=None): return x\n"
            "class MPI:\n"
            "    COMM_WORLD = _Comm()\n"
            "    SUM = 'sum'\n")
    os.makedirs(os.path.join(d, "matplotlib"))
    open(os.path.join(d, "matplotlib", "__init__.py"), "w").close()
    with open(os.path.join(d, "matplotlib", "pyplot.py"), "w") as f:
        f.write("def plot(*a, **k): pass\ndef show(*a, **k): pass\n")
    sys.path.insert(0, d)
    sys.path.insert(0, REF)


class FakeEnv:
    distance_threshold = 0.05

    def compute_reward(self, ag, g, info):
        d = np.linalg.norm(ag - g, axis=-1)
        return -(d > self.distance_threshold).astype(np.float32)


def synth_episodes(rng, E, T=100, Do=27, Dg=3, Da=4):
    """random-walk episodes so that relabelled goals land on both sides of the 5 cm threshold"""
    obs = rng.standard_normal((E, T + 1, Do)) * 0.3
    ag = 0.3 + np.cumsum(rng.standard_normal((E, T + 1, Dg)) * 0.01, axis=1)
    obs[:, :, 12:15] = ag
    g = np.repeat(0.3 + rng.standard_normal((E, 1, Dg)) * 0.05, T, axis=1)
    acts = rng.uniform(-0.5, 0.5, (E, T, Da))
    return obs, ag, g, acts


def main():
    install_stubs()
    import her as ref_her
    import replay_buffer as ref_rb
    import normalizer as ref_norm
    rng = np.random.RandomState(7)

    # ---- HER + buffer ------------------------------------------------------------------
    env = FakeEnv()
    hs = ref_her.her_sampler('future', 4, env.compute_reward)
    params = {'obs': 27, 'goal': 3, 'action': 4, 'action_max': 0.5, 'max_timesteps': 100}
    rb = ref_rb.replay_buffer(params, 12 * 100, hs.sample_her_transitions)   # 12 episode slots
    store_log = []
    np.random.seed(125)
    eps_all = []
    for n in (5, 1, 4, 5, 3, 1):   # append, reach capacity mid-batch, then random overwrite
        ep = synth_episodes(rng, n)
        eps_all.append(ep)
        before = rb.current_size
        state = np.random.get_state()
        rb.store_episode(list(ep))
        store_log.append((n, before, rb.current_size))
    demo = np.load(os.path.join(REF, "bmirobot_1000_push_demo.npz"), allow_pickle=True)
    # overwrite two slots with real recorded episodes (actual env statistics incl. successes)
    rb.buffers['obs'][:2] = demo['obs'][:2]
    rb.buffers['ag'][:2] = demo['ag'][:2]
    rb.buffers['g'][:2] = demo['g'][:2]
    rb.buffers['actions'][:2] = demo['acs'][:2]
    np.random.seed(2024)
    tr = rb.sample(256)
    np.savez_compressed(
        os.path.join(OUT, "learner_her.npz"),
        buf_obs=rb.buffers['obs'], buf_ag=rb.buffers['ag'], buf_g=rb.buffers['g'], buf_actions=rb.buffers['actions'],
        current_size=rb.current_size, seed=2024, batch=256, future_p=hs.future_p,
        **{"tr_" + k: v for k, v in tr.items()})
    # storage-index trace: replay the same sequence and record idx
    np.random.seed(125)
    cs, idx_trace = 0, []
    rb2 = ref_rb.replay_buffer(params, 12 * 100, None)
    for n in (5, 1, 4, 5, 3, 1):
        idx = rb2._get_storage_idx(inc=n)
        idx_trace.append(np.atleast_1d(idx))
    np.savez_compressed(os.path.join(OUT, "learner_storage_idx.npz"), incs=np.array([5, 1, 4, 5, 3, 1]), size=12, seed=125,
                        idx=np.concatenate(idx_trace), final_size=rb2.current_size)

    # ---- normaliser ----------------------------------------------------------------------
    on = ref_norm.normalizer(27, default_clip_range=5)
    gn = ref_norm.normalizer(3, default_clip_range=5)
    feeds = []
    for i in range(4):
        v = np.clip(rng.standard_normal((100, 27)) * (0.1 + i) + 0.05 * i, -200, 200)
        w = 0.3 + rng.standard_normal((100, 3)) * 0.1
        feeds.append((v, w))
        on.update(v)
        gn.update(w)
        on.recompute_stats()
        gn.recompute_stats()
    probe = rng.standard_normal((16, 27)) * 3
    np.savez_compressed(
        os.path.join(OUT, "learner_norm.npz"),
        feeds_o=np.stack([f[0] for f in feeds]), feeds_g=np.stack([f[1] for f in feeds]),
        o_mean=on.mean, o_std=np.asarray(on.std), g_mean=gn.mean, g_std=np.asarray(gn.std),
        o_total_sum=on.total_sum, o_total_sumsq=on.total_sumsq, o_total_count=on.total_count,
        probe=probe, probe_norm=on.normalize(probe))

    # ---- full _update_network through the reference agent ---------------------------------
    import torch
    import arguments as ref_args
    import ddpg_agent as ref_agent
    args = ref_args.Args()
    args.add_demo = False
    args.save_dir = tempfile.mkdtemp(prefix="bmi_saved_") + "/"
    torch.manual_seed(3)
    torch.set_num_threads(1)
    ag = ref_agent.ddpg_agent(args, env, params)
    wr = np.random.RandomState(11)
    for net in (ag.actor_network, ag.critic_network):
        for _, p in net.named_parameters():
            bound = 1.0 / np.sqrt(p.shape[-1] if p.dim() > 1 else 256)
            p.data.copy_(torch.tensor(wr.uniform(-bound, bound, tuple(p.shape)).astype(np.float32)))
    ag.actor_target_network.load_state_dict(ag.actor_network.state_dict())
    ag.critic_target_network.load_state_dict(ag.critic_network.state_dict())
    ep = synth_episodes(rng, 8)
    ag.buffer.store_episode(list(ep))
    ag.buffer.buffers['obs'][:2] = demo['obs'][2:4]
    ag.buffer.buffers['ag'][:2] = demo['ag'][2:4]
    ag.buffer.buffers['g'][:2] = demo['g'][2:4]
    ag.buffer.buffers['actions'][:2] = demo['acs'][2:4]
    np.random.seed(99)
    ag._update_normalizer([ag.buffer.buffers[k][:2].copy() for k in ('obs', 'ag', 'g', 'actions')])
    from utils import _get_flat_params
    snaps = []
    for i in range(3):
        ag._update_network()
        snaps.append(np.concatenate([_get_flat_params(ag.actor_network)[0], _get_flat_params(ag.critic_network)[0]]))
    ag._soft_update_target_network(ag.actor_target_network, ag.actor_network)
    ag._soft_update_target_network(ag.critic_target_network, ag.critic_network)
    tgt = np.concatenate([_get_flat_params(ag.actor_target_network)[0], _get_flat_params(ag.critic_target_network)[0]])
    pick = np.random.RandomState(5).choice(snaps[0].shape[0], 4096, replace=False)
    np.savez_compressed(
        os.path.join(OUT, "learner_update.npz"),
        buf_obs=ag.buffer.buffers['obs'][:8], buf_ag=ag.buffer.buffers['ag'][:8], buf_g=ag.buffer.buffers['g'][:8],
        buf_actions=ag.buffer.buffers['actions'][:8], weight_seed=11, np_seed=99,
        o_mean=ag.o_norm.mean, o_std=np.asarray(ag.o_norm.std), g_mean=ag.g_norm.mean, g_std=np.asarray(ag.g_norm.std),
        pick=pick, params_after=np.stack([s[pick] for s in snaps]), target_after=tgt[pick],
        param_sums=np.array([s.astype(np.float64).sum() for s in snaps]),
        param_abs_sums=np.array([np.abs(s.astype(np.float64)).sum() for s in snaps]))
    print("wrote", sorted(f for f in os.listdir(OUT) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
