"""DDPG + HER agent over the vectorised CUDA env.

Mirror of the reference ``ddpg_agent.py:17-304``: ``ddpg_agent(args, env, env_params)``, ``learn()``,
``plot_success_rate()``, ``success_rates`` and the private steps ``_init_demo_buffer``,
``_update_normalizer``, ``_update_network``, ``_soft_update_target_network``, ``_eval_agent``; the
checkpoint tuple ``[o_mean, o_std, g_mean, g_std, actor.state_dict()]`` is the reference's
(ddpg_agent.py:155-161) so ``demo_push.py`` can load it.

What differs is where the arithmetic runs: ``num_rollouts_per_mpi`` episodes are rolled out
SIMULTANEOUSLY (one env instance per CUDA block), the whole T-step rollout is one CUDA graph
(normalise -> actor GEMMs -> exploration noise -> record -> physics), and the ``n_batches`` updates
are one CUDA graph (HER gather -> 5 forward / 2 backward passes -> NCCL gradient sum -> Adam).
Nothing in this file touches host memory inside those graphs.
"""
import ctypes
import os
import time
from datetime import datetime

import numpy as np
import torch

from . import _lib, utils
from .her import her_sampler
from .models import actor, critic
from .normalizer import normalizer
from .replay_buffer import replay_buffer


class ddpg_agent:
    def __init__(self, args, env, env_params):
        self.savetime = 0
        self.args = args
        self.env = env
        self.vec = getattr(env, "_vec", env)          # BmiVecEnv behind the gym-style wrapper
        self.env_params = env_params
        self.device = self.vec.device
        self.R = self.vec.n_envs
        self.T = int(env_params['max_timesteps'])
        Do, Dg, Da = env_params['obs'], env_params['goal'], env_params['action']
        # networks: built with torch's CPU generator exactly like the reference (ddpg_agent.py:24-25),
        # then flattened onto the device; rank 0's weights are broadcast (utils.py:6-15)
        self.actor_network = actor(env_params, device=self.device)
        self.critic_network = critic(env_params, device=self.device)
        utils.sync_networks(self.actor_network)
        utils.sync_networks(self.critic_network)
        self.actor_target_network = actor(env_params, device=self.device)
        self.critic_target_network = critic(env_params, device=self.device)
        self.actor_target_network.flat.copy_(self.actor_network.flat)
        self.critic_target_network.flat.copy_(self.critic_network.flat)
        cfg = _lib.DdpgConfig(Do, Dg, Da, 256, int(args.batch_size), max(self.R, 1), float(env_params['action_max']),
                              float(args.gamma), float(args.action_l2), float(args.lr_actor), float(args.lr_critic),
                              float(args.polyak), 0.9, 0.999, 1e-8, float(1.0 / (1.0 - args.gamma)),
                              float(1.0 - args.polyak), 0.0)
        self._cfg = cfg
        h = ctypes.c_void_p()
        _lib.call("bmi_ddpg_create", ctypes.byref(h), ctypes.byref(cfg), _lib.ptr(self.actor_network.flat),
                  _lib.ptr(self.critic_network.flat), _lib.ptr(self.actor_target_network.flat),
                  _lib.ptr(self.critic_target_network.flat))
        self._h = h
        gp, gn = ctypes.c_void_p(), ctypes.c_int64()
        _lib.call("bmi_ddpg_grad_buffer", h, ctypes.byref(gp), ctypes.byref(gn))
        self._grad_ptr, self._grad_n = gp, int(gn.value)
        self._p2p = False
        if utils.world_size() > 1 and getattr(args, "p2p_adam", True):
            self._p2p = self._attach_peers()
        # her sampler / replay buffer / demos / normalisers (ddpg_agent.py:44-53)
        self.her_module = her_sampler(args.replay_strategy, args.replay_k, self.vec.compute_reward,
                                      distance_threshold=self.vec.distance_threshold)
        dt = torch.float64 if getattr(args, "buffer_dtype", "float32") == "float64" else torch.float32
        self.buffer = replay_buffer(env_params, args.buffer_size, self.her_module.sample_her_transitions, dtype=dt,
                                    device=self.device, verbose=getattr(args, "verbose", True))
        if args.add_demo:
            self._init_demo_buffer()
        self.o_norm = normalizer(size=Do, default_clip_range=args.clip_range, device=self.device)
        self.g_norm = normalizer(size=Dg, default_clip_range=args.clip_range, device=self.device)
        self.success_rates = []
        # ---- device work buffers ------------------------------------------------------------------
        f32 = lambda *s: torch.zeros(s, dtype=torch.float32, device=self.device)
        R, T, B = self.R, self.T, int(args.batch_size)
        self.ep = {'obs': f32(R, T + 1, Do), 'ag': f32(R, T + 1, Dg), 'g': f32(R, T, Dg), 'actions': f32(R, T, Da)}
        self._x_pol, self._pi, self._act = f32(R, Do + Dg), f32(R, Da), f32(R, Da)
        self._x, self._xn, self._a, self._r = f32(B, Do + Dg), f32(B, Do + Dg), f32(B, Da), f32(B)
        self._losses = f32(2)
        self._draw = (torch.zeros(B, dtype=torch.int64, device=self.device), torch.zeros(B, dtype=torch.int64, device=self.device),
                      torch.zeros(B, dtype=torch.float64, device=self.device), torch.zeros(B, dtype=torch.float64, device=self.device))
        self._actor_t = torch.zeros_like(self.actor_network.flat)   # transposed copy read by the fused rollout
        self._ctr_her = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._ctr_explore = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._seed = int(args.seed) + utils.rank()
        self._graphs = {}
        self.env_steps = 0
        self.updates = 0
        if utils.rank() == 0:
            os.makedirs(args.save_dir, exist_ok=True)
            self.model_path = os.path.join(args.save_dir, args.env_name)
            os.makedirs(self.model_path, exist_ok=True)

    def _attach_peers(self):
        """Map every rank's gradient buffer into this process (CUDA IPC over NVLink) for the fused
        sum-over-ranks + Adam kernel; returns False (NCCL allreduce + Adam is used instead) if the mapping fails."""
        import torch.distributed as dist
        buf = (ctypes.c_uint8 * 128)()
        ok = True
        try:
            _lib.call("bmi_ddpg_p2p_export", self._h, ctypes.cast(buf, ctypes.c_void_p))
        except _lib.BmiError:
            ok = False
        gathered = [None] * utils.world_size()
        dist.all_gather_object(gathered, bytes(buf) if ok else None)
        if any(g is None for g in gathered):
            return False
        allh = (ctypes.c_uint8 * (128 * utils.world_size())).from_buffer_copy(b"".join(gathered))
        try:
            _lib.call("bmi_ddpg_p2p_attach", self._h, utils.rank(), utils.world_size(), ctypes.cast(allh, ctypes.c_void_p))
        except _lib.BmiError:
            ok = False
        flags = [None] * utils.world_size()
        dist.all_gather_object(flags, ok)
        return all(flags)

    def p2p_timed_out(self):
        t = ctypes.c_int32(0)
        _lib.call("bmi_ddpg_p2p_status", self._h, ctypes.byref(t))
        return bool(t.value)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.call("bmi_ddpg_destroy", h)
            except Exception:
                pass
            self._h = None

    # ------------------------------------------------------------------------------------------------
    def plot_success_rate(self):
        """ddpg_agent.py:73-80 (the plot needs matplotlib, which is optional here)."""
        saved_dir = 'test_rates/'
        os.makedirs(saved_dir, exist_ok=True)
        np.save(saved_dir + str(self.args.seed) + '_' + str(self.args.add_demo) + '_success_rates.npy',
                np.array(self.success_rates))
        try:
            import matplotlib.pyplot as plt
            plt.plot(self.success_rates)
            plt.show()
        except ImportError:
            pass

    def _resolve_demo(self):
        """args.demo_name as given (cwd, like the reference), else next to the package / the repo root"""
        name = self.args.demo_name
        here = os.path.dirname(os.path.abspath(__file__))
        for cand in (name, os.path.join(here, name), os.path.join(os.path.dirname(here), name)):
            if os.path.exists(cand):
                return cand
        raise FileNotFoundError(
            "add_demo=True but the demo file %r was not found (cwd, %s, %s).  Copy the reference's .npz there or "
            "generate one with `python -m rl_arm_under_sparse_reward_b200.get_demo_data --task push --n 1000`; "
            "Args.add_demo=False trains without demonstrations." % (name, here, os.path.dirname(here)))

    def _init_demo_buffer(self):
        """ddpg_agent.py:82-90: preload the expert episodes (normalisers are NOT updated from them)."""
        demo = np.load(self._resolve_demo(), allow_pickle=True)
        self.buffer.store_episode([np.array(demo['obs']), np.array(demo['ag']), np.array(demo['g']), np.array(demo['acs'])])

    # ---- rollout -------------------------------------------------------------------------------------
    def _episodes_struct(self):
        e = self.ep
        return _lib.Episodes(_lib.ptr(e['obs']), _lib.ptr(e['ag']), _lib.ptr(e['g']), _lib.ptr(e['actions']), self.R,
                             self.T, e['obs'].shape[2], e['ag'].shape[2], e['actions'].shape[2], _lib.BMI_F32, 0)

    def _policy(self, obs, g, explore, late_clip):
        """ddpg_agent.py:113-119: _preproc_inputs -> actor -> _select_actions (+ late action clip)."""
        p, st = self.env_params, _lib.stream_ptr()
        _lib.call("bmi_preproc_inputs", _lib.ptr(obs), _lib.ptr(g), self.R, p['obs'], p['goal'], _lib.BMI_F32,
                  _lib.ptr(self.o_norm.mean_dev), _lib.ptr(self.o_norm.std_dev), _lib.ptr(self.g_norm.mean_dev),
                  _lib.ptr(self.g_norm.std_dev), float(self.args.clip_range), _lib.ptr(self._x_pol), st)
        _lib.call("bmi_ddpg_act", self._h, _lib.ptr(self._x_pol), self.R, 0, _lib.ptr(self._pi), st)
        if not explore:
            return self._pi
        _lib.call("bmi_select_actions", _lib.ptr(self._pi), self.R, p['action'], float(p['action_max']),
                  float(self.args.noise_eps), float(self.args.random_eps), float(late_clip),
                  ctypes.c_uint64(self._seed), _lib.ptr(self._ctr_explore), _lib.ptr(self._act), st)
        return self._act

    def _rollout_body(self, late_clip):
        eps = self._episodes_struct()
        st = _lib.stream_ptr()
        obs, ag, g = self.vec.reset()
        for t in range(self.T):
            act = self._policy(obs, g, True, late_clip)
            _lib.call("bmi_rollout_record", ctypes.byref(eps), t, _lib.ptr(obs), _lib.ptr(ag), _lib.ptr(g), _lib.ptr(act), st)
            obs, ag, _, _ = self.vec.step(act)
        _lib.call("bmi_rollout_record", ctypes.byref(eps), self.T, _lib.ptr(obs), _lib.ptr(ag), None, None, st)

    def _run_graphed(self, key, body):
        """Run `body` (a sequence of stream-ordered launches) through a cached CUDA graph."""
        if not getattr(self.args, "use_cuda_graphs", True):
            body()
            return
        g = self._graphs.get(key)
        if g is None:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):   # eager warm-up (creates cuBLASLt plans) off the default stream
                body()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                body()
            self._graphs[key] = g
            return   # the warm-up already did this call's work
        g.replay()

    def _fused_rollout_body(self, explore, late_clip):
        p = self.env_params
        _lib.call("bmi_actor_transpose", _lib.ptr(self.actor_network.flat), p['obs'], p['goal'], p['action'], 256,
                  _lib.ptr(self._actor_t), _lib.stream_ptr())
        self.vec.rollout(self.T, self._actor_t, self.o_norm, self.g_norm, self.args.clip_range, explore,
                         noise_eps=self.args.noise_eps, random_eps=self.args.random_eps, late_clip=late_clip,
                         seed=self._seed, counter=self._ctr_explore, episodes=self.ep if explore else None)

    def rollout(self, epoch=0):
        """One batch of R simultaneous episodes (ddpg_agent.py:103-141); fills self.ep.  Default: the fused
        per-env rollout kernel (one launch per episode batch); args.fused_rollout=False runs the step-wise
        pipeline (cuBLASLt actor + one env kernel per step) inside a CUDA graph."""
        late = 0.15 if epoch >= getattr(self.args, "late_clip_epoch", 100) else 0.0
        if getattr(self.args, "fused_rollout", True):
            self._fused_rollout_body(True, late)
        else:
            self._run_graphed(("rollout", late), lambda: self._rollout_body(late))
        self.env_steps += self.R * self.T

    # ---- normaliser ------------------------------------------------------------------------------------
    def _update_normalizer(self, episode_batch=None):
        """ddpg_agent.py:187-212: HER-sample from the NEW episodes, clip +-clip_obs, update, recompute.
        The reference draws T transitions from its 2 new episodes (50 % of them); with R simultaneous
        episodes the same fraction is kept: T * max(1, R // 2) transitions."""
        if episode_batch is None:
            obs, ag, g, act = self.ep['obs'], self.ep['ag'], self.ep['g'], self.ep['actions']
        else:
            obs, ag, g, act = [torch.as_tensor(np.ascontiguousarray(a)).to(self.device) if isinstance(a, np.ndarray)
                               else a.contiguous() for a in episode_batch]
        R = obs.shape[0]
        n = self.T * max(1, R // 2)
        if getattr(self.args, "device_rng", True):
            draws = self._device_draws(n, R)
        else:
            draws = self.her_module.draw(R, self.T, n)
        out = {"obs": torch.empty((n, obs.shape[2]), dtype=obs.dtype, device=self.device),
               "g": torch.empty((n, g.shape[2]), dtype=obs.dtype, device=self.device)}
        self.her_module.sample_device(obs, ag, g, act, R, draws, out=out)
        self.o_norm.update(out["obs"], pre_clip=self.args.clip_obs)
        self.g_norm.update(out["g"], pre_clip=self.args.clip_obs)
        self.o_norm.recompute_stats()
        self.g_norm.recompute_stats()

    def _device_draws(self, n, n_valid):
        d = (torch.empty(n, dtype=torch.int64, device=self.device), torch.empty(n, dtype=torch.int64, device=self.device),
             torch.empty(n, dtype=torch.float64, device=self.device), torch.empty(n, dtype=torch.float64, device=self.device))
        nv = torch.full((1,), int(n_valid), dtype=torch.int64, device=self.device)
        _lib.call("bmi_her_draw", ctypes.c_uint64(self._seed), _lib.ptr(self._ctr_her), n, _lib.ptr(nv), self.T,
                  _lib.ptr(d[0]), _lib.ptr(d[1]), _lib.ptr(d[2]), _lib.ptr(d[3]), _lib.stream_ptr())
        return d

    def _preproc_og(self, o, g):
        """ddpg_agent.py:214-217 (kept for API parity; the kernels fuse this clip)."""
        return torch.clamp(o, -self.args.clip_obs, self.args.clip_obs), torch.clamp(g, -self.args.clip_obs, self.args.clip_obs)

    # ---- updates ---------------------------------------------------------------------------------------
    def _soft_update_target_network(self, target=None, source=None):
        """ddpg_agent.py:220-222.  Without arguments: both target nets in one fused launch (what learn() uses).  With the
        reference's (target, source) arguments: that one net, target <- (1 - polyak) source + polyak target with the
        reference's rounding (two products, one sum)."""
        if target is None and source is None:
            _lib.call("bmi_ddpg_soft_update", self._h, _lib.stream_ptr())
            return
        if target is None or source is None or target.flat.shape != source.flat.shape:
            raise ValueError("_soft_update_target_network(target, source): both networks of the same architecture, or neither")
        p = float(self.args.polyak)
        target.flat.copy_((1 - p) * source.flat + p * target.flat)

    def _update_body(self, draws=None):
        a, st, b = self.args, _lib.stream_ptr(), self.buffer
        bufs = b.buffers
        eps = _lib.Episodes(_lib.ptr(bufs['obs']), _lib.ptr(bufs['ag']), _lib.ptr(bufs['g']), _lib.ptr(bufs['actions']),
                            b.size, b.T, bufs['obs'].shape[2], bufs['ag'].shape[2], bufs['actions'].shape[2],
                            _lib.dtype_code(bufs['obs'].dtype), 0)
        B = int(a.batch_size)
        if draws is None:
            d = self._draw
            _lib.call("bmi_her_draw", ctypes.c_uint64(self._seed), _lib.ptr(self._ctr_her), B, _lib.ptr(b.current_size_dev),
                      b.T, _lib.ptr(d[0]), _lib.ptr(d[1]), _lib.ptr(d[2]), _lib.ptr(d[3]), st)
        else:
            d = draws
        _lib.call("bmi_her_sample_inputs", ctypes.byref(eps), -1, _lib.ptr(d[0]), _lib.ptr(d[1]), _lib.ptr(d[2]),
                  _lib.ptr(d[3]), B, float(self.her_module.future_p), float(self.her_module.distance_threshold),
                  float(a.clip_obs), float(a.clip_range), _lib.ptr(self.o_norm.mean_dev), _lib.ptr(self.o_norm.std_dev),
                  _lib.ptr(self.g_norm.mean_dev), _lib.ptr(self.g_norm.std_dev), _lib.ptr(self._x), _lib.ptr(self._xn),
                  _lib.ptr(self._a), _lib.ptr(self._r), st)
        self._learn_from(self._x, self._xn, self._a, self._r)

    def _update_network(self):
        """ddpg_agent.py:225-277, one update.  With device_rng=False the four HER arrays come from numpy's
        global stream in the reference order, so the sampled batch is bit-identical to the reference's."""
        if self.buffer.current_size == 0:
            raise ValueError("cannot update from an empty replay buffer")
        if getattr(self.args, "device_rng", True):
            self._update_body()
        else:
            h = self.her_module.draw(self.buffer.current_size, self.T, int(self.args.batch_size))
            d = (torch.as_tensor(h[0]).to(self.device), torch.as_tensor(h[1]).to(self.device),
                 torch.as_tensor(h[2]).to(self.device), torch.as_tensor(h[3]).to(self.device))
            self._update_body(d)
        self.updates += 1

    def update_many(self, n):
        """n_batches updates as one CUDA graph (device_rng path)."""
        if not getattr(self.args, "device_rng", True):
            for _ in range(n):
                self._update_network()
            return
        if self.buffer.current_size == 0:
            raise ValueError("cannot update from an empty replay buffer")

        def body():
            self._sample_batches(n)
            for i in range(n):
                self._learn_from(self._XA[i], self._XNA[i], self._AA[i], self._RA[i])
        self._run_graphed(("update", n), body)
        self.updates += n

    def _sample_batches(self, n):
        """HER-sample ALL n batches of the cycle in ONE draw launch + ONE gather launch (n x batch_size transitions;
        SURVEY 8d: a batch-256 gather is launch-bound).  Equivalent to n separate her.py:13-41 calls: the replay buffer
        does not change between the updates of a cycle, the sampler does not depend on the networks, and the
        counter-based Philox draws of sample k are the same whichever launch produces them."""
        a, st, b = self.args, _lib.stream_ptr(), self.buffer
        B = int(a.batch_size)
        if getattr(self, "_XA", None) is None or self._XA.shape[0] != n:
            f32 = lambda *s: torch.zeros(s, dtype=torch.float32, device=self.device)
            Dx = self.env_params['obs'] + self.env_params['goal']
            self._XA, self._XNA, self._AA, self._RA = f32(n, B, Dx), f32(n, B, Dx), f32(n, B, self.env_params['action']), f32(n, B)
            self._draw_all = (torch.zeros(n * B, dtype=torch.int64, device=self.device), torch.zeros(n * B, dtype=torch.int64, device=self.device),
                              torch.zeros(n * B, dtype=torch.float64, device=self.device), torch.zeros(n * B, dtype=torch.float64, device=self.device))
        bufs, d = b.buffers, self._draw_all
        eps = _lib.Episodes(_lib.ptr(bufs['obs']), _lib.ptr(bufs['ag']), _lib.ptr(bufs['g']), _lib.ptr(bufs['actions']),
                            b.size, b.T, bufs['obs'].shape[2], bufs['ag'].shape[2], bufs['actions'].shape[2],
                            _lib.dtype_code(bufs['obs'].dtype), 0)
        _lib.call("bmi_her_draw", ctypes.c_uint64(self._seed), _lib.ptr(self._ctr_her), n * B, _lib.ptr(b.current_size_dev),
                  b.T, _lib.ptr(d[0]), _lib.ptr(d[1]), _lib.ptr(d[2]), _lib.ptr(d[3]), st)
        _lib.call("bmi_her_sample_inputs", ctypes.byref(eps), -1, _lib.ptr(d[0]), _lib.ptr(d[1]), _lib.ptr(d[2]),
                  _lib.ptr(d[3]), n * B, float(self.her_module.future_p), float(self.her_module.distance_threshold),
                  float(a.clip_obs), float(a.clip_range), _lib.ptr(self.o_norm.mean_dev), _lib.ptr(self.o_norm.std_dev),
                  _lib.ptr(self.g_norm.mean_dev), _lib.ptr(self.g_norm.std_dev), _lib.ptr(self._XA), _lib.ptr(self._XNA),
                  _lib.ptr(self._AA), _lib.ptr(self._RA), st)

    def _learn_from(self, x, xn, act, r):
        """ddpg_agent.py:250-277 on one prepared batch: 5 forward + 2 backward passes, gradient sum over ranks, Adam"""
        st = _lib.stream_ptr()
        _lib.call("bmi_ddpg_backward", self._h, _lib.ptr(x), _lib.ptr(xn), _lib.ptr(act), _lib.ptr(r), _lib.ptr(self._losses), st)
        if self._p2p:                # sync_grads of both nets (SUM, utils.py:43-48) fused into the Adam kernel over NVLink
            _lib.call("bmi_ddpg_adam_step_p2p", self._h, st)
            return
        if utils.world_size() > 1:   # NCCL variant: both nets in ONE collective, then Adam
            _lib.call("bmi_comm_allreduce_sum_f32", utils._state["comm"], self._grad_ptr, self._grad_n, st)
        _lib.call("bmi_ddpg_adam_step", self._h, st)

    def check_p2p(self):
        """utils.py:43-48 semantics must hold on every update: a timed-out peer-memory gradient sum is fatal (the kernel
        froze this replica's parameters instead of applying a partial sum)."""
        if self._p2p and self.p2p_timed_out():
            raise RuntimeError("fused peer-memory gradient sum timed out: a peer rank stopped responding; the "
                               "replicas are no longer synchronised (restart, or set Args.p2p_adam=False for NCCL)")

    def release_graphs(self):
        """Drop the captured CUDA graphs (they hold references to the NCCL communicator)."""
        torch.cuda.synchronize()
        self._graphs.clear()

    def losses(self):
        return self._losses.cpu().numpy().copy()

    # ---- training loop -----------------------------------------------------------------------------------
    def learn(self):
        a = self.args
        if getattr(a, "verbose", True):
            print("initial buffer size:", self.buffer.current_size)
        for epoch in range(a.n_epochs):
            start_time = time.time()
            for _ in range(a.n_cycles):
                self.rollout(epoch)
                self.buffer.store_episode([self.ep['obs'], self.ep['ag'], self.ep['g'], self.ep['actions']])
                self._update_normalizer()
                self.update_many(a.n_batches)
                self._soft_update_target_network()
                self.check_p2p()
            torch.cuda.synchronize()
            if getattr(a, "verbose", True):
                print(str(time.time() - start_time))
            success_rate = self._eval_agent()
            self.success_rates.append(success_rate)
            if utils.rank() == 0:
                if getattr(a, "verbose", True):
                    print('[{}] epoch is: {}, eval success rate is: {:.3f}'.format(datetime.now(), epoch, success_rate))
                self.save_checkpoint()

    def save_checkpoint(self):
        """ddpg_agent.py:155-161: same tuple, CPU tensors / numpy arrays so the reference's loader works."""
        self.savetime += 1
        sd = {k: v.detach().cpu().clone() for k, v in self.actor_network.state_dict().items()}
        path = self.model_path + '/' + str(self.args.seed) + '_' + str(self.args.add_demo) + str(self.savetime) + '_model.pt'
        torch.save([self.o_norm.mean, self.o_norm.std, self.g_norm.mean, self.g_norm.std, sd], path)
        return path

    def _eval_body(self):
        obs, ag, g = self.vec.reset()
        for _ in range(self.T):
            act = self._policy(obs, g, False, 0.0)
            obs, ag, _, _ = self.vec.step(act)

    def _eval_agent(self):
        """ddpg_agent.py:280-304: noise-free episodes, success = is_success at the LAST step, averaged over
        ranks.  ceil(n_test_rollouts / R) batches of R simultaneous episodes."""
        n_test = int(self.args.n_test_rollouts)
        if getattr(self.args, "eval_all_envs", False):     # vectorised runs: every env of the batch is an eval episode
            n_test = max(n_test, self.R)
        n_batches = max(1, -(-n_test // self.R))
        total = torch.zeros((), dtype=torch.float32, device=self.device)
        left = n_test
        for _ in range(n_batches):
            if getattr(self.args, "fused_rollout", True):
                self._fused_rollout_body(False, 0.0)
            else:
                self._run_graphed(("eval",), self._eval_body)
            k = min(left, self.R)              # exactly n_test_rollouts episodes count (the reference runs 25, not 26)
            total += self.vec.success[:k].sum()
            left -= k
        local = torch.stack([total / n_test]).contiguous()
        utils.allreduce_sum_(local)
        return float(local.item()) / utils.world_size()
