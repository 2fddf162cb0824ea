#!/usr/bin/env python
"""Benchmark of the rollout -> HER-relabel -> DDPG-update cycle (BASELINE.json metric: env-steps/s,
push task, 4096 envs per GPU).

One "step" = one training CYCLE of the reference loop (ddpg_agent.py:101-150) with
num_rollouts_per_mpi = 4096 simultaneous episodes: 4096 x 100 env-steps (policy + IK + 20 physics
sub-steps + observation each), store_episode, normaliser update, n_batches = 40 HER-relabelled DDPG
updates of batch 256, Polyak.  Update:data ratio = 40 updates per 409 600 env-steps per rank (the
reference's loop shape at R = 4096; its R = 2 default gives 1 update per 5 env-steps — see DESIGN.md).

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (CPU arm: the oracle port on all host cores)
    python bench.py --workload pick ...                      (BASELINE config 3: pick-and-place, 4096 envs, add_demo)
    python bench.py --workload replay-stress ...             (BASELINE config 5: 5e5 stored transitions, HER batch sweep)

The buffer is pre-filled with 1000 scripted-controller demonstrations of the task (add_demo=True, the reference's
bmirobot_1000_{push,pick}_demo.npz role; generated on the device before the timed region because the reference's
18 MB files do not travel).  Multi-rank runs end with a parity probe: MD5 of every rank's parameters (must be
identical) and the fused peer-memory gradient sum + Adam against the NCCL allreduce path on identical data.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ENVS = 4096
T = 100
ALGO_BYTES_PER_ENV_STEP = 412          # state in+out 2x136 + action 16 + episode write 124
# dram__bytes_read.sum + dram__bytes_write.sum of ONE rollout_kernel launch (4096 envs x 100 steps), from the
# `ncu --set full` capture summarised in profiles/r02_rollout_kernel_ncu.md (bench.py cannot run ncu itself)
NCU_DRAM_BYTES_PER_LAUNCH = 108596736  # 17.92 MB read + 90.68 MB written
# sm__warps_active / smsp__issue_active / FMA pipe of the same capture; numeric so that the roofline object says what bounds
# this FP32-issue kernel (the HBM fraction cannot): 28 of 64 warp slots resident by design (shared memory), issue slots 75 % busy
# (T = 100 capture.  The final kernel -- cheaper motor / contact-normal rows, block island started cold -- was re-captured at
# T = 10 only: 3.156e10 warp instructions against 3.153e10 for the T = 10 capture of the same table, i.e. the same 36.6 k)
NCU_ROLLOUT = {"warps_active_pct": 42.2, "issue_slot_util_pct": 75.3, "fma_pipe_pct": 28.6, "lsu_pipe_pct": 60.6,
               "warp_instructions_per_env_substep": 36600}
METRIC = "env-steps/s (push, 4096 envs) at 1/2/4/8 B200 vs CPU PyBullet+MPI"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------ CPU arm
def _oracle_worker(args):
    seed, n_steps = args
    from oracle.physics_oracle import OracleEnv
    import random
    rng = np.random.RandomState(seed)
    random.seed(seed)
    e = OracleEnv(0)
    done = 0
    t0 = time.perf_counter()
    while done < n_steps:
        while True:
            x, y = 0.15 + 0.2 * random.random(), random.random() * 0.3 + 0.2
            ang = 3.14 * 0.5 + 3.1415925438 * random.random()
            xt, yt = 0.35 * random.random(), random.random() * 0.3 + 0.2
            if ((x - xt) ** 2 + (y - yt) ** 2) ** 0.5 >= 0.15:
                break
        e.reset([x, y, 0.2, ang, xt, yt, 0.2, 0.0])
        for _ in range(min(T, n_steps - done)):
            # exploration-like actions: sigma 0.005 around 0 with 30 % uniform (ddpg_agent.py:174-184)
            a = rng.uniform(-0.5, 0.5, 4) if rng.uniform() < 0.3 else 0.005 * rng.standard_normal(4)
            e.step(a)
            done += 1
    return done, time.perf_counter() - t0


def cpu_env_steps_per_s(steps_per_core, cores):
    """A bounded sample of the cycle on the host: `cores` processes each run `steps_per_core` env-steps of the
    oracle port (independent envs, like the reference's one-env-per-MPI-rank layout), then the torch-CPU
    restatement of _update_network runs the proportional number of updates (40 per 409 600 env-steps).
    Throughput = env-steps / (slowest worker + updates); process start-up is not timed."""
    import multiprocessing as mp
    from oracle import physics_oracle
    physics_oracle.build()
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        res = pool.map(_oracle_worker, [(1000 + i, steps_per_core) for i in range(cores)])
    total = sum(r[0] for r in res)
    t_env = max(r[1] for r in res)
    n_upd = max(1, int(round(total * 40.0 / (N_ENVS * T))))
    import torch
    from oracle import ddpg_oracle
    torch.set_num_threads(cores)
    L = ddpg_oracle.Learner()
    g = torch.Generator().manual_seed(0)
    x, xn = torch.randn(256, 30, generator=g), torch.randn(256, 30, generator=g)
    act, r = torch.rand(256, 4, generator=g) - 0.5, -torch.ones(256, 1)
    L.update(x, xn, act, r)                      # warm-up (autograd graph caches, allocator)
    t0 = time.perf_counter()
    for _ in range(n_upd):
        L.update(x, xn, act, r)
    t_upd = time.perf_counter() - t0
    wall = t_env + t_upd
    return total / wall, total, wall


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_core = 300          # ~1.5 s per step at ~200 env-steps/s/core (GJK + EPA self-collision, 150 solver iterations)
    for _ in range(args.warmup):
        cpu_env_steps_per_s(20, cores)
    vals, t_ms = [], []
    for _ in range(args.steps):
        v, total, wall = cpu_env_steps_per_s(per_core, cores)
        vals.append(v)
        t_ms.append(wall * 1e3)
    value = float(np.mean(vals))
    sample = ("per step: %d processes x %d env-steps of the C oracle port (restated env step, NOT PyBullet: pybullet/gym/"
              "mpi4py are not installable in this image; faithful mode: GJK + EPA self-collision on the full hulls, Bullet's plain "
              "150-iteration solver loop) + the proportional DDPG updates on torch CPU" % (cores, per_core))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(np.mean(t_ms)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "push task, one env per host core, cycle shape of the GPU arm (100-step episodes, 40 HER DDPG updates "
                                   "of batch 256 per 409600 env-steps); bounded sample per step", "envs": cores,
                       "updates_per_env_step": 40.0 / (N_ENVS * T)},
            "reference_arm": "C port of the reference env step (oracle/), NOT PyBullet + mpi4py: not installable here",
            "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm
def _demo_file(task, rank, n_demo, device_index):
    """add_demo=True (ddpg_agent.py:82-90): `n_demo` successful scripted episodes in the reference's .npz format,
    generated on the device with the reference's scripted controllers (get_demo_data.py) -- the stand-in for
    bmirobot_1000_{push,pick}_demo.npz, which do not travel to the GPU box.  Not timed."""
    from rl_arm_under_sparse_reward_b200 import get_demo_data
    path = "/tmp/bmi_bench_demo_%s_%d.npz" % (task, rank)
    demo, rate = get_demo_data.get_demo(task, n_demo, n_envs=2048, seed=1000 + rank, max_batches=24, verbose=False)
    np.savez_compressed(path, acs=demo["acs"], obs=demo["obs"], info=demo["info"], g=demo["g"], ag=demo["ag"])
    return path, int(demo["acs"].shape[0]), float(rate)


def _digest(agent):
    import hashlib
    return hashlib.md5(agent.actor_network.flat.cpu().numpy().tobytes() + agent.critic_network.flat.cpu().numpy().tobytes()).hexdigest()


def parity_probe(world, rank, seed):
    """Multi-rank proof carried by the bench line (the 1-GPU test box cannot run tests/test_gpu_multi.py): two small
    agents with identical seeds and data, one with the fused peer-memory gradient sum + Adam kernel, one with the NCCL
    allreduce (sum) + Adam path (utils.py:43-48 semantics), 1 rollout + 6 updates each.  Returns whether every rank
    ends with identical parameters on each path and how far the two paths are apart (2 ranks: a + b is order-free, so
    bit-identical; more ranks: NCCL's reduction order differs from the kernel's rank order by float rounding)."""
    import torch
    import torch.distributed as dist
    from rl_arm_under_sparse_reward_b200.arguments import Args
    from rl_arm_under_sparse_reward_b200.bmirobot_env.vec_env import BmiVecEnv
    from rl_arm_under_sparse_reward_b200.ddpg_agent import ddpg_agent
    from rl_arm_under_sparse_reward_b200.train import get_env_params
    out = {}
    flats = {}
    for name, p2p in (("p2p", True), ("nccl", False)):
        a = Args()
        a.add_demo, a.verbose, a.n_envs, a.buffer_size, a.save_dir = False, False, 64, 256 * 100, "/tmp/bmi_probe_%d/" % rank
        a.p2p_adam = p2p
        torch.manual_seed(seed)
        env = BmiVecEnv(a.n_envs, seed=seed + rank)
        ag = ddpg_agent(a, env, get_env_params(env))
        ag.rollout(0)
        ag.buffer.store_episode([ag.ep['obs'], ag.ep['ag'], ag.ep['g'], ag.ep['actions']])
        ag._update_normalizer()
        a.use_cuda_graphs = False
        ag.update_many(6)
        torch.cuda.synchronize()
        digests = [None] * world
        dist.all_gather_object(digests, _digest(ag))
        out[name + "_params_identical_across_ranks"] = len(set(digests)) == 1
        out[name + "_attached"] = bool(ag._p2p)
        if p2p:
            out["p2p_timed_out"] = bool(ag.p2p_timed_out()) if ag._p2p else False
        flats[name] = torch.cat([ag.actor_network.flat, ag.critic_network.flat]).clone()
        ag.release_graphs()
        del ag, env
    d = (flats["p2p"] - flats["nccl"]).abs().max().item()
    out["p2p_equals_nccl"] = bool(d == 0.0)
    out["p2p_vs_nccl_max_abs_diff"] = d
    return out


def run_replay_stress(args):
    """BASELINE config 5: 5e5 stored transitions per rank, HER 'future' (k = 4) relabel + network-input assembly, batch
    sweep.  Each rank owns its buffer (the reference design, train.py:34-39): weak scaling, no collective on the data
    path.  value = transitions/s at the largest batch, summed over ranks (max-over-ranks time)."""
    import ctypes
    import torch
    import torch.distributed as dist
    from rl_arm_under_sparse_reward_b200 import _lib, utils
    rank, world = utils.init_comm()
    if world == 1:
        torch.cuda.set_device(0)
    dev = torch.device("cuda", torch.cuda.current_device())
    if os.environ.get("BMI_L2_FETCH"):       # experiment knob: DRAM bytes fetched per L2 miss (32 / 64 / 128)
        _lib.call("bmi_set_l2_fetch_granularity", int(os.environ["BMI_L2_FETCH"]))
    E = args.stress_episodes
    g = torch.Generator(device=dev).manual_seed(125 + rank)
    obs = torch.randn(E, T + 1, 27, device=dev, generator=g)
    ag = 0.3 + 0.1 * torch.randn(E, T + 1, 3, device=dev, generator=g)
    gg = 0.3 + 0.1 * torch.randn(E, T, 3, device=dev, generator=g)
    act = torch.rand(E, T, 4, device=dev, generator=g) - 0.5
    eps = _lib.Episodes(_lib.ptr(obs), _lib.ptr(ag), _lib.ptr(gg), _lib.ptr(act), E, T, 27, 3, 4, _lib.BMI_F32, 0)
    stats = [torch.zeros(27, device=dev), torch.ones(27, device=dev), torch.zeros(3, device=dev), torch.ones(3, device=dev)]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    flush_sink = torch.zeros((), dtype=torch.float32, device=dev)
    ctr = torch.zeros(1, dtype=torch.int64, device=dev)
    nv = torch.tensor([E], dtype=torch.int64, device=dev)
    peak, peak_src = peaks()
    n0 = _lib.launch_count()
    rows = []
    sampler = ClockSampler(torch.cuda.current_device())
    if rank == 0:
        sampler.start()
    batches = (256, 1024, 4096, 16384, 65536, 262144, 1048576)
    for B in batches:
        d = (torch.empty(B, dtype=torch.int64, device=dev), torch.empty(B, dtype=torch.int64, device=dev),
             torch.empty(B, dtype=torch.float64, device=dev), torch.empty(B, dtype=torch.float64, device=dev))
        X, XN, A, R = (torch.empty(B, 30, device=dev), torch.empty(B, 30, device=dev), torch.empty(B, 4, device=dev), torch.empty(B, device=dev))

        def draw():
            _lib.call("bmi_her_draw", ctypes.c_uint64(125 + rank), _lib.ptr(ctr), B, _lib.ptr(nv), T, _lib.ptr(d[0]), _lib.ptr(d[1]),
                      _lib.ptr(d[2]), _lib.ptr(d[3]), _lib.stream_ptr())

        def fused():
            _lib.call("bmi_her_sample_inputs", ctypes.byref(eps), E, _lib.ptr(d[0]), _lib.ptr(d[1]), _lib.ptr(d[2]), _lib.ptr(d[3]), B,
                      0.8, 0.05, 200.0, 5.0, _lib.ptr(stats[0]), _lib.ptr(stats[1]), _lib.ptr(stats[2]), _lib.ptr(stats[3]),
                      _lib.ptr(X), _lib.ptr(XN), _lib.ptr(A), _lib.ptr(R), _lib.stream_ptr())
        for _ in range(max(args.warmup, 3)):
            draw(); fused()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = []
        for _ in range(args.steps):
            draw()
            flush.fill_(0.0)               # L2 flush between timed launches ...
            flush_sink.copy_(flush.sum())  # ... then a 256 MiB read pass: the lines the timed kernel evicts are clean, so the
                                           # flush's own write-backs are not charged to it
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fused(); e.record()
            torch.cuda.synchronize()
            ms.append(s.elapsed_time(e))
        t = torch.tensor([float(np.mean(ms)) * 1e-3], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t = float(t.item())
        rows.append({"batch": B, "us": t * 1e6, "transitions_per_s": world * B / t, "GB_s_per_gpu": B * 516 / t / 1e9,
                     "frac_of_hbm_peak": B * 516 / t / 1e9 / peak})
    # the protocol's own floor: the same flush + event pair around a one-element fill kernel
    tiny = torch.zeros(1, device=dev)
    fl = []
    for _ in range(max(args.steps, 10)):
        flush.fill_(0.0)
        flush_sink.copy_(flush.sum())
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); tiny.fill_(1.0); e.record()
        torch.cuda.synchronize()
        fl.append(s.elapsed_time(e))
    floor_us = float(np.median(fl)) * 1e3
    sampler.stop_flag = True
    if rank == 0:
        sampler.join(timeout=2)
        top = rows[-1]
        line = {"metric": "HER-relabelled transitions/s (replay-buffer stress: 5e5 stored transitions per GPU, future k=4, fused network-input assembly)",
                "value": top["transitions_per_s"], "unit": "transitions/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": top["us"] * 1e-3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": "replay-stress: %d episodes x 100 steps" % E + " of N(0,1) obs / N(0.3,0.1) goals / U(-0.5,0.5) actions per GPU, "
                                       "her_inputs_lane_kernel batch sweep, value = batch %d" % top["batch"],
                           "l2": "flushed before every timed launch (256 MiB fill, then a 256 MiB read so that the evicted lines are clean); the 75 MB float32 buffer itself fits the 126 MB L2",
                           "parallelism": "dp%d (one buffer per rank, no data-path collective)" % world},
                "roofline": {"bound": "hbm", "achieved": top["GB_s_per_gpu"], "peak": peak, "unit": "GB/s", "frac": top["frac_of_hbm_peak"],
                             "traffic": None, "kernel": "her_inputs_lane_kernel", "algorithmic_bytes_per_launch": 516 * top["batch"],
                             "peak_source": peak_src, "sweep": rows, "launch_floor_us": floor_us,
                             "launch_floor_note": "same flush + CUDA-event pair around a one-element fill kernel: what a launch costs in this protocol before any sample is gathered"},
                "cpu_baseline": None,
                "e2e": None, "gpu_launches": int(_lib.launch_count() - n0), "clocks": sampler.summary()}
        print(json.dumps(line), flush=True)
    utils.shutdown_comm()


def run_gpu(args):
    import torch
    from rl_arm_under_sparse_reward_b200 import _lib, utils
    from rl_arm_under_sparse_reward_b200.arguments import Args
    from rl_arm_under_sparse_reward_b200.bmirobot_env.vec_env import BmiVecEnv
    from rl_arm_under_sparse_reward_b200.ddpg_agent import ddpg_agent
    from rl_arm_under_sparse_reward_b200.train import get_env_params
    import torch.distributed as dist

    task = "pick" if args.workload == "pick" else "push"
    rank, world = utils.init_comm()
    if world == 1:
        torch.cuda.set_device(0)
    dev = torch.device("cuda", torch.cuda.current_device())
    a = Args()
    a.verbose = False
    a.n_envs = args.envs
    a.n_batches = args.n_batches
    a.fused_rollout = not args.stepwise
    a.buffer_size = args.buffer_episodes * T
    a.save_dir = "/tmp/bmi_bench_%d/" % rank
    a.add_demo = not args.no_demo
    n_demo, demo_rate = 0, None
    if a.add_demo:
        a.demo_name, n_demo, demo_rate = _demo_file(task, rank, args.demos, torch.cuda.current_device())
        a.add_demo = n_demo > 0
    env = BmiVecEnv(a.n_envs, task=task, seed=a.seed + rank)
    np.random.seed(a.seed + rank)
    torch.manual_seed(a.seed + rank)
    agent = ddpg_agent(a, env, get_env_params(env))
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    roll_ev = []

    def cycle(timed=False):
        if timed:   # CUDA events around the rollout launch of THIS cycle (on the launching stream)
            rs, re_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            rs.record()
        agent.rollout(0)
        if timed:
            re_.record()
            roll_ev.append((rs, re_))
        agent.buffer.store_episode([agent.ep['obs'], agent.ep['ag'], agent.ep['g'], agent.ep['actions']])
        agent._update_normalizer()
        agent.update_many(a.n_batches)
        agent._soft_update_target_network()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):   # first call captures the graphs, later ones replay them
        cycle()
    barrier()
    env.contact_drops(reset=True)
    sampler = ClockSampler(torch.cuda.current_device())
    if rank == 0:          # one nvidia-smi poller per job is enough (rank 0 prints the line)
        sampler.start()
    # launches per cycle: graph replays do not pass through the host launch counter, so count one eager cycle
    a.use_cuda_graphs = False
    n0 = _lib.launch_count()
    cycle()
    launches_per_cycle = _lib.launch_count() - n0
    a.use_cuda_graphs = True
    barrier()
    env.contact_drops(reset=True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for s, e in ev:
        flush.fill_(1.0)               # L2 flush between timed iterations (outside the timed event pair)
        s.record()
        cycle(timed=True)
        e.record()
    barrier()
    drops = env.contact_drops()
    ms = np.array([s.elapsed_time(e) for s, e in ev])
    tot = torch.tensor([float(ms.sum())], device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    total_ms = float(tot.item())
    env_steps_per_cycle = a.n_envs * T
    value = world * env_steps_per_cycle * args.steps / (total_ms * 1e-3)

    # ---- dominant kernel: CUDA events recorded around the rollout launch inside each of the K timed cycles -------------
    # (the fused rollout kernel is ONE launch per batch of episodes: T env-steps for each of the n_envs envs; the
    # events also bracket the four tiny weight-transpose launches, < 0.1 % of the interval)
    kern_ms = float(np.mean([s.elapsed_time(e) for s, e in roll_ev]))
    kernel_name = "rollout_kernel" if getattr(a, "fused_rollout", True) else "env_step_kernel x %d" % T
    peak, peak_src = peaks()
    achieved = ALGO_BYTES_PER_ENV_STEP * a.n_envs * T / (kern_ms * 1e-3) / 1e9
    kernel_share = kern_ms / (total_ms / args.steps)

    # ---- e2e: the same cycle with the HOST in the loop -------------------------------------------------------------------
    # Every env-step: observation D2H into pinned memory -> sync -> the host hands the observation to the agent's policy
    # (its arrays are host arrays, as in the reference loop ddpg_agent.py:111-120: H2D, normalise + actor + exploration
    # noise on the device, action D2H -> sync) -> the host hands the action to env.step (H2D) -> step.  Two host syncs per
    # env-step, policy inference inside the timed region.  The episode record accumulated on the host is then uploaded
    # into store_episode / _update_normalizer, the updates run, the loss is read back.
    e2e = e2e_async = None
    if not args.no_e2e:
        obs_host = torch.empty((T + 1, a.n_envs, 27), dtype=torch.float32).pin_memory()
        ag_host = torch.empty((T + 1, a.n_envs, 3), dtype=torch.float32).pin_memory()
        act_host = torch.empty((T, a.n_envs, 4), dtype=torch.float32).pin_memory()
        rs_host = torch.empty((2, a.n_envs), dtype=torch.float32).pin_memory()
        loss_host = torch.empty(2, dtype=torch.float32).pin_memory()
        obs_dev = torch.empty((a.n_envs, 27), dtype=torch.float32, device=dev)
        act_dev = torch.empty((a.n_envs, 4), dtype=torch.float32, device=dev)
        h2d = d2h = 0

        def e2e_cycle(host_in_loop):
            nonlocal h2d, d2h
            h2d = d2h = 0
            obs, ag, g = env.reset()
            obs_host[0].copy_(obs, non_blocking=True)
            ag_host[0].copy_(ag, non_blocking=True)
            g_host = g.cpu()
            d2h += obs.numel() * 4 + ag.numel() * 4 + g.numel() * 4
            for t in range(T):
                if host_in_loop:
                    torch.cuda.synchronize()                                   # the host now owns obs_host[t]
                    obs_dev.copy_(obs_host[t], non_blocking=True)              # host observation -> policy
                    act = agent._policy(obs_dev, g, True, 0.0)
                    act_host[t].copy_(act, non_blocking=True)
                    torch.cuda.synchronize()                                   # the host now owns the action
                    h2d += obs_dev.numel() * 4
                    d2h += act.numel() * 4
                act_dev.copy_(act_host[t], non_blocking=True)                  # host action -> env.step
                obs, ag, r, s = env.step(act_dev)
                obs_host[t + 1].copy_(obs, non_blocking=True)
                ag_host[t + 1].copy_(ag, non_blocking=True)
                rs_host[0].copy_(r, non_blocking=True)
                rs_host[1].copy_(s, non_blocking=True)
                h2d += act_dev.numel() * 4
                d2h += (obs.numel() + ag.numel() + r.numel() + s.numel()) * 4
            # The host now holds the cycle's time-major record.  Episode batch for store_episode / _update_normalizer: ONE
            # pinned H2D copy per array, transposed to the reference's (R, T+1, dim) layout on the device.
            torch.cuda.synchronize()
            mb_obs = obs_host.to(dev, non_blocking=True).permute(1, 0, 2).contiguous()
            mb_ag = ag_host.to(dev, non_blocking=True).permute(1, 0, 2).contiguous()
            mb_g = g_host.to(dev)[:, None, :].expand(a.n_envs, T, 3).contiguous()
            mb_act = act_host.to(dev, non_blocking=True).permute(1, 0, 2).contiguous()
            h2d += (obs_host.numel() + ag_host.numel() + g_host.numel() + act_host.numel()) * 4
            agent.buffer.store_episode([mb_obs, mb_ag, mb_g, mb_act])
            agent._update_normalizer([mb_obs, mb_ag, mb_g, mb_act])
            agent.update_many(a.n_batches)
            agent._soft_update_target_network()
            loss_host.copy_(agent._losses, non_blocking=True)
            d2h += 8
            torch.cuda.synchronize()

        def time_e2e(host_in_loop):
            e2e_cycle(host_in_loop)
            barrier()
            t0 = time.perf_counter()
            n_e2e = max(1, min(args.steps, 3))
            for _ in range(n_e2e):
                e2e_cycle(host_in_loop)
            barrier()
            el = torch.tensor([time.perf_counter() - t0], device=dev)
            if world > 1:
                dist.all_reduce(el, op=dist.ReduceOp.MAX)
            return world * env_steps_per_cycle * n_e2e / float(el.item()), n_e2e

        v, n_e2e = time_e2e(True)
        e2e = {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": n_e2e,
               "note": "host in the loop: per env-step obs D2H (pinned) -> sync -> agent._policy on the host's observation (H2D, "
                       "normalise + actor + exploration noise, action D2H) -> sync -> env.step on the host's action (H2D); the "
                       "host-side episode record uploaded into store_episode / _update_normalizer; loss read back"}
        v2, _ = time_e2e(False)
        e2e_async = {"value": v2, "unit": "env-steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                     "note": "round-1 figure: host (pinned) buffers every env-step but pre-recorded actions and no per-step sync"}
    sampler.stop_flag = True
    if rank == 0:
        sampler.join(timeout=2)

    p2p_flag = bool(agent.p2p_timed_out()) if (world > 1 and agent._p2p) else False
    digests = [None] * world
    if world > 1:
        dist.all_gather_object(digests, _digest(agent))
    grad_sync = ("none (1 rank)" if world == 1 else ("fused peer-memory sum + Adam kernel (CUDA IPC over NVLink)"
                 if agent._p2p else "NCCL allreduce (sum) + Adam"))
    agent.release_graphs()
    probe = parity_probe(world, rank, a.seed) if (world > 1 and not args.no_probe) else None

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, total, wall = cpu_env_steps_per_s(args.cpu_steps_per_core, cores)
        cpu = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
               "sample": "%d env-steps of the C oracle port (restated env step in its faithful mode, NOT PyBullet) on %d processes + "
                         "proportional torch-CPU updates, %.1f s" % (total, cores, wall)}
    if rank == 0:
        metric = METRIC if task == "push" else "env-steps/s (pick-and-place, 4096 envs, add_demo) on B200"
        multi = None
        if world > 1:
            multi = {"params_identical_across_ranks": len(set(digests)) == 1, "probe": probe}
        line = {"metric": metric, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": "%s task, %d vectorised envs per GPU, one cycle = %d env-steps + store + normaliser + %d HER "
                                       "(future k=4) DDPG updates of batch %d + Polyak" % (task, a.n_envs, env_steps_per_cycle, a.n_batches, a.batch_size),
                           "envs_per_gpu": a.n_envs, "updates_per_env_step": a.n_batches / env_steps_per_cycle,
                           "buffer_episodes": args.buffer_episodes, "l2": "flushed between timed iterations (256 MiB fill)",
                           "add_demo": bool(a.add_demo), "demo_episodes": n_demo, "demo_kept_fraction": demo_rate,
                           "physics": "arm self-collision on (baked pair tables), IK joint damping 0.5, solver schedule %d iterations "
                                      "(2-fold compressed ramp of Bullet's 150)" % 80,
                           "contacts_dropped_per_env_substep": drops / float(args.steps * a.n_envs * T * 20),
                           "parallelism": "dp%d" % world, "grad_sync": grad_sync, "p2p_timed_out": p2p_flag},
                "reference_arm": "bench.py --impl reference = C port of the reference env step (oracle/), NOT PyBullet + mpi4py",
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": NCU_DRAM_BYTES_PER_LAUNCH,
                             "kernel": kernel_name, "algorithmic_bytes_per_launch": ALGO_BYTES_PER_ENV_STEP * a.n_envs * T,
                             "kernel_ms": kern_ms, "kernel_share_of_step": kernel_share, "peak_source": peak_src,
                             "warps_active_pct": NCU_ROLLOUT["warps_active_pct"], "issue_slot_util_pct": NCU_ROLLOUT["issue_slot_util_pct"],
                             "fma_pipe_pct": NCU_ROLLOUT["fma_pipe_pct"], "lsu_pipe_pct": NCU_ROLLOUT["lsu_pipe_pct"],
                             "warp_instructions_per_env_substep": NCU_ROLLOUT["warp_instructions_per_env_substep"],
                             "note": "FP32-issue bound kernel (SURVEY 8d): the HBM fraction is small by construction; what bounds it is the issue-slot "
                                     "utilisation above (ncu capture in profiles/)"},
                "cpu_baseline": cpu, "e2e": e2e, "e2e_async": e2e_async, "multi_rank": multi,
                "gpu_launches": int(launches_per_cycle * args.steps), "clocks": sampler.summary()}
        print(json.dumps(line), flush=True)
    utils.shutdown_comm()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="push", choices=["push", "pick", "replay-stress"])
    ap.add_argument("--envs", type=int, default=N_ENVS)
    ap.add_argument("--n-batches", type=int, default=40, help="DDPG updates per cycle (reference: 40)")
    ap.add_argument("--buffer-episodes", type=int, default=65536)
    ap.add_argument("--demos", type=int, default=1000, help="scripted demonstration episodes pre-loaded into the buffer (add_demo)")
    ap.add_argument("--no-demo", action="store_true")
    ap.add_argument("--cpu-steps-per-core", type=int, default=2500)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-probe", action="store_true", help="skip the multi-rank parity probe")
    ap.add_argument("--stress-episodes", type=int, default=5000, help="replay-stress: stored episodes per GPU (5000 = 5e5 transitions)")
    ap.add_argument("--stepwise", action="store_true", help="step-wise rollout pipeline instead of the fused kernel")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "replay-stress":
        run_replay_stress(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
