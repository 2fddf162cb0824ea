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
# `ncu --set full` capture summarised in profiles/r01_rollout_kernel_ncu.md (bench.py cannot run ncu itself)
NCU_DRAM_BYTES_PER_LAUNCH = 80482304   # 4.774 MB read + 75.708 MB written
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
    per_core = 2000
    for _ in range(args.warmup):
        cpu_env_steps_per_s(50, cores)
    vals, t_ms = [], []
    for _ in range(args.steps):
        v, total, wall = cpu_env_steps_per_s(per_core, cores)
        vals.append(v)
        t_ms.append(wall * 1e3)
    value = float(np.mean(vals))
    sample = ("per step: %d processes x %d env-steps of the C oracle port (restated env step, NOT PyBullet: pybullet/gym/"
              "mpi4py are not installable in this image) + the proportional DDPG updates on torch CPU" % (cores, per_core))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(np.mean(t_ms)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "push task, one env per host core, cycle shape of the GPU arm (100-step episodes, 40 HER DDPG updates "
                                   "of batch 256 per 409600 env-steps); bounded sample per step", "envs": cores,
                       "updates_per_env_step": 40.0 / (N_ENVS * T)},
            "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    from rl_arm_under_sparse_reward_b200 import _lib, utils
    from rl_arm_under_sparse_reward_b200.arguments import Args
    from rl_arm_under_sparse_reward_b200.bmirobot_env.vec_env import BmiVecEnv
    from rl_arm_under_sparse_reward_b200.ddpg_agent import ddpg_agent
    from rl_arm_under_sparse_reward_b200.train import get_env_params
    import torch.distributed as dist

    rank, world = utils.init_comm()
    if world == 1:
        torch.cuda.set_device(0)
    dev = torch.device("cuda", torch.cuda.current_device())
    a = Args()
    a.add_demo, a.verbose = False, False
    a.n_envs = args.envs
    a.fused_rollout = not args.stepwise
    a.buffer_size = args.buffer_episodes * T
    a.save_dir = "/tmp/bmi_bench_%d/" % rank
    env = BmiVecEnv(a.n_envs, task="push", seed=a.seed + rank)
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

    for _ in range(max(args.warmup, 2)):   # first call captures the graphs, later ones replay them
        cycle()
    barrier()
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
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for s, e in ev:
        flush.fill_(1.0)               # L2 flush between timed iterations (outside the timed event pair)
        s.record()
        cycle(timed=True)
        e.record()
    barrier()
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

    # ---- e2e: the same cycle driven through the reference-facing calls with HOST (pinned) buffers -----------
    e2e = None
    if not args.no_e2e:
        act_host = agent.ep['actions'].permute(1, 0, 2).contiguous().cpu().pin_memory()        # [T][n][4]
        obs_host = torch.empty((T + 1, a.n_envs, 27), dtype=torch.float32).pin_memory()
        ag_host = torch.empty((T + 1, a.n_envs, 3), dtype=torch.float32).pin_memory()
        rs_host = torch.empty((2, a.n_envs), dtype=torch.float32).pin_memory()
        loss_host = torch.empty(2, dtype=torch.float32).pin_memory()
        act_dev = torch.empty((a.n_envs, 4), dtype=torch.float32, device=dev)
        h2d = d2h = 0

        def e2e_cycle():
            nonlocal h2d, d2h
            h2d = d2h = 0
            obs, ag, g = env.reset()
            obs_host[0].copy_(obs, non_blocking=True)
            ag_host[0].copy_(ag, non_blocking=True)
            g_host = g.cpu()
            d2h += obs.numel() * 4 + ag.numel() * 4 + g.numel() * 4
            for t in range(T):
                act_dev.copy_(act_host[t], non_blocking=True)                  # host policy output -> device
                obs, ag, r, s = env.step(act_dev)
                obs_host[t + 1].copy_(obs, non_blocking=True)
                ag_host[t + 1].copy_(ag, non_blocking=True)
                rs_host[0].copy_(r, non_blocking=True)
                rs_host[1].copy_(s, non_blocking=True)
                h2d += act_dev.numel() * 4
                d2h += (obs.numel() + ag.numel() + r.numel() + s.numel()) * 4
            # The host now holds the cycle's time-major record (what a host-driven loop accumulates step by step).
            # Episode batch for store_episode / _update_normalizer: ONE pinned H2D copy per array, transposed to the
            # reference's (R, T+1, dim) layout on the device (a 60 MB strided transpose on the host costs more than
            # the whole update phase).
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

        e2e_cycle()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 3))
        for _ in range(n_e2e):
            e2e_cycle()
        barrier()
        el = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
        e2e = {"value": world * env_steps_per_cycle * n_e2e / float(el.item()), "unit": "env-steps/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": n_e2e,
               "note": "host (pinned) action/obs buffers every env-step through env.step, the host-side episode record uploaded (pinned H2D) into store_episode / _update_normalizer, loss read back"}
    sampler.stop_flag = True
    if rank == 0:
        sampler.join(timeout=2)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, total, wall = cpu_env_steps_per_s(args.cpu_steps_per_core, cores)
        cpu = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
               "sample": "%d env-steps of the C oracle port (restated env step, NOT PyBullet) on %d processes + proportional "
                         "torch-CPU updates, %.1f s" % (total, cores, wall)}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": "push task, %d vectorised envs per GPU, one cycle = %d env-steps + store + normaliser + %d HER "
                                       "(future k=4) DDPG updates of batch %d + Polyak" % (a.n_envs, env_steps_per_cycle, a.n_batches, a.batch_size),
                           "envs_per_gpu": a.n_envs, "updates_per_env_step": a.n_batches / env_steps_per_cycle,
                           "buffer_episodes": args.buffer_episodes, "l2": "flushed between timed iterations (256 MiB fill)",
                           "parallelism": "dp%d" % world,
                           "grad_sync": ("none (1 rank)" if world == 1 else ("fused peer-memory sum + Adam kernel (CUDA IPC over NVLink)"
                                         if agent._p2p else "NCCL allreduce (sum) + Adam")),
                           "p2p_timed_out": bool(agent.p2p_timed_out()) if (world > 1 and agent._p2p) else False},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": NCU_DRAM_BYTES_PER_LAUNCH,
                             "kernel": kernel_name, "algorithmic_bytes_per_launch": ALGO_BYTES_PER_ENV_STEP * a.n_envs * T,
                             "kernel_ms": kern_ms, "kernel_share_of_step": kernel_share, "peak_source": peak_src,
                             "note": "FP32-issue/latency-bound kernel, the HBM fraction is small by construction (SURVEY 8d); ncu (profiles/r01_rollout_kernel_ncu.md): 22.9 k warp instructions per env sub-step, issue slots busy 63-68 % of active cycles, the launch ends with its slowest env (median env: half the launch)"},
                "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches_per_cycle * args.steps),
                "clocks": sampler.summary()}
        print(json.dumps(line), flush=True)
    agent.release_graphs()
    utils.shutdown_comm()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--envs", type=int, default=N_ENVS)
    ap.add_argument("--buffer-episodes", type=int, default=65536)
    ap.add_argument("--cpu-steps-per-core", type=int, default=12000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--stepwise", action="store_true", help="step-wise rollout pipeline instead of the fused kernel")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
