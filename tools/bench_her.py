"""Replay-buffer stress (BASELINE.json config 5): 5e5 stored transitions (5000 episodes), HER 'future' relabel,
batch sweep 256 ... 65536.  Reports sampled transitions/s and achieved GB/s against the HBM roofline using the
algorithmic 516 B / transition (fused network-input kernel: 268 B read + 248 B written) for the fused
kernel and 540 B for the plain gather that also returns ag / ag_next.  Draws are device-side (Philox), inputs stay
resident; the 75 MB float32 buffer is SMALLER than the 126 MB L2, so this is an L2-resident gather unless
--episodes is raised."""
import argparse, ctypes, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_arm_under_sparse_reward_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--episodes", type=int, default=5000)
ap.add_argument("--dtype", default="float32")
ap.add_argument("--iters", type=int, default=50)
args = ap.parse_args()
dev = torch.device("cuda")
dt = torch.float32 if args.dtype == "float32" else torch.float64
E, T = args.episodes, 100
g = torch.Generator(device=dev).manual_seed(125)
obs = torch.randn(E, T + 1, 27, device=dev, generator=g).to(dt)
ag = (0.3 + 0.1 * torch.randn(E, T + 1, 3, device=dev, generator=g)).to(dt)
gg = (0.3 + 0.1 * torch.randn(E, T, 3, device=dev, generator=g)).to(dt)
act = (torch.rand(E, T, 4, device=dev, generator=g) - 0.5).to(dt)
eps = _lib.Episodes(_lib.ptr(obs), _lib.ptr(ag), _lib.ptr(gg), _lib.ptr(act), E, T, 27, 3, 4, _lib.dtype_code(dt), 0)
stats = [torch.zeros(27, device=dev), torch.ones(27, device=dev), torch.zeros(3, device=dev), torch.ones(3, device=dev)]
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
ctr = torch.zeros(1, dtype=torch.int64, device=dev)
nv = torch.tensor([E], dtype=torch.int64, device=dev)
rows = []
for B in (256, 1024, 4096, 16384, 65536, 262144, 1048576):
    d = (torch.empty(B, dtype=torch.int64, device=dev), torch.empty(B, dtype=torch.int64, device=dev),
         torch.empty(B, dtype=torch.float64, device=dev), torch.empty(B, dtype=torch.float64, device=dev))
    X, XN, A, R = (torch.empty(B, 30, device=dev), torch.empty(B, 30, device=dev), torch.empty(B, 4, device=dev), torch.empty(B, device=dev))
    def draw():
        _lib.call("bmi_her_draw", ctypes.c_uint64(125), _lib.ptr(ctr), B, _lib.ptr(nv), T, _lib.ptr(d[0]), _lib.ptr(d[1]),
                  _lib.ptr(d[2]), _lib.ptr(d[3]), _lib.stream_ptr())
    def fused():
        _lib.call("bmi_her_sample_inputs", ctypes.byref(eps), E, _lib.ptr(d[0]), _lib.ptr(d[1]), _lib.ptr(d[2]), _lib.ptr(d[3]), B,
                  0.8, 0.05, 200.0, 5.0, _lib.ptr(stats[0]), _lib.ptr(stats[1]), _lib.ptr(stats[2]), _lib.ptr(stats[3]),
                  _lib.ptr(X), _lib.ptr(XN), _lib.ptr(A), _lib.ptr(R), _lib.stream_ptr())
    for _ in range(3):
        draw(); fused()
    torch.cuda.synchronize()
    ms = []
    for _ in range(args.iters):
        draw()
        flush.fill_(0.0)      # L2 flush between timed launches
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fused(); e.record()
        torch.cuda.synchronize()
        ms.append(s.elapsed_time(e))
    t = float(np.median(ms)) * 1e-3
    bytes_per = 516 if dt == torch.float32 else 516 + 268      # float64 storage doubles the gathered bytes
    rows.append({"batch": B, "us": t * 1e6, "transitions_per_s": B / t, "GB_s": B * bytes_per / t / 1e9, "frac_of_hbm_peak": B * bytes_per / t / 1e9 / peak})
    print(json.dumps(rows[-1]))
print(json.dumps({"workload": "replay-buffer stress, %d episodes x 100 (%s), HER future k=4, fused network-input kernel, L2 flushed before each launch"
                  % (E, args.dtype), "hbm_peak_GB_s": peak, "rows": rows}))
