"""Tuning aid: device time of the headline batch with its long instances (more than 60 IPM iterations, ~0.1 % of the batch)
replaced by copies of instance 0 -- what the kernel sustains per trip when no single instance sets the end of the launch.
    python scripts/throughput_probe.py [B] [N] [reps]"""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from kiss_mpc_b200 import BatchedMotionPlanner, PlannerConfig
from kiss_mpc_b200.synthetic import make_batch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
b = make_batch(B, seed=1000)
pl = BatchedMotionPlanner(PlannerConfig(N=N), max_batch=B)
pl.set_timing(True)
x = torch.tensor(b["x_cur"], device="cuda"); g = torch.tensor(b["goal"], device="cuda")
r = pl.solve(x, g); torch.cuda.synchronize()
full = []
for _ in range(reps):
    pl.solve(x, g); torch.cuda.synchronize(); full.append(pl.stats()["last_kernel_ms"])
long_ = r.iters > 60
x[long_] = x[0]; g[long_] = g[0]
ms = []
for _ in range(reps):
    pl.solve(x, g); torch.cuda.synchronize()
    s = pl.stats(); ms.append(s["last_kernel_ms"])
print(f"B {B} N {N} full {min(full):.3f} ms | without {int(long_.sum())} long instances {min(ms):.3f} ms, trips/solve {s['trips'] / B:.3f}, us per slot-trip {min(ms) * 1e3 / (s['trips'] / (148 * (16 if N <= 31 else 12))):.3f}")
