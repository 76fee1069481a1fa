import sys, numpy as np, torch
sys.path.insert(0, ".")
from kiss_mpc_b200 import BatchedMotionPlanner, PlannerConfig
from kiss_mpc_b200.synthetic import make_batch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
b = make_batch(B, seed=1000)
pl = BatchedMotionPlanner(PlannerConfig(N=N), max_batch=B)
pl.set_timing(True)
x = torch.tensor(b["x_cur"], device="cuda"); g = torch.tensor(b["goal"], device="cuda")
for _ in range(reps):
    r = pl.solve(x, g); torch.cuda.synchronize()
    s = pl.stats()
    print(B, N, "ms", s["last_kernel_ms"], "solves/s", B / s["last_kernel_ms"] * 1e3, "trips", s["trips"] / B, "conv", (r.status == 0).float().mean().item())
