import sys, numpy as np, torch
sys.path.insert(0, ".")
from kiss_mpc_b200 import BatchedMotionPlanner, PlannerConfig
from kiss_mpc_b200.synthetic import make_batch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
O = int(sys.argv[2]) if len(sys.argv) > 2 else 10
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
b = make_batch(B, seed=1004, O=O)
pl = BatchedMotionPlanner(PlannerConfig(N=30, O_max=O), max_batch=B)
pl.set_timing(True)
x = torch.tensor(b["x_cur"], device="cuda"); g = torch.tensor(b["goal"], device="cuda"); ob = torch.tensor(b["obs"], device="cuda")
for _ in range(reps):
    r = pl.solve(x, g, obstacles=ob, obstacle_radius=0.3, inflation_radius=0.5); torch.cuda.synchronize()
    s = pl.stats()
    print(B, "O", O, "ms", s["last_kernel_ms"], "solves/s", B / s["last_kernel_ms"] * 1e3, "trips", s["trips"] / B, "iters", r.iters.float().mean().item(),
          "max iters", r.iters.max().item(), "conv", (r.status == 0).float().mean().item(), "host trips", s["blocks"])
