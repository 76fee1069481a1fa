import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from kiss_mpc_b200 import BatchedMotionPlanner, PlannerConfig
from kiss_mpc_b200.synthetic import make_batch
for B, N in ((4096, 30), (65536, 30), (65536, 50)):
    b = make_batch(B, seed=1000)
    pl = BatchedMotionPlanner(PlannerConfig(N=N), max_batch=B)
    pl.set_timing(True)
    x = torch.tensor(b["x_cur"], device="cuda"); g = torch.tensor(b["goal"], device="cuda")
    for i in range(3):
        r = pl.solve(x, g); torch.cuda.synchronize()
        s = pl.stats()
        print(B, N, "ms", s["last_kernel_ms"], "solves/s", B / s["last_kernel_ms"] * 1e3, "trips", s["trips"] / B, "iters", r.iters.float().mean().item(), "conv", (r.status == 0).float().mean().item(), s["slots"], s["blocks"], flush=True)
    if B == 4096: print("fp64 peak TF", pl.measure_fp64_peak())
    pl.close()
