"""B = 1 kernel latency (events, median over instances) for the current library (KMPC_LIB may point at a variant build)."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from kiss_mpc_b200 import BatchedMotionPlanner, PlannerConfig
from kiss_mpc_b200.synthetic import make_batch
N = int(sys.argv[1]) if len(sys.argv) > 1 else 30
b = make_batch(64, seed=1000)
pl = BatchedMotionPlanner(PlannerConfig(N=N), max_batch=64)
pl.set_timing(True)
lat = []
for i in range(48):
    x = torch.tensor(b["x_cur"][i:i + 1], device="cuda"); g = torch.tensor(b["goal"][i:i + 1], device="cuda")
    best = 1e9
    for _ in range(3):
        pl.solve(x, g); torch.cuda.synchronize(); best = min(best, pl.stats()["last_kernel_ms"])
    lat.append((best * 1e3, pl.stats()["trips"]))
print("N", N, "B1 kernel us p50", np.median([l[0] for l in lat]), "us/trip p50", np.median([l[0] / l[1] for l in lat]))
