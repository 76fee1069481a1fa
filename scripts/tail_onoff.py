"""Tail mode on / off (KMPC_NO_TAIL=1) for small batches: kernel time of B in {1 (median over 48 instances), 64, 592, 2368, 4096, 8192}."""
import sys, json, numpy as np, torch
sys.path.insert(0, ".")
from kiss_mpc_b200 import BatchedMotionPlanner, PlannerConfig
from kiss_mpc_b200.synthetic import make_batch
N = int(sys.argv[1]) if len(sys.argv) > 1 else 30
b = make_batch(65536, seed=1000)
pl = BatchedMotionPlanner(PlannerConfig(N=N), max_batch=8192)
pl.set_timing(True)
X = torch.tensor(b["x_cur"], device="cuda"); G = torch.tensor(b["goal"], device="cuda")
def t_of(lo, hi, reps=3):
    best = 1e9
    for _ in range(reps):
        pl.solve(X[lo:hi].contiguous(), G[lo:hi].contiguous()); torch.cuda.synchronize(); best = min(best, pl.stats()["last_kernel_ms"])
    return best
out = {"B1_p50_us": float(np.median([t_of(i, i + 1) for i in range(48)]) * 1e3)}
for B in (64, 592, 2368, 4096, 8192):
    out[f"B{B}_ms"] = t_of(24576, 24576 + B)
print(json.dumps(out))
