"""Tuning aid: the library under KMPC_LIB against the oracle on the first B instances of the headline batch (statuses, iteration counts, controls)."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from kiss_mpc_b200 import BatchedMotionPlanner, PlannerConfig
from kiss_mpc_b200.synthetic import make_batch
from oracle import oracle as ok
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
b = make_batch(B, seed=1000)
pl = BatchedMotionPlanner(PlannerConfig(N=N), max_batch=B)
r = pl.solve(torch.tensor(b["x_cur"], device="cuda"), torch.tensor(b["goal"], device="cuda")); torch.cuda.synchronize()
ref = ok.solve(ok.OracleConfig(N=N, linsolve="riccati"), b["x_cur"], b["goal"])
st = r.status.cpu().numpy(); it = r.iters.cpu().numpy(); U = r.controls.cpu().numpy()
conv = st == 0
print("B", B, "status equal", float((st == ref.status).mean()), "iters equal", float((it == ref.iters).mean()), "max|dU|", float(np.abs(U - ref.U)[conv].max()))
