import os, sys, numpy as np, torch
sys.path.insert(0, ".")
from kiss_mpc_b200 import BatchedMotionPlanner, PlannerConfig
from kiss_mpc_b200.synthetic import make_batch
from oracle import oracle as ok
B = 64
b = make_batch(B, seed=1000)
ref = ok.solve(ok.OracleConfig(linsolve="riccati"), b["x_cur"], b["goal"])
retries = ref.diag["n_factor"] - 1 - ref.iters
pl = BatchedMotionPlanner(PlannerConfig(), max_batch=B)
x = torch.tensor(b["x_cur"], device="cuda"); g = torch.tensor(b["goal"], device="cuda")
rows = []
for i in range(B):
    os.environ["KMPC_NO_TAIL"] = "1"
    r0 = pl.solve(x[i:i+1].contiguous(), g[i:i+1].contiguous()); torch.cuda.synchronize()
    del os.environ["KMPC_NO_TAIL"]
    r1 = pl.solve(x[i:i+1].contiguous(), g[i:i+1].contiguous()); torch.cuda.synchronize()
    r2 = pl.solve(x[i:i+1].contiguous(), g[i:i+1].contiguous()); torch.cuda.synchronize()
    rows.append((i, int(retries[i]), (r0.controls - r1.controls).abs().max().item(), (r1.controls - r2.controls).abs().max().item()))
print("inst retries |notail-tail| |tail-tail|")
for r in rows: print(r)
nz = [r for r in rows if r[2] > 0]
print("differ:", len(nz), "of which retries==0:", sum(1 for r in nz if r[1] == 0), "| instances with retries==0:", sum(1 for r in rows if r[1] == 0))
