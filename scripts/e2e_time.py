import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from kiss_mpc_b200 import BatchedMotionPlanner, PlannerConfig
from kiss_mpc_b200.synthetic import make_batch
B = 65536
b = make_batch(B, seed=1000)
pl = BatchedMotionPlanner(PlannerConfig(N=30), max_batch=B)
xh, gh = b["x_cur"], b["goal"]
for copy in (False, True):
    pl.solve(xh, gh, copy=copy)
    t0 = time.perf_counter()
    for _ in range(5):
        r = pl.solve(xh, gh, copy=copy); chk = float(r.objective[0])
    dt = (time.perf_counter() - t0) / 5
    print("copy", copy, "e2e ms", dt * 1e3, "solves/s", B / dt, "conv", float((r.status == 0).mean()))
