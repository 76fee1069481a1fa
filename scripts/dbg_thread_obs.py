import sys, numpy as np, torch
sys.path.insert(0, ".")
from kiss_mpc_b200 import BatchedMotionPlanner, PlannerConfig
from kiss_mpc_b200.synthetic import make_batch
for O, base in ((10, None), (20, 500.0), (200, 500.0)):
    B = 6
    if base is None:
        b = make_batch(B, seed=1004, O=O); obs = b["obs"]
    else:
        b = make_batch(B, seed=3); obs = np.tile(np.array([[[base, base]]]), (B, O, 1)) + np.arange(O)[None, :, None]
    pl = BatchedMotionPlanner(PlannerConfig(N=30, O_max=O), max_batch=B)
    r = pl.solve(torch.tensor(b["x_cur"], device="cuda"), torch.tensor(b["goal"], device="cuda"), obstacles=torch.tensor(obs, device="cuda"),
                 obstacle_radius=0.3, inflation_radius=0.5)
    torch.cuda.synchronize()
    print("O", O, "status", r.status.tolist(), "iters", r.iters.tolist(), "stats", pl.stats()["blocks"])
