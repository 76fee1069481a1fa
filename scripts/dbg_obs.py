import sys, numpy as np, torch
sys.path.insert(0, ".")
from kiss_mpc_b200 import BatchedMotionPlanner, PlannerConfig
from kiss_mpc_b200.synthetic import make_batch
from oracle import oracle as O
b = make_batch(512, seed=1004, O=10)
cfg = O.OracleConfig(O=10, linsolve="riccati")
ro = O.solve(cfg, b["x_cur"], b["goal"], obs=b["obs"])
pl = BatchedMotionPlanner(PlannerConfig(O_max=10), max_batch=512)
d = lambda a: torch.tensor(a, device="cuda")
for rep in range(2):
    r = pl.solve(d(b["x_cur"]), d(b["goal"]), obstacles=d(b["obs"]), obstacle_radius=0.3, inflation_radius=0.5)
    st = r.status.cpu().numpy(); it = r.iters.cpu().numpy()
    bad = np.where(st != ro.status)[0]
    print("rep", rep, "bad", bad, st[bad], it[bad], ro.iters[bad], "iters differ", (it != ro.iters).sum(), "maxdU", np.abs(r.controls.cpu().numpy() - ro.U)[st == 0].max())
    print("obj", r.objective.cpu().numpy()[bad], ro.obj[bad])
