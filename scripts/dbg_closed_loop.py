import sys, numpy as np, torch
sys.path.insert(0, ".")
from kiss_mpc_b200 import BatchedMotionPlanner, PlannerConfig
from kiss_mpc_b200.synthetic import make_batch
B, steps = 16384, 200
b = make_batch(B, seed=1005)
pl = BatchedMotionPlanner(PlannerConfig(N=30, T=0.1), max_batch=B)
g = torch.tensor(b["goal"], device="cuda"); x = torch.tensor(b["x_cur"], device="cuda")
X, U, applied, iters, status = pl.closed_loop(x, g, steps)
torch.cuda.synchronize()
st = status.cpu().numpy(); it = iters.cpu().numpy()
vals, cnt = np.unique(st, return_counts=True)
print("status counts", dict(zip(vals.tolist(), cnt.tolist())))
print("max iters per step", it.max(1)[::5])
print("mean iters per step", it.mean(1)[::10])
bad = np.argwhere(st != 0)
print("first bad (step, agent)", bad[:10].tolist())
for s, a in bad[:3]:
    print("step", s, "agent", a, "status", st[s, a], "iters", it[s, a], "prev iters", it[max(0, s - 3):s, a])
