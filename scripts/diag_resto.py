"""Diagnostic: which instances of an obstacle batch leave the warp solver for the restoration finisher, and what they cost.
usage: python scripts/diag_resto.py B O tracks(0|1) [seed]"""
import sys, json, numpy as np, torch
sys.path.insert(0, ".")
from kiss_mpc_b200 import BatchedMotionPlanner, PlannerConfig
from kiss_mpc_b200.synthetic import make_batch, make_tracks
B = int(sys.argv[1]); O = int(sys.argv[2]); tr = int(sys.argv[3]); seed = int(sys.argv[4]) if len(sys.argv) > 4 else 1004
N = 30
b = make_batch(B, seed=seed, O=O)
if tr: b["obs"] = make_tracks(b["obs"], N, seed=seed)
pl = BatchedMotionPlanner(PlannerConfig(N=N, O_max=O), max_batch=B)
x = torch.tensor(b["x_cur"], device="cuda"); g = torch.tensor(b["goal"], device="cuda"); ob = torch.tensor(b["obs"], device="cuda")
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = pl.solve(x, g, obstacles=ob, obstacle_radius=0.3, inflation_radius=0.5); e1.record(); torch.cuda.synchronize()
st = r.status.cpu().numpy(); it = r.iters.cpu().numpy()
bad = np.nonzero(st != 0)[0]
print(json.dumps({"B": B, "O": O, "tracks": tr, "ms": e0.elapsed_time(e1), "status_hist": {int(k): int(v) for k, v in zip(*np.unique(st, return_counts=True))},
                  "iters_of_nonconverged": sorted(it[bad].tolist())[-20:], "iters_top": sorted(it.tolist())[-10:]}))
