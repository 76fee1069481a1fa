"""Strong-scaling model on ONE GPU: the 65,536-instance headline batch cut into the contiguous slices `shard_range` gives 1, 2, 4
and 8 ranks; every slice is solved alone and the slowest slice of a split is what an N-GPU run waits for (no gather here).
Also: the longest instance of the batch solved alone (B = 1) -- the floor no split can go below.
usage: [KMPC_LIB=variant.so] python scripts/strong_slices.py [N]"""
import sys, json, numpy as np, torch
sys.path.insert(0, ".")
from kiss_mpc_b200 import BatchedMotionPlanner, PlannerConfig, shard_range
from kiss_mpc_b200.synthetic import make_batch
N = int(sys.argv[1]) if len(sys.argv) > 1 else 30
B = 65536
b = make_batch(B, seed=1000)
pl = BatchedMotionPlanner(PlannerConfig(N=N), max_batch=B)
pl.set_timing(True)
X = torch.tensor(b["x_cur"], device="cuda"); G = torch.tensor(b["goal"], device="cuda")
def t_of(lo, hi, reps=3):
    best = 1e9
    for _ in range(reps):
        r = pl.solve(X[lo:hi].contiguous(), G[lo:hi].contiguous()); torch.cuda.synchronize()
        best = min(best, pl.stats()["last_kernel_ms"])
    return best, r
out = {}
_, r = t_of(0, B, 1)
it = r.iters.cpu().numpy()
worst = int(np.argmax(it))
for world in (1, 2, 4, 8):
    ts = [t_of(*shard_range(B, k, world))[0] for k in range(world)]
    out[f"x{world}"] = {"max_ms": max(ts), "min_ms": min(ts), "slices_ms": [round(t, 3) for t in ts]}
tw, rw = t_of(worst, worst + 1, 5)
out["longest_instance"] = {"index": worst, "iterations": int(it[worst]), "trips": int(pl.stats()["trips"]), "alone_ms": tw}
print(json.dumps(out))
