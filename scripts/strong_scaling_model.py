"""Strong-scaling prediction on ONE GPU: the 65,536-instance headline batch cut into G contiguous slices (shard_range), every
slice solved on its own and timed; max over the slices = what G GPUs would need (no communication in the solve).  Also B = 1 and
B = 4,096 timings.  Usage: python scripts/strong_scaling_model.py [reps]"""
import json, sys
import numpy as np, torch
sys.path.insert(0, ".")
from kiss_mpc_b200 import BatchedMotionPlanner, PlannerConfig, shard_range
from kiss_mpc_b200.synthetic import make_batch
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = 65536
b = make_batch(B, seed=1000)
pl = BatchedMotionPlanner(PlannerConfig(), max_batch=B)
pl.set_timing(True)
out = {}
def best(x, g):
    ms = []
    for _ in range(reps):
        r = pl.solve(x, g); torch.cuda.synchronize(); ms.append(pl.stats()["last_kernel_ms"])
    return min(ms), pl.stats()["trips"]
for G in (1, 2, 4, 8):
    per = []
    for r in range(G):
        lo, hi = shard_range(B, r, G)
        x = torch.tensor(b["x_cur"][lo:hi], device="cuda"); g = torch.tensor(b["goal"][lo:hi], device="cuda")
        per.append(best(x, g)[0])
    out[f"G{G}"] = {"max_ms": max(per), "per_slice_ms": per, "speedup": out["G1"]["max_ms"] / max(per) if G > 1 else 1.0}
b2 = make_batch(4096, seed=1002)
out["B4096_ms"] = best(torch.tensor(b2["x_cur"], device="cuda"), torch.tensor(b2["goal"], device="cuda"))[0]
lat = []
for i in range(32):
    ms, tr = best(torch.tensor(b["x_cur"][i:i + 1], device="cuda"), torch.tensor(b["goal"][i:i + 1], device="cuda"))
    lat.append((ms * 1e3, tr))
out["B1_kernel_us_p50"] = float(np.median([l[0] for l in lat])); out["B1_us_per_trip_p50"] = float(np.median([l[0] / l[1] for l in lat]))
for i in (30520, 25585):
    ms, tr = best(torch.tensor(b["x_cur"][i:i + 1], device="cuda"), torch.tensor(b["goal"][i:i + 1], device="cuda"))
    out[f"instance_{i}_alone"] = {"ms": ms, "trips": int(tr)}
print(json.dumps(out))
