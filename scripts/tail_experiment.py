"""How much of the headline batch time is queue tail?  Times the 65,536 x N=30 batch in its natural order, with the instances
sorted by descending / ascending iteration count (longest-first is the best any static order can do), and prints the iteration
histogram.  Device-resident, CUDA events, best of 5."""
import json, sys
import numpy as np, torch
sys.path.insert(0, ".")
from kiss_mpc_b200 import BatchedMotionPlanner, PlannerConfig
from kiss_mpc_b200.synthetic import make_batch

B = 65536
b = make_batch(B, seed=1000)
pl = BatchedMotionPlanner(PlannerConfig(), max_batch=B)
x = torch.tensor(b["x_cur"], device="cuda"); g = torch.tensor(b["goal"], device="cuda")


def timed(xx, gg):
    best = 1e30
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); r = pl.solve(xx, gg); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, r


pl.set_queue_order(True)
tp, rp = timed(x, g)
pl.set_queue_order(False)
t0, r = timed(x, g)
assert torch.equal(r.controls, rp.controls) and torch.equal(r.iters, rp.iters)
it = r.iters
out = {"prior_order_ms": tp, "natural_ms": t0, "iters_mean": it.float().mean().item(), "iters_pcts": np.percentile(it.cpu().numpy(), [50, 90, 99, 99.9, 100]).tolist()}
for name, desc in (("longest_first", True), ("shortest_first", False)):
    o = torch.argsort(it, descending=desc, stable=True)
    out[name + "_ms"] = timed(x[o].contiguous(), g[o].contiguous())[0]
# feature heuristics for a static order: goal distance, |bearing error| of the goal as seen from the start heading
d = g[:, :2] - x[:, :2]
dist = d.norm(dim=1)
bear = torch.atan2(d[:, 1], d[:, 0]) - x[:, 2]
bear = torch.atan2(torch.sin(bear), torch.cos(bear)).abs()
for name, key in (("by_bearing", bear), ("by_dist", dist), ("by_bearing_x_dist", bear * dist)):
    o = torch.argsort(key, descending=True)
    out[name + "_ms"] = timed(x[o].contiguous(), g[o].contiguous())[0]
    out[name + "_corr"] = float(np.corrcoef(key.cpu().numpy(), it.cpu().numpy())[0, 1])
# the library's prior, applied by physically permuting the inputs (natural queue order): what the in-kernel order should give
import re
tab = torch.tensor([float(v) for v in re.findall(r"([0-9.]+)f", open("kiss_mpc_b200/csrc/kmpc_order_prior.h").read().split("{")[1])], device="cuda")
bs = torch.atan2(d[:, 1], d[:, 0]) - x[:, 2]
bs = bs - 2 * np.pi * torch.round(bs / (2 * np.pi))
dt = torch.where(bs < 0, -1.0, 1.0) * (g[:, 2] - x[:, 2])
ia = torch.clamp(torch.floor(bs.abs() / np.pi * 12), 0, 11).long()
idt = torch.clamp(torch.floor((dt + 2 * np.pi) / (4 * np.pi) * 24), 0, 23).long()
ir = torch.clamp(torch.floor(dist / 6.0 * 4), 0, 3).long()
key = tab[(ia * 24 + idt) * 4 + ir]
o = torch.argsort(key, descending=True, stable=True)
out["prior_permuted_inputs_ms"] = timed(x[o].contiguous(), g[o].contiguous())[0]
out["prior_corr"] = float(np.corrcoef(key.cpu().numpy(), it.cpu().numpy())[0, 1])
print(json.dumps(out))
