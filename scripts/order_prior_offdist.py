"""Does the queue-order prior (fitted on T = 0.1, N = 30 batches) help off its own distribution?  65,536 cold instances per case,
prior order vs index order, kernel time (events), best of 3.  Cases: the headline, N = 50 / T = 0.041 (agent.py:99-100 defaults),
the ROS node's N = 7 / T = 0.8 / +-0.3 bounds (ros2interface.py:28-38), a warm-started batch (second solve from the first's solution)."""
import json, sys, numpy as np, torch
sys.path.insert(0, ".")
from kiss_mpc_b200 import BatchedMotionPlanner, PlannerConfig
from kiss_mpc_b200.synthetic import make_batch
B = 65536
out = {}
for name, kw, seed in (("N30_T0.1", dict(N=30, T=0.1), 2001), ("N50_T0.041", dict(N=50, T=0.041), 2002),
                       ("N7_T0.8_ros", dict(N=7, T=0.8, v_bounds=(-0.3, 0.3), w_bounds=(-0.3, 0.3)), 2003)):
    b = make_batch(B, seed=seed)
    pl = BatchedMotionPlanner(PlannerConfig(**kw), max_batch=B)
    pl.set_timing(True)
    x = torch.tensor(b["x_cur"], device="cuda"); g = torch.tensor(b["goal"], device="cuda")
    res = {}
    for mode in (True, False):
        pl.set_queue_order(mode)
        best = 1e9
        for _ in range(3):
            r = pl.solve(x, g); torch.cuda.synchronize(); best = min(best, pl.stats()["last_kernel_ms"])
        res["prior_ms" if mode else "natural_ms"] = best
    if name == "N30_T0.1":   # warm start: the previous solution as the start, one control interval later
        x2 = r.states[:, :, 1].contiguous()
        for mode in (True, False):
            pl.set_queue_order(mode)
            best = 1e9
            for _ in range(3):
                r2 = pl.solve(x2, g, r.states, r.controls); torch.cuda.synchronize(); best = min(best, pl.stats()["last_kernel_ms"])
            res["warm_prior_ms" if mode else "warm_natural_ms"] = best
    res["mean_iters"] = float(r.iters.float().mean()); res["max_iters"] = int(r.iters.max())
    out[name] = res
    pl.close()
print(json.dumps(out))
