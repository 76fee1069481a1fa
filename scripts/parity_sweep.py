"""Randomised parity sweep: the CUDA path (through the C ABI) against the CPU oracle over several seeds and problem forms.
Prints one JSON line per case: status agreement, max |dU| and max relative objective difference on instances both converge on,
fraction with identical iteration counts.  Tolerances of BASELINE.json: objective 1e-6 rel., controls 1e-5."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from kiss_mpc_b200 import BatchedMotionPlanner, PlannerConfig
from kiss_mpc_b200.synthetic import make_batch
from oracle import oracle as ok

ok.build()
nthr = os.cpu_count()
dev = lambda a: None if a is None else torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")
cases = []
for seed in (11, 12, 13, 14):
    cases += [("box_N30", dict(), 8192, 0, seed), ("N50", dict(N=50), 2048, 0, seed), ("obs_O10", dict(O=10), 2048, 10, seed),
              ("code_literal", dict(cost_mode="code_literal", goal_range="code", y_bounds=(-ok.INF, ok.INF)), 2048, 0, seed),
              ("N7_T0.8_ros", dict(N=7, T=0.8, v_bounds=(-0.3, 0.3), w_bounds=(-0.3, 0.3)), 2048, 0, seed)]
worst = {}
for name, kw, B, O, seed in cases:
    ocfg = ok.OracleConfig(linsolve="riccati", **kw)
    pk = {k: v for k, v in kw.items() if k != "O"}
    if O: pk["O_max"] = O
    if "y_bounds" in pk: pk["y_bounds"] = (-1e30, 1e30)
    b = make_batch(B, seed=seed, O=O)
    pl = BatchedMotionPlanner(PlannerConfig(**pk), max_batch=B)
    for warm in (False, True):
        if warm:
            X0, U0, x = ref.X, ref.U, ref.X[:, :, 1].copy()
        else:
            X0, U0, x = None, None, b["x_cur"]
        ref = ok.solve(ocfg, x, b["goal"], X0=X0, U0=U0, obs=b["obs"], nthreads=nthr)
        res = pl.solve(dev(x), dev(b["goal"]), dev(X0), dev(U0), obstacles=dev(b["obs"]), obstacle_radius=ocfg.obs_radius,
                       inflation_radius=ocfg.inflation if O else 0.0)
        st = res.status.cpu().numpy(); U = res.controls.cpu().numpy(); obj = res.objective.cpu().numpy(); it = res.iters.cpu().numpy()
        conv = (st == 0) & (ref.status == 0)
        row = {"case": name, "seed": seed, "warm_start": warm, "B": B, "status_equal": float((st == ref.status).mean()),
               "converged": float(conv.mean()), "max_abs_dU": float(np.abs(U - ref.U)[conv].max()),
               "max_rel_dobj": float((np.abs(obj - ref.obj) / np.maximum(1.0, np.abs(ref.obj)))[conv].max()),
               "iters_equal": float((it == ref.iters).mean()), "statuses": {int(k): int(v) for k, v in zip(*np.unique(st, return_counts=True))}}
        print(json.dumps(row), flush=True)
        w = worst.setdefault(name, {"status_equal": 1.0, "max_abs_dU": 0.0, "max_rel_dobj": 0.0, "iters_equal": 1.0, "instances": 0})
        w["status_equal"] = min(w["status_equal"], row["status_equal"]); w["max_abs_dU"] = max(w["max_abs_dU"], row["max_abs_dU"])
        w["max_rel_dobj"] = max(w["max_rel_dobj"], row["max_rel_dobj"]); w["iters_equal"] = min(w["iters_equal"], row["iters_equal"]); w["instances"] += B
    pl.close()
print(json.dumps({"summary_worst_over_seeds": worst}))
