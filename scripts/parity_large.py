"""Large parity run: the headline problem (N = 30, box bounds, cold start) on `n_batches` x 65,536 fresh instances (seeds disjoint
from the tests' and the benchmark's) against the CPU oracle on all host cores.  One JSON line: status agreement, iteration-count
agreement, max |dU| / relative objective difference over the instances both sides converge on, and the status histogram."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from kiss_mpc_b200 import BatchedMotionPlanner, PlannerConfig
from kiss_mpc_b200.synthetic import make_batch
from oracle import oracle as ok

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 16
B = 65536
ok.build()
pl = BatchedMotionPlanner(PlannerConfig(), max_batch=B)
ocfg = ok.OracleConfig(linsolve="riccati")
tot = dict(instances=0, status_equal=0, iters_equal=0, converged_both=0)
worst_dU = worst_dobj = 0.0
hist = {}
t_gpu = t_cpu = 0.0
for k in range(nb):
    b = make_batch(B, seed=500000 + k)
    x = torch.tensor(b["x_cur"], device="cuda"); g = torch.tensor(b["goal"], device="cuda")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = pl.solve(x, g); torch.cuda.synchronize(); t_gpu += time.perf_counter() - t0
    t0 = time.perf_counter(); ref = ok.solve(ocfg, b["x_cur"], b["goal"], nthreads=os.cpu_count()); t_cpu += time.perf_counter() - t0
    st = r.status.cpu().numpy(); it = r.iters.cpu().numpy(); U = r.controls.cpu().numpy(); obj = r.objective.cpu().numpy()
    conv = (st == 0) & (ref.status == 0)
    tot["instances"] += B; tot["status_equal"] += int((st == ref.status).sum()); tot["iters_equal"] += int((it == ref.iters).sum())
    tot["converged_both"] += int(conv.sum())
    worst_dU = max(worst_dU, float(np.abs(U - ref.U)[conv].max())); worst_dobj = max(worst_dobj, float((np.abs(obj - ref.obj) / np.abs(ref.obj))[conv].max()))
    for s, c in zip(*np.unique(st, return_counts=True)):
        hist[int(s)] = hist.get(int(s), 0) + int(c)
n = tot["instances"]
print(json.dumps({"instances": n, "status_equal": tot["status_equal"] / n, "iters_equal": tot["iters_equal"] / n, "converged_both": tot["converged_both"] / n,
                  "max_abs_dU": worst_dU, "max_rel_dobj": worst_dobj, "gpu_status_histogram": hist, "gpu_seconds": t_gpu, "cpu_oracle_seconds": t_cpu,
                  "cpu_cores": os.cpu_count()}))
