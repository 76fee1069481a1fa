"""Obstacle batches at full size against the oracle on EVERY instance (the restoration phase included): 65,536 x N = 30 x O = 10 with
static circles and with circles on constant-velocity tracks.  Prints one JSON line per case."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from kiss_mpc_b200 import BatchedMotionPlanner, PlannerConfig
from kiss_mpc_b200.synthetic import make_batch, make_tracks
from oracle import oracle as ok
B, N, O = int(sys.argv[1]) if len(sys.argv) > 1 else 65536, 30, 10
for tracks in (False, True):
    b = make_batch(B, seed=1004, O=O)
    if tracks:
        b["obs"] = make_tracks(b["obs"], N, seed=1004)
    pl = BatchedMotionPlanner(PlannerConfig(N=N, O_max=O), max_batch=B)
    x = torch.tensor(b["x_cur"], device="cuda"); g = torch.tensor(b["goal"], device="cuda"); ob = torch.tensor(b["obs"], device="cuda")
    r = pl.solve(x, g, obstacles=ob, obstacle_radius=0.3, inflation_radius=0.5); torch.cuda.synchronize()
    t0 = time.perf_counter()
    ref = ok.solve(ok.OracleConfig(N=N, O=O, linsolve="riccati", obs_stagewise=tracks), b["x_cur"], b["goal"], obs=b["obs"], nthreads=os.cpu_count())
    dt = time.perf_counter() - t0
    st = r.status.cpu().numpy(); it = r.iters.cpu().numpy(); U = r.controls.cpu().numpy(); obj = r.objective.cpu().numpy()
    conv = (st == 0) & (ref.status == 0)
    mism = np.nonzero(st != ref.status)[0]
    print(json.dumps({"case": "tracks" if tracks else "static", "B": B, "O": O, "oracle_seconds": dt,
                      "gpu_status_hist": {int(k): int(v) for k, v in zip(*np.unique(st, return_counts=True))},
                      "oracle_status_hist": {int(k): int(v) for k, v in zip(*np.unique(ref.status, return_counts=True))},
                      "status_equal": float((st == ref.status).mean()), "status_mismatches": [(int(i), int(st[i]), int(ref.status[i])) for i in mism[:10]],
                      "iters_equal": float((it == ref.iters).mean()), "iters_equal_among_nonzero_status": float((it == ref.iters)[ref.status != 0].mean()) if (ref.status != 0).any() else None,
                      "max_abs_dU_converged": float(np.abs(U - ref.U)[conv].max()), "n_dU_above_1e-5": int((np.abs(U - ref.U).max(axis=(1, 2))[conv] > 1e-5).sum()),
                      "max_rel_dobj_converged": float((np.abs(obj - ref.obj) / np.maximum(1.0, np.abs(ref.obj)))[conv].max())}))
    pl.close()
