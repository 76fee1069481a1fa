"""All five BASELINE.json configurations on one GPU (device-resident timing, CUDA events, best of `reps` after a warm-up).
Prints one JSON object; the headline metric itself is bench.py's job.  CPU arm (oracle port, all cores) beside each."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from kiss_mpc_b200 import BatchedMotionPlanner, MotionPlanner, PlannerConfig
from kiss_mpc_b200.synthetic import cfg1_instance, make_batch, make_tracks
from oracle import oracle as ok

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
out = {}


def timed(fn):
    best = 1e30
    r = None
    for _ in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, r


def cpu(cfg, b, n, **kw):
    t0 = time.perf_counter()
    r = ok.solve(cfg, b["x_cur"][:n], b["goal"][:n], obs=None if b["obs"] is None else b["obs"][:n], nthreads=os.cpu_count(), **kw)
    return n / (time.perf_counter() - t0), r


def batch_case(name, B, N, O, seed, cpu_n, tracks=False):
    b = make_batch(B, seed=seed, O=O)
    if tracks:   # every circle drifts along a constant-velocity track
        b["obs"] = make_tracks(b["obs"], N, seed=seed)
    pl = BatchedMotionPlanner(PlannerConfig(N=N, O_max=O), max_batch=B)
    x = torch.tensor(b["x_cur"], device="cuda"); g = torch.tensor(b["goal"], device="cuda")
    ob = torch.tensor(b["obs"], device="cuda") if O else None
    ms, r = timed(lambda: pl.solve(x, g, obstacles=ob, obstacle_radius=0.3, inflation_radius=0.5 if O else 0.0))
    rate, ref = cpu(ok.OracleConfig(N=N, O=O, linsolve="riccati", obs_stagewise=tracks), b, cpu_n)
    st = r.status.cpu().numpy()[:cpu_n]; U = r.controls.cpu().numpy()[:cpu_n]; obj = r.objective.cpu().numpy()[:cpu_n]
    conv = (st == 0) & (ref.status == 0)
    out[name] = {"B": B, "N": N, "O": O, "ms": ms, "solves_per_sec": B / ms * 1e3, "mean_iters": r.iters.float().mean().item(),
                 "max_iters": int(r.iters.max().item()), "converged": (r.status == 0).float().mean().item(),
                 "cpu_oracle_solves_per_sec": rate, "cpu_cores": os.cpu_count(), "cpu_sample": cpu_n,
                 "parity_status_equal": float((st == ref.status).mean()), "parity_max_abs_dU": float(np.abs(U - ref.U)[conv].max()),
                 "parity_max_rel_dobj": float((np.abs(obj - ref.obj) / np.abs(ref.obj))[conv].max())}
    pl.close()


# cfg 1: single agent through the drop-in MotionPlanner (B = 1 latency, NumPy in/out, H2D/D2H inside)
xc, gl = cfg1_instance(); N = 30
mp = MotionPlanner(time_step=0.1, horizon=N)
X0 = np.tile(xc.reshape(3, 1), (1, N + 1)); U0 = np.zeros((2, N))
kw = dict(current_state=xc[0], current_linear_velocity=0.0, current_angular_velocity=0.0, goal_state=gl[0], states_matrix=X0,
          controls_matrix=U0, state_bounds=(-20, 20), linear_velocity_bounds=(-0.2, 0.5), angular_velocity_bounds=(-0.5, 0.5))
mp.solve(**kw)
lat = []
for _ in range(20):
    t0 = time.perf_counter(); mp.solve(**kw); lat.append(time.perf_counter() - t0)
t0 = time.perf_counter(); ref1 = ok.solve(ok.OracleConfig(linsolve="dense"), xc, gl, X0=X0[None], U0=U0[None]); cpu1 = time.perf_counter() - t0
out["cfg1_single_agent_dropin"] = {"p50_ms": float(np.median(lat) * 1e3), "iters": mp.last_iterations, "status": mp.last_status,
                                   "cpu_oracle_ms": cpu1 * 1e3, "objective": mp.last_objective, "oracle_objective": float(ref1.obj[0])}
batch_case("cfg2_4096_N30", 4096, 30, 0, 1002, 4096)
batch_case("headline_65536_N30", 65536, 30, 0, 1000, 8192)
batch_case("cfg3_65536_N50", 65536, 50, 0, 1003, 4096)
batch_case("cfg4_4096_N30_O10", 4096, 30, 10, 1004, 4096)
batch_case("cfg4_65536_N30_O10", 65536, 30, 10, 1004, 2048)
batch_case("tracks_4096_N30_O10", 4096, 30, 10, 1004, 2048, tracks=True)
batch_case("tracks_65536_N30_O10", 65536, 30, 10, 1004, 2048, tracks=True)
# cfg 5: closed loop, 16,384 agents x 200 steps, warm-started, stop at goal (agent.py:65 goal radius 0.5)
B, steps = 16384, 200
b = make_batch(B, seed=1005)
pl = BatchedMotionPlanner(PlannerConfig(N=30), max_batch=B)
g = torch.tensor(b["goal"], device="cuda")
best = 1e30
for _ in range(2):
    x = torch.tensor(b["x_cur"], device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    X, U, applied, iters, status = pl.closed_loop(x, g, steps, goal_radius=0.5)
    e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
solved = status != 1000
n = int(solved.sum().item())
out["cfg5_closed_loop_16384x200"] = {"ms": best, "solves": n, "solves_per_sec": n / best * 1e3,
                                      "mean_iters_step0": (iters[0].float().mean()).item(),
                                      "mean_iters_warm": float(((iters[1:].float() * solved[1:]).sum() / solved[1:].sum()).item()),
                                      "agents_at_goal_after_200": float((~solved[-1]).float().mean().item()),
                                      "converged_of_solved": float(((status == 0) & solved).sum().item() / n)}
print(json.dumps(out))
