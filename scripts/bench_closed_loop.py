"""BASELINE configs[4]: closed-loop receding-horizon rollout, 200 steps, warm-started, 16,384 agents (N=30, T=0.1, seed 1005).
Device-resident (kmpc_closed_loop): solve -> x <- X[:,1] -> unshifted warm start -> repeat.  Prints one JSON line."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from kiss_mpc_b200 import BatchedMotionPlanner, PlannerConfig
from kiss_mpc_b200.synthetic import make_batch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
goal_radius = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5    # agent.py:65 (0 = keep solving at the goal)
agent_radius = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0   # 0.4 = the literal at_goal distance of geometry.py:44
b = make_batch(B, seed=1005)
pl = BatchedMotionPlanner(PlannerConfig(N=30, T=0.1), max_batch=B)
g = torch.tensor(b["goal"], device="cuda")
for rep in range(2):
    x = torch.tensor(b["x_cur"], device="cuda")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    X, U, applied, iters, status = pl.closed_loop(x, g, steps, goal_radius=goal_radius, agent_radius=agent_radius)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
solved = status != 1000
n_solved = int(solved.sum().item())
it = (iters.float() * solved).sum(1) / solved.sum(1).clamp(min=1)
it = it.cpu().numpy()
dist = (x[:, :2] - g[:, :2]).norm(dim=1)
print(json.dumps({"workload": f"closed loop, {B} agents x {steps} steps, N=30, warm-started (unshifted), seed 1005", "ms_total": ms,
                  "goal_radius": goal_radius, "agent_radius": agent_radius, "solves": n_solved, "solves_per_sec": n_solved / (ms * 1e-3),
                  "active_agents_last_step": int(solved[-1].sum().item()), "mean_iters_step0": float(it[0]), "mean_iters_steps_1_10": float(it[1:11].mean()),
                  "mean_iters_last_10": float(it[-10:].mean()), "converged_fraction_of_solved": float(((status == 0) & solved).sum().item() / max(1, n_solved)),
                  "agents_within_0.5m_of_goal": float((dist < 0.5).float().mean().item()),
                  "status_counts_of_solved": {int(k): int(v) for k, v in zip(*[t.tolist() for t in torch.unique(status[solved], return_counts=True)])}}))
