"""Tuning aid: per-phase cycle budget of the warp solver (needs a -DKMPC_PHASE_TIMING build passed as KMPC_LIB)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from kiss_mpc_b200 import BatchedMotionPlanner, PlannerConfig, _lib
from kiss_mpc_b200.synthetic import make_batch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
b = make_batch(B, seed=1000)
pl = BatchedMotionPlanner(PlannerConfig(N=N), max_batch=B)
pl.set_timing(True)
x = torch.tensor(b["x_cur"], device="cuda"); g = torch.tensor(b["goal"], device="cuda")
L = _lib.load()
out = (C.c_double * 48)()
for rep in range(2):
    r = pl.solve(x, g); torch.cuda.synchronize()
    n = L.kmpc_debug_phase_cycles(out)
s = pl.stats()
names16 = ["n_inertia_retry", "n_soc_started", "n_accept", "n_backtrack", "n_accept_in_soc", "-"]
names = ["fetch/init", "wait0 (block_any)", "assemble", "wait1", "serial(warp0)/idle", "wait2", "step+logic", "trial", "decide/accept/begin_iter", "output"]
tot = sum(out[i] for i in range(10))
print("ms", s["last_kernel_ms"], "trips/inst", s["trips"] / B)
for i in range(10, 15):
    print(f"{names16[i - 10]:20s} {out[i] / B:8.3f} per instance")
if n > 17 and out[16]:
    print(f"serial window (all the block's recursions, on the serial warp): {out[17] / out[16]:.0f} cycles per block trip, {out[16]:.0f} block trips")
for i in range(10):
    print(f"{names[i]:28s} {out[i] / tot * 100:6.2f}%   {out[i] / s['trips']:10.0f} cycles per trip")
