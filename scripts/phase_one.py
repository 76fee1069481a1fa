"""Tuning aid (needs a -DKMPC_PHASE_TIMING build as KMPC_LIB): phase cycle budget of ONE instance of the headline batch solved alone."""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, ".")
from kiss_mpc_b200 import BatchedMotionPlanner, PlannerConfig, _lib
from kiss_mpc_b200.synthetic import make_batch
i = int(sys.argv[1])
b = make_batch(65536, seed=1000)
pl = BatchedMotionPlanner(PlannerConfig(), max_batch=1)
pl.set_timing(True)
x = torch.tensor(b["x_cur"][i:i + 1], device="cuda"); g = torch.tensor(b["goal"][i:i + 1], device="cuda")
L = _lib.load()
out = (C.c_double * 48)()
for rep in range(2):
    r = pl.solve(x, g); torch.cuda.synchronize()
    L.kmpc_debug_phase_cycles(out)
s = pl.stats()
names = ["fetch/init", "wait0", "assemble", "wait1", "serial/idle", "wait2", "step+logic", "trial", "decide/accept/begin_iter", "output"]
print("instance", i, "ms", s["last_kernel_ms"], "trips", s["trips"], "iters", int(r.iters[0]), "counts: retry %d soc %d accept %d backtrack %d" % (out[10], out[11], out[12], out[13]))
print("block trips", out[16], "serial window cycles/blocktrip", out[17] / max(1, out[16]))
for k in range(10):
    print(f"{names[k]:28s} {out[k] / 1.965e3:10.1f} us total over the 4 warps")
