"""Tuning aid: per-instance schedule of the persistent kernel (needs the -DKMPC_SCHED_TRACE build, KMPC_LIB=...):
when each instance was taken and finished, its trips, its SM.  Natural and prior queue order on the headline batch."""
import ctypes as C, json, sys
import numpy as np, torch
sys.path.insert(0, ".")
from kiss_mpc_b200 import BatchedMotionPlanner, PlannerConfig, _lib
from kiss_mpc_b200.synthetic import make_batch

B = 65536
b = make_batch(B, seed=1000)
pl = BatchedMotionPlanner(PlannerConfig(), max_batch=B)
L = _lib.load()
x = torch.tensor(b["x_cur"], device="cuda"); g = torch.tensor(b["goal"], device="cuda")
buf = torch.zeros((B, 4), dtype=torch.int64, device="cuda")
L.kmpc_debug_sched_trace.argtypes = [C.c_void_p]
out = {}
for prior in (False, True):
    pl.set_queue_order(prior)
    pl.solve(x, g); torch.cuda.synchronize()
    buf.zero_(); L.kmpc_debug_sched_trace(C.c_void_p(buf.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = pl.solve(x, g); e1.record(); torch.cuda.synchronize()
    t = buf.cpu().numpy()
    t0 = t[:, 0].min()
    st = (t[:, 0] - t0) * 1e-6; en = (t[:, 1] - t0) * 1e-6; trips = t[:, 2]
    dur = en - st
    name = "prior" if prior else "natural"
    np.save(f"gpurun_out/sched_{name}.npy", np.stack([st, en, trips, t[:, 3] >> 32, r.iters.cpu().numpy(), t[:, 3] & 0xffffffff], 1))
    edges = np.linspace(0, en.max(), 21)
    conc = [int(((st <= e) & (en > e)).sum()) for e in edges]
    late = np.argsort(-en)[:8]
    out[name] = {"ms": e0.elapsed_time(e1), "end_ms": float(en.max()), "trips_sum": int(trips.sum()), "us_per_trip_mean": float((dur * 1e3).sum() / trips.sum()),
                 "us_per_trip_by_quartile_of_start": [float((dur[q] * 1e3).sum() / trips[q].sum()) for q in np.array_split(np.argsort(st), 4)],
                 "concurrency_at_5pct_steps": conc,
                 "longest": [(int(i), round(float(st[i]), 2), round(float(en[i]), 2), int(trips[i]), int(t[i, 3] >> 32)) for i in np.argsort(-trips)[:6]],
                 "last_finishers": [(round(float(st[i]), 2), round(float(en[i]), 2), int(trips[i])) for i in late]}
print(json.dumps(out))
