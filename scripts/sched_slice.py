"""Tuning aid (needs a -DKMPC_SCHED_TRACE build as KMPC_LIB): schedule of one contiguous slice of the headline batch -- when the
longest instances were taken / finished and what a trip cost them.  usage: python scripts/sched_slice.py lo hi"""
import ctypes as C, json, sys
import numpy as np, torch
sys.path.insert(0, ".")
from kiss_mpc_b200 import BatchedMotionPlanner, PlannerConfig, _lib
from kiss_mpc_b200.synthetic import make_batch
lo, hi = int(sys.argv[1]), int(sys.argv[2])
b = make_batch(65536, seed=1000)
B = hi - lo
pl = BatchedMotionPlanner(PlannerConfig(), max_batch=B)
L = _lib.load()
x = torch.tensor(b["x_cur"][lo:hi], device="cuda"); g = torch.tensor(b["goal"][lo:hi], device="cuda")
buf = torch.zeros((B, 4), dtype=torch.int64, device="cuda")
L.kmpc_debug_sched_trace.argtypes = [C.c_void_p]
pl.solve(x, g); torch.cuda.synchronize()
buf.zero_(); L.kmpc_debug_sched_trace(C.c_void_p(buf.data_ptr()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); r = pl.solve(x, g); e1.record(); torch.cuda.synchronize()
t = buf.cpu().numpy()
t0 = t[:, 0].min()
st = (t[:, 0] - t0) * 1e-6; en = (t[:, 1] - t0) * 1e-6; trips = t[:, 2]; sm = t[:, 3] >> 32; blk = (t[:, 3] >> 8) & 0xffffff
top = np.argsort(-trips)[:5]
res = {"ms": e0.elapsed_time(e1), "end_ms": float(en.max()), "p99_end_ms": float(np.percentile(en, 99)), "p999_end_ms": float(np.percentile(en, 99.9)),
       "longest": [{"i": int(i + lo), "start": round(float(st[i]), 3), "end": round(float(en[i]), 3), "trips": int(trips[i]), "iters": int(r.iters[i]),
                    "us_per_trip": round(float((en[i] - st[i]) * 1e3 / trips[i]), 2), "block": int(blk[i]),
                    "others_in_block_ending_after_start+1ms": int(((blk == blk[i]) & (en > st[i] + 1.0)).sum()) - 1,
                    "block_mates_end_max": round(float(np.append(en[(blk == blk[i]) & (np.arange(B) != i)], 0.0).max()), 3)} for i in top]}
print(json.dumps(res))
