"""Where one B = 1 solve through the drop-in spends its time: MotionPlanner.solve (NumPy in/out) vs the bare kmpc_solve_host call
vs the solver kernel alone (CUDA events).  usage: python scripts/b1_breakdown.py [N] [T]"""
import ctypes as C, sys, time
import numpy as np
sys.path.insert(0, ".")
from kiss_mpc_b200 import MotionPlanner
from kiss_mpc_b200.synthetic import cfg1_instance
N = int(sys.argv[1]) if len(sys.argv) > 1 else 30
T = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
vb, wb = ((-0.2, 0.5), (-0.5, 0.5)) if N == 30 else ((-0.3, 0.3), (-0.3, 0.3))
x, g = cfg1_instance()
mp = MotionPlanner(time_step=T, horizon=N, on_failure="ignore")
X0 = np.tile(x[0], (N + 1, 1)).T; U0 = np.zeros((2, N))
kw = dict(current_state=x[0], goal_state=g[0], states_matrix=X0, controls_matrix=U0, linear_velocity_bounds=vb, angular_velocity_bounds=wb)
def p50(fn, n=200):
    for _ in range(10): fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return float(np.median(ts) * 1e6)
full = p50(lambda: mp.solve(**kw))
pl = mp._planner
xs = np.ascontiguousarray(x[:1]); gs = np.ascontiguousarray(g[:1]); X0b = np.ascontiguousarray(X0[None]); U0b = np.ascontiguousarray(U0[None])
Xo = np.empty((1, 3, N + 1)); Uo = np.empty((1, 2, N)); ob = np.empty(1); st = np.empty(1, np.int32); it = np.empty(1, np.int32)
p = lambda a: C.c_void_p(a.ctypes.data)
args = (pl._h, 1, p(xs), p(gs), p(X0b), p(U0b), None, 0, 0.0, None, 0.0, p(Xo), p(Uo), p(ob), p(st), p(it))
raw = p50(lambda: pl._L.kmpc_solve_host(*args))
batched = p50(lambda: pl.solve(xs, gs, X0b, U0b))
pl.set_timing(True)
ks = []
for _ in range(50):
    pl._L.kmpc_solve_host(*args); ks.append(pl.stats()["last_kernel_ms"] * 1e3)
pl.set_timing(False)
print(f"N {N} T {T}: MotionPlanner.solve p50 {full:.1f} us | BatchedMotionPlanner.solve(numpy) {batched:.1f} us | kmpc_solve_host (ctypes, prebuilt args) {raw:.1f} us | "
      f"kernel alone (events) p50 {np.median(ks):.1f} us | iterations {int(it[0])} status {int(st[0])}")
