"""ctypes front end of tests/host_emul (g++ build of the per-thread CUDA solver source).  TEST HARNESS ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "host_emul", "emul.cpp")
_SO = os.path.join(_HERE, "host_emul", "libkmpc_emul.so")
_CORE = os.path.join(_HERE, "..", "kiss_mpc_b200", "csrc", "kmpc_core.cuh")
_WARP = os.path.join(_HERE, "..", "kiss_mpc_b200", "csrc", "kmpc_warp.cuh")


class KmpcConfig(C.Structure):
    _fields_ = [("N", C.c_int32), ("O_max", C.c_int32), ("cost_mode", C.c_int32), ("goal_k_lo", C.c_int32),
                ("goal_k_hi", C.c_int32), ("max_iter", C.c_int32), ("B_max", C.c_int32), ("layout", C.c_int32),
                ("device", C.c_int32), ("reserved", C.c_int32), ("T", C.c_double), ("W", C.c_double * 3),
                ("Wv_neg", C.c_double), ("Wv_pos", C.c_double), ("Ww", C.c_double), ("lo", C.c_double * 4),
                ("hi", C.c_double * 4), ("tol", C.c_double)]


def build():
    csrc = os.path.dirname(_CORE)
    deps = [_SRC, os.path.join(_HERE, "host_emul", "simt.h")] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".h"))]
    newest = max(os.path.getmtime(p) for p in deps)
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < newest:
        subprocess.check_call(["g++", "-O2", "-fopenmp", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-I", os.path.join(_HERE, "host_emul"), "-o", _SO, _SRC])
    return _SO


def cfg_from_oracle(ocfg, B_max=1, layout=0):
    c = KmpcConfig()
    c.N, c.O_max = ocfg.N, ocfg.O
    c.cost_mode = {"readme": 0, "code_literal": 1}[ocfg.cost_mode]
    c.goal_k_lo, c.goal_k_hi = 1, (ocfg.N if ocfg.goal_range == "readme" else ocfg.N - 1)
    c.max_iter, c.B_max, c.layout, c.device = ocfg.max_iter, B_max, layout, 0
    c.T = ocfg.T
    c.W = (C.c_double * 3)(*ocfg.W)
    c.Wv_neg, c.Wv_pos, c.Ww = ocfg.Wv_neg, ocfg.Wv_pos, ocfg.Ww
    lo = [ocfg.x_bounds[0], ocfg.y_bounds[0], ocfg.v_bounds[0], ocfg.w_bounds[0]]
    hi = [ocfg.x_bounds[1], ocfg.y_bounds[1], ocfg.v_bounds[1], ocfg.w_bounds[1]]
    c.lo = (C.c_double * 4)(*[max(float(v), -1e20) for v in lo])
    c.hi = (C.c_double * 4)(*[min(float(v), 1e20) for v in hi])
    c.tol = ocfg.tol
    return c


def solve(ocfg, x_cur, goal, X0=None, U0=None, obs=None, layout=0, warp=False, obs_rad=None, warps=1, chunk=0):
    L = C.CDLL(build())
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    L.emul_solve.restype = C.c_int
    L.emul_solve.argtypes = [C.POINTER(KmpcConfig), C.c_int, dp, dp, dp, dp, dp, dp, C.c_int, C.c_int, C.c_double, C.c_double, dp, dp,
                             dp, ip, ip, ip]
    x_cur = np.ascontiguousarray(x_cur, float); goal = np.ascontiguousarray(goal, float)
    B, N, O = x_cur.shape[0], ocfg.N, ocfg.O
    c = cfg_from_oracle(ocfg, B, layout)
    sw = int(bool(getattr(ocfg, "obs_stagewise", False)) and O > 0)

    def tr_in(a, shape):
        if a is None:
            return None
        a = np.ascontiguousarray(a, float).reshape((B,) + shape)
        return np.ascontiguousarray(np.moveaxis(a, 0, -1)) if layout else a

    xi, gi = tr_in(x_cur, (3,)), tr_in(goal, (3,))
    X0i, U0i, obi = tr_in(X0, (3, N + 1)), tr_in(U0, (2, N)), tr_in(obs if O else None, (O, N, 2) if sw else (O, 2))
    ori = tr_in(None if (obs_rad is None or not O) else np.broadcast_to(np.asarray(obs_rad, float), (B, O)), (O,))
    Xo = np.full((3, N + 1, B) if layout else (B, 3, N + 1), np.nan); Uo = np.full((2, N, B) if layout else (B, 2, N), np.nan)
    obj = np.empty(B); st = np.empty(B, np.int32); it = np.empty(B, np.int32); tp = np.empty(B, np.int32)
    p = lambda a, t=C.c_double: None if a is None else a.ctypes.data_as(C.POINTER(t))
    if warp:
        L.emul_solve_warp.restype = C.c_int
        L.emul_solve_warp.argtypes = [C.POINTER(KmpcConfig), C.c_int, dp, dp, dp, dp, dp, dp, C.c_int, C.c_int, C.c_double, C.c_double, dp, dp,
                                      dp, ip, ip, ip, C.c_int, C.c_int]
        rc = L.emul_solve_warp(C.byref(c), B, p(xi), p(gi), p(X0i), p(U0i), p(obi), p(ori), O, sw, ocfg.obs_radius, ocfg.inflation, p(Xo), p(Uo),
                               p(obj), p(st, C.c_int32), p(it, C.c_int32), p(tp, C.c_int32), int(warps), int(chunk))
    else:
        rc = L.emul_solve(C.byref(c), B, p(xi), p(gi), p(X0i), p(U0i), p(obi), p(ori), O, sw, ocfg.obs_radius, ocfg.inflation, p(Xo), p(Uo),
                          p(obj), p(st, C.c_int32), p(it, C.c_int32), p(tp, C.c_int32))
    assert rc == 0
    if layout:
        Xo = np.ascontiguousarray(np.moveaxis(Xo, -1, 0)); Uo = np.ascontiguousarray(np.moveaxis(Uo, -1, 0))
    return Xo, Uo, obj, st, it, tp
