"""ctypes front end of the CPU oracle (oracle/kmpc_oracle.c).

TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product (kiss_mpc_b200) never does.

The oracle restates the reference's per-step solve, mpc/optimizer.py:319-400 (MotionPlanner.solve -> CasADi nlpsol
"ipopt", optimizer.py:354/:375-391).  PARITY UNPINNED: the reference has no golden vectors and CasADi/IPOPT cannot be
installed in this image (see the header of kmpc_oracle.c and DESIGN.md).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libkmpc_oracle.so")

INF = 1e20  # IPOPT treats |b| >= 1e19 as "no bound"


class _Cfg(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("O", C.c_int32), ("cost_mode", C.c_int32), ("goal_k_lo", C.c_int32),
        ("goal_k_hi", C.c_int32), ("max_iter", C.c_int32), ("linsolve", C.c_int32), ("obs_stagewise", C.c_int32),
        ("T", C.c_double), ("W", C.c_double * 3), ("Wv_neg", C.c_double), ("Wv_pos", C.c_double), ("Ww", C.c_double),
        ("lo", C.c_double * 5), ("hi", C.c_double * 5), ("obs_radius", C.c_double), ("inflation", C.c_double),
        ("tol", C.c_double),
    ]


class _Diag(C.Structure):
    _fields_ = [
        ("n_factor", C.c_int32), ("n_trials", C.c_int32), ("n_soc", C.c_int32), ("max_filter", C.c_int32),
        ("mu", C.c_double), ("err", C.c_double), ("obj_scaling", C.c_double), ("max_delta_w", C.c_double),
        ("n_resto", C.c_int32), ("reserved", C.c_int32),
    ]


DIAG_DTYPE = np.dtype([("n_factor", "i4"), ("n_trials", "i4"), ("n_soc", "i4"), ("max_filter", "i4"),
                       ("mu", "f8"), ("err", "f8"), ("obj_scaling", "f8"), ("max_delta_w", "f8"), ("n_resto", "i4"), ("reserved", "i4")])


@dataclass
class OracleConfig:
    """Problem + solver options.  Defaults = the headline (README) form with EgoAgent's bounds (agent.py:104-106)."""
    N: int = 30
    T: float = 0.1
    W: Sequence[float] = (100.0, 100.0, 50.0)      # optimizer.py:57
    Wv_neg: float = 300.0                            # optimizer.py:59
    Wv_pos: float = 0.0                              # README.md:24
    Ww: float = 10.0                                 # optimizer.py:60
    cost_mode: str = "readme"                        # "readme" | "code_literal" (optimizer.py:91-96)
    goal_range: str = "readme"                       # "readme": k=1..N | "code": k=1..N-1 (optimizer.py:80)
    x_bounds: Sequence[float] = (-20.0, 20.0)        # optimizer.py:114-115
    y_bounds: Sequence[float] = (-20.0, 20.0)        # README.md:61-66 (code: none -> (-INF, INF))
    v_bounds: Sequence[float] = (-0.2, 0.5)
    w_bounds: Sequence[float] = (-0.5, 0.5)
    O: int = 0
    obs_radius: float = 0.3                          # dynamic_obstacle.py:9
    inflation: float = 0.5                           # agent.py:149 (radius + 0.1)
    tol: float = 1e-8
    max_iter: int = 2000                             # optimizer.py:346
    linsolve: str = "dense"                          # "dense" | "riccati"
    obs_stagewise: bool = False                      # centres per obstacle AND stage, obs[B,O,N,2] (dynamic_obstacle.py:47-56)

    def to_c(self) -> _Cfg:
        c = _Cfg()
        c.N, c.O = int(self.N), int(self.O)
        c.cost_mode = {"readme": 0, "code_literal": 1}[self.cost_mode]
        c.goal_k_lo = 1
        c.goal_k_hi = self.N if self.goal_range == "readme" else self.N - 1
        c.max_iter = int(self.max_iter)
        c.linsolve = {"dense": 0, "riccati": 1}[self.linsolve]
        c.obs_stagewise = int(bool(self.obs_stagewise))
        c.T = float(self.T)
        c.W = (C.c_double * 3)(*[float(v) for v in self.W])
        c.Wv_neg, c.Wv_pos, c.Ww = float(self.Wv_neg), float(self.Wv_pos), float(self.Ww)
        lo = [self.x_bounds[0], self.y_bounds[0], -INF, self.v_bounds[0], self.w_bounds[0]]
        hi = [self.x_bounds[1], self.y_bounds[1], INF, self.v_bounds[1], self.w_bounds[1]]
        c.lo = (C.c_double * 5)(*[float(max(v, -INF)) for v in lo])
        c.hi = (C.c_double * 5)(*[float(min(v, INF)) for v in hi])
        c.obs_radius, c.inflation, c.tol = float(self.obs_radius), float(self.inflation), float(self.tol)
        return c


_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only, no GPU)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "kmpc_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        L.kmo_version.restype = C.c_int
        L.kmo_duals_len.restype = C.c_int
        L.kmo_duals_len.argtypes = [C.POINTER(_Cfg)]
        L.kmo_solve.restype = C.c_int
        L.kmo_solve.argtypes = [C.POINTER(_Cfg), C.c_int, dp, dp, dp, dp, dp, dp, dp, dp, dp, ip, ip, dp, C.c_void_p, C.c_int]
        L.kmo_solve_trace.restype = C.c_int
        L.kmo_solve_trace.argtypes = [C.POINTER(_Cfg), dp, dp, dp, dp, dp, dp, dp, dp, dp, ip, ip, dp, C.c_int, ip]
        _lib = L
    return _lib


def _p(a, t=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


@dataclass
class OracleResult:
    X: np.ndarray          # [B,3,N+1]
    U: np.ndarray          # [B,2,N]
    obj: np.ndarray        # [B] unscaled objective
    status: np.ndarray     # [B] IPOPT ApplicationReturnStatus numbering
    iters: np.ndarray      # [B]
    diag: np.ndarray       # [B] DIAG_DTYPE
    duals: Optional[np.ndarray] = None  # [B, duals_len] multipliers of the SCALED problem
    meta: dict = field(default_factory=dict)


def solve(cfg: OracleConfig, x_cur, goal, X0=None, U0=None, obs=None, want_duals=False, nthreads=0, obs_rad=None) -> OracleResult:
    """Solve B independent instances.  x_cur,goal: [B,3]; X0: [B,3,N+1]; U0: [B,2,N]; obs: [B,O,2], or [B,O,N,2] with
    cfg.obs_stagewise (column t of an obstacle's track is paired with X_{t+1}, dynamic_obstacle.py:47-56); obs_rad: [B,O] radius per
    obstacle (optimizer.py:231-250 keeps one per obstacle class; None: cfg.obs_radius for all)."""
    L = lib()
    c = cfg.to_c()
    x_cur = np.ascontiguousarray(np.atleast_2d(x_cur), dtype=np.float64)
    goal = np.ascontiguousarray(np.atleast_2d(goal), dtype=np.float64)
    B, N, O = x_cur.shape[0], cfg.N, cfg.O
    assert goal.shape == (B, 3) and x_cur.shape == (B, 3)
    X0 = None if X0 is None else np.ascontiguousarray(X0, dtype=np.float64).reshape(B, 3, N + 1)
    U0 = None if U0 is None else np.ascontiguousarray(U0, dtype=np.float64).reshape(B, 2, N)
    if O:
        obs = np.ascontiguousarray(obs, dtype=np.float64).reshape((B, O, N, 2) if cfg.obs_stagewise else (B, O, 2))
    obs_rad = None if (not O or obs_rad is None) else np.ascontiguousarray(np.broadcast_to(np.asarray(obs_rad, dtype=np.float64), (B, O)))
    X = np.empty((B, 3, N + 1)); U = np.empty((B, 2, N)); obj = np.empty(B)
    status = np.empty(B, np.int32); iters = np.empty(B, np.int32)
    diag = np.zeros(B, DIAG_DTYPE)
    dl = L.kmo_duals_len(C.byref(c))
    duals = np.empty((B, dl)) if want_duals else None
    rc = L.kmo_solve(C.byref(c), B, _p(x_cur), _p(goal), _p(X0), _p(U0), _p(obs) if O else None, _p(obs_rad), _p(X), _p(U), _p(obj),
                     _p(status, C.c_int32), _p(iters, C.c_int32), _p(duals), diag.ctypes.data_as(C.c_void_p), int(nthreads))
    if rc != 0:
        raise ValueError(f"kmo_solve rejected its arguments (rc={rc})")
    return OracleResult(X, U, obj, status, iters, diag, duals, {"df": diag["obj_scaling"].copy()})


def solve_trace(cfg: OracleConfig, x_cur, goal, X0=None, U0=None, obs=None, cap=2048, obs_rad=None):
    """Single instance; returns (X,U,obj,status,iters, trace[len,8]) with rows mu,alpha_pr,alpha_du,delta_w,theta,phi,E0,f."""
    L = lib(); c = cfg.to_c(); N = cfg.N
    x_cur = np.ascontiguousarray(x_cur, dtype=np.float64).reshape(3); goal = np.ascontiguousarray(goal, dtype=np.float64).reshape(3)
    X0 = None if X0 is None else np.ascontiguousarray(X0, dtype=np.float64).reshape(3, N + 1)
    U0 = None if U0 is None else np.ascontiguousarray(U0, dtype=np.float64).reshape(2, N)
    obs = None if not cfg.O else np.ascontiguousarray(obs, dtype=np.float64).reshape((cfg.O, N, 2) if cfg.obs_stagewise else (cfg.O, 2))
    X = np.empty((3, N + 1)); U = np.empty((2, N)); obj = np.empty(1); st = np.empty(1, np.int32); it = np.empty(1, np.int32)
    rows = np.zeros((cap, 8)); ln = np.zeros(1, np.int32)
    obs_rad = None if (not cfg.O or obs_rad is None) else np.ascontiguousarray(np.broadcast_to(np.asarray(obs_rad, dtype=np.float64), (cfg.O,)))
    rc = L.kmo_solve_trace(C.byref(c), _p(x_cur), _p(goal), _p(X0), _p(U0), _p(obs), _p(obs_rad), _p(X), _p(U), _p(obj),
                           _p(st, C.c_int32), _p(it, C.c_int32), _p(rows), cap, _p(ln, C.c_int32))
    if rc != 0:
        raise ValueError("kmo_solve_trace rejected its arguments")
    return X, U, float(obj[0]), int(st[0]), int(it[0]), rows[: int(ln[0])]


def split_duals(cfg: OracleConfig, duals: np.ndarray):
    """Split one row of OracleResult.duals into (yc[3(N+1)], zL[n], zU[n], s[NO], yd[NO], vL[NO])."""
    N, O = cfg.N, cfg.O
    mc, n, ns = 3 * (N + 1), 5 * N + 3, N * O
    o = 0
    out = []
    for ln in (mc, n, n, ns, ns, ns):
        out.append(duals[..., o:o + ln]); o += ln
    return tuple(out)
