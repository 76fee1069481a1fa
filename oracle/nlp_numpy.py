"""NumPy statement of the reference NLP, independent of both the C oracle and the CUDA kernels.

TEST INFRASTRUCTURE ONLY (same rule as oracle.py).  Used for
  * solver-independent KKT certificates of any returned (X, U),
  * the SciPy SLSQP "polish" cross-check,
  * finite-difference checks of the derivatives the kernels use.

Problem (all reference line numbers are /root/reference/mpc/optimizer.py unless noted):
  variables   z = [vec(X) col-major ; vec(U) col-major]                        :74-77
  cost        goal tracking :79-83, v penalty :91-96 / README.md:23-24, omega penalty :97-101
  equalities  g = [x_0 - x_cur ; x_{k+1} - f(x_k,u_k)], explicit-Euler unicycle   :163-196
  obstacles   |p_k - c_o| - r_o >= I, k = 1..N, obstacle-major rows (intended form, README.md:78-81; :217-258)
  bounds      :111-156 (+ y per README.md:61-66)
"""
from __future__ import annotations

import numpy as np


class NLP:
    def __init__(self, cfg, x_cur, goal, obs=None, obs_rad=None):
        self.cfg = cfg
        self.N = N = int(cfg.N)
        self.O = int(cfg.O)
        self.T = float(cfg.T)
        self.W = np.asarray(cfg.W, float)
        self.x_cur = np.asarray(x_cur, float).reshape(3)
        self.goal = np.asarray(goal, float).reshape(3)
        # centre of obstacle o at stage k = self.cen[o, k - 1]: constant per obstacle (optimizer.py:217-221), or the obstacle's
        # own track, column t paired with X_{t+1} (dynamic_obstacle.py:47-56), when obs has a stage axis
        obs = None if not self.O else np.asarray(obs, float)
        self.obs = obs
        # radius per obstacle (optimizer.py:231-250: one per obstacle class); default: cfg.obs_radius for all
        self.rad = np.full(self.O, float(cfg.obs_radius)) if obs_rad is None else np.broadcast_to(np.asarray(obs_rad, float), (self.O,)).copy()
        if obs is not None:
            self.cen = obs.reshape(self.O, N, 2) if obs.ndim == 3 else np.repeat(obs.reshape(self.O, 1, 2), N, axis=1)
        self.n = 5 * N + 3
        self.k_lo = 1
        self.k_hi = N if cfg.goal_range == "readme" else N - 1
        lo = np.empty(self.n); hi = np.empty(self.n)
        lo[0:3 * (N + 1):3], hi[0:3 * (N + 1):3] = cfg.x_bounds
        lo[1:3 * (N + 1):3], hi[1:3 * (N + 1):3] = cfg.y_bounds
        lo[2:3 * (N + 1):3], hi[2:3 * (N + 1):3] = -np.inf, np.inf
        lo[3 * (N + 1)::2], hi[3 * (N + 1)::2] = cfg.v_bounds
        lo[3 * (N + 1) + 1::2], hi[3 * (N + 1) + 1::2] = cfg.w_bounds
        lo[lo <= -1e19] = -np.inf; hi[hi >= 1e19] = np.inf
        self.lo, self.hi = lo, hi

    # -- packing ---------------------------------------------------------
    def pack(self, X, U):
        return np.concatenate([np.asarray(X, float).T.reshape(-1), np.asarray(U, float).T.reshape(-1)])

    def unpack(self, z):
        N = self.N
        return z[:3 * (N + 1)].reshape(N + 1, 3).T, z[3 * (N + 1):].reshape(N, 2).T

    # -- objective --------------------------------------------------------
    def f(self, z):
        X, U = self.unpack(z)
        e = X[:, self.k_lo:self.k_hi + 1] - self.goal[:, None]
        f = float(np.sum(self.W[:, None] * e * e))
        v, om = U[0], U[1]
        if self.cfg.cost_mode == "readme":
            f += float(np.sum(self.cfg.Wv_neg * np.minimum(v, 0) ** 2 + self.cfg.Wv_pos * np.maximum(v, 0) ** 2))
        else:
            f += float(self.cfg.Wv_neg * np.sum(np.minimum(v, 0)))
        return f + float(self.cfg.Ww * np.sum(om * om))

    def grad(self, z):
        N = self.N
        X, U = self.unpack(z)
        gX = np.zeros((3, N + 1)); gU = np.zeros((2, N))
        gX[:, self.k_lo:self.k_hi + 1] = 2 * self.W[:, None] * (X[:, self.k_lo:self.k_hi + 1] - self.goal[:, None])
        v = U[0]
        if self.cfg.cost_mode == "readme":
            gU[0] = 2 * self.cfg.Wv_neg * np.minimum(v, 0) + 2 * self.cfg.Wv_pos * np.maximum(v, 0)
        else:
            gU[0] = self.cfg.Wv_neg * np.where(v < 0, 1.0, np.where(v == 0, 0.5, 0.0))
        gU[1] = 2 * self.cfg.Ww * U[1]
        return self.pack(gX, gU)

    # -- constraints -------------------------------------------------------
    def c(self, z):
        X, U = self.unpack(z)
        T = self.T
        nxt = X[:, :-1] + T * np.stack([U[0] * np.cos(X[2, :-1]), U[0] * np.sin(X[2, :-1]), U[1]])
        return np.concatenate([(X[:, 0] - self.x_cur), (X[:, 1:] - nxt).T.reshape(-1)])

    def jac_c(self, z):
        N, T = self.N, self.T
        X, U = self.unpack(z)
        J = np.zeros((3 * (N + 1), self.n))
        J[0:3, 0:3] = np.eye(3)
        for k in range(N):
            r = 3 * (k + 1); cx = 3 * k; cu = 3 * (N + 1) + 2 * k
            th, v = X[2, k], U[0, k]
            J[r:r + 3, r:r + 3] = np.eye(3)
            J[r:r + 3, cx:cx + 3] = -np.eye(3)
            J[r, cx + 2] = T * v * np.sin(th); J[r + 1, cx + 2] = -T * v * np.cos(th)
            J[r, cu] = -T * np.cos(th); J[r + 1, cu] = -T * np.sin(th); J[r + 2, cu + 1] = -T
        return J

    def d(self, z):
        """obstacle rows, obstacle-major, k=1..N: |p_k - c_o| - r_o"""
        if not self.O:
            return np.zeros(0)
        X, _ = self.unpack(z)
        diff = X[None, :2, 1:] - self.cen.transpose(0, 2, 1)
        return (np.sqrt((diff ** 2).sum(1)) - self.rad[:, None]).reshape(-1)

    def jac_d(self, z):
        N, O = self.N, self.O
        J = np.zeros((N * O, self.n))
        if not O:
            return J
        X, _ = self.unpack(z)
        for o in range(O):
            for k in range(1, N + 1):
                e = X[:2, k] - self.cen[o, k - 1]
                r = np.linalg.norm(e)
                J[o * N + k - 1, 3 * k:3 * k + 2] = e / r
        return J

    # -- Hessian of the Lagrangian  f + yc^T c + yd^T d ---------------------
    def hess_lag(self, z, yc, yd=None, obj_factor=1.0):
        N, T = self.N, self.T
        X, U = self.unpack(z)
        H = np.zeros((self.n, self.n))
        for k in range(self.k_lo, self.k_hi + 1):
            for j in range(3):
                H[3 * k + j, 3 * k + j] += obj_factor * 2 * self.W[j]
        for k in range(N):
            iv = 3 * (N + 1) + 2 * k; it = 3 * k + 2
            v, th = U[0, k], X[2, k]
            if self.cfg.cost_mode == "readme":
                a = 1.0 if v < 0 else (0.5 if v == 0 else 0.0); b = 1.0 if v > 0 else (0.5 if v == 0 else 0.0)
                H[iv, iv] += obj_factor * (2 * self.cfg.Wv_neg * a * a + 2 * self.cfg.Wv_pos * b * b)
            H[iv + 1, iv + 1] += obj_factor * 2 * self.cfg.Ww
            lx, ly = yc[3 * (k + 1)], yc[3 * (k + 1) + 1]
            H[it, it] += T * v * (lx * np.cos(th) + ly * np.sin(th))
            hv = T * (lx * np.sin(th) - ly * np.cos(th))
            H[it, iv] += hv; H[iv, it] += hv
        if self.O and yd is not None:
            for o in range(self.O):
                for k in range(1, N + 1):
                    e = X[:2, k] - self.cen[o, k - 1]; r = np.linalg.norm(e); nn = e / r
                    H[3 * k:3 * k + 2, 3 * k:3 * k + 2] += yd[o * N + k - 1] * (np.eye(2) - np.outer(nn, nn)) / r
        return H


def kkt_certificate(nlp: NLP, X, U, inflation=None):
    """Solver-independent first-order certificate for a returned point, computed on the UNSCALED problem.

    Multipliers are recovered by bounded least squares, so only (X, U) are needed:
      minimise |grad f + Jc^T yc + Jd^T yd - zL + zU|  with zL,zU >= 0 supported on (near-)active bounds and
      yd <= 0 supported on (near-)active obstacle rows.
    Returns dict(stationarity, primal, bound_violation, obstacle_violation).
    """
    from scipy.optimize import lsq_linear

    z = nlp.pack(X, U)
    g = nlp.grad(z)
    Jc = nlp.jac_c(z)
    cols = [Jc.T]; lb = [np.full(Jc.shape[0], -np.inf)]; ub = [np.full(Jc.shape[0], np.inf)]
    act_tol = 1e-5
    aL = np.where(np.isfinite(nlp.lo) & (z - nlp.lo <= act_tol))[0]
    aU = np.where(np.isfinite(nlp.hi) & (nlp.hi - z <= act_tol))[0]
    EL = np.zeros((nlp.n, len(aL))); EL[aL, np.arange(len(aL))] = -1.0
    EU = np.zeros((nlp.n, len(aU))); EU[aU, np.arange(len(aU))] = 1.0
    cols += [EL, EU]; lb += [np.zeros(len(aL)), np.zeros(len(aU))]; ub += [np.full(len(aL), np.inf), np.full(len(aU), np.inf)]
    obst_viol = 0.0
    if nlp.O:
        I = nlp.cfg.inflation if inflation is None else inflation
        d = nlp.d(z); Jd = nlp.jac_d(z)
        act = np.where(d - I <= act_tol)[0]
        cols.append(Jd[act].T); lb.append(np.full(len(act), -np.inf)); ub.append(np.zeros(len(act)))
        obst_viol = float(max(0.0, np.max(I - d)))
    A = np.hstack(cols)
    res = lsq_linear(A, -g, bounds=(np.concatenate(lb), np.concatenate(ub)), method="bvls" if A.shape[1] < 400 else "trf",
                     tol=1e-14, max_iter=2000)
    stat = float(np.max(np.abs(A @ res.x + g)))
    return {
        "stationarity": stat,
        "stationarity_rel": stat / max(1.0, float(np.max(np.abs(g)))),
        "primal": float(np.max(np.abs(nlp.c(z)))),
        "bound_violation": float(max(0.0, np.max(nlp.lo - z), np.max(z - nlp.hi))),
        "obstacle_violation": obst_viol,
    }


def slsqp_polish(nlp: NLP, X, U, maxiter=200):
    """SciPy SLSQP started from (X,U) with analytic derivatives.  Returns (X, U, f)."""
    from scipy.optimize import minimize

    z0 = nlp.pack(X, U)
    cons = [{"type": "eq", "fun": nlp.c, "jac": nlp.jac_c}]
    if nlp.O:
        I = nlp.cfg.inflation
        cons.append({"type": "ineq", "fun": lambda z: nlp.d(z) - I, "jac": nlp.jac_d})
    bnds = [(None if not np.isfinite(l) else l, None if not np.isfinite(h) else h) for l, h in zip(nlp.lo, nlp.hi)]
    r = minimize(nlp.f, z0, jac=nlp.grad, bounds=bnds, constraints=cons, method="SLSQP",
                 options={"maxiter": maxiter, "ftol": 1e-15})
    Xp, Up = nlp.unpack(r.x)
    return Xp.copy(), Up.copy(), float(r.fun)


def dual_certificate(nlp: NLP, X, U, yc, zL, zU, df, inflation=None, yd=None, vL=None, relax=1e-8):
    """First-order certificate of a returned primal-dual point on the UNSCALED problem: the oracle's multipliers belong to the
    objective-scaled problem (factor df), so they are divided by df here.  Complementarity is measured against the bounds IPOPT
    works with (relaxed by bound_relax_factor = 1e-8 max(1, |b|)).  Returns dict(stationarity, primal, complementarity,
    bound_violation, dual_sign) -- all infinity norms."""
    z = nlp.pack(X, U)
    lo_r = nlp.lo - relax * np.maximum(1.0, np.abs(np.where(np.isfinite(nlp.lo), nlp.lo, 0.0)))
    hi_r = nlp.hi + relax * np.maximum(1.0, np.abs(np.where(np.isfinite(nlp.hi), nlp.hi, 0.0)))
    r = nlp.grad(z) + (nlp.jac_c(z).T @ yc - zL + zU) / df
    comp = 0.0
    if nlp.O:
        r = r + nlp.jac_d(z).T @ yd / df
        I = nlp.cfg.inflation if inflation is None else inflation
        I_r = I - relax * max(1.0, abs(I))
        comp = float(np.max(np.abs((nlp.d(z) - I_r) * vL))) / df
    fl, fu = np.isfinite(nlp.lo), np.isfinite(nlp.hi)
    comp = max(comp, float(np.max(np.abs((z - lo_r)[fl] * zL[fl]), initial=0.0)) / df, float(np.max(np.abs((hi_r - z)[fu] * zU[fu]), initial=0.0)) / df)
    return {"stationarity": float(np.max(np.abs(r))), "primal": float(np.max(np.abs(nlp.c(z)))), "complementarity": comp,
            "bound_violation": float(max(0.0, np.max(nlp.lo - z), np.max(z - nlp.hi))),
            "dual_sign": float(max(0.0, -min(zL.min(), zU.min())))}


def reduced_hessian_min_eig(nlp: NLP, X, U, yc, df, yd=None, act_tol=1e-6, inflation=None):
    """Second-order sufficiency check: smallest eigenvalue of the Hessian of the (unscaled) Lagrangian projected onto the null
    space of the equality Jacobian and of the active bounds / active obstacle rows.  > 0 at a strict local minimum."""
    z = nlp.pack(X, U)
    H = nlp.hess_lag(z, np.asarray(yc) / df, None if yd is None else np.asarray(yd) / df, 1.0)
    rows = [nlp.jac_c(z)]
    act = np.where((np.isfinite(nlp.lo) & (z - nlp.lo <= act_tol)) | (np.isfinite(nlp.hi) & (nlp.hi - z <= act_tol)))[0]
    E = np.zeros((len(act), nlp.n)); E[np.arange(len(act)), act] = 1.0
    rows.append(E)
    if nlp.O:
        I = nlp.cfg.inflation if inflation is None else inflation
        d = nlp.d(z)
        rows.append(nlp.jac_d(z)[d - I <= act_tol])
    A = np.vstack(rows)
    _, sv, Vt = np.linalg.svd(A, full_matrices=True)
    rank = int((sv > 1e-10 * sv[0]).sum())
    Z = Vt[rank:].T
    if Z.shape[1] == 0:
        return np.inf
    return float(np.linalg.eigvalsh(Z.T @ H @ Z).min())
