// kmpc_resto.cuh -- feasibility restoration phase of the interior-point solver, one CUDA thread per instance.
//
// What it replaces: the part of IPOPT (behind mpc/optimizer.py:354 / :375-391) that takes over when the filter line search of
// the regular algorithm runs below its minimal step size -- MinC_1NrmRestorationPhase (Waechter & Biegler 2006, section 3.3):
//     min  rho (sum n + p) + eta/2 |D_R (w - w_ref)|^2        rho = 1000, eta = sqrt(mu), D_R = diag 1 / max(1, |w_ref|)
//     s.t. c(w) + n_c - p_c = 0,  d(w) - s + n_d - p_d = 0,   w, s within their bounds,  n, p >= 0
// solved by the same interior-point iteration (own filter, barrier parameter from max(mu, |c|_inf, |d - s|_inf)) until an iterate
// reduces the ORIGINAL constraint violation to 0.9 of its value at entry and is acceptable to the original filter and current
// point (back to the regular algorithm), or until it converges on its own problem: Infeasible_Problem_Detected (2).
//
// This is the RARE path (none of the 65,536 instances of the headline batch enters it; infeasible starts and degenerate warm
// starts do), so it is written for clarity, not speed: one thread per instance, all state in the instance's column of a
// workspace in HBM (rows of kmpc_core.cuh's layout + the rows below).  n and p are eliminated from the step system, which leaves
// -(n/z_n + p/z_p) on the diagonal of the constraint blocks; the multiplier step is eliminated next, and what remains is a
// symmetric positive definite block-tridiagonal system in the stage variables (x_k, u_k), factored by a block Cholesky recursion
// (5x5 blocks; positive pivots <=> the inertia IPOPT asks for).
#pragma once
// (included by kmpc_core.cuh in front of phase_trial: everything it uses is defined above that point)

namespace kmpc {

#define K_RESTO_RHO 1000.0        /* resto_penalty_parameter */
#define K_RESTO_KAPPA 0.9         /* required_infeasibility_reduction */
#define K_RESTO_THETA_MAX 1e8     /* resto.theta_max_fact */
#define K_BOUND_MULT_RESET 1000.0 /* bound_mult_reset_threshold */
enum { ST_INFEASIBLE = 2 };
enum { Z_N = 0, Z_P, Z_ZN, Z_ZP, Z_DN, Z_DP, Z_NT, Z_PT, Z_NF };   // per constraint row: n, p, their multipliers, their steps, the trial values
enum { RB_NF = 24 };                                                // per stage: Riccati feedback, cost-to-go, row-block rhs and diagonal (RB_* below)

struct RestoRows { int rRc, rRd, rXref, rBand, total; };
KMPC_HD RestoRows make_resto_rows(const Rows &L) {
    RestoRows R;
    int r = L.total;
    R.rRc = r; r += Z_NF * 3 * (L.N + 1);
    R.rRd = r; r += Z_NF * L.N * L.O;
    R.rXref = r; r += 5 * (L.N + 1);
    R.rBand = r; r += RB_NF * (L.N + 1);
    R.total = r;
    return R;
}

struct RestoPoint { double theta_r, pinf_r, phi_r, theta_o, pinf_o, f_o, bar_o, damp_o; bool ok; };

// One sweep over the stages at the point (state buffer `buf`, n / p fields fn / fp): constraint values of the original problem and of
// the restoration problem, both merit functions.  store_c: keep c + n - p (rows of rCsoc / rDsoc are NOT touched; it goes to the
// step buffer 1's D_Y rows and obstacle dyd rows, which the caller owns at that moment).
template <bool OBS>
KMPC_HDN inline RestoPoint resto_eval(const Cfg &c, const Ctx &t, const RestoRows &RR, double *wsp, size_t S, int buf, int fn, int fp, double mu_r,
                                      double eta, bool store_c) {
    const int N = c.N, O = OBS ? c.O : 0;
    const Rows &L = c.L;
    const double T = c.T, df = t.df;
    const double *sc = wsp + (size_t)L.rSc * S;
    RestoPoint P;
    double theta_r = 0, theta_o = 0, pinf = 0, pinf_r = 0, f = 0, bar = 0, damp = 0, bar_r = 0, damp_r = 0, sumnp = 0, prox = 0;
    bool ok = true;
    double pr0 = FD(sc, 0), pr1 = FD(sc, 1), pr2 = FD(sc, 2);
    for (int k = 0; k <= N; ++k) {
        const double *ps = wsp + ((size_t)L.rState[buf] + (size_t)NSTATE * k) * S;
        const double x[5] = {FD(ps, F_X0), FD(ps, F_X1), FD(ps, F_X2), FD(ps, F_V), FD(ps, F_OM)};
        const double cv[3] = {x[0] - pr0, x[1] - pr1, x[2] - pr2};
        for (int j = 0; j < 3; ++j) {
            double *pz = wsp + ((size_t)RR.rRc + (size_t)Z_NF * (3 * k + j)) * S;
            const double n = FD(pz, fn), p = FD(pz, fp), r = cv[j] + n - p;
            if (!(n > 0) || !(p > 0)) ok = false;
            theta_r += fabs(r); theta_o += fabs(cv[j]); pinf = maxabs_nan(pinf, cv[j]); pinf_r = maxabs_nan(pinf_r, r);
            bar_r += log(n) + log(p); damp_r += n + p; sumnp += n + p;
            if (store_c) FD(wsp + ((size_t)L.rStep[1] + (size_t)NSTEP * k) * S, D_Y0 + j) = r;
        }
        if (k >= c.gk_lo && k <= c.gk_hi) for (int j = 0; j < 3; ++j) { const double e = x[j] - FD(sc, 3 + j); f += c.W[j] * e * e; }
        const int nv = k < N ? 5 : 3;
        for (int j = 0; j < nv; ++j) {
            const double xr = FD(wsp + ((size_t)RR.rXref + 5 * k + j) * S, 0), dr = 1.0 / fmax(1.0, fabs(xr)), e = x[j] - xr;
            prox += dr * dr * e * e;
            const int bi = j < 2 ? j : j - 1;                      // bound index: x, y, (theta: none), v, omega
            if (j == 2) continue;
            if (c.hasL[bi]) { const double sl = x[j] - c.lb[bi]; if (!(sl > 0)) ok = false; bar += log(sl); if (!c.hasU[bi]) damp += sl; }
            if (c.hasU[bi]) { const double su = c.ub[bi] - x[j]; if (!(su > 0)) ok = false; bar += log(su); if (!c.hasL[bi]) damp += su; }
        }
        if (OBS && k >= 1)
            for (int o = 0; o < O; ++o) {
                const int i = (k - 1) * O + o;
                const double *po = wsp + ((size_t)L.rState[buf] + L.sObs + (size_t)3 * i) * S;
                double *pz = wsp + ((size_t)RR.rRd + (size_t)Z_NF * i) * S;
                const double ex = x[0] - FD(sc, CEN_ROW(o, k, 0)), ey = x[1] - FD(sc, CEN_ROW(o, k, 1));
                const double s = FD(po, 0), dm = (sqrt(ex * ex + ey * ey) - FD(sc, RAD_ROW(o))) - s;
                const double n = FD(pz, fn), p = FD(pz, fp), r = dm + n - p, sl = s - c.dL;
                if (!(n > 0) || !(p > 0) || !(sl > 0)) ok = false;
                theta_r += fabs(r); theta_o += fabs(dm); pinf = maxabs_nan(pinf, dm); pinf_r = maxabs_nan(pinf_r, r);
                bar += log(sl); damp += sl;
                bar_r += log(n) + log(p); damp_r += n + p; sumnp += n + p;
                if (store_c) FD(wsp + ((size_t)L.rStep[1] + L.dObs + (size_t)2 * i) * S, 1) = r;
            }
        if (k < N) {
            const double v = x[3], om = x[4];
            if (c.cost_mode == 0) { const double vm = fmin(v, 0.0), vp = fmax(v, 0.0); f += c.Wvn * vm * vm + c.Wvp * vp * vp; }
            else f += c.Wvn * fmin(v, 0.0);
            f += c.Ww * om * om;
            double sn, cs;
            sincos_(x[2], &sn, &cs);
            pr0 = x[0] + T * v * cs; pr1 = x[1] + T * v * sn; pr2 = x[2] + T * om;
        }
    }
    f *= df;
    const double f_r = K_RESTO_RHO * sumnp + 0.5 * eta * prox;
    P.theta_r = theta_r; P.pinf_r = pinf_r; P.phi_r = f_r - mu_r * (bar + bar_r) + K_KAPPA_D * mu_r * (damp + damp_r);
    P.theta_o = theta_o; P.pinf_o = pinf; P.f_o = f; P.bar_o = bar; P.damp_o = damp;
    P.ok = ok && isfinite(P.phi_r) && isfinite(theta_r);
    return P;
}

// Soft transition of the Riccati recursion.  Row block k+1 of the restoration system reads  dx+ - (A dx + B du + e) - D dy = 0  with
// D = n/z_n + p/z_p > 0 (diagonal) where the regular system has D = 0.  With the cost-to-go (P, p) of stage k+1 and dy = -(P dx+ + p):
//     (I + D P) dx+ = z - D p,  z = A dx + B du + e,      cost-to-go seen from z:  Pt = (P^-1 + D)^-1,  pt = (I + P D)^-1 p.
// Everything goes through G = I + D^1/2 P D^1/2 (Cholesky; positive definite <=> the right inertia of this block), so nothing is
// divided by D: as D -> 0 the formulas turn into the hard-constraint recursion, which is what keeps the multiplier steps accurate
// when n, p -> mu / rho (a condensed form J^T D^-1 J loses them: tried first, the dual residual stalled at 1e-7).
struct SoftT { double Li[3][3]; double sd[3]; };   // Cholesky factor of G, sqrt(D)
KMPC_HD bool soft_factor(const double P[3][3], const double D[3], SoftT &F) {
    double G[3][3];
    for (int i = 0; i < 3; ++i) F.sd[i] = sqrt(D[i]);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) G[i][j] = F.sd[i] * P[i][j] * F.sd[j] + (i == j ? 1.0 : 0.0);
    for (int j = 0; j < 3; ++j) {
        double d = G[j][j];
        for (int k = 0; k < j; ++k) d -= F.Li[j][k] * F.Li[j][k];
        if (!(d > 0.0)) return false;
        d = sqrt(d); F.Li[j][j] = d;
        for (int i = j + 1; i < 3; ++i) { double v = G[i][j]; for (int k = 0; k < j; ++k) v -= F.Li[i][k] * F.Li[j][k]; F.Li[i][j] = v / d; }
    }
    return true;
}
// v <- D^1/2 G^-1 D^1/2 v
KMPC_HD void soft_apply(const SoftT &F, double v[3]) {
    double t[3];
    for (int i = 0; i < 3; ++i) t[i] = F.sd[i] * v[i];
    for (int i = 0; i < 3; ++i) { double a = t[i]; for (int k = 0; k < i; ++k) a -= F.Li[i][k] * t[k]; t[i] = a / F.Li[i][i]; }
    for (int i = 2; i >= 0; --i) { double a = t[i]; for (int k = i + 1; k < 3; ++k) a -= F.Li[k][i] * t[k]; t[i] = a / F.Li[i][i]; }
    for (int i = 0; i < 3; ++i) v[i] = F.sd[i] * t[i];
}

// X <- M^-1 X for a general 3x3 matrix and nrhs columns (Gaussian elimination with partial pivoting).  The transition is evaluated
// as (I + D P)^-1 v with this, not through G's Woodbury form I - D^1/2 G^-1 D^1/2 P: when a state is pressed against its bound P is
// stiff (1e8) and the Woodbury form subtracts two nearly equal vectors -- seven digits of dx+ were gone, and with them the multiplier
// step (P dx+ + p cancels to the size of dy).  G's Cholesky factor is still what decides the inertia.
KMPC_HD void solve3(double M[3][3], double *X, int nrhs) {
    for (int c0 = 0; c0 < 3; ++c0) {
        int pr = c0;
        for (int r = c0 + 1; r < 3; ++r) if (fabs(M[r][c0]) > fabs(M[pr][c0])) pr = r;
        if (pr != c0) {
            for (int j = 0; j < 3; ++j) { const double tmp = M[c0][j]; M[c0][j] = M[pr][j]; M[pr][j] = tmp; }
            for (int j = 0; j < nrhs; ++j) { const double tmp = X[c0 * nrhs + j]; X[c0 * nrhs + j] = X[pr * nrhs + j]; X[pr * nrhs + j] = tmp; }
        }
        for (int r = c0 + 1; r < 3; ++r) {
            const double f = M[r][c0] / M[c0][c0];
            for (int j = c0; j < 3; ++j) M[r][j] -= f * M[c0][j];
            for (int j = 0; j < nrhs; ++j) X[r * nrhs + j] -= f * X[c0 * nrhs + j];
        }
    }
    for (int r = 2; r >= 0; --r)
        for (int j = 0; j < nrhs; ++j) {
            double v = X[r * nrhs + j];
            for (int k = r + 1; k < 3; ++k) v -= M[r][k] * X[k * nrhs + j];
            X[r * nrhs + j] = v / M[r][r];
        }
}

// rows of one stage in the band area: feedback K (2x3), feed-forward, cost-to-go P (symmetric, 6), p, rhs e and diagonal D of the row block
// that ENDS in this stage's state
enum { RB_K = 0, RB_KF = 6, RB_P = 8, RB_PV = 14, RB_E = 17, RB_D = 20 };

// quantities of row block k (x_k - pred) at the current iterate: D = n/zn + p/zp, e = b_c = -(c + n - p) + rn n/zn - rp p/zp
template <bool OBS>
KMPC_HDN inline void resto_row_block(const Cfg &c, const RestoRows &RR, double *wsp, size_t S, int k, const double cv[3], const double y[3], double mu, int rhs_c,
                                     double D[3], double e[3]) {
    const double rho = K_RESTO_RHO;
    for (int j = 0; j < 3; ++j) {
        const double *pz = wsp + ((size_t)RR.rRc + (size_t)Z_NF * (3 * k + j)) * S;
        const double n = FD(pz, Z_N), p = FD(pz, Z_P), zn = FD(pz, Z_ZN), zp = FD(pz, Z_ZP);
        const double rn = rho + y[j] - mu / n + K_KAPPA_D * mu, rp = rho - y[j] - mu / p + K_KAPPA_D * mu;
        const double rc = rhs_c ? FD(wsp + ((size_t)c.L.rCsoc + 3 * k + j) * S, 0) : cv[j] + n - p;
        D[j] = n / zn + p / zp;
        e[j] = -rc + rn * n / zn - rp * p / zp;
    }
}

// The Newton step of the restoration problem at the current iterate (buffer t.cur): returns false on wrong inertia.
//   rhs_c: 0 = the constraint values c + n - p of the current point, 1 = the second-order-correction rhs (rCsoc / rDsoc rows)
//   out:   step records of step buffer `sel` (dx, du, dy, obstacle ds / dyd) and the Z_DN / Z_DP fields
template <bool OBS>
KMPC_HDN inline bool resto_step(const Cfg &c, const Ctx &t, const RestoRows &RR, double *wsp, size_t S, double mu, double eta, double delta, int rhs_c,
                                int sel) {
    const int N = c.N, O = OBS ? c.O : 0, cur = t.cur;
    const Rows &L = c.L;
    const double T = c.T, rho = K_RESTO_RHO;
    const double *sc = wsp + (size_t)L.rSc * S;
    // ---- backward sweep ----
    double P[3][3] = {{0}}, pv[3] = {0, 0, 0};
    for (int k = N; k >= 0; --k) {
        const double *ps = wsp + ((size_t)L.rState[cur] + (size_t)NSTATE * k) * S;
        const double x[5] = {FD(ps, F_X0), FD(ps, F_X1), FD(ps, F_X2), FD(ps, F_V), FD(ps, F_OM)};
        const double zl[5] = {FD(ps, F_ZLX), FD(ps, F_ZLY), 0.0, FD(ps, F_ZLV), FD(ps, F_ZLW)}, zu[5] = {FD(ps, F_ZUX), FD(ps, F_ZUY), 0.0, FD(ps, F_ZUV), FD(ps, F_ZUW)};
        const double y[3] = {FD(ps, F_Y0), FD(ps, F_Y1), FD(ps, F_Y2)};
        const int m = k < N ? 5 : 3;
        // stage Hessian (proximity + barrier + delta_w; curvature of the dynamics below) and q = gradient of the barrier Lagrangian
        double H[5][5] = {{0}}, q[5] = {0, 0, 0, 0, 0};
        for (int j = 0; j < m; ++j) {
            const double xr = FD(wsp + ((size_t)RR.rXref + 5 * k + j) * S, 0), dr = 1.0 / fmax(1.0, fabs(xr));
            double g = eta * dr * dr * (x[j] - xr), sig = 0.0;
            const int bi = j < 2 ? j : j - 1;
            if (j != 2) {
                if (c.hasL[bi]) { const double sl = x[j] - c.lb[bi]; sig += zl[j] / sl; g -= mu / sl; if (!c.hasU[bi]) g += K_KAPPA_D * mu; }
                if (c.hasU[bi]) { const double su = c.ub[bi] - x[j]; sig += zu[j] / su; g += mu / su; if (!c.hasL[bi]) g -= K_KAPPA_D * mu; }
            }
            H[j][j] = eta * dr * dr + sig + delta;
            q[j] = g;
        }
        for (int j = 0; j < 3; ++j) q[j] += y[j];   // J^T y, row block k
        if (OBS && k >= 1)
            for (int o = 0; o < O; ++o) {   // obstacle rows: slack, n_d, p_d and the multiplier step condensed into the x-y block
                const int i = (k - 1) * O + o;
                const double *po = wsp + ((size_t)L.rState[cur] + L.sObs + (size_t)3 * i) * S, *pz = wsp + ((size_t)RR.rRd + (size_t)Z_NF * i) * S;
                const double ex = x[0] - FD(sc, CEN_ROW(o, k, 0)), ey = x[1] - FD(sc, CEN_ROW(o, k, 1)), rr = sqrt(ex * ex + ey * ey);
                const double nx = ex / rr, ny = ey / rr, s = FD(po, 0), yd = FD(po, 1), vL = FD(po, 2), sl = s - c.dL;
                const double n = FD(pz, Z_N), p = FD(pz, Z_P), zn = FD(pz, Z_ZN), zp = FD(pz, Z_ZP);
                const double rn = rho + yd - mu / n + K_KAPPA_D * mu, rp = rho - yd - mu / p + K_KAPPA_D * mu;
                const double rd = rhs_c ? FD(wsp + ((size_t)L.rDsoc + i) * S, 0) : ((rr - FD(sc, RAD_ROW(o))) - s) + n - p;
                const double Ds = vL / sl + delta, bs = yd + mu / sl - K_KAPPA_D * mu;        // slack block: Ds ds - dyd = bs
                const double bd = -rd + rn * n / zn - rp * p / zp, Dd = n / zn + p / zp;      // row: n^T dx - ds - Dd dyd = bd
                const double Ed = 1.0 / (1.0 / Ds + Dd), eb = Ed * (bd + bs / Ds);            // dyd = Ed (n^T dx - bd - bs / Ds)
                const double h = yd / rr;
                H[0][0] += h * (1.0 - nx * nx) + Ed * nx * nx; H[1][0] += h * (-nx * ny) + Ed * nx * ny; H[0][1] = H[1][0]; H[1][1] += h * (1.0 - ny * ny) + Ed * ny * ny;
                q[0] -= nx * (eb - yd); q[1] -= ny * (eb - yd);
            }
        double *pb = wsp + ((size_t)RR.rBand + (size_t)RB_NF * k) * S;
        if (k < N) {
            const double *pn = wsp + ((size_t)L.rState[cur] + (size_t)NSTATE * (k + 1)) * S;
            const double yn[3] = {FD(pn, F_Y0), FD(pn, F_Y1), FD(pn, F_Y2)}, xn[3] = {FD(pn, F_X0), FD(pn, F_X1), FD(pn, F_X2)};
            double sn, cs;
            sincos_(x[2], &sn, &cs);
            const double v = x[3];
            const double A[3][3] = {{1, 0, -T * v * sn}, {0, 1, T * v * cs}, {0, 0, 1}}, B[3][2] = {{T * cs, 0}, {T * sn, 0}, {0, T}};
            // curvature of the dynamics in the Lagrangian, and -[A B]^T y of row block k+1
            H[2][2] += T * v * (yn[0] * cs + yn[1] * sn);
            H[3][2] += T * (yn[0] * sn - yn[1] * cs); H[2][3] = H[3][2];
            for (int a = 0; a < 3; ++a) { double acc = 0; for (int j = 0; j < 3; ++j) acc += A[j][a] * yn[j]; q[a] -= acc; }
            for (int a = 0; a < 2; ++a) { double acc = 0; for (int j = 0; j < 3; ++j) acc += B[j][a] * yn[j]; q[3 + a] -= acc; }
            // row block k+1 and the soft transition through it: (P, pv) are stage k+1's
            const double cn[3] = {xn[0] - (x[0] + T * v * cs), xn[1] - (x[1] + T * v * sn), xn[2] - (x[2] + T * x[4])};
            double D[3], e[3];
            resto_row_block<OBS>(c, RR, wsp, S, k + 1, cn, yn, mu, rhs_c, D, e);
            SoftT F;
            if (!soft_factor(P, D, F)) return false;
            double Pt[3][3], pt[3];
            {   // (I + P D) [Pt | pt] = [P | p]
                double Mx[3][3], X[12];
                for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Mx[i][j] = P[i][j] * D[j] + (i == j ? 1.0 : 0.0);
                for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) X[i * 4 + j] = P[i][j]; X[i * 4 + 3] = pv[i]; }
                solve3(Mx, X, 4);
                for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) Pt[i][j] = X[i * 4 + j]; pt[i] = X[i * 4 + 3]; }
            }
            for (int i = 0; i < 3; ++i) for (int j = i + 1; j < 3; ++j) { const double sy = 0.5 * (Pt[i][j] + Pt[j][i]); Pt[i][j] = sy; Pt[j][i] = sy; }
            double Pe[3], PA[3][3], PB[3][2];
            for (int i = 0; i < 3; ++i) {
                Pe[i] = Pt[i][0] * e[0] + Pt[i][1] * e[1] + Pt[i][2] * e[2] + pt[i];
                for (int j = 0; j < 3; ++j) PA[i][j] = Pt[i][0] * A[0][j] + Pt[i][1] * A[1][j] + Pt[i][2] * A[2][j];
                for (int j = 0; j < 2; ++j) PB[i][j] = Pt[i][0] * B[0][j] + Pt[i][1] * B[1][j] + Pt[i][2] * B[2][j];
            }
            double Qxx[3][3], Qux[2][3], Quu[2][2], qx[3], qu[2];
            for (int i = 0; i < 3; ++i) {
                for (int j = 0; j < 3; ++j) Qxx[i][j] = H[i][j] + A[0][i] * PA[0][j] + A[1][i] * PA[1][j] + A[2][i] * PA[2][j];
                qx[i] = q[i] + A[0][i] * Pe[0] + A[1][i] * Pe[1] + A[2][i] * Pe[2];
            }
            for (int i = 0; i < 2; ++i) {
                for (int j = 0; j < 3; ++j) Qux[i][j] = H[3 + i][j] + B[0][i] * PA[0][j] + B[1][i] * PA[1][j] + B[2][i] * PA[2][j];
                for (int j = 0; j < 2; ++j) Quu[i][j] = H[3 + i][3 + j] + B[0][i] * PB[0][j] + B[1][i] * PB[1][j] + B[2][i] * PB[2][j];
                qu[i] = q[3 + i] + B[0][i] * Pe[0] + B[1][i] * Pe[1] + B[2][i] * Pe[2];
            }
            const double qa = Quu[0][0], qb = 0.5 * (Quu[0][1] + Quu[1][0]), qc = Quu[1][1], det = qa * qc - qb * qb;
            if (!(qa > 0.0) || !(det > 0.0)) return false;
            const double i00 = qc / det, i01 = -qb / det, i11 = qa / det;
            double K[2][3], kf[2];
            for (int j = 0; j < 3; ++j) { K[0][j] = -(i00 * Qux[0][j] + i01 * Qux[1][j]); K[1][j] = -(i01 * Qux[0][j] + i11 * Qux[1][j]); }
            kf[0] = -(i00 * qu[0] + i01 * qu[1]); kf[1] = -(i01 * qu[0] + i11 * qu[1]);
            for (int i = 0; i < 3; ++i) {
                for (int j = 0; j < 3; ++j) P[i][j] = Qxx[i][j] + Qux[0][i] * K[0][j] + Qux[1][i] * K[1][j];
                pv[i] = qx[i] + Qux[0][i] * kf[0] + Qux[1][i] * kf[1];
            }
            for (int i = 0; i < 3; ++i) for (int j = i + 1; j < 3; ++j) { const double sy = 0.5 * (P[i][j] + P[j][i]); P[i][j] = sy; P[j][i] = sy; }
            for (int j = 0; j < 3; ++j) { FD(pb, RB_K + j) = K[0][j]; FD(pb, RB_K + 3 + j) = K[1][j]; }
            FD(pb, RB_KF) = kf[0]; FD(pb, RB_KF + 1) = kf[1];
            // e and D of row block k+1 are needed again by the roll-out: kept with stage k+1
            double *pbn = wsp + ((size_t)RR.rBand + (size_t)RB_NF * (k + 1)) * S;
            for (int j = 0; j < 3; ++j) { FD(pbn, RB_E + j) = e[j]; FD(pbn, RB_D + j) = D[j]; }
        } else {
            for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) P[i][j] = H[i][j]; pv[i] = q[i]; }
        }
        FD(pb, RB_P) = P[0][0]; FD(pb, RB_P + 1) = P[1][0]; FD(pb, RB_P + 2) = P[1][1]; FD(pb, RB_P + 3) = P[2][0]; FD(pb, RB_P + 4) = P[2][1]; FD(pb, RB_P + 5) = P[2][2];
        for (int j = 0; j < 3; ++j) FD(pb, RB_PV + j) = pv[j];
        if (k == 0) {   // row block 0: x_0 - x_cur
            const double cv[3] = {x[0] - FD(sc, 0), x[1] - FD(sc, 1), x[2] - FD(sc, 2)};
            double D[3], e[3];
            resto_row_block<OBS>(c, RR, wsp, S, 0, cv, y, mu, rhs_c, D, e);
            SoftT F;
            if (!soft_factor(P, D, F)) return false;
            for (int j = 0; j < 3; ++j) { FD(pb, RB_E + j) = e[j]; FD(pb, RB_D + j) = D[j]; }
        }
    }
    // ---- roll-out: dx+ = (I + D P)^-1 (z - D p), dy = -(P dx+ + p); n / p steps; obstacle slack and multiplier steps ----
    double dx[3] = {0, 0, 0}, z[3];
    for (int k = 0; k <= N; ++k) {
        const double *pb = wsp + ((size_t)RR.rBand + (size_t)RB_NF * k) * S;
        const double *ps = wsp + ((size_t)L.rState[cur] + (size_t)NSTATE * k) * S;
        double *pd = wsp + ((size_t)L.rStep[sel] + (size_t)NSTEP * k) * S;
        const double x[5] = {FD(ps, F_X0), FD(ps, F_X1), FD(ps, F_X2), FD(ps, F_V), FD(ps, F_OM)};
        double Pk[3][3] = {{FD(pb, RB_P), FD(pb, RB_P + 1), FD(pb, RB_P + 3)}, {FD(pb, RB_P + 1), FD(pb, RB_P + 2), FD(pb, RB_P + 4)}, {FD(pb, RB_P + 3), FD(pb, RB_P + 4), FD(pb, RB_P + 5)}};
        const double pk[3] = {FD(pb, RB_PV), FD(pb, RB_PV + 1), FD(pb, RB_PV + 2)}, D[3] = {FD(pb, RB_D), FD(pb, RB_D + 1), FD(pb, RB_D + 2)};
        if (k == 0) for (int j = 0; j < 3; ++j) z[j] = FD(pb, RB_E + j);
        {   // (I + D P) dx+ = z - D p
            double Mx[3][3];
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Mx[i][j] = D[i] * Pk[i][j] + (i == j ? 1.0 : 0.0);
            for (int j = 0; j < 3; ++j) dx[j] = z[j] - D[j] * pk[j];
            solve3(Mx, dx, 1);
        }
        FD(pd, D_X0) = dx[0]; FD(pd, D_X1) = dx[1]; FD(pd, D_X2) = dx[2];
        for (int j = 0; j < 3; ++j) {
            // dy = -(P dx+ + p) (the multiplier is the cost-to-go gradient) = (dx+ - z) / D (the row itself).  The first form is the
            // accurate one for a nearly hard row (D -> 0: the second would divide rounding noise by D), the second for a soft row under
            // a stiff cost-to-go (D |P| > 1: P dx+ and p cancel down to dy there, e.g. D = p^2 / mu ~ 1e10 on a row that stays violated
            // while the state it ends in is pressed against its bound, P ~ z / slack ~ 1e13)
            const double dy = D[j] * fabs(Pk[j][j]) > 1.0 ? (dx[j] - z[j]) / D[j] : -(Pk[j][0] * dx[0] + Pk[j][1] * dx[1] + Pk[j][2] * dx[2] + pk[j]);
            FD(pd, D_Y0 + j) = dy;
            double *pz = wsp + ((size_t)RR.rRc + (size_t)Z_NF * (3 * k + j)) * S;
            const double n = FD(pz, Z_N), p = FD(pz, Z_P), zn = FD(pz, Z_ZN), zp = FD(pz, Z_ZP), y = FD(ps, F_Y0 + j);
            const double rn = rho + y - mu / n + K_KAPPA_D * mu, rp = rho - y - mu / p + K_KAPPA_D * mu;
            FD(pz, Z_DN) = -(rn + dy) * n / zn; FD(pz, Z_DP) = -(rp - dy) * p / zp;
        }
        if (OBS && k >= 1)
            for (int o = 0; o < O; ++o) {
                const int i = (k - 1) * O + o;
                const double *po = wsp + ((size_t)L.rState[cur] + L.sObs + (size_t)3 * i) * S;
                double *pz = wsp + ((size_t)RR.rRd + (size_t)Z_NF * i) * S, *pdo = wsp + ((size_t)L.rStep[sel] + L.dObs + (size_t)2 * i) * S;
                const double ex = x[0] - FD(sc, CEN_ROW(o, k, 0)), ey = x[1] - FD(sc, CEN_ROW(o, k, 1)), rr = sqrt(ex * ex + ey * ey);
                const double nx = ex / rr, ny = ey / rr, s = FD(po, 0), yd = FD(po, 1), vL = FD(po, 2), sl = s - c.dL;
                const double n = FD(pz, Z_N), p = FD(pz, Z_P), zn = FD(pz, Z_ZN), zp = FD(pz, Z_ZP);
                const double rn = rho + yd - mu / n + K_KAPPA_D * mu, rp = rho - yd - mu / p + K_KAPPA_D * mu;
                const double rd = rhs_c ? FD(wsp + ((size_t)L.rDsoc + i) * S, 0) : ((rr - FD(sc, RAD_ROW(o))) - s) + n - p;
                const double Ds = vL / sl + delta, bs = yd + mu / sl - K_KAPPA_D * mu, bd = -rd + rn * n / zn - rp * p / zp, Dd = n / zn + p / zp;
                const double dyd = (nx * dx[0] + ny * dx[1] - bd - bs / Ds) / (1.0 / Ds + Dd), ds = (bs + dyd) / Ds;
                FD(pdo, 0) = ds; FD(pdo, 1) = dyd;
                FD(pz, Z_DN) = -(rn + dyd) * n / zn; FD(pz, Z_DP) = -(rp - dyd) * p / zp;
            }
        if (k < N) {
            const double du0 = FD(pb, RB_K) * dx[0] + FD(pb, RB_K + 1) * dx[1] + FD(pb, RB_K + 2) * dx[2] + FD(pb, RB_KF);
            const double du1 = FD(pb, RB_K + 3) * dx[0] + FD(pb, RB_K + 4) * dx[1] + FD(pb, RB_K + 5) * dx[2] + FD(pb, RB_KF + 1);
            FD(pd, D_U0) = du0; FD(pd, D_U1) = du1;
            const double *pbn = wsp + ((size_t)RR.rBand + (size_t)RB_NF * (k + 1)) * S;
            double sn, cs;
            sincos_(x[2], &sn, &cs);
            const double v = x[3];
            z[0] = dx[0] - T * v * sn * dx[2] + T * cs * du0 + FD(pbn, RB_E);
            z[1] = dx[1] + T * v * cs * dx[2] + T * sn * du0 + FD(pbn, RB_E + 1);
            z[2] = dx[2] + T * du1 + FD(pbn, RB_E + 2);
        } else { FD(pd, D_U0) = 0.0; FD(pd, D_U1) = 0.0; }
    }
    return true;
}

// everything the line search needs from a computed step: primal / dual fraction-to-the-boundary limits and the directional derivative
// of the restoration barrier objective
template <bool OBS>
KMPC_HDN inline void resto_limits(const Cfg &c, const Ctx &t, const RestoRows &RR, double *wsp, size_t S, double mu, double eta, double tau, int sel,
                                  double *apr, double *adu, double *gbd) {
    const int N = c.N, O = OBS ? c.O : 0, cur = t.cur;
    const Rows &L = c.L;
    const double rho = K_RESTO_RHO;
    double a = 1.0, ad = 1.0, g = 0.0;
#define R_BND(val, d, lo, hi, hL, hU, zL, zU) do { \
        if (hL) { const double sl = (val) - (lo); if ((d) < 0) a = fmin(a, -tau * sl / (d)); const double dz = mu / sl - (zL) - (zL) / sl * (d); if (dz < 0) ad = fmin(ad, -tau * (zL) / dz); } \
        if (hU) { const double su = (hi) - (val); if ((d) > 0) a = fmin(a, tau * su / (d)); const double dz = mu / su - (zU) + (zU) / su * (d); if (dz < 0) ad = fmin(ad, -tau * (zU) / dz); } } while (0)
#define R_NP(pz) do { \
        const double n = FD(pz, Z_N), p = FD(pz, Z_P), zn = FD(pz, Z_ZN), zp = FD(pz, Z_ZP), dn = FD(pz, Z_DN), dp = FD(pz, Z_DP); \
        if (dn < 0) a = fmin(a, -tau * n / dn); if (dp < 0) a = fmin(a, -tau * p / dp); \
        const double dzn = mu / n - zn - zn / n * dn, dzp = mu / p - zp - zp / p * dp; \
        if (dzn < 0) ad = fmin(ad, -tau * zn / dzn); if (dzp < 0) ad = fmin(ad, -tau * zp / dzp); \
        g += (rho - mu / n + K_KAPPA_D * mu) * dn + (rho - mu / p + K_KAPPA_D * mu) * dp; } while (0)
    for (int k = 0; k <= N; ++k) {
        const double *ps = wsp + ((size_t)L.rState[cur] + (size_t)NSTATE * k) * S, *pd = wsp + ((size_t)L.rStep[sel] + (size_t)NSTEP * k) * S;
        const double x[5] = {FD(ps, F_X0), FD(ps, F_X1), FD(ps, F_X2), FD(ps, F_V), FD(ps, F_OM)};
        const double d[5] = {FD(pd, D_X0), FD(pd, D_X1), FD(pd, D_X2), FD(pd, D_U0), FD(pd, D_U1)};
        const double zl[5] = {FD(ps, F_ZLX), FD(ps, F_ZLY), 0.0, FD(ps, F_ZLV), FD(ps, F_ZLW)}, zu[5] = {FD(ps, F_ZUX), FD(ps, F_ZUY), 0.0, FD(ps, F_ZUV), FD(ps, F_ZUW)};
        const int m = k < N ? 5 : 3;
        for (int j = 0; j < m; ++j) {
            const double xr = FD(wsp + ((size_t)RR.rXref + 5 * k + j) * S, 0), dr = 1.0 / fmax(1.0, fabs(xr));
            double gp = eta * dr * dr * (x[j] - xr);
            if (j != 2) {
                const int bi = j < 2 ? j : j - 1;
                R_BND(x[j], d[j], c.lb[bi], c.ub[bi], c.hasL[bi], c.hasU[bi], zl[j], zu[j]);
                if (c.hasL[bi]) { gp -= mu / (x[j] - c.lb[bi]); if (!c.hasU[bi]) gp += K_KAPPA_D * mu; }
                if (c.hasU[bi]) { gp += mu / (c.ub[bi] - x[j]); if (!c.hasL[bi]) gp -= K_KAPPA_D * mu; }
            }
            g += gp * d[j];
        }
        for (int j = 0; j < 3; ++j) { double *pz = wsp + ((size_t)RR.rRc + (size_t)Z_NF * (3 * k + j)) * S; R_NP(pz); }
        if (OBS && k >= 1)
            for (int o = 0; o < O; ++o) {
                const int i = (k - 1) * O + o;
                const double *po = wsp + ((size_t)L.rState[cur] + L.sObs + (size_t)3 * i) * S, *pdo = wsp + ((size_t)L.rStep[sel] + L.dObs + (size_t)2 * i) * S;
                const double s = FD(po, 0), vL = FD(po, 2), ds = FD(pdo, 0);
                R_BND(s, ds, c.dL, 0.0, 1, 0, vL, 0.0);
                g += (-mu / (s - c.dL) + K_KAPPA_D * mu) * ds;
                double *pz = wsp + ((size_t)RR.rRd + (size_t)Z_NF * i) * S; R_NP(pz);
            }
    }
#undef R_BND
#undef R_NP
    *apr = a; *adu = ad; *gbd = g;
}

// trial point w + alpha d into the other state buffer / the Z_NT, Z_PT fields
template <bool OBS>
KMPC_HDN inline void resto_trial_point(const Cfg &c, const Ctx &t, const RestoRows &RR, double *wsp, size_t S, double alpha, int sel) {
    const int N = c.N, O = OBS ? c.O : 0, cur = t.cur;
    const Rows &L = c.L;
    for (int k = 0; k <= N; ++k) {
        const double *ps = wsp + ((size_t)L.rState[cur] + (size_t)NSTATE * k) * S, *pd = wsp + ((size_t)L.rStep[sel] + (size_t)NSTEP * k) * S;
        double *pn = wsp + ((size_t)L.rState[cur ^ 1] + (size_t)NSTATE * k) * S;
        FD(pn, F_X0) = FD(ps, F_X0) + alpha * FD(pd, D_X0); FD(pn, F_X1) = FD(ps, F_X1) + alpha * FD(pd, D_X1); FD(pn, F_X2) = FD(ps, F_X2) + alpha * FD(pd, D_X2);
        FD(pn, F_V) = FD(ps, F_V) + alpha * FD(pd, D_U0); FD(pn, F_OM) = FD(ps, F_OM) + alpha * FD(pd, D_U1);
        for (int j = 0; j < 3; ++j) { double *pz = wsp + ((size_t)RR.rRc + (size_t)Z_NF * (3 * k + j)) * S; FD(pz, Z_NT) = FD(pz, Z_N) + alpha * FD(pz, Z_DN); FD(pz, Z_PT) = FD(pz, Z_P) + alpha * FD(pz, Z_DP); }
        if (OBS && k >= 1)
            for (int o = 0; o < O; ++o) {
                const int i = (k - 1) * O + o;
                FD(wsp + ((size_t)L.rState[cur ^ 1] + L.sObs + (size_t)3 * i) * S, 0) =
                    FD(wsp + ((size_t)L.rState[cur] + L.sObs + (size_t)3 * i) * S, 0) + alpha * FD(wsp + ((size_t)L.rStep[sel] + L.dObs + (size_t)2 * i) * S, 0);
                double *pz = wsp + ((size_t)RR.rRd + (size_t)Z_NF * i) * S; FD(pz, Z_NT) = FD(pz, Z_N) + alpha * FD(pz, Z_DN); FD(pz, Z_PT) = FD(pz, Z_P) + alpha * FD(pz, Z_DP);
            }
    }
}

KMPC_HD bool resto_filter_ok(const double *filt, size_t FS, int fn, double theta, double phi) {
    for (int i = 0; i < fn; ++i)
        if (!(theta <= filt[(size_t)(2 * i) * FS] || phi <= filt[(size_t)(2 * i + 1) * FS])) return false;
    return true;
}

// The restoration phase proper.  In: the iterate of buffer t.cur at which the regular line search failed, t.c its statistics, the
// original filter (already augmented by the caller).  Returns 0 when it hands a point back to the regular algorithm -- the state
// buffer t.cur holds it (y = 0, bound multipliers as IPOPT resets them), t.c its statistics, t.iter counts the restoration
// iterations -- or a final status.  It needs a private filter: the last 2 K_FILTER_CAP band rows of stage 0 .. are not used for
// that; a small local array is (the restoration filter rarely holds more than a few entries).
template <bool OBS>
KMPC_HDN inline int resto_phase(const Cfg &c, Ctx &t, const RestoRows &RR, double *wsp, size_t S) {
    const int N = c.N, O = OBS ? c.O : 0;
    const Rows &L = c.L;
    const double rho = K_RESTO_RHO, mu_orig = t.mu;
    const double theta_ref = t.c.theta, phi_ref = phi_of(t.c, t.mu);
    const double *filt_o = wsp + (size_t)L.rFilt * S;
    const double *sc = wsp + (size_t)L.rSc * S;
    const double eta = sqrt(mu_orig);
    // ---- RestoIterateInitializer ----
    double cmax = 0.0;
    {
        double pr0 = FD(sc, 0), pr1 = FD(sc, 1), pr2 = FD(sc, 2);
        for (int k = 0; k <= N; ++k) {
            const double *ps = wsp + ((size_t)L.rState[t.cur] + (size_t)NSTATE * k) * S;
            const double x[5] = {FD(ps, F_X0), FD(ps, F_X1), FD(ps, F_X2), FD(ps, F_V), FD(ps, F_OM)};
            for (int j = 0; j < 5; ++j) FD(wsp + ((size_t)RR.rXref + 5 * k + j) * S, 0) = x[j];
            const double cv[3] = {x[0] - pr0, x[1] - pr1, x[2] - pr2};
            for (int j = 0; j < 3; ++j) { cmax = maxabs_nan(cmax, cv[j]); FD(wsp + ((size_t)RR.rRc + (size_t)Z_NF * (3 * k + j)) * S, Z_NT) = cv[j]; }
            if (OBS && k >= 1)
                for (int o = 0; o < O; ++o) {
                    const int i = (k - 1) * O + o;
                    const double ex = x[0] - FD(sc, CEN_ROW(o, k, 0)), ey = x[1] - FD(sc, CEN_ROW(o, k, 1));
                    const double dm = (sqrt(ex * ex + ey * ey) - FD(sc, RAD_ROW(o))) - FD(wsp + ((size_t)L.rState[t.cur] + L.sObs + (size_t)3 * i) * S, 0);
                    cmax = maxabs_nan(cmax, dm); FD(wsp + ((size_t)RR.rRd + (size_t)Z_NF * i) * S, Z_NT) = dm;
                }
            if (k < N) { double sn, cs; sincos_(x[2], &sn, &cs); pr0 = x[0] + c.T * x[3] * cs; pr1 = x[1] + c.T * x[3] * sn; pr2 = x[2] + c.T * x[4]; }
        }
    }
    double mu = fmax(mu_orig, cmax), tau = fmax(K_TAU_MIN, 1.0 - mu);
    const int nrows_c = 3 * (N + 1), nrows_d = N * O;
    for (int i = 0; i < nrows_c + nrows_d; ++i) {
        double *pz = i < nrows_c ? wsp + ((size_t)RR.rRc + (size_t)Z_NF * i) * S : wsp + ((size_t)RR.rRd + (size_t)Z_NF * (i - nrows_c)) * S;
        const double cv = FD(pz, Z_NT), a = (mu - rho * cv) / (2.0 * rho), n = a + sqrt(a * a + mu * cv / (2.0 * rho)), p = cv + n;
        FD(pz, Z_N) = n; FD(pz, Z_P) = p; FD(pz, Z_ZN) = mu / n; FD(pz, Z_ZP) = mu / p;
    }
    for (int k = 0; k <= N; ++k) {   // bound multipliers capped at rho, equality multipliers zero
        double *ps = wsp + ((size_t)L.rState[t.cur] + (size_t)NSTATE * k) * S;
        for (int f = F_ZLX; f <= F_ZUW; ++f) FD(ps, f) = fmin(rho, FD(ps, f));
        FD(ps, F_Y0) = 0.0; FD(ps, F_Y1) = 0.0; FD(ps, F_Y2) = 0.0;
    }
    for (int i = 0; i < nrows_d; ++i) { double *po = wsp + ((size_t)L.rState[t.cur] + L.sObs + (size_t)3 * i) * S; FD(po, 1) = 0.0; FD(po, 2) = fmin(rho, FD(po, 2)); }

    double filt[2 * K_FILTER_CAP];
    int fn = 0;
    double theta_max = -1.0, theta_min = -1.0, delta_last = 0.0;
    bool first = true;
    int st = -100;
    for (;;) {
        // ---- optimality error of the restoration problem at the current iterate (dual residuals need the Jacobian transpose) ----
        RestoPoint cp = resto_eval<OBS>(c, t, RR, wsp, S, t.cur, Z_N, Z_P, mu, eta, false);
        double dinf = 0, mn = INFINITY, mx = 0, sumz = 0, sumy = 0, wmax = 0;
        int nb = 0;
        {
#define R_CP(slack, z) do { const double p_ = (slack) * (z); mn = fmin(mn, p_); mx = fmax(mx, p_); sumz += fabs(z); nb++; } while (0)
            for (int k = 0; k <= N; ++k) {
                const double *ps = wsp + ((size_t)L.rState[t.cur] + (size_t)NSTATE * k) * S;
                const double x[5] = {FD(ps, F_X0), FD(ps, F_X1), FD(ps, F_X2), FD(ps, F_V), FD(ps, F_OM)};
                const double zl[5] = {FD(ps, F_ZLX), FD(ps, F_ZLY), 0.0, FD(ps, F_ZLV), FD(ps, F_ZLW)}, zu[5] = {FD(ps, F_ZUX), FD(ps, F_ZUY), 0.0, FD(ps, F_ZUV), FD(ps, F_ZUW)};
                const double y[3] = {FD(ps, F_Y0), FD(ps, F_Y1), FD(ps, F_Y2)};
                const int m = k < N ? 5 : 3;
                double r[5];
                for (int j = 0; j < m; ++j) {
                    const double xr = FD(wsp + ((size_t)RR.rXref + 5 * k + j) * S, 0), dr = 1.0 / fmax(1.0, fabs(xr));
                    r[j] = eta * dr * dr * (x[j] - xr);
                    wmax = fmax(wmax, fabs(x[j]));
                    if (j != 2) {
                        const int bi = j < 2 ? j : j - 1;
                        if (c.hasL[bi]) { r[j] -= zl[j]; R_CP(x[j] - c.lb[bi], zl[j]); }
                        if (c.hasU[bi]) { r[j] += zu[j]; R_CP(c.ub[bi] - x[j], zu[j]); }
                    }
                }
                for (int j = 0; j < 3; ++j) { r[j] += y[j]; sumy += fabs(y[j]); }
                if (k < N) {
                    const double *pn = wsp + ((size_t)L.rState[t.cur] + (size_t)NSTATE * (k + 1)) * S;
                    const double yn[3] = {FD(pn, F_Y0), FD(pn, F_Y1), FD(pn, F_Y2)};
                    double sn, cs;
                    sincos_(x[2], &sn, &cs);
                    const double v = x[3], T = c.T;
                    r[0] -= yn[0]; r[1] -= yn[1]; r[2] -= (-T * v * sn) * yn[0] + (T * v * cs) * yn[1] + yn[2];
                    r[3] -= T * cs * yn[0] + T * sn * yn[1]; r[4] -= T * yn[2];
                }
                if (OBS && k >= 1)
                    for (int o = 0; o < O; ++o) {
                        const int i = (k - 1) * O + o;
                        const double *po = wsp + ((size_t)L.rState[t.cur] + L.sObs + (size_t)3 * i) * S;
                        const double *pz = wsp + ((size_t)RR.rRd + (size_t)Z_NF * i) * S;
                        const double ex = x[0] - FD(sc, CEN_ROW(o, k, 0)), ey = x[1] - FD(sc, CEN_ROW(o, k, 1)), rr = sqrt(ex * ex + ey * ey);
                        const double yd = FD(po, 1), vL = FD(po, 2);
                        r[0] += ex / rr * yd; r[1] += ey / rr * yd;
                        dinf = maxabs_nan(dinf, -yd - vL); sumy += fabs(yd);
                        R_CP(FD(po, 0) - c.dL, vL);
                        dinf = maxabs_nan(maxabs_nan(dinf, rho + yd - FD(pz, Z_ZN)), rho - yd - FD(pz, Z_ZP));
                        R_CP(FD(pz, Z_N), FD(pz, Z_ZN)); R_CP(FD(pz, Z_P), FD(pz, Z_ZP));
                    }
                for (int j = 0; j < m; ++j) dinf = maxabs_nan(dinf, r[j]);
                for (int j = 0; j < 3; ++j) {
                    const double *pz = wsp + ((size_t)RR.rRc + (size_t)Z_NF * (3 * k + j)) * S;
                    dinf = maxabs_nan(maxabs_nan(dinf, rho + y[j] - FD(pz, Z_ZN)), rho - y[j] - FD(pz, Z_ZP));
                    R_CP(FD(pz, Z_N), FD(pz, Z_ZN)); R_CP(FD(pz, Z_P), FD(pz, Z_ZP));
                }
            }
#undef R_CP
        }
        const double pinf_r = cp.pinf_r;
        const double sd = fmax(K_S_MAX, (sumy + sumz) / (double)(nrows_c + nrows_d + nb)) / K_S_MAX, scl = fmax(K_S_MAX, sumz / (double)nb) / K_S_MAX;
#define R_COMPL(m_) fmax(fabs(mx - (m_)), fabs(mn - (m_)))
#define R_ERR(m_) fmax(dinf / sd, fmax(pinf_r, R_COMPL(m_) / scl))
        const double E0 = R_ERR(0.0);
        if (!cp.ok || !isfinite(E0) || !isfinite(dinf)) { st = ST_INVALID; break; }
        // ---- RestoFilterConvergenceCheck (not at the starting point) ----
        if (!first) {
            if (cp.theta_o <= K_RESTO_KAPPA * theta_ref) {
                const double phi_o = cp.f_o - mu_orig * cp.bar_o + K_KAPPA_D * mu_orig * cp.damp_o;
                if (isfinite(phi_o) && resto_filter_ok(filt_o, S, t.fn, cp.theta_o, phi_o) &&
                    (cmp_le(cp.theta_o, (1.0 - K_GAMMA_THETA) * theta_ref, theta_ref) || cmp_le(phi_o - phi_ref, -K_GAMMA_PHI * theta_ref, phi_ref))) { st = 0; break; }
            }
            if (E0 <= c.tol && dinf <= K_DUAL_INF_TOL && pinf_r <= K_CONSTR_VIOL_TOL && R_COMPL(0.0) <= K_COMPL_INF_TOL) {
                st = cp.pinf_o <= 1e2 * c.tol ? (int)ST_RESTORATION : (int)ST_INFEASIBLE;
                break;
            }
        }
        first = false;
        if (t.iter >= c.max_iter) { st = ST_MAXITER; break; }
        if (wmax > K_DIVERGING) { st = ST_DIVERGING; break; }
        {   // monotone barrier update
            bool done = false;
            while (!done && R_ERR(mu) <= K_KAPPA_EPS * mu) {
                const double nm = fmax(fmin(K_MU_LIN * mu, mu * sqrt(mu)), fmin(c.tol, K_COMPL_INF_TOL) / (K_KAPPA_EPS + 1.0));
                const bool changed = nm != mu;
                mu = nm; tau = fmax(K_TAU_MIN, 1.0 - mu);
                if (changed) fn = 0; else done = true;
            }
        }
#undef R_ERR
#undef R_COMPL
        // ---- step with inertia correction ----
        double delta = 0.0;
        bool ok = false;
        for (;;) {
            ok = resto_step<OBS>(c, t, RR, wsp, S, mu, eta, delta, 0, 0);
            if (ok) break;
            delta = inertia_next_delta(delta, delta_last);
            if (delta > K_DW_MAX) break;
        }
        if (!ok) { st = ST_STEP_ERROR; break; }
        if (delta > 0.0) delta_last = delta;
        // ---- filter line search with second-order corrections ----
        cp = resto_eval<OBS>(c, t, RR, wsp, S, t.cur, Z_N, Z_P, mu, eta, false);   // (phi with the barrier parameter of this iteration)
        double amax, adu, gBD;
        resto_limits<OBS>(c, t, RR, wsp, S, mu, eta, tau, 0, &amax, &adu, &gBD);
        if (theta_max < 0) { theta_max = K_RESTO_THETA_MAX * fmax(1.0, cp.theta_r); theta_min = K_THETA_MIN_FACT * fmax(1.0, cp.theta_r); }
        double alpha_min = K_GAMMA_THETA;
        if (gBD < 0) {
            alpha_min = fmin(K_GAMMA_THETA, K_GAMMA_PHI * cp.theta_r / (-gBD));
            if (cp.theta_r <= theta_min) alpha_min = fmin(alpha_min, K_DELTA_LS * pow(cp.theta_r, K_S_THETA) / pow(-gBD, K_S_PHI));
        }
        alpha_min *= K_ALPHA_MIN_FRAC;
        double alpha = amax, alpha_test = amax, adu_acc = adu;
        int sel = 0, nsteps = 0;
        bool accept = false;
        RestoPoint tp;
        auto acceptable = [&](const RestoPoint &q, double at) {
            if (q.theta_r > theta_max) return false;
            bool acc;
            const bool ftype = gBD < 0 && at * pow(-gBD, K_S_PHI) > K_DELTA_LS * pow(cp.theta_r, K_S_THETA);
            if (ftype && cp.theta_r <= theta_min) acc = cmp_le(q.phi_r - cp.phi_r, K_ETA_PHI * at * gBD, cp.phi_r);
            else {
                acc = true;
                if (q.phi_r > cp.phi_r) { const double bas = fabs(cp.phi_r) > 10.0 ? log10(fabs(cp.phi_r)) : 1.0; if (log10(q.phi_r - cp.phi_r) > K_OBJ_MAX_INC + bas) acc = false; }
                if (acc) acc = cmp_le(q.theta_r, (1.0 - K_GAMMA_THETA) * cp.theta_r, cp.theta_r) || cmp_le(q.phi_r - cp.phi_r, -K_GAMMA_PHI * cp.theta_r, cp.phi_r);
            }
            if (acc) acc = resto_filter_ok(filt, 1, fn, q.theta_r, q.phi_r);
            return acc;
        };
        while (alpha > alpha_min || nsteps == 0) {
            resto_trial_point<OBS>(c, t, RR, wsp, S, alpha, 0);
            tp = resto_eval<OBS>(c, t, RR, wsp, S, t.cur ^ 1, Z_NT, Z_PT, mu, eta, true);
            alpha_test = alpha;
            if (tp.ok) accept = acceptable(tp, alpha_test);
            if (accept) break;
            if (tp.ok && alpha == amax && cp.theta_r <= tp.theta_r) {   // second-order correction
                double theta_soc_old = 0, theta_trial = tp.theta_r, alpha_soc = alpha;
                // c_soc starts from the constraint values of the current point
                resto_eval<OBS>(c, t, RR, wsp, S, t.cur, Z_N, Z_P, mu, eta, true);
                for (int i = 0; i < nrows_c; ++i) FD(wsp + ((size_t)L.rCsoc + i) * S, 0) = FD(wsp + ((size_t)L.rStep[1] + (size_t)NSTEP * (i / 3)) * S, D_Y0 + i % 3);
                for (int i = 0; i < nrows_d; ++i) FD(wsp + ((size_t)L.rDsoc + i) * S, 0) = FD(wsp + ((size_t)L.rStep[1] + L.dObs + (size_t)2 * i) * S, 1);
                resto_eval<OBS>(c, t, RR, wsp, S, t.cur ^ 1, Z_NT, Z_PT, mu, eta, true);   // constraint values of the trial point again
                int count = 0;
                while (count < K_MAX_SOC && !accept && (count == 0 || theta_trial <= K_KAPPA_SOC * theta_soc_old)) {
                    theta_soc_old = theta_trial;
                    for (int i = 0; i < nrows_c; ++i) { double *q = wsp + ((size_t)L.rCsoc + i) * S; FD(q, 0) = alpha_soc * FD(q, 0) + FD(wsp + ((size_t)L.rStep[1] + (size_t)NSTEP * (i / 3)) * S, D_Y0 + i % 3); }
                    for (int i = 0; i < nrows_d; ++i) { double *q = wsp + ((size_t)L.rDsoc + i) * S; FD(q, 0) = alpha_soc * FD(q, 0) + FD(wsp + ((size_t)L.rStep[1] + L.dObs + (size_t)2 * i) * S, 1); }
                    if (!resto_step<OBS>(c, t, RR, wsp, S, mu, eta, delta, 1, 1)) break;    // (same matrix: cannot fail; the step lands in buffer 1)
                    double a2, ad2, g2;
                    resto_limits<OBS>(c, t, RR, wsp, S, mu, eta, tau, 1, &a2, &ad2, &g2);
                    alpha_soc = a2;
                    resto_trial_point<OBS>(c, t, RR, wsp, S, alpha_soc, 1);
                    // (the corrected step sits in step buffer 1, whose D_Y rows the evaluation below overwrites with constraint values:
                    //  keep dy of the corrected step in the rows of step buffer 0 first -- the original step is recomputed if needed)
                    for (int k = 0; k <= N; ++k) for (int j = 0; j < 3; ++j) FD(wsp + ((size_t)L.rStep[0] + (size_t)NSTEP * k) * S, D_Y0 + j) = FD(wsp + ((size_t)L.rStep[1] + (size_t)NSTEP * k) * S, D_Y0 + j);
                    for (int i = 0; i < nrows_d; ++i) FD(wsp + ((size_t)L.rStep[0] + L.dObs + (size_t)2 * i) * S, 1) = FD(wsp + ((size_t)L.rStep[1] + L.dObs + (size_t)2 * i) * S, 1);
                    const RestoPoint ts = resto_eval<OBS>(c, t, RR, wsp, S, t.cur ^ 1, Z_NT, Z_PT, mu, eta, true);
                    if (!ts.ok) break;
                    if (acceptable(ts, alpha_test)) { accept = true; tp = ts; alpha = alpha_soc; adu_acc = ad2; sel = 1; }
                    else { count++; theta_trial = ts.theta_r; }
                }
                if (accept) break;
                resto_step<OBS>(c, t, RR, wsp, S, mu, eta, delta, 0, 0);   // back to the original step (n / p steps and dy were overwritten)
            }
            alpha *= K_ALPHA_RED; nsteps++;
        }
        if (!accept) { st = ST_RESTORATION; break; }
        {
            const bool ftype = gBD < 0 && alpha_test * pow(-gBD, K_S_PHI) > K_DELTA_LS * pow(cp.theta_r, K_S_THETA);
            if (!ftype || !cmp_le(tp.phi_r - cp.phi_r, K_ETA_PHI * alpha_test * gBD, cp.phi_r)) {
                const double th = (1.0 - K_GAMMA_THETA) * cp.theta_r, ph = cp.phi_r - K_GAMMA_PHI * cp.theta_r;
                int m = 0;
                for (int i = 0; i < fn; ++i) if (!(filt[2 * i] >= th && filt[2 * i + 1] >= ph)) { filt[2 * m] = filt[2 * i]; filt[2 * m + 1] = filt[2 * i + 1]; ++m; }
                fn = m;
                if (fn < K_FILTER_CAP) { filt[2 * fn] = th; filt[2 * fn + 1] = ph; fn++; }
            }
        }
        // ---- accept: the trial buffer becomes current; multipliers: y with alpha, every bound multiplier with alpha_dual + kappa_sigma clamp ----
        // (sel == 1: x, u, ds come from step buffer 1, dy / dyd of the corrected step were parked in step buffer 0)
        for (int k = 0; k <= N; ++k) {
            const double *ps = wsp + ((size_t)L.rState[t.cur] + (size_t)NSTATE * k) * S, *pd = wsp + ((size_t)L.rStep[sel] + (size_t)NSTEP * k) * S;
            const double *pdy = wsp + ((size_t)L.rStep[0] + (size_t)NSTEP * k) * S;
            double *pn = wsp + ((size_t)L.rState[t.cur ^ 1] + (size_t)NSTATE * k) * S;
            const double xo[5] = {FD(ps, F_X0), FD(ps, F_X1), FD(ps, F_X2), FD(ps, F_V), FD(ps, F_OM)};
            const double xn[5] = {FD(pn, F_X0), FD(pn, F_X1), FD(pn, F_X2), FD(pn, F_V), FD(pn, F_OM)};
            const double d[5] = {FD(pd, D_X0), FD(pd, D_X1), FD(pd, D_X2), FD(pd, D_U0), FD(pd, D_U1)};
            for (int j = 0; j < 3; ++j) FD(pn, F_Y0 + j) = FD(ps, F_Y0 + j) + alpha * FD(pdy, D_Y0 + j);
            const int zf[5][2] = {{F_ZLX, F_ZUX}, {F_ZLY, F_ZUY}, {-1, -1}, {F_ZLV, F_ZUV}, {F_ZLW, F_ZUW}};
            for (int j = 0; j < 5; ++j) {
                if (j == 2) continue;
                const int bi = j < 2 ? j : j - 1;
                const bool has = j < 3 || k < N;
                double zl = 0.0, zu = 0.0;
                if (has && c.hasL[bi]) { const double sl = xo[j] - c.lb[bi], z = FD(ps, zf[j][0]), zn = z + adu_acc * (mu / sl - z - z / sl * d[j]), sn = xn[j] - c.lb[bi]; zl = fmax(fmin(zn, K_KAPPA_SIGMA * mu / sn), mu / (K_KAPPA_SIGMA * sn)); }
                if (has && c.hasU[bi]) { const double su = c.ub[bi] - xo[j], z = FD(ps, zf[j][1]), zn = z + adu_acc * (mu / su - z + z / su * d[j]), sn = c.ub[bi] - xn[j]; zu = fmax(fmin(zn, K_KAPPA_SIGMA * mu / sn), mu / (K_KAPPA_SIGMA * sn)); }
                FD(pn, zf[j][0]) = zl; FD(pn, zf[j][1]) = zu;
            }
            double sn = 0.0, cs = 1.0;
            if (k < N) sincos_(xn[2], &sn, &cs);
            FD(pn, F_CS) = cs; FD(pn, F_SN) = sn;
            for (int j = 0; j < 3; ++j) {
                double *pz = wsp + ((size_t)RR.rRc + (size_t)Z_NF * (3 * k + j)) * S;
                const double n = FD(pz, Z_N), p = FD(pz, Z_P), zn = FD(pz, Z_ZN), zp = FD(pz, Z_ZP), nn = FD(pz, Z_NT), pp = FD(pz, Z_PT);
                const double a1 = zn + adu_acc * (mu / n - zn - zn / n * FD(pz, Z_DN)), a2 = zp + adu_acc * (mu / p - zp - zp / p * FD(pz, Z_DP));
                FD(pz, Z_ZN) = fmax(fmin(a1, K_KAPPA_SIGMA * mu / nn), mu / (K_KAPPA_SIGMA * nn)); FD(pz, Z_ZP) = fmax(fmin(a2, K_KAPPA_SIGMA * mu / pp), mu / (K_KAPPA_SIGMA * pp));
                FD(pz, Z_N) = nn; FD(pz, Z_P) = pp;
            }
            if (OBS && k >= 1)
                for (int o = 0; o < O; ++o) {
                    const int i = (k - 1) * O + o;
                    const double *po = wsp + ((size_t)L.rState[t.cur] + L.sObs + (size_t)3 * i) * S;
                    double *pno = wsp + ((size_t)L.rState[t.cur ^ 1] + L.sObs + (size_t)3 * i) * S, *pz = wsp + ((size_t)RR.rRd + (size_t)Z_NF * i) * S;
                    const double so = FD(po, 0), vL = FD(po, 2), ds = FD(wsp + ((size_t)L.rStep[sel] + L.dObs + (size_t)2 * i) * S, 0), dyd = FD(wsp + ((size_t)L.rStep[0] + L.dObs + (size_t)2 * i) * S, 1);
                    const double slo = so - c.dL, sln = FD(pno, 0) - c.dL, zv = vL + adu_acc * (mu / slo - vL - vL / slo * ds);
                    FD(pno, 1) = FD(po, 1) + alpha * dyd;
                    FD(pno, 2) = fmax(fmin(zv, K_KAPPA_SIGMA * mu / sln), mu / (K_KAPPA_SIGMA * sln));
                    const double n = FD(pz, Z_N), p = FD(pz, Z_P), zn = FD(pz, Z_ZN), zp = FD(pz, Z_ZP), nn = FD(pz, Z_NT), pp = FD(pz, Z_PT);
                    const double a1 = zn + adu_acc * (mu / n - zn - zn / n * FD(pz, Z_DN)), a2 = zp + adu_acc * (mu / p - zp - zp / p * FD(pz, Z_DP));
                    FD(pz, Z_ZN) = fmax(fmin(a1, K_KAPPA_SIGMA * mu / nn), mu / (K_KAPPA_SIGMA * nn)); FD(pz, Z_ZP) = fmax(fmin(a2, K_KAPPA_SIGMA * mu / pp), mu / (K_KAPPA_SIGMA * pp));
                    FD(pz, Z_N) = nn; FD(pz, Z_P) = pp;
                }
        }
        t.cur ^= 1;
        t.iter++;
    }
    if (st == 0) {
        // back to the regular algorithm: bound multipliers kept unless one exceeds bound_mult_reset_threshold (then all 1); equality multipliers zero
        double zm = 0.0;
        for (int k = 0; k <= N; ++k) { const double *ps = wsp + ((size_t)L.rState[t.cur] + (size_t)NSTATE * k) * S; for (int f = F_ZLX; f <= F_ZUW; ++f) zm = fmax(zm, FD(ps, f)); }
        for (int i = 0; i < nrows_d; ++i) zm = fmax(zm, FD(wsp + ((size_t)L.rState[t.cur] + L.sObs + (size_t)3 * i) * S, 2));
        const bool reset = zm > K_BOUND_MULT_RESET;
        for (int k = 0; k <= N; ++k) {
            double *ps = wsp + ((size_t)L.rState[t.cur] + (size_t)NSTATE * k) * S;
            FD(ps, F_Y0) = 0.0; FD(ps, F_Y1) = 0.0; FD(ps, F_Y2) = 0.0;
            if (reset) {
                FD(ps, F_ZLX) = c.hasL[0] ? 1.0 : 0.0; FD(ps, F_ZUX) = c.hasU[0] ? 1.0 : 0.0; FD(ps, F_ZLY) = c.hasL[1] ? 1.0 : 0.0; FD(ps, F_ZUY) = c.hasU[1] ? 1.0 : 0.0;
                FD(ps, F_ZLV) = (k < N && c.hasL[2]) ? 1.0 : 0.0; FD(ps, F_ZUV) = (k < N && c.hasU[2]) ? 1.0 : 0.0;
                FD(ps, F_ZLW) = (k < N && c.hasL[3]) ? 1.0 : 0.0; FD(ps, F_ZUW) = (k < N && c.hasU[3]) ? 1.0 : 0.0;
            }
        }
        for (int i = 0; i < nrows_d; ++i) { double *po = wsp + ((size_t)L.rState[t.cur] + L.sObs + (size_t)3 * i) * S; FD(po, 1) = 0.0; if (reset) FD(po, 2) = 1.0; }
    }
    return st;
}

// The regular algorithm's line search failed at the current iterate (trial_decide returned ST_RESTORATION): what IPOPT's
// BacktrackingLineSearch does next.  Returns 100 to continue with a Newton step at the restored point, else a final status.
template <bool OBS>
KMPC_HDN inline int resto_enter(const Cfg &c, Ctx &t, const RestoRows &RR, double *wsp, size_t S) {
    if (t.c.theta <= 1e-2 * c.tol) return ST_RESTORATION;   // "Restoration phase called, but point is almost feasible": Restoration_Failed
    double *filt = wsp + (size_t)c.L.rFilt * S;
    if (!filter_add(t, filt, S, (1.0 - K_GAMMA_THETA) * t.c.theta, phi_of(t.c, t.mu) - K_GAMMA_PHI * t.c.theta)) return ST_INTERNAL;   // PrepareRestoPhaseStart
    const int rs = resto_phase<OBS>(c, t, RR, wsp, S);
    if (rs != 0) return rs;
    // statistics of the restored point with its new multipliers: a zero step evaluated by the regular trial pass into the other buffer
    for (int k = 0; k <= c.N; ++k) for (int f = 0; f < NSTEP; ++f) FD(wsp + ((size_t)c.L.rStep[0] + (size_t)NSTEP * k) * S, f) = 0.0;
    for (int i = 0; i < 2 * c.N * c.O; ++i) FD(wsp + ((size_t)c.L.rStep[0] + c.L.dObs + i) * S, 0) = 0.0;
    Stats st;
    pass_trial<OBS>(c, t, wsp, S, 0, TU_INIT, 0.0, 0.0, 0.0, &st);
    t.c = st; t.cur ^= 1;
    t.iter++;
    return begin_iteration(c, t);
}

}  // namespace kmpc
