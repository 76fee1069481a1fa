// kmpc_warp.cuh -- warp-per-instance form of the interior-point solver with a block-cooperative Riccati phase.
//
// One warp owns one problem instance for the whole solve: stage s lives in lane s / SPL, slot s % SPL, the iterate
// (primal, duals) stays in registers, the rest of the per-instance state in shared memory; nothing but the problem data
// and the result touches HBM.  A trip of the solver state machine (kmpc_core.cuh) is cut into block-synchronous phases:
//   1a  ASSEMBLE  (owner warps, stage-parallel)   KKT stage blocks of every instance that needs a factorisation
//                                                 -> shared "coop" area  coop[instance][field][stage]
//   1b  RICCATI   (warp 0, lane = instance)       the two serial recursions -- backward Riccati sweep with the inertia
//                                                 test, forward roll-out -- of ALL the block's instances at once, one
//                                                 lane per instance, reading/writing the coop area
//   2   STEP      (owner warps, stage-parallel)   multiplier step, fraction-to-the-boundary limits, directional derivative
//   3   TRIAL     (owner warps, stage-parallel)   trial point, residual norms (shuffle butterflies), filter logic
// The serial recursions are thus executed once per instance (by one lane) instead of once per lane, and the lanes of
// the Riccati warp are filled with the block's instances.  The scalar IPOPT logic (filter, barrier update, inertia
// correction, termination; kmpc_core.cuh functions) runs on lane 0 of the owner warp on a Ctx kept in shared memory.
// What it replaces in the reference is the IPOPT solve behind mpc/optimizer.py:354/:375-391.
#pragma once
#include "kmpc_core.cuh"
#include "kmpc_warp_prims.cuh"

namespace kmpc {

template <int SPL>
struct WState {  // one iterate: this lane's SPL stages
    double x0[SPL], x1[SPL], x2[SPL], v[SPL], om[SPL], y0[SPL], y1[SPL], y2[SPL];
    double zLx[SPL], zUx[SPL], zLy[SPL], zUy[SPL], zLv[SPL], zUv[SPL], zLw[SPL], zUw[SPL], cs[SPL], sn[SPL];
};
template <int SPL>
struct WStep { double dx0[SPL], dx1[SPL], dx2[SPL], du0[SPL], du1[SPL], dy0[SPL], dy1[SPL], dy2[SPL]; };

// ---- shared-memory layout ----------------------------------------------------------------------------------------
// coop area of one instance: C_NF fields x NSTG stages (field-major: the owner lanes touch consecutive stages, the
// Riccati lanes -- one per instance -- are COOP doubles apart, COOP odd => both patterns are bank-conflict free).
// Fields 7..17 are the stage blocks going in; the backward sweep overwrites them (and fields 18..23) with K, k_ff, P, p;
// the forward roll-out overwrites K with (dx, du).
enum { C_A13 = 0, C_A23, C_B11, C_B21, C_E0, C_E1, C_E2,
       C_Q00 = 7, C_Q11, C_Q22, C_Q0, C_Q1, C_Q2, C_QV, C_QW, C_DV, C_DW, C_HTV,
       C_NF = 24 };
enum { C_K00 = 7, C_K01, C_K02, C_K10, C_K11, C_K12, C_KF0, C_KF1,
       C_P00 = 15, C_P10, C_P11, C_P20, C_P21, C_P22, C_PV0, C_PV1, C_PV2 };
enum { C_DX0 = 7, C_DX1, C_DX2, C_DU0, C_DU1 };
// private area of one instance (owner warp only): the kept Newton step (back-tracking / failed corrections return to
// it), the second-order-correction rhs, the constraint values of the last trial point
enum { V_DX0 = 0, V_DX1, V_DX2, V_DU0, V_DU1, V_DY0, V_DY1, V_DY2, V_CS0, V_CS1, V_CS2, V_CT0, V_CT1, V_CT2, V_NF };

struct WScal {  // warp-uniform per-instance scalars
    Ctx t;
    double filt[2 * K_FILTER_CAP];
    double xc[3], gl[3], d0[3];
    int flag, ok, r, status;
};

template <int SPL>
struct WLay {
    static constexpr int NSTG = 32 * SPL;
    static constexpr int COOP = C_NF * NSTG + 1;
    static constexpr int PRIV = V_NF * NSTG;
    static size_t bytes(int warps) { return (size_t)warps * ((COOP + PRIV) * sizeof(double) + sizeof(WScal)); }
};

// value of the next / previous stage (neighbouring slot, or the neighbouring lane's edge slot)
template <int SPL>
KMPC_W void w_next(const double (&a)[SPL], double (&n)[SPL]) {
    const double h = w_down(a[0], 1);
#pragma unroll
    for (int j = 0; j < SPL - 1; ++j) n[j] = a[j + 1];
    n[SPL - 1] = h;
}
template <int SPL>
KMPC_W void w_prev(const double (&a)[SPL], double (&p)[SPL]) {
    const double h = w_up(a[SPL - 1], 1);
    p[0] = h;
#pragma unroll
    for (int j = 1; j < SPL; ++j) p[j] = a[j - 1];
}

// ---- starting point: optimizer.py:375-385 (warm start) / agent.py:59-60 (cold start); IPOPT initialisation ----
template <int SPL>
KMPC_WN inline void w_init(const Cfg &c, WScal *sc, const IO &io, int b, WState<SPL> &w) {
    const int N = c.N, lane = w_lane();
    double xc[3], gl[3];
    for (int j = 0; j < 3; ++j) { xc[j] = io.x_cur[io_vec3(c, b, j)]; gl[j] = io.goal[io_vec3(c, b, j)]; }
    double gm = 0.0;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        const int s = lane * SPL + j;
        double x[3] = {xc[0], xc[1], xc[2]}, u[2] = {0.0, 0.0};
        const bool valid = s <= N, hasu = s < N;
        if (valid && io.X0) for (int i = 0; i < 3; ++i) x[i] = io.X0[io_X(c, b, i, s)];
        if (hasu && io.U0) for (int i = 0; i < 2; ++i) u[i] = io.U0[io_U(c, b, i, s)];
        if (valid && s >= c.gk_lo && s <= c.gk_hi)
            for (int i = 0; i < 3; ++i) gm = maxabs_nan(gm, 2.0 * c.W[i] * (x[i] - gl[i]));
        x[0] = push_in(x[0], c.lb[0], c.ub[0], c.hasL[0], c.hasU[0]);
        x[1] = push_in(x[1], c.lb[1], c.ub[1], c.hasL[1], c.hasU[1]);
        double cs = 1.0, sn = 0.0;
        if (hasu) {
            double gv, hv;
            vcost(c, 1.0, u[0], &gv, &hv);
            gm = maxabs_nan(gm, gv); gm = maxabs_nan(gm, 2.0 * c.Ww * u[1]);
            u[0] = push_in(u[0], c.lb[2], c.ub[2], c.hasL[2], c.hasU[2]);
            u[1] = push_in(u[1], c.lb[3], c.ub[3], c.hasL[3], c.hasU[3]);
            sincos_(x[2], &sn, &cs);
        }
        w.x0[j] = x[0]; w.x1[j] = x[1]; w.x2[j] = x[2]; w.v[j] = u[0]; w.om[j] = u[1];
        w.y0[j] = 0.0; w.y1[j] = 0.0; w.y2[j] = 0.0;
        w.zLx[j] = (valid && c.hasL[0]) ? 1.0 : 0.0; w.zUx[j] = (valid && c.hasU[0]) ? 1.0 : 0.0;
        w.zLy[j] = (valid && c.hasL[1]) ? 1.0 : 0.0; w.zUy[j] = (valid && c.hasU[1]) ? 1.0 : 0.0;
        w.zLv[j] = (hasu && c.hasL[2]) ? 1.0 : 0.0; w.zUv[j] = (hasu && c.hasU[2]) ? 1.0 : 0.0;
        w.zLw[j] = (hasu && c.hasL[3]) ? 1.0 : 0.0; w.zUw[j] = (hasu && c.hasU[3]) ? 1.0 : 0.0;
        w.cs[j] = cs; w.sn[j] = sn;
    }
    gm = w_maxabs_nan(gm);
    if (lane == 0) {
        Ctx &t = sc->t;
        for (int j = 0; j < 3; ++j) { sc->xc[j] = xc[j]; sc->gl[j] = gl[j]; }
        t.df = gm > K_SCALING_MAX_GRAD ? fmax(K_SCALING_MAX_GRAD / gm, K_SCALING_MIN) : 1.0;
        t.inst = b; t.cur = 0; t.iter = 0; t.mu = K_MU_INIT; t.tau = fmax(K_TAU_MIN, 1.0 - K_MU_INIT);
        t.delta = 0.0; t.delta_last = 0.0; t.theta_max = -1.0; t.theta_min = -1.0; t.fn = 0;
        t.nsteps = 0; t.soc_count = 0; t.trips = 0; t.sel = 0; t.tu = TU_INIT;
        t.alpha = t.alpha_test = t.alpha_min = t.alpha_du0 = t.alpha_soc = t.gBD = t.theta_soc_old = t.theta_trial = 0.0;
        t.a_pr = t.a_y = t.a_du = 0.0; t.pw_g = t.pw_t = 0.0;
        t.c.f = t.c.bar = t.c.damp = t.c.theta = t.c.dinf = t.c.pinf = t.c.mn = t.c.mx = t.c.sumy = t.c.sumz = t.c.wmax = 0.0;
        t.mode = M_LSQ;
    }
    w_sync();
}

// ---- phase 1a, ASSEMBLE: stage blocks of the KKT system -> coop area (all stages at once) ----
// Stages without a control (the terminal stage N) become pass-through steps of the recursion: zero dynamics, unit Q_uu,
// zero rhs -> P_out = P_in + Q, p_out = p_in + q.
template <int SPL>
KMPC_WN inline void w_assemble(const Cfg &c, WScal *sc, const WState<SPL> &w, const double *priv, double *coop) {
    constexpr int NSTG = WLay<SPL>::NSTG;
    const int N = c.N, lane = w_lane();
    const int mode = sc->t.mode;
    const bool lsq = mode == M_LSQ, soc = mode == M_SOC;
    const double mu = sc->t.mu, delta = sc->t.delta, df = sc->t.df, T = c.T;
    const double gl0 = sc->gl[0], gl1 = sc->gl[1], gl2 = sc->gl[2];
    double yn0[SPL], yn1[SPL], yn2[SPL], xn0[SPL], xn1[SPL], xn2[SPL];
    w_next<SPL>(w.y0, yn0); w_next<SPL>(w.y1, yn1); w_next<SPL>(w.y2, yn2);
    w_next<SPL>(w.x0, xn0); w_next<SPL>(w.x1, xn1); w_next<SPL>(w.x2, xn2);
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        const int s = lane * SPL + j;
        if (s > N) continue;
        const double x0 = w.x0[j], x1 = w.x1[j], x2 = w.x2[j];
        const bool ing = s >= c.gk_lo && s <= c.gk_hi;
        double gx0 = 0, gx1 = 0, gx2 = 0, h0 = 0, h1 = 0, h2 = 0;
        if (ing) {
            gx0 = df * 2.0 * c.W[0] * (x0 - gl0); gx1 = df * 2.0 * c.W[1] * (x1 - gl1); gx2 = df * 2.0 * c.W[2] * (x2 - gl2);
            h0 = df * 2.0 * c.W[0]; h1 = df * 2.0 * c.W[1]; h2 = df * 2.0 * c.W[2];
        }
        double q0, q1, q2, Q00, Q11, Q22;
        if (lsq) {
            q0 = -(gx0 - w.zLx[j] + w.zUx[j]); q1 = -(gx1 - w.zLy[j] + w.zUy[j]); q2 = -gx2;
            Q00 = 1.0; Q11 = 1.0; Q22 = 1.0;
        } else {
            double sg0, rb0, sg1, rb1;
            bound_terms(x0, c.lb[0], c.ub[0], c.hasL[0], c.hasU[0], w.zLx[j], w.zUx[j], mu, &sg0, &rb0);
            bound_terms(x1, c.lb[1], c.ub[1], c.hasL[1], c.hasU[1], w.zLy[j], w.zUy[j], mu, &sg1, &rb1);
            q0 = gx0 + w.y0[j] + rb0; q1 = gx1 + w.y1[j] + rb1; q2 = gx2 + w.y2[j];
            Q00 = h0 + sg0 + delta; Q11 = h1 + sg1 + delta; Q22 = h2 + delta;
        }
        const double v = w.v[j], om = w.om[j], cs = w.cs[j], sn = w.sn[j];
        double a13 = -T * v * sn, a23 = T * v * cs, b11 = T * cs, b21 = T * sn;
        double gv, hvv;
        vcost(c, df, v, &gv, &hvv);
        const double gw = df * 2.0 * c.Ww * om;
        const double hww = df * 2.0 * c.Ww;
        double htv = 0.0, qv, qw, dv, dw, e0, e1, e2;
        if (lsq) {
            qv = -(gv - w.zLv[j] + w.zUv[j]); qw = -(gw - w.zLw[j] + w.zUw[j]);
            dv = 1.0; dw = 1.0;
            e0 = e1 = e2 = 0.0;
        } else {
            double sgv, rbv, sgw, rbw;
            bound_terms(v, c.lb[2], c.ub[2], c.hasL[2], c.hasU[2], w.zLv[j], w.zUv[j], mu, &sgv, &rbv);
            bound_terms(om, c.lb[3], c.ub[3], c.hasL[3], c.hasU[3], w.zLw[j], w.zUw[j], mu, &sgw, &rbw);
            if (s < N) {
                // J^T y of dynamics row s+1 and the curvature of the dynamics in the Lagrangian
                q0 -= yn0[j]; q1 -= yn1[j]; q2 -= a13 * yn0[j] + a23 * yn1[j] + yn2[j];
                Q22 += T * v * (yn0[j] * cs + yn1[j] * sn);
                htv = T * (yn0[j] * sn - yn1[j] * cs);
            }
            qv = gv - (b11 * yn0[j] + b21 * yn1[j]) + rbv;
            qw = gw - T * yn2[j] + rbw;
            dv = hvv + (sgv + delta); dw = hww + (sgw + delta);
            if (soc) {  // rhs of the dynamics row s+1 = -c_soc of stage s+1
                const double *pn = priv + (s + 1 < NSTG ? s + 1 : s);
                e0 = -pn[V_CS0 * NSTG]; e1 = -pn[V_CS1 * NSTG]; e2 = -pn[V_CS2 * NSTG];
            } else { e0 = -(xn0[j] - (x0 + T * v * cs)); e1 = -(xn1[j] - (x1 + T * v * sn)); e2 = -(xn2[j] - (x2 + T * om)); }
        }
        if (s >= N) {
            a13 = a23 = b11 = b21 = 0.0; qv = qw = htv = 0.0; dv = dw = 1.0;
            e0 = e1 = e2 = 0.0;
        }
        double *q = coop + s;
        q[C_A13 * NSTG] = a13; q[C_A23 * NSTG] = a23; q[C_B11 * NSTG] = b11; q[C_B21 * NSTG] = b21;
        q[C_E0 * NSTG] = e0; q[C_E1 * NSTG] = e1; q[C_E2 * NSTG] = e2;
        q[C_Q00 * NSTG] = Q00; q[C_Q11 * NSTG] = Q11; q[C_Q22 * NSTG] = Q22;
        q[C_Q0 * NSTG] = q0; q[C_Q1 * NSTG] = q1; q[C_Q2 * NSTG] = q2;
        q[C_QV * NSTG] = qv; q[C_QW * NSTG] = qw; q[C_DV * NSTG] = dv; q[C_DW * NSTG] = dw; q[C_HTV * NSTG] = htv;
        if (s == 0) {  // dx of stage 0 (the rhs of the initial-state row)
            if (lsq) { sc->d0[0] = sc->d0[1] = sc->d0[2] = 0.0; }
            else if (soc) { sc->d0[0] = -priv[V_CS0 * NSTG]; sc->d0[1] = -priv[V_CS1 * NSTG]; sc->d0[2] = -priv[V_CS2 * NSTG]; }
            else { sc->d0[0] = -(x0 - sc->xc[0]); sc->d0[1] = -(x1 - sc->xc[1]); sc->d0[2] = -(x2 - sc->xc[2]); }
        }
    }
}

// ---- phase 1b, RICCATI: the two serial recursions of ONE instance, executed by ONE lane of the block's Riccati warp ----
// Backward sweep (K, k_ff, P, p of every stage; false = some Q_uu not positive definite = wrong inertia), then the
// forward substitution dx+ = A dx + B du + e, du = K dx + k_ff.
KMPC_WN inline bool w_serial(const Cfg &c, double *coop, const int NSTG, const double *d0) {
    const int N = c.N;
    const double T = c.T;
    double P00 = 0, P10 = 0, P11 = 0, P20 = 0, P21 = 0, P22 = 0, p0 = 0, p1 = 0, p2 = 0;
#pragma unroll 1
    for (int s = N; s >= 0; --s) {
        double *q = coop + s;
        const double a13 = q[C_A13 * NSTG], a23 = q[C_A23 * NSTG], b11 = q[C_B11 * NSTG], b21 = q[C_B21 * NSTG];
        const double e0 = q[C_E0 * NSTG], e1 = q[C_E1 * NSTG], e2 = q[C_E2 * NSTG];
        const double Q00 = q[C_Q00 * NSTG], Q11 = q[C_Q11 * NSTG], Q22 = q[C_Q22 * NSTG];
        const double q0 = q[C_Q0 * NSTG], q1 = q[C_Q1 * NSTG], q2 = q[C_Q2 * NSTG];
        const double qv = q[C_QV * NSTG], qw = q[C_QW * NSTG], dv = q[C_DV * NSTG], dw = q[C_DW * NSTG], htv = q[C_HTV * NSTG];
        RicK rk;
        if (!riccati_step(P00, P10, P11, P20, P21, P22, p0, p1, p2, a13, a23, b11, b21, T, Q00, 0.0, Q11, Q22, q0, q1, q2, qv, qw,
                          dv, dw, htv, e0, e1, e2, rk))
            return false;
        q[C_K00 * NSTG] = rk.K00; q[C_K01 * NSTG] = rk.K01; q[C_K02 * NSTG] = rk.K02;
        q[C_K10 * NSTG] = rk.K10; q[C_K11 * NSTG] = rk.K11; q[C_K12 * NSTG] = rk.K12;
        q[C_KF0 * NSTG] = rk.kf0; q[C_KF1 * NSTG] = rk.kf1;
        q[C_P00 * NSTG] = P00; q[C_P10 * NSTG] = P10; q[C_P11 * NSTG] = P11;
        q[C_P20 * NSTG] = P20; q[C_P21 * NSTG] = P21; q[C_P22 * NSTG] = P22;
        q[C_PV0 * NSTG] = p0; q[C_PV1 * NSTG] = p1; q[C_PV2 * NSTG] = p2;
    }
    double x0 = d0[0], x1 = d0[1], x2 = d0[2];
#pragma unroll 1
    for (int s = 0; s <= N; ++s) {
        double *q = coop + s;
        const double K00 = q[C_K00 * NSTG], K01 = q[C_K01 * NSTG], K02 = q[C_K02 * NSTG];
        const double K10 = q[C_K10 * NSTG], K11 = q[C_K11 * NSTG], K12 = q[C_K12 * NSTG];
        const double kf0 = q[C_KF0 * NSTG], kf1 = q[C_KF1 * NSTG];
        const double a13 = q[C_A13 * NSTG], a23 = q[C_A23 * NSTG], b11 = q[C_B11 * NSTG], b21 = q[C_B21 * NSTG];
        const double e0 = q[C_E0 * NSTG], e1 = q[C_E1 * NSTG], e2 = q[C_E2 * NSTG];
        const double du0 = fma(K00, x0, fma(K01, x1, fma(K02, x2, kf0)));
        const double du1 = fma(K10, x0, fma(K11, x1, fma(K12, x2, kf1)));
        q[C_DX0 * NSTG] = x0; q[C_DX1 * NSTG] = x1; q[C_DX2 * NSTG] = x2; q[C_DU0 * NSTG] = du0; q[C_DU1 * NSTG] = du1;
        const double n0 = x0 + a13 * x2 + b11 * du0 + e0;
        const double n1 = x1 + a23 * x2 + b21 * du0 + e1;
        const double n2 = x2 + T * du1 + e2;
        x0 = n0; x1 = n1; x2 = n2;
    }
    return true;
}

// ---- phase 2, STEP: multiplier step dy = -(P dx + p), step-size limits, directional derivative (all stages at once) ----
template <int SPL>
KMPC_WN inline void w_step(const Cfg &c, const WScal *sc, const WState<SPL> &w, const double *coop, WStep<SPL> &d,
                           double *alpha_pr, double *alpha_du, double *gBD, double *ymax) {
    constexpr int NSTG = WLay<SPL>::NSTG;
    const int N = c.N, lane = w_lane();
    const bool lsq = sc->t.mode == M_LSQ;
    const double mu = sc->t.mu, df = sc->t.df, tau = sc->t.tau;
    const double gl0 = sc->gl[0], gl1 = sc->gl[1], gl2 = sc->gl[2];
    double apr = 1.0, adu = 1.0, gbd = 0.0, ym = 0.0;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        const int s = lane * SPL + j;
        if (s > N) { d.dx0[j] = d.dx1[j] = d.dx2[j] = d.du0[j] = d.du1[j] = d.dy0[j] = d.dy1[j] = d.dy2[j] = 0.0; continue; }
        const double *q = coop + s;
        const double d0 = q[C_DX0 * NSTG], d1 = q[C_DX1 * NSTG], d2 = q[C_DX2 * NSTG];
        const double du0 = s < N ? q[C_DU0 * NSTG] : 0.0, du1 = s < N ? q[C_DU1 * NSTG] : 0.0;
        const double P00 = q[C_P00 * NSTG], P10 = q[C_P10 * NSTG], P11 = q[C_P11 * NSTG], P20 = q[C_P20 * NSTG],
                     P21 = q[C_P21 * NSTG], P22 = q[C_P22 * NSTG];
        const double dy0 = -(P00 * d0 + P10 * d1 + P20 * d2 + q[C_PV0 * NSTG]);
        const double dy1 = -(P10 * d0 + P11 * d1 + P21 * d2 + q[C_PV1 * NSTG]);
        const double dy2 = -(P20 * d0 + P21 * d1 + P22 * d2 + q[C_PV2 * NSTG]);
        d.dx0[j] = d0; d.dx1[j] = d1; d.dx2[j] = d2; d.du0[j] = du0; d.du1[j] = du1;
        d.dy0[j] = dy0; d.dy1[j] = dy1; d.dy2[j] = dy2;
        ym = maxabs_nan(maxabs_nan(maxabs_nan(ym, dy0), dy1), dy2);
        if (lsq) continue;
        const double x0 = w.x0[j], x1 = w.x1[j], x2 = w.x2[j];
        double sg, rb;
        bound_ftb(x0, d0, c.lb[0], c.ub[0], c.hasL[0], c.hasU[0], w.zLx[j], w.zUx[j], mu, tau, &apr, &adu);
        bound_ftb(x1, d1, c.lb[1], c.ub[1], c.hasL[1], c.hasU[1], w.zLy[j], w.zUy[j], mu, tau, &apr, &adu);
        const bool ing = s >= c.gk_lo && s <= c.gk_hi;
        bound_terms(x0, c.lb[0], c.ub[0], c.hasL[0], c.hasU[0], w.zLx[j], w.zUx[j], mu, &sg, &rb);
        gbd += ((ing ? df * 2.0 * c.W[0] * (x0 - gl0) : 0.0) + rb) * d0;
        bound_terms(x1, c.lb[1], c.ub[1], c.hasL[1], c.hasU[1], w.zLy[j], w.zUy[j], mu, &sg, &rb);
        gbd += ((ing ? df * 2.0 * c.W[1] * (x1 - gl1) : 0.0) + rb) * d1;
        gbd += (ing ? df * 2.0 * c.W[2] * (x2 - gl2) : 0.0) * d2;
        if (s < N) {
            const double v = w.v[j], om = w.om[j];
            bound_ftb(v, du0, c.lb[2], c.ub[2], c.hasL[2], c.hasU[2], w.zLv[j], w.zUv[j], mu, tau, &apr, &adu);
            bound_ftb(om, du1, c.lb[3], c.ub[3], c.hasL[3], c.hasU[3], w.zLw[j], w.zUw[j], mu, tau, &apr, &adu);
            double gv, hv;
            vcost(c, df, v, &gv, &hv);
            bound_terms(v, c.lb[2], c.ub[2], c.hasL[2], c.hasU[2], w.zLv[j], w.zUv[j], mu, &sg, &rb);
            gbd += (gv + rb) * du0;
            bound_terms(om, c.lb[3], c.ub[3], c.hasL[3], c.hasU[3], w.zLw[j], w.zUw[j], mu, &sg, &rb);
            gbd += (df * 2.0 * c.Ww * om + rb) * du1;
        }
    }
    *alpha_pr = w_min(apr); *alpha_du = w_min(adu); *gBD = w_sum(gbd); *ymax = w_maxabs_nan(ym);
}

// kept Newton step <-> private area
template <int SPL>
KMPC_W void w_step_store(const WStep<SPL> &d, double *priv) {
    constexpr int NSTG = WLay<SPL>::NSTG;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        double *p = priv + w_lane() * SPL + j;
        p[V_DX0 * NSTG] = d.dx0[j]; p[V_DX1 * NSTG] = d.dx1[j]; p[V_DX2 * NSTG] = d.dx2[j]; p[V_DU0 * NSTG] = d.du0[j];
        p[V_DU1 * NSTG] = d.du1[j]; p[V_DY0 * NSTG] = d.dy0[j]; p[V_DY1 * NSTG] = d.dy1[j]; p[V_DY2 * NSTG] = d.dy2[j];
    }
}
template <int SPL>
KMPC_W void w_step_load(WStep<SPL> &d, const double *priv) {
    constexpr int NSTG = WLay<SPL>::NSTG;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        const double *p = priv + w_lane() * SPL + j;
        d.dx0[j] = p[V_DX0 * NSTG]; d.dx1[j] = p[V_DX1 * NSTG]; d.dx2[j] = p[V_DX2 * NSTG]; d.du0[j] = p[V_DU0 * NSTG];
        d.du1[j] = p[V_DU1 * NSTG]; d.dy0[j] = p[V_DY0 * NSTG]; d.dy1[j] = p[V_DY1 * NSTG]; d.dy2[j] = p[V_DY2 * NSTG];
    }
}

// ---- phase 3, TRIAL: trial point + speculative multiplier update + residual norms (all stages at once) ----
template <int SPL>
KMPC_WN inline bool w_trial(const Cfg &c, const WScal *sc, const WState<SPL> &w, const WStep<SPL> &d, double alpha, double ay,
                            double adu, bool clamp, WState<SPL> &n, double *priv, Stats *out) {
    constexpr int NSTG = WLay<SPL>::NSTG;
    const int N = c.N, lane = w_lane();
    const double mu = sc->t.mu, df = sc->t.df, T = c.T;
    const double xc0 = sc->xc[0], xc1 = sc->xc[1], xc2 = sc->xc[2], gl0 = sc->gl[0], gl1 = sc->gl[1], gl2 = sc->gl[2];
    Stats st;
    st.f = 0; st.bar = 0; st.damp = 0; st.theta = 0; st.dinf = 0; st.pinf = 0; st.mn = INFINITY; st.mx = 0; st.sumy = 0;
    st.sumz = 0; st.wmax = 0;
    bool valid = true;
    double xp0[SPL], xp1[SPL], xp2[SPL];  // state predicted from this stage
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        const int s = lane * SPL + j;
        n.x0[j] = w.x0[j] + alpha * d.dx0[j]; n.x1[j] = w.x1[j] + alpha * d.dx1[j]; n.x2[j] = w.x2[j] + alpha * d.dx2[j];
        n.v[j] = w.v[j] + alpha * d.du0[j]; n.om[j] = w.om[j] + alpha * d.du1[j];
        n.y0[j] = w.y0[j] + ay * d.dy0[j]; n.y1[j] = w.y1[j] + ay * d.dy1[j]; n.y2[j] = w.y2[j] + ay * d.dy2[j];
        n.zLx[j] = n.zUx[j] = n.zLy[j] = n.zUy[j] = n.zLv[j] = n.zUv[j] = n.zLw[j] = n.zUw[j] = 0.0;
        double sn = 0.0, cs = 1.0;
        if (s < N) sincos_(n.x2[j], &sn, &cs);
        n.cs[j] = cs; n.sn[j] = sn;
        xp0[j] = n.x0[j] + T * n.v[j] * cs; xp1[j] = n.x1[j] + T * n.v[j] * sn; xp2[j] = n.x2[j] + T * n.om[j];
    }
    double pp0[SPL], pp1[SPL], pp2[SPL], yn0[SPL], yn1[SPL], yn2[SPL];
    w_prev<SPL>(xp0, pp0); w_prev<SPL>(xp1, pp1); w_prev<SPL>(xp2, pp2);
    w_next<SPL>(n.y0, yn0); w_next<SPL>(n.y1, yn1); w_next<SPL>(n.y2, yn2);
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        const int s = lane * SPL + j;
        if (s > N) continue;
        const double x0 = n.x0[j], x1 = n.x1[j], x2 = n.x2[j];
        const double c0 = x0 - (s == 0 ? xc0 : pp0[j]), c1 = x1 - (s == 0 ? xc1 : pp1[j]), c2 = x2 - (s == 0 ? xc2 : pp2[j]);
        double *pv = priv + s;
        pv[V_CT0 * NSTG] = c0; pv[V_CT1 * NSTG] = c1; pv[V_CT2 * NSTG] = c2;
        st.theta += fabs(c0) + fabs(c1) + fabs(c2);
        st.pinf = maxabs_nan(maxabs_nan(maxabs_nan(st.pinf, c0), c1), c2);
        st.sumy += fabs(n.y0[j]) + fabs(n.y1[j]) + fabs(n.y2[j]);
        st.wmax = fmax(st.wmax, fmax(fabs(x0), fmax(fabs(x1), fabs(x2))));
        double r0 = n.y0[j], r1 = n.y1[j], r2 = n.y2[j];
        if (s >= c.gk_lo && s <= c.gk_hi) {
            const double e0 = x0 - gl0, e1 = x1 - gl1, e2 = x2 - gl2;
            st.f += c.W[0] * e0 * e0; st.f += c.W[1] * e1 * e1; st.f += c.W[2] * e2 * e2;
            r0 += df * 2.0 * c.W[0] * e0; r1 += df * 2.0 * c.W[1] * e1; r2 += df * 2.0 * c.W[2] * e2;
        }
        double prod = 1.0, zLn, zUn;
        valid &= bound_trial(w.x0[j], d.dx0[j], x0, c.lb[0], c.ub[0], c.hasL[0], c.hasU[0], w.zLx[j], w.zUx[j], mu, adu, clamp,
                             &zLn, &zUn, &prod, &st.damp, &st);
        n.zLx[j] = zLn; n.zUx[j] = zUn; r0 += zUn - zLn;
        valid &= bound_trial(w.x1[j], d.dx1[j], x1, c.lb[1], c.ub[1], c.hasL[1], c.hasU[1], w.zLy[j], w.zUy[j], mu, adu, clamp,
                             &zLn, &zUn, &prod, &st.damp, &st);
        n.zLy[j] = zLn; n.zUy[j] = zUn; r1 += zUn - zLn;
        if (s < N) {
            const double v = n.v[j], om = n.om[j], cs = n.cs[j], sn = n.sn[j];
            st.wmax = fmax(st.wmax, fmax(fabs(v), fabs(om)));
            const double a13 = -T * v * sn, a23 = T * v * cs;
            r0 -= yn0[j]; r1 -= yn1[j]; r2 -= a13 * yn0[j] + a23 * yn1[j] + yn2[j];
            double gv, hv;
            vcost(c, df, v, &gv, &hv);
            double rv = gv - (T * cs * yn0[j] + T * sn * yn1[j]), rw = df * 2.0 * c.Ww * om - T * yn2[j];
            if (c.cost_mode == 0) { const double vm = fmin(v, 0.0), vp = fmax(v, 0.0); st.f += c.Wvn * vm * vm + c.Wvp * vp * vp; }
            else st.f += c.Wvn * fmin(v, 0.0);
            st.f += c.Ww * om * om;
            valid &= bound_trial(w.v[j], d.du0[j], v, c.lb[2], c.ub[2], c.hasL[2], c.hasU[2], w.zLv[j], w.zUv[j], mu, adu, clamp,
                                 &zLn, &zUn, &prod, &st.damp, &st);
            n.zLv[j] = zLn; n.zUv[j] = zUn; rv += zUn - zLn;
            valid &= bound_trial(w.om[j], d.du1[j], om, c.lb[3], c.ub[3], c.hasL[3], c.hasU[3], w.zLw[j], w.zUw[j], mu, adu, clamp,
                                 &zLn, &zUn, &prod, &st.damp, &st);
            n.zLw[j] = zLn; n.zUw[j] = zUn; rw += zUn - zLn;
            st.dinf = maxabs_nan(maxabs_nan(st.dinf, rv), rw);
        } else {
            n.zLv[j] = n.zUv[j] = n.zLw[j] = n.zUw[j] = 0.0;
        }
        st.dinf = maxabs_nan(maxabs_nan(maxabs_nan(st.dinf, r0), r1), r2);
        st.bar += log(prod);
    }
    Stats g;
    g.f = w_sum(st.f) * df; g.bar = w_sum(st.bar); g.damp = w_sum(st.damp); g.theta = w_sum(st.theta);
    g.dinf = w_maxabs_nan(st.dinf); g.pinf = w_maxabs_nan(st.pinf); g.mn = w_min(st.mn); g.mx = w_max(st.mx);
    g.sumy = w_sum(st.sumy); g.sumz = w_sum(st.sumz); g.wmax = w_max(st.wmax);
    if (c.nb == 0) g.mn = 0.0;
    *out = g;
    const double phi = g.f - mu * g.bar + K_KAPPA_D * mu * g.damp;
    return w_all(valid) && isfinite(phi) && isfinite(g.theta);
}

// c_soc <- al * base + c(trial); base = c(current) for the first correction, else the previous c_soc
template <int SPL>
KMPC_WN inline void w_soc_rhs(const Cfg &c, const WScal *sc, const WState<SPL> &w, double al, bool first, double *priv) {
    constexpr int NSTG = WLay<SPL>::NSTG;
    const int N = c.N, lane = w_lane();
    const double T = c.T;
    double xp0[SPL], xp1[SPL], xp2[SPL], pp0[SPL], pp1[SPL], pp2[SPL];
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        xp0[j] = w.x0[j] + T * w.v[j] * w.cs[j]; xp1[j] = w.x1[j] + T * w.v[j] * w.sn[j]; xp2[j] = w.x2[j] + T * w.om[j];
    }
    w_prev<SPL>(xp0, pp0); w_prev<SPL>(xp1, pp1); w_prev<SPL>(xp2, pp2);
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        const int s = lane * SPL + j;
        if (s > N) continue;
        double *pv = priv + s;
        double b0, b1, b2;
        if (first) { b0 = w.x0[j] - (s == 0 ? sc->xc[0] : pp0[j]); b1 = w.x1[j] - (s == 0 ? sc->xc[1] : pp1[j]); b2 = w.x2[j] - (s == 0 ? sc->xc[2] : pp2[j]); }
        else { b0 = pv[V_CS0 * NSTG]; b1 = pv[V_CS1 * NSTG]; b2 = pv[V_CS2 * NSTG]; }
        pv[V_CS0 * NSTG] = al * b0 + pv[V_CT0 * NSTG]; pv[V_CS1 * NSTG] = al * b1 + pv[V_CT1 * NSTG]; pv[V_CS2 * NSTG] = al * b2 + pv[V_CT2 * NSTG];
    }
    w_sync();
}

// ---- persistent worker: one warp pulls instances from a queue and solves each start to finish; the warps of a block
// walk through the phases of a trip in step (block barriers) so that warp 0 can run every instance's serial recursions.
// smem: WLay<SPL>::bytes(warps per block) bytes of block-shared scratch.
template <int SPL>
KMPC_WN inline void w_worker(const Cfg &c, const IO &io, double *smem, int *queue, unsigned long long *trips_total) {
    typedef WLay<SPL> LY;
    const int N = c.N, lane = w_lane(), wid = w_warp(), W = w_warps();
    double *coop = smem + (size_t)wid * LY::COOP;
    double *priv = smem + (size_t)W * LY::COOP + (size_t)wid * LY::PRIV;
    WScal *scal0 = (WScal *)(smem + (size_t)W * (LY::COOP + LY::PRIV));
    WScal *sc = scal0 + wid;
    Ctx &t = sc->t;
    WState<SPL> cur;
    WStep<SPL> act;
    bool have = false;
    int b = -1;
    if (lane == 0) { t.mode = M_DONE; sc->flag = 0; sc->ok = 0; }
    w_sync();
#pragma unroll 1
    for (;;) {
        if (!have) {
            b = w_fetch(queue);
            if (b < c.B) { w_init<SPL>(c, sc, io, b, cur); have = true; }
        }
        if (!w_block_any(have)) break;
        int status = 100;
        // ---- phase 1a: assemble the stage blocks ----
        const int mode = t.mode;
        const bool do_sweep = have && mode != M_TRIAL;
        if (do_sweep) w_assemble<SPL>(c, sc, cur, priv, coop);
        if (lane == 0) { sc->flag = do_sweep ? 1 : 0; if (do_sweep) t.trips++; }
        w_block_sync();
        // ---- phase 1b: the serial recursions of all the block's instances, one lane each ----
        if (wid == 0 && lane < W) {
            WScal *so = scal0 + lane;
            if (so->flag) so->ok = w_serial(c, smem + (size_t)lane * LY::COOP, LY::NSTG, so->d0) ? 1 : 0;
        }
        w_block_sync();
        // ---- phase 2: search direction, step sizes, line-search set-up (or inertia correction) ----
        bool go_trial = have && !do_sweep;
        if (do_sweep) {
            if (!sc->ok) {
                if (lane == 0) sc->status = mode != M_NEWTON ? (int)ST_STEP_ERROR : inertia_update(t);  // R_RETRY: sweep again next trip
                w_sync();
                status = sc->status;
            } else {
                double apr, adu, gbd, ym;
                w_step<SPL>(c, sc, cur, coop, act, &apr, &adu, &gbd, &ym);
                if (lane == 0) rollout_logic(t, apr, adu, gbd, ym);
                w_sync();
                if (t.sel == 0) w_step_store<SPL>(act, priv);
                go_trial = true;
            }
        } else if (have) {
            if (lane == 0) trial_setup(t);
            w_sync();
            w_step_load<SPL>(act, priv);
        }
        // ---- phase 3: trial point + acceptance logic ----
        if (go_trial) {
            Stats ts;
            WState<SPL> tri;
            const bool evok = w_trial<SPL>(c, sc, cur, act, t.a_pr, t.a_y, t.a_du, t.tu == TU_STEP, tri, priv, &ts);
            if (lane == 0) {
                bool aug; double ath, aph;
                const int r = trial_decide(t, sc->filt, 1, ts, evok, &aug, &ath, &aph);
                if (aug) filter_add(t, sc->filt, 1, ath, aph);
                sc->r = r;
            }
            w_sync();
            const int r = sc->r;
            if (r == R_SOC1 || r == R_SOC2) w_soc_rhs<SPL>(c, sc, cur, t.alpha_soc, r == R_SOC1, priv);
            else if (r == R_ACCEPT) {
                cur = tri;
                if (lane == 0) { t.c = ts; sc->status = begin_iteration(c, t); }
                w_sync();
                status = sc->status;
            } else if (r != R_BACKTRACK) status = r;
        }
        if (have && status != 100 && status != R_RETRY) {
            // returned matrices (optimizer.py:392-400): every lane writes its stages
#pragma unroll
            for (int j = 0; j < SPL; ++j) {
                const int s = lane * SPL + j;
                if (s <= N) { io.X_out[io_X(c, b, 0, s)] = cur.x0[j]; io.X_out[io_X(c, b, 1, s)] = cur.x1[j]; io.X_out[io_X(c, b, 2, s)] = cur.x2[j]; }
                if (s < N) { io.U_out[io_U(c, b, 0, s)] = cur.v[j]; io.U_out[io_U(c, b, 1, s)] = cur.om[j]; }
            }
            if (lane == 0) {
                if (io.obj) io.obj[b] = t.c.f / t.df;
                if (io.status) io.status[b] = status;
                if (io.iters) io.iters[b] = t.iter;
                w_count_trips(trips_total, t.trips);
                t.mode = M_DONE;
            }
            have = false;
        }
    }
}

}  // namespace kmpc
