// kmpc_warp.cuh -- warp-per-instance form of the interior-point solver: the stages of ONE problem instance are spread
// over the 32 lanes of a warp (stage s lives in lane s / SPL, slot s % SPL), the whole iterate (primal, duals, step,
// Riccati factors) stays in registers for the entire solve, and nothing but the problem data and the result touches HBM.
//   - stage-parallel work (linearisation, barrier terms, trial-point evaluation, multiplier updates, residual norms)
//     runs on all lanes at once, reductions are shuffle butterflies;
//   - the two serial recursions (backward Riccati, forward roll-out) walk lane by lane, handing the 3x3 cost-to-go /
//     the 3-vector state step to the neighbour with register shuffles;
//   - the scalar IPOPT logic (filter, barrier update, inertia correction, termination) is computed redundantly and
//     identically by all lanes (kmpc_core.cuh functions), so control flow is warp-uniform.
// Same algorithm and formulas as the thread-per-instance passes in kmpc_core.cuh (which it replaces for problems
// without obstacle rows and N < 32*SPL); what it replaces in the reference is the same: the IPOPT solve behind
// mpc/optimizer.py:354/:375-391.
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
template <int SPL>
struct WFact { double K00[SPL], K01[SPL], K02[SPL], K10[SPL], K11[SPL], K12[SPL], kf0[SPL], kf1[SPL];
               double P00[SPL], P10[SPL], P11[SPL], P20[SPL], P21[SPL], P22[SPL], pv0[SPL], pv1[SPL], pv2[SPL]; };
template <int SPL>
struct WVec3 { double a[SPL], b[SPL], c[SPL]; };
template <int SPL>
struct WLin { double a13[SPL], a23[SPL], b11[SPL], b21[SPL]; };  // non-trivial entries of A_k, B_k (zero for stages >= N)

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
KMPC_WN inline void w_init(const Cfg &c, Ctx &t, const IO &io, int b, WState<SPL> &w, double (&xc)[3], double (&gl)[3]) {
    const int N = c.N, lane = w_lane();
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
    t.df = gm > K_SCALING_MAX_GRAD ? fmax(K_SCALING_MAX_GRAD / gm, K_SCALING_MIN) : 1.0;
    t.inst = b; t.cur = 0; t.iter = 0; t.mu = K_MU_INIT; t.tau = fmax(K_TAU_MIN, 1.0 - K_MU_INIT);
    t.delta = 0.0; t.delta_last = 0.0; t.theta_max = -1.0; t.theta_min = -1.0; t.fn = 0;
    t.nsteps = 0; t.soc_count = 0; t.trips = 0; t.sel = 0; t.tu = TU_INIT;
    t.alpha = t.alpha_test = t.alpha_min = t.alpha_du0 = t.alpha_soc = t.gBD = t.theta_soc_old = t.theta_trial = 0.0;
    t.a_pr = t.a_y = t.a_du = 0.0; t.pw_g = t.pw_t = 0.0;
    t.c.f = t.c.bar = t.c.damp = t.c.theta = t.c.dinf = t.c.pinf = t.c.mn = t.c.mx = t.c.sumy = t.c.sumz = t.c.wmax = 0.0;
    t.mode = M_LSQ;
}

// ---- SWEEP: stage-parallel assembly of the KKT blocks, then the backward Riccati recursion lane by lane ----
// Returns false on wrong inertia (some Q_uu not positive definite).  e = bc of the NEXT stage's dynamics row, kept for
// the roll-out.
template <int SPL>
KMPC_WN inline bool w_sweep(const Cfg &c, const Ctx &t, const WState<SPL> &w, const WVec3<SPL> &csoc, const double (&gl)[3],
                            WFact<SPL> &f, WVec3<SPL> &e, WLin<SPL> &lin) {
    const int N = c.N, lane = w_lane();
    const bool lsq = t.mode == M_LSQ, soc = t.mode == M_SOC;
    const double mu = t.mu, delta = t.delta, df = t.df, T = c.T;
    double yn0[SPL], yn1[SPL], yn2[SPL], xn0[SPL], xn1[SPL], xn2[SPL], cn0[SPL], cn1[SPL], cn2[SPL];
    w_next<SPL>(w.y0, yn0); w_next<SPL>(w.y1, yn1); w_next<SPL>(w.y2, yn2);
    w_next<SPL>(w.x0, xn0); w_next<SPL>(w.x1, xn1); w_next<SPL>(w.x2, xn2);
    w_next<SPL>(csoc.a, cn0); w_next<SPL>(csoc.b, cn1); w_next<SPL>(csoc.c, cn2);
    // per-stage blocks that do not depend on the cost-to-go
    double q0[SPL], q1[SPL], q2[SPL], Q00[SPL], Q11[SPL], Q22[SPL], qv[SPL], qw[SPL], dv[SPL], dw[SPL], htv[SPL];
    double (&a13)[SPL] = lin.a13, (&a23)[SPL] = lin.a23, (&b11)[SPL] = lin.b11, (&b21)[SPL] = lin.b21;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        const int s = lane * SPL + j;
        const double x0 = w.x0[j], x1 = w.x1[j], x2 = w.x2[j];
        const bool ing = s >= c.gk_lo && s <= c.gk_hi;
        double gx0 = 0, gx1 = 0, gx2 = 0, h0 = 0, h1 = 0, h2 = 0;
        if (ing) {
            gx0 = df * 2.0 * c.W[0] * (x0 - gl[0]); gx1 = df * 2.0 * c.W[1] * (x1 - gl[1]); gx2 = df * 2.0 * c.W[2] * (x2 - gl[2]);
            h0 = df * 2.0 * c.W[0]; h1 = df * 2.0 * c.W[1]; h2 = df * 2.0 * c.W[2];
        }
        if (lsq) {
            q0[j] = -(gx0 - w.zLx[j] + w.zUx[j]); q1[j] = -(gx1 - w.zLy[j] + w.zUy[j]); q2[j] = -gx2;
            Q00[j] = 1.0; Q11[j] = 1.0; Q22[j] = 1.0;
        } else {
            double sg0, rb0, sg1, rb1;
            bound_terms(x0, c.lb[0], c.ub[0], c.hasL[0], c.hasU[0], w.zLx[j], w.zUx[j], mu, &sg0, &rb0);
            bound_terms(x1, c.lb[1], c.ub[1], c.hasL[1], c.hasU[1], w.zLy[j], w.zUy[j], mu, &sg1, &rb1);
            q0[j] = gx0 + w.y0[j] + rb0; q1[j] = gx1 + w.y1[j] + rb1; q2[j] = gx2 + w.y2[j];
            Q00[j] = h0 + sg0 + delta; Q11[j] = h1 + sg1 + delta; Q22[j] = h2 + delta;
        }
        const double v = w.v[j], om = w.om[j], cs = w.cs[j], sn = w.sn[j];
        a13[j] = -T * v * sn; a23[j] = T * v * cs; b11[j] = T * cs; b21[j] = T * sn;
        double gv, hvv;
        vcost(c, df, v, &gv, &hvv);
        const double gw = df * 2.0 * c.Ww * om;
        double hww = df * 2.0 * c.Ww;
        htv[j] = 0.0;
        if (lsq) {
            qv[j] = -(gv - w.zLv[j] + w.zUv[j]); qw[j] = -(gw - w.zLw[j] + w.zUw[j]);
            dv[j] = 1.0; dw[j] = 1.0;
            e.a[j] = e.b[j] = e.c[j] = 0.0;
        } else {
            double sgv, rbv, sgw, rbw;
            bound_terms(v, c.lb[2], c.ub[2], c.hasL[2], c.hasU[2], w.zLv[j], w.zUv[j], mu, &sgv, &rbv);
            bound_terms(om, c.lb[3], c.ub[3], c.hasL[3], c.hasU[3], w.zLw[j], w.zUw[j], mu, &sgw, &rbw);
            if (s < N) {
                // J^T y of dynamics row s+1 and the curvature of the dynamics in the Lagrangian
                q0[j] -= yn0[j]; q1[j] -= yn1[j]; q2[j] -= a13[j] * yn0[j] + a23[j] * yn1[j] + yn2[j];
                Q22[j] += T * v * (yn0[j] * cs + yn1[j] * sn);
                htv[j] = T * (yn0[j] * sn - yn1[j] * cs);
            }
            qv[j] = gv - (b11[j] * yn0[j] + b21[j] * yn1[j]) + rbv;
            qw[j] = gw - T * yn2[j] + rbw;
            dv[j] = hvv + (sgv + delta); dw[j] = hww + (sgw + delta);
            if (soc) { e.a[j] = -cn0[j]; e.b[j] = -cn1[j]; e.c[j] = -cn2[j]; }
            else { e.a[j] = -(xn0[j] - (x0 + T * v * cs)); e.b[j] = -(xn1[j] - (x1 + T * v * sn)); e.c[j] = -(xn2[j] - (x2 + T * om)); }
        }
    }
    // Stages without a control (the terminal stage N and the padding stages behind it) become pass-through steps:
    // zero dynamics, unit Quu, zero rhs -> P_out = P_in + Q, p_out = p_in + q; the padding stages carry Q = q = 0.
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        const int s = lane * SPL + j;
        if (s >= N) {
            a13[j] = a23[j] = b11[j] = b21[j] = 0.0; qv[j] = qw[j] = htv[j] = 0.0; dv[j] = dw[j] = 1.0;
            e.a[j] = e.b[j] = e.c[j] = 0.0;
        }
        if (s > N) { q0[j] = q1[j] = q2[j] = 0.0; Q00[j] = Q11[j] = Q22[j] = 0.0; }
    }
    // Backward recursion as a warp-uniform fixed-point loop: every lane applies its stage step(s) to whatever its right
    // neighbour produced in the previous round.  The dependency is triangular, so lane top-i is exact from round i on and
    // stays exact: after top+1 rounds every lane holds its true K, k_ff, P, p -- no branches, no predicated commits.
    // (The last lane's last slot is always a padding stage -- the launcher guarantees N + 1 < 32 * SPL -- so the chain
    // end feeds zeros.)
    double C00 = 0, C10 = 0, C11 = 0, C20 = 0, C21 = 0, C22 = 0, c0 = 0, c1 = 0, c2 = 0;  // this lane's slot-0 (P, p)
    const int top = N / SPL;
    bool pd = true;
#pragma unroll 1
    for (int it = 0; it <= top; ++it) {
        double P00 = w_down(C00, 1), P10 = w_down(C10, 1), P11 = w_down(C11, 1), P20 = w_down(C20, 1), P21 = w_down(C21, 1),
               P22 = w_down(C22, 1), p0 = w_down(c0, 1), p1 = w_down(c1, 1), p2 = w_down(c2, 1);
        pd = true;
#pragma unroll
        for (int j = SPL - 1; j >= 0; --j) {
            RicK rk;
            pd &= riccati_step(P00, P10, P11, P20, P21, P22, p0, p1, p2, a13[j], a23[j], b11[j], b21[j], T, Q00[j], 0.0, Q11[j], Q22[j],
                               q0[j], q1[j], q2[j], qv[j], qw[j], dv[j], dw[j], htv[j], e.a[j], e.b[j], e.c[j], rk);
            f.K00[j] = rk.K00; f.K01[j] = rk.K01; f.K02[j] = rk.K02; f.K10[j] = rk.K10; f.K11[j] = rk.K11; f.K12[j] = rk.K12;
            f.kf0[j] = rk.kf0; f.kf1[j] = rk.kf1;
            f.P00[j] = P00; f.P10[j] = P10; f.P11[j] = P11; f.P20[j] = P20; f.P21[j] = P21; f.P22[j] = P22;
            f.pv0[j] = p0; f.pv1[j] = p1; f.pv2[j] = p2;
        }
        C00 = P00; C10 = P10; C11 = P11; C20 = P20; C21 = P21; C22 = P22; c0 = p0; c1 = p1; c2 = p2;
        // wrong inertia detected by a lane that is already exact: stop early (looked at every 8th round)
        if ((it & 7) == 7 && w_any(!pd && lane >= top - it)) return false;
    }
    return w_all(pd);
}

// ---- ROLL-OUT: forward substitution lane by lane, then stage-parallel step-size limits ----
template <int SPL>
KMPC_WN inline void w_rollout(const Cfg &c, const Ctx &t, const WState<SPL> &w, const WFact<SPL> &f, const WVec3<SPL> &e,
                              const WLin<SPL> &lin, const WVec3<SPL> &csoc, const double (&xc)[3], const double (&gl)[3], WStep<SPL> &d,
                              double *alpha_pr, double *alpha_du, double *gBD, double *ymax) {
    const int N = c.N, lane = w_lane();
    const bool lsq = t.mode == M_LSQ, soc = t.mode == M_SOC;
    const double mu = t.mu, df = t.df, T = c.T, tau = t.tau;
    // forward substitution, again as a warp-uniform fixed-point loop (lane i is exact from round i on)
    double i0, i1, i2;  // dx of stage 0
    if (lsq) { i0 = i1 = i2 = 0.0; }
    else if (soc) { i0 = -csoc.a[0]; i1 = -csoc.b[0]; i2 = -csoc.c[0]; }
    else { i0 = -(w.x0[0] - xc[0]); i1 = -(w.x1[0] - xc[1]); i2 = -(w.x2[0] - xc[2]); }
    double r0 = 0, r1 = 0, r2 = 0;  // dx of the stage after this lane's last slot
    const int top = N / SPL;
#pragma unroll 1
    for (int it = 0; it <= top; ++it) {
        const double u0 = w_up(r0, 1), u1 = w_up(r1, 1), u2 = w_up(r2, 1);
        double d0 = lane == 0 ? i0 : u0, d1 = lane == 0 ? i1 : u1, d2 = lane == 0 ? i2 : u2;
#pragma unroll
        for (int j = 0; j < SPL; ++j) {
            d.dx0[j] = d0; d.dx1[j] = d1; d.dx2[j] = d2;
            const double du0 = fma(f.K00[j], d0, fma(f.K01[j], d1, fma(f.K02[j], d2, f.kf0[j])));
            const double du1 = fma(f.K10[j], d0, fma(f.K11[j], d1, fma(f.K12[j], d2, f.kf1[j])));
            d.du0[j] = du0; d.du1[j] = du1;
            const double n0 = d0 + lin.a13[j] * d2 + lin.b11[j] * du0 + e.a[j];
            const double n1 = d1 + lin.a23[j] * d2 + lin.b21[j] * du0 + e.b[j];
            const double n2 = d2 + T * du1 + e.c[j];
            d0 = n0; d1 = n1; d2 = n2;
        }
        r0 = d0; r1 = d1; r2 = d2;
    }
    double apr = 1.0, adu = 1.0, gbd = 0.0, ym = 0.0;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        const int s = lane * SPL + j;
        if (s > N) { d.dx0[j] = d.dx1[j] = d.dx2[j] = d.du0[j] = d.du1[j] = d.dy0[j] = d.dy1[j] = d.dy2[j] = 0.0; continue; }
        const double d0 = d.dx0[j], d1 = d.dx1[j], d2 = d.dx2[j];
        const double dy0 = -(f.P00[j] * d0 + f.P10[j] * d1 + f.P20[j] * d2 + f.pv0[j]);
        const double dy1 = -(f.P10[j] * d0 + f.P11[j] * d1 + f.P21[j] * d2 + f.pv1[j]);
        const double dy2 = -(f.P20[j] * d0 + f.P21[j] * d1 + f.P22[j] * d2 + f.pv2[j]);
        d.dy0[j] = dy0; d.dy1[j] = dy1; d.dy2[j] = dy2;
        ym = maxabs_nan(maxabs_nan(maxabs_nan(ym, dy0), dy1), dy2);
        if (lsq) continue;
        const double x0 = w.x0[j], x1 = w.x1[j], x2 = w.x2[j];
        double sg, rb;
        bound_ftb(x0, d0, c.lb[0], c.ub[0], c.hasL[0], c.hasU[0], w.zLx[j], w.zUx[j], mu, tau, &apr, &adu);
        bound_ftb(x1, d1, c.lb[1], c.ub[1], c.hasL[1], c.hasU[1], w.zLy[j], w.zUy[j], mu, tau, &apr, &adu);
        const bool ing = s >= c.gk_lo && s <= c.gk_hi;
        bound_terms(x0, c.lb[0], c.ub[0], c.hasL[0], c.hasU[0], w.zLx[j], w.zUx[j], mu, &sg, &rb);
        gbd += ((ing ? df * 2.0 * c.W[0] * (x0 - gl[0]) : 0.0) + rb) * d0;
        bound_terms(x1, c.lb[1], c.ub[1], c.hasL[1], c.hasU[1], w.zLy[j], w.zUy[j], mu, &sg, &rb);
        gbd += ((ing ? df * 2.0 * c.W[1] * (x1 - gl[1]) : 0.0) + rb) * d1;
        gbd += (ing ? df * 2.0 * c.W[2] * (x2 - gl[2]) : 0.0) * d2;
        if (s < N) {
            const double v = w.v[j], om = w.om[j], du0 = d.du0[j], du1 = d.du1[j];
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

// ---- TRIAL + speculative update (all stages at once) ----
template <int SPL>
KMPC_WN inline bool w_trial(const Cfg &c, const Ctx &t, const WState<SPL> &w, const WStep<SPL> &d, const double (&xc)[3],
                            const double (&gl)[3], double alpha, double ay, double adu, bool clamp, WState<SPL> &n,
                            WVec3<SPL> &ct, Stats *out) {
    const int N = c.N, lane = w_lane();
    const double mu = t.mu, df = t.df, T = c.T;
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
        if (s > N) { ct.a[j] = ct.b[j] = ct.c[j] = 0.0; continue; }
        const double x0 = n.x0[j], x1 = n.x1[j], x2 = n.x2[j];
        const double c0 = x0 - (s == 0 ? xc[0] : pp0[j]), c1 = x1 - (s == 0 ? xc[1] : pp1[j]), c2 = x2 - (s == 0 ? xc[2] : pp2[j]);
        ct.a[j] = c0; ct.b[j] = c1; ct.c[j] = c2;
        st.theta += fabs(c0) + fabs(c1) + fabs(c2);
        st.pinf = maxabs_nan(maxabs_nan(maxabs_nan(st.pinf, c0), c1), c2);
        st.sumy += fabs(n.y0[j]) + fabs(n.y1[j]) + fabs(n.y2[j]);
        st.wmax = fmax(st.wmax, fmax(fabs(x0), fmax(fabs(x1), fabs(x2))));
        double r0 = n.y0[j], r1 = n.y1[j], r2 = n.y2[j];
        if (s >= c.gk_lo && s <= c.gk_hi) {
            const double e0 = x0 - gl[0], e1 = x1 - gl[1], e2 = x2 - gl[2];
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
KMPC_WN inline void w_soc_rhs(const Cfg &c, const WState<SPL> &w, const WVec3<SPL> &ct, const double (&xc)[3], double al, bool first,
                              WVec3<SPL> &csoc) {
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
        if (s > N) { csoc.a[j] = csoc.b[j] = csoc.c[j] = 0.0; continue; }
        double b0, b1, b2;
        if (first) { b0 = w.x0[j] - (s == 0 ? xc[0] : pp0[j]); b1 = w.x1[j] - (s == 0 ? xc[1] : pp1[j]); b2 = w.x2[j] - (s == 0 ? xc[2] : pp2[j]); }
        else { b0 = csoc.a[j]; b1 = csoc.b[j]; b2 = csoc.c[j]; }
        csoc.a[j] = al * b0 + ct.a[j]; csoc.b[j] = al * b1 + ct.b[j]; csoc.c[j] = al * b2 + ct.c[j];
    }
}

template <int SPL>
KMPC_W void w_step_select(WStep<SPL> &o, const WStep<SPL> &a, const WStep<SPL> &b, bool pick_b) {
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        o.dx0[j] = pick_b ? b.dx0[j] : a.dx0[j]; o.dx1[j] = pick_b ? b.dx1[j] : a.dx1[j]; o.dx2[j] = pick_b ? b.dx2[j] : a.dx2[j];
        o.du0[j] = pick_b ? b.du0[j] : a.du0[j]; o.du1[j] = pick_b ? b.du1[j] : a.du1[j];
        o.dy0[j] = pick_b ? b.dy0[j] : a.dy0[j]; o.dy1[j] = pick_b ? b.dy1[j] : a.dy1[j]; o.dy2[j] = pick_b ? b.dy2[j] : a.dy2[j];
    }
}

// ---- persistent worker: one warp pulls instances from a queue and solves each start to finish. ----
// filt: 2*K_FILTER_CAP doubles of warp-private scratch.  The warps of a block walk through the three phases of a trip in
// step (block barriers): at any time they execute the same few KB of code, so the instruction cache is shared instead of
// being thrashed by warps that sit in different phases of a ~140 KB kernel.
template <int SPL>
KMPC_WN inline void w_worker(const Cfg &c, const IO &io, double *filt, int *queue, unsigned long long *trips_total) {
    const int N = c.N, lane = w_lane();
    Ctx t;
    WState<SPL> cur, tri;
    WStep<SPL> st0, st1, act;
    WFact<SPL> fact;
    WVec3<SPL> e, csoc, ct;
    WLin<SPL> lin;
    double xc[3] = {0, 0, 0}, gl[3] = {0, 0, 0};
    bool have = false;
    int b = -1;
    t.mode = M_DONE;
#pragma unroll 1
    for (;;) {
        if (!have) {
            b = w_fetch(queue);
            if (b < c.B) {
#pragma unroll
                for (int j = 0; j < SPL; ++j) { csoc.a[j] = csoc.b[j] = csoc.c[j] = 0.0; ct.a[j] = ct.b[j] = ct.c[j] = 0.0; }
                w_init<SPL>(c, t, io, b, cur, xc, gl);
                have = true;
            }
        }
        if (!w_block_any(have)) break;  // also the barrier in front of phase 1
        int status = 100;
        // ---- phase 1: backward sweep ----
        const bool do_sweep = have && t.mode != M_TRIAL;
        bool ok = false;
        if (do_sweep) { t.trips++; ok = w_sweep<SPL>(c, t, cur, csoc, gl, fact, e, lin); }
        w_block_sync();
        // ---- phase 2: roll-out + line-search set-up (or inertia correction) ----
        bool go_trial = have && !do_sweep;
        if (do_sweep) {
            if (!ok) status = t.mode != M_NEWTON ? (int)ST_STEP_ERROR : inertia_update(t);  // R_RETRY: sweep again next trip
            else {
                double apr, adu, gbd, ym;
                w_rollout<SPL>(c, t, cur, fact, e, lin, csoc, xc, gl, act, &apr, &adu, &gbd, &ym);
                rollout_logic(t, apr, adu, gbd, ym);
                if (t.sel) st1 = act; else st0 = act;
                go_trial = true;
            }
        } else if (have) {
            trial_setup(t);
            w_step_select<SPL>(act, st0, st1, false);
        }
        w_block_sync();
        // ---- phase 3: trial point + acceptance logic ----
        if (go_trial) {
            Stats ts;
            const bool evok = w_trial<SPL>(c, t, cur, act, xc, gl, t.a_pr, t.a_y, t.a_du, t.tu == TU_STEP, tri, ct, &ts);
            bool aug; double ath, aph;
            const int r = trial_decide(t, filt, 1, ts, evok, &aug, &ath, &aph);
            if (aug) {  // one lane edits the warp's filter, everybody learns the new length
                if (lane == 0) filter_add(t, filt, 1, ath, aph);
                t.fn = w_bcast_i(t.fn, 0);
            }
            w_sync();
            if (r == R_SOC1 || r == R_SOC2) w_soc_rhs<SPL>(c, cur, ct, xc, t.alpha_soc, r == R_SOC1, csoc);
            else if (r == R_ACCEPT) { cur = tri; t.c = ts; status = begin_iteration(c, t); }
            else if (r != R_BACKTRACK) status = r;
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
            }
            have = false;
        }
    }
}

}  // namespace kmpc
