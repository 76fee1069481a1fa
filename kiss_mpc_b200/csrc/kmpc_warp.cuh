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
// coop area of one instance: C_NF fields x NSTG stages (field-major: the owner lanes touch consecutive stages; the
// Riccati lanes -- one per instance -- are COOP doubles apart and read / write an even-odd pair of stages with one 128-bit
// access: COOP = 2 mod 32 doubles puts 8 lanes x 16 bytes on the 32 banks, the minimum of two wavefronts per access).
// Fields 7..17 (+ Q01) are the stage blocks going in.  The backward sweep adds K and P and overwrites the blocks its vector
// part has consumed (q, qv, qw) with p and k_ff; the forward roll-out overwrites K with (dx, du).  The matrix blocks
// (Q00 Q11 Q22 Q01 dv dw htv) are never overwritten: the speculative inertia candidates read them at their own pace.
enum { C_A13 = 0, C_A23, C_B11, C_B21, C_E0, C_E1, C_E2,
       C_Q00 = 7, C_Q11, C_Q22, C_Q0, C_Q1, C_Q2, C_QV, C_QW, C_DV, C_DW, C_HTV,
       C_P00 = 18, C_P10, C_P11, C_P20, C_P21, C_P22,
       C_K00 = 24, C_K01, C_K02, C_K10, C_K11, C_K12, C_NF_BOX = 30,   // (problems without obstacle rows end here: 30 fields)
       C_Q01 = 30,                                                      // x-y coupling of the obstacle rows
       C_S00 = 31, C_S01, C_S11, C_NF = 34 };   // sum of n n^T over the obstacle rows of a stage: d(x-y block) / d(delta_w) - I
enum { C_PV0 = C_Q0, C_PV1 = C_Q1, C_PV2 = C_Q2, C_KF0 = C_QV, C_KF1 = C_QW };
#ifndef KMPC_NCAND
#define KMPC_NCAND 4  /* inertia candidates tried at once: the current delta_w and the next ones of IPOPT's sequence */
#endif
enum { C_DX0 = C_K00, C_DX1 = C_K01, C_DX2 = C_K02, C_DU0 = C_K10, C_DU1 = C_K11 };
// private area of one instance (owner warp only): the kept Newton step (back-tracking / failed corrections return to
// it), the second-order-correction rhs, the constraint values of the last trial point
// + the reciprocal slacks 1/(x - l), 1/(u - x) of the current iterate's four bounded variables (x, y, v, omega)
enum { V_RL0 = 0, V_RL1, V_RL2, V_RL3, V_RU0, V_RU1, V_RU2, V_RU3, V_NF };
// ... and its rarely-read part in GLOBAL memory (one slot per resident warp, L2-resident; written with fire-and-forget
// stores, read only by back-tracking trials and second-order corrections): the kept Newton step, the correction rhs, the
// constraint values of the last trial point.  Keeping it out of shared memory is what lets more instances fit per SM.
enum { G_DX0 = 0, G_DX1, G_DX2, G_DU0, G_DU1, G_DY0, G_DY1, G_DY2, G_CS0, G_CS1, G_CS2, G_CT0, G_CT1, G_CT2, G_NF };

// obstacle area of one instance (owner warp only; circular obstacle-distance rows of optimizer.py:198-258, README.md:78-81):
// per (field, obstacle, stage) the slack s of the row d(x_k) - s = 0, its multiplier yd, the multiplier vL of s >= I, the
// second-order-correction rhs, the row residual at the last trial point; then the circle centres: (x, y) per obstacle
// (optimizer.py:217-221), or -- stage-wise centres, dynamic_obstacle.py:47-56 -- an x plane and a y plane of O * NSTG.
enum { B_S = 0, B_YD, B_VL, B_DSOC, B_DM, B_NF };

struct WScal {  // warp-uniform per-instance scalars
    Ctx t;
    double xc[3], gl[3], d0[3];
    double dshift[KMPC_NCAND];  // delta_w of candidate k minus the delta_w the stage blocks were assembled with (NaN: none)
    int pdc[KMPC_NCAND];        // candidate k has the right inertia
    int flag, ok, r, status;
    int tinfo, ncand;           // tail mode (w_worker): slot borrowing of this trip, full candidates assembled this trip
    int pred, pstat;            // inertia prediction (w_worker): element of IPOPT's perturbation sequence this sweep was assembled with (0: none), status once it is judged
    double fnear[2 * KMPC_FILTER_NEAR];   // the first entries of the instance's filter (FiltSplit); the others in its global scratch slot
};

// NST = stage slots allocated per field (>= N + 1, <= 32 * SPL): a smaller NST than 32 * SPL lets more instances fit
// OBS = false: the four fields only obstacle rows use are not allocated (1.0 / 1.7 / 2.0 KB per instance at 32 / 52 / 64 stage slots: at
// N <= 51 that is the 13th instance per SM)
template <int SPL, int NST = 32 * SPL, bool OBS = true>
struct WLay {
    static constexpr int NSTG = NST;
    static constexpr int COOP = (OBS ? C_NF : C_NF_BOX) * NSTG + 2;   // (even: the Riccati lanes move two stages per 128-bit access, w_ld2 / w_st2)
    static constexpr int PRIV = V_NF * NSTG;
    static constexpr int GFILT = G_NF * NSTG;                  // the filter of the instance sits behind the per-stage fields of the global scratch slot
    static constexpr int GPRIV = G_NF * NSTG + 2 * K_FILTER_CAP;   // doubles of global scratch per resident warp
    KMPC_HD static int obs_doubles(int O, int sw = 0) { return O > 0 ? B_NF * O * NSTG + 2 * O * (sw ? NSTG : 1) + O : 0; }   // rows, centres, radii
    static size_t bytes(int warps, int O = 0, int sw = 0) { return (size_t)warps * ((COOP + PRIV + obs_doubles(O, sw)) * sizeof(double) + sizeof(WScal)) + 16; }   // + the block's live-instance masks
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


// ---- bound helpers of the warp solver.  FULL = every bound of x, y, v, omega exists (compile-time), so the per-side
// tests fold away; otherwise hL / hU say which sides exist.  The reciprocal slacks rsL = 1/(val - l), rsU = 1/(u - val)
// are computed once per assembled iterate and reused by the step-size and trial-point passes. ----
template <bool FULL>
KMPC_W void wb_terms(double val, double lb, double ub, bool hL, bool hU, double zL, double zU, double mu, double &rsL,
                     double &rsU, double &sigma, double &rb) {
    double sg = 0.0, r = 0.0;
    rsL = 0.0; rsU = 0.0;
    if (FULL || hL) { rsL = KRCPF(val - lb); sg = zL * rsL; r = -(mu * rsL); if (!FULL && !hU) r += K_KAPPA_D * mu; }
    if (FULL || hU) { rsU = KRCPF(ub - val); sg = fma(zU, rsU, sg); r = fma(mu, rsU, r); if (!FULL && !hL) r -= K_KAPPA_D * mu; }
    sigma = sg; rb = r;
}
template <bool FULL>
KMPC_W double wb_rb(bool hL, bool hU, double rsL, double rsU, double mu) {
    double r = 0.0;
    if (FULL || hL) { r = -(mu * rsL); if (!FULL && !hU) r += K_KAPPA_D * mu; }
    if (FULL || hU) { r = fma(mu, rsU, r); if (!FULL && !hL) r -= K_KAPPA_D * mu; }
    return r;
}
// fraction-to-the-boundary as "largest relative decrease": rpr = max(-d / slack), rdu = max(-dz / z); the step limits are
// min(1, tau / rpr), min(1, tau / rdu) (one division per instance instead of one per bound side).
template <bool FULL>
KMPC_W void wb_ftb(double d, bool hL, bool hU, double zL, double zU, double rsL, double rsU, double mu, double &rpr, double &rdu) {
    if (FULL || hL) {
        rpr = kmax(-d * rsL, rpr);
        const double dz = fma(rsL, fma(-zL, d, mu), -zL);  // mu/s - z - z d/s
        rdu = kmax(-dz * KRCPF(zL), rdu);
    }
    if (FULL || hU) {
        rpr = kmax(d * rsU, rpr);
        const double dz = fma(rsU, fma(zU, d, mu), -zU);   // mu/s - z + z d/s
        rdu = kmax(-dz * KRCPF(zU), rdu);
    }
}
KMPC_W double w_ftb_alpha(double rmax, double tau) { return rmax > tau ? tau / rmax : 1.0; }
// trial-point treatment of one bounded variable (cf. bound_trial): new multipliers with the kappa_sigma safeguard,
// barrier product, damping, complementarity stats.  False if the trial value is not strictly inside its bounds.
template <bool FULL>
KMPC_W bool wb_trial(double d, double vt, double lb, double ub, bool hL, bool hU, double zL, double zU, double rsL, double rsU,
                     double mu, double adu, bool clamp, double &zLn, double &zUn, double &prod, Stats &st) {
    bool ok = true;
    zLn = 0.0; zUn = 0.0;
    if (FULL || hL) {
        const double sn = vt - lb;
        ok = sn > 0;
        prod *= sn;
        if (!FULL && !hU) st.damp += sn;
        double z = fma(adu, fma(rsL, fma(-zL, d, mu), -zL), zL);
        if (clamp) { const double mr = mu * KRCPF(sn); z = kmax(kmin(z, K_KAPPA_SIGMA * mr), mr * (1.0 / K_KAPPA_SIGMA)); }
        zLn = z;
        const double p = sn * z;
        st.mn = kmin(p, st.mn); st.mx = kmax(p, st.mx); st.sumz += fabs(z);
    }
    if (FULL || hU) {
        const double sn = ub - vt;
        ok = ok && sn > 0;
        prod *= sn;
        if (!FULL && !hL) st.damp += sn;
        double z = fma(adu, fma(rsU, fma(zU, d, mu), -zU), zU);
        if (clamp) { const double mr = mu * KRCPF(sn); z = kmax(kmin(z, K_KAPPA_SIGMA * mr), mr * (1.0 / K_KAPPA_SIGMA)); }
        zUn = z;
        const double p = sn * z;
        st.mn = kmin(p, st.mn); st.mx = kmax(p, st.mx); st.sumz += fabs(z);
    }
    return ok;
}


// ---- obstacle rows (stage-parallel: each lane loops over the O obstacles of its stage(s), s = 1..N) ----
struct WObsT { double nx, ny, ir, rr, Ds, bd, bs, rsl; };  // unit normal, 1/|p-c|, |p-c|, condensed slack block, rhs terms, 1/(s - I)
// terms of one (stage, obstacle) row at the current iterate (cf. obs_terms of kmpc_core.cuh); kind: lsq / soc / Newton
// circle centres seen from stage s of this lane: one (x, y) per obstacle, or the obstacle's own track at that stage
struct WCen { const double *p, *rp; int so, yo; };
KMPC_W WCen w_cen(const Cfg &c, const double *ob, int O, int NSTG, int s) {
    const double *base = ob + B_NF * O * NSTG;
    WCen r;
    r.p = c.obs_sw ? base + s : base; r.so = c.obs_sw ? NSTG : 2; r.yo = c.obs_sw ? O * NSTG : 1;
    r.rp = base + 2 * O * (c.obs_sw ? NSTG : 1);   // one radius per obstacle slot (optimizer.py:231-250)
    return r;
}
KMPC_W double w_cx(const WCen &cn, int o) { return cn.p[o * cn.so]; }
KMPC_W double w_cy(const WCen &cn, int o) { return cn.p[o * cn.so + cn.yo]; }
KMPC_W double w_cr(const WCen &cn, int o) { return cn.rp[o]; }

KMPC_W WObsT w_obs_terms(const Cfg &c, double px, double py, double cx, double cy, double rad, double s, double yd, double vL, double mu,
                         double delta, bool lsq, bool soc, double dsoc) {
    WObsT r;
    const double ex = px - cx, ey = py - cy;
    r.rr = sqrt(ex * ex + ey * ey);
    r.ir = KRCPF(r.rr);
    r.nx = ex * r.ir; r.ny = ey * r.ir;
    r.rsl = KRCPF(s - c.dL);
    if (lsq) { r.Ds = 1.0; r.bd = 0.0; r.bs = -yd - vL; }
    else {
        r.Ds = fma(vL, r.rsl, delta);
        r.bs = yd + mu * r.rsl - K_KAPPA_D * mu;
        r.bd = soc ? -dsoc : -((r.rr - rad) - s);
    }
    return r;
}
// slack / multiplier values of one row at the trial point (used by the trial pass and, again, to commit an accepted point)
struct WObsV { double s, yd, z, sln; };
KMPC_W WObsV w_obs_vals(const Cfg &c, const WObsT &ot, double s, double yd, double vL, double dxn, double mu, double alpha,
                        double ay, double adu, bool clamp) {
    WObsV v;
    const double ds = dxn - ot.bd;                 // n^T dx - bd
    const double dyd = fma(ot.Ds, ds, -ot.bs);
    v.s = fma(alpha, ds, s);
    v.sln = v.s - c.dL;
    v.yd = fma(ay, dyd, yd);
    double z = fma(adu, fma(ot.rsl, fma(-vL, ds, mu), -vL), vL);   // vL + adu (mu/sl - vL - vL ds/sl)
    if (clamp) { const double mr = mu * KRCPF(v.sln); z = kmax(kmin(z, K_KAPPA_SIGMA * mr), mr * (1.0 / K_KAPPA_SIGMA)); }
    v.z = z;
    return v;
}

// ---- starting point: optimizer.py:375-385 (warm start) / agent.py:59-60 (cold start); IPOPT initialisation ----
template <int SPL, int NST, bool OBS>
KMPC_WN inline void w_init(const Cfg &c, WScal *sc, const IO &io, int b, WState<SPL> &w, double *ob) {
    constexpr int NSTG = WLay<SPL, NST>::NSTG;
    const int N = c.N, lane = w_lane(), O = OBS ? c.O : 0;
    if (OBS) {  // circle centres -> shared memory (optimizer.py:217-221)
        double *cxy = ob + B_NF * O * NSTG;
        if (c.obs_sw) {  // track of every obstacle, column t for stage t + 1 (dynamic_obstacle.py:47-56)
            for (int o = 0; o < O; ++o)
                for (int t = lane; t < N; t += 32) {
                    cxy[o * NSTG + t + 1] = io.obs[io_obs_sw(c, b, o, t, 0)];
                    cxy[(O + o) * NSTG + t + 1] = io.obs[io_obs_sw(c, b, o, t, 1)];
                }
        } else {
            for (int i = lane; i < 2 * O; i += 32) cxy[i] = io.obs[io_obs(c, b, i >> 1, i & 1)];
        }
        double *rad = cxy + 2 * O * (c.obs_sw ? NSTG : 1);
        for (int o = lane; o < O; o += 32) rad[o] = io.orad ? io.orad[io_orad(c, b, o)] : c.obs_radius;
        w_sync();
    }
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
        if (OBS && s >= 1 && s <= N) {  // slacks pushed inside their bound, yd = 0, vL = 1
            const WCen cen = w_cen(c, ob, O, NSTG, s);
            const double dLpush = c.dL + K_BOUND_PUSH * fmax(1.0, fabs(c.dL));
            for (int o = 0; o < O; ++o) {
                const double ex = x[0] - w_cx(cen, o), ey = x[1] - w_cy(cen, o);
                const double d = sqrt(ex * ex + ey * ey) - w_cr(cen, o);
                double *po = ob + o * NSTG + s;
                po[B_S * O * NSTG] = fmax(d, dLpush); po[B_YD * O * NSTG] = 0.0; po[B_VL * O * NSTG] = 1.0;
            }
        }
    }
    gm = w_max_nn(gm);
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
// zero rhs -> P_out = P_in + Q, p_out = p_in + q.  Also leaves the reciprocal slacks of the iterate in the private area.
template <int SPL, int NST, bool FULL, bool OBS>
KMPC_WN inline void w_assemble(const Cfg &c, WScal *sc, const WState<SPL> &w, double *priv, const double *gp, double *coop, const double *ob) {
    constexpr int NSTG = WLay<SPL, NST>::NSTG;
    const int N = c.N, lane = w_lane(), O = OBS ? c.O : 0;
    const int mode = sc->t.mode;
    const bool lsq = mode == M_LSQ, soc = mode == M_SOC;
    const double mu = sc->t.mu, delta = sc->t.delta, df = sc->t.df, T = c.T;
    const double gl0 = sc->gl[0], gl1 = sc->gl[1], gl2 = sc->gl[2];
    const bool hL0 = c.hasL[0], hU0 = c.hasU[0], hL1 = c.hasL[1], hU1 = c.hasU[1], hL2 = c.hasL[2], hU2 = c.hasU[2], hL3 = c.hasL[3], hU3 = c.hasU[3];
    double yn0[SPL], yn1[SPL], yn2[SPL], xn0[SPL], xn1[SPL], xn2[SPL];
    w_next<SPL>(w.y0, yn0); w_next<SPL>(w.y1, yn1); w_next<SPL>(w.y2, yn2);
    w_next<SPL>(w.x0, xn0); w_next<SPL>(w.x1, xn1); w_next<SPL>(w.x2, xn2);
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        const int s = lane * SPL + j;
        if (s > N) continue;
        const double x0 = w.x0[j], x1 = w.x1[j], x2 = w.x2[j];
        const double v = w.v[j], om = w.om[j], cs = w.cs[j], sn = w.sn[j];
        const bool ing = s >= c.gk_lo && s <= c.gk_hi;
        double gx0 = 0, gx1 = 0, gx2 = 0, h0 = 0, h1 = 0, h2 = 0;
        if (ing) {
            gx0 = df * 2.0 * c.W[0] * (x0 - gl0); gx1 = df * 2.0 * c.W[1] * (x1 - gl1); gx2 = df * 2.0 * c.W[2] * (x2 - gl2);
            h0 = df * 2.0 * c.W[0]; h1 = df * 2.0 * c.W[1]; h2 = df * 2.0 * c.W[2];
        }
        // barrier terms of the four bounded variables (their reciprocal slacks are kept for the later passes)
        double rl0, ru0, rl1, ru1, rl2, ru2, rl3, ru3, sg0, rb0, sg1, rb1, sgv, rbv, sgw, rbw;
        wb_terms<FULL>(x0, c.lb[0], c.ub[0], hL0, hU0, w.zLx[j], w.zUx[j], mu, rl0, ru0, sg0, rb0);
        wb_terms<FULL>(x1, c.lb[1], c.ub[1], hL1, hU1, w.zLy[j], w.zUy[j], mu, rl1, ru1, sg1, rb1);
        wb_terms<FULL>(v, c.lb[2], c.ub[2], hL2, hU2, w.zLv[j], w.zUv[j], mu, rl2, ru2, sgv, rbv);
        wb_terms<FULL>(om, c.lb[3], c.ub[3], hL3, hU3, w.zLw[j], w.zUw[j], mu, rl3, ru3, sgw, rbw);
        double *pv = priv + s;
        pv[V_RL0 * NSTG] = rl0; pv[V_RU0 * NSTG] = ru0; pv[V_RL1 * NSTG] = rl1; pv[V_RU1 * NSTG] = ru1;
        pv[V_RL2 * NSTG] = rl2; pv[V_RU2 * NSTG] = ru2; pv[V_RL3 * NSTG] = rl3; pv[V_RU3 * NSTG] = ru3;
        double q0, q1, q2, Q00, Q11, Q22;
        if (lsq) {
            q0 = -(gx0 - w.zLx[j] + w.zUx[j]); q1 = -(gx1 - w.zLy[j] + w.zUy[j]); q2 = -gx2;
            Q00 = 1.0; Q11 = 1.0; Q22 = 1.0;
        } else {
            q0 = gx0 + w.y0[j] + rb0; q1 = gx1 + w.y1[j] + rb1; q2 = gx2 + w.y2[j];
            Q00 = KADD(KADD(h0, sg0), delta); Q11 = KADD(KADD(h1, sg1), delta); Q22 = KADD(h2, delta);   // (uncontracted: w_assemble_cands forms the same sums)
        }
        double Q01 = 0.0, S00 = 0.0, S01 = 0.0, S11 = 0.0;
        if (OBS && s >= 1) {  // slacks of the obstacle rows condensed into the x-y block
            const WCen cen = w_cen(c, ob, O, NSTG, s);
            for (int o = 0; o < O; ++o) {
                const double *po = ob + o * NSTG + s;
                const double yd = po[B_YD * O * NSTG];
                const WObsT ot = w_obs_terms(c, x0, x1, w_cx(cen, o), w_cy(cen, o), w_cr(cen, o), po[B_S * O * NSTG], yd, po[B_VL * O * NSTG], mu, delta,
                                             lsq, soc, soc ? po[B_DSOC * O * NSTG] : 0.0);
                if (!lsq) {
                    const double h = yd * ot.ir;
                    Q00 += h * (1.0 - ot.nx * ot.nx); Q01 += h * (-ot.nx * ot.ny); Q11 += h * (1.0 - ot.ny * ot.ny);
                    q0 += ot.nx * yd; q1 += ot.ny * yd;
                }
                Q00 += ot.Ds * ot.nx * ot.nx; Q01 += ot.Ds * ot.nx * ot.ny; Q11 += ot.Ds * ot.ny * ot.ny;
                S00 += ot.nx * ot.nx; S01 += ot.nx * ot.ny; S11 += ot.ny * ot.ny;   // delta_w enters every Ds
                const double tt = ot.Ds * ot.bd + ot.bs;
                q0 -= ot.nx * tt; q1 -= ot.ny * tt;
            }
        }
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
            if (s < N) {
                // J^T y of dynamics row s+1 and the curvature of the dynamics in the Lagrangian
                q0 -= yn0[j]; q1 -= yn1[j]; q2 -= a13 * yn0[j] + a23 * yn1[j] + yn2[j];
                Q22 = KADD(Q22, KMUL(KMUL(T, v), fma(yn0[j], cs, KMUL(yn1[j], sn))));
                htv = T * (yn0[j] * sn - yn1[j] * cs);
            }
            qv = gv - (b11 * yn0[j] + b21 * yn1[j]) + rbv;
            qw = gw - T * yn2[j] + rbw;
            dv = KADD(hvv, KADD(sgv, delta)); dw = KADD(hww, KADD(sgw, delta));
            if (soc) {  // rhs of the dynamics row s+1 = -c_soc of stage s+1
                const double *pn = gp + (s + 1 < NSTG ? s + 1 : s);
                e0 = -pn[G_CS0 * NSTG]; e1 = -pn[G_CS1 * NSTG]; e2 = -pn[G_CS2 * NSTG];
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
        if (OBS) { q[C_Q01 * NSTG] = Q01; q[C_S00 * NSTG] = S00; q[C_S01 * NSTG] = S01; q[C_S11 * NSTG] = S11; }
        q[C_Q0 * NSTG] = q0; q[C_Q1 * NSTG] = q1; q[C_Q2 * NSTG] = q2;
        q[C_QV * NSTG] = qv; q[C_QW * NSTG] = qw; q[C_DV * NSTG] = dv; q[C_DW * NSTG] = dw; q[C_HTV * NSTG] = htv;
        if (s == 0) {  // dx of stage 0 (the rhs of the initial-state row)
            if (lsq) { sc->d0[0] = sc->d0[1] = sc->d0[2] = 0.0; }
            else if (soc) { sc->d0[0] = -gp[G_CS0 * NSTG]; sc->d0[1] = -gp[G_CS1 * NSTG]; sc->d0[2] = -gp[G_CS2 * NSTG]; }
            else { sc->d0[0] = -(x0 - sc->xc[0]); sc->d0[1] = -(x1 - sc->xc[1]); sc->d0[2] = -(x2 - sc->xc[2]); }
        }
    }
}

// Full-solve inertia candidates (tail mode, w_worker): the same Newton system with the NEXT perturbations delta_w of IPOPT's
// sequence, written next to the base system into instance slots that are free, so that the Riccati warp solves all of them side
// by side.  Runs after w_assemble on the same iterate: the entries that do not depend on delta_w are copied from the base
// blocks, the ones that do are formed again exactly as w_assemble forms them (same operands -- the reciprocal slacks it left in
// the private area -- same order), hence a candidate's solve has the bits of the retry trip it replaces.  No obstacle rows.
struct WCands { int n; double *coop[KMPC_NCAND > 1 ? KMPC_NCAND - 1 : 1]; double delta[KMPC_NCAND > 1 ? KMPC_NCAND - 1 : 1]; WScal *sc[KMPC_NCAND > 1 ? KMPC_NCAND - 1 : 1]; };
template <int SPL, int NST, bool FULL>
KMPC_WN inline void w_assemble_cands(const Cfg &c, const WScal *sc, const WState<SPL> &w, const double *priv, const double *coop, const WCands &cand) {
    constexpr int NSTG = WLay<SPL, NST>::NSTG;
    const int N = c.N, lane = w_lane();
    const double df = sc->t.df, T = c.T;
    const bool hL0 = c.hasL[0], hU0 = c.hasU[0], hL1 = c.hasL[1], hU1 = c.hasU[1], hL2 = c.hasL[2], hU2 = c.hasU[2], hL3 = c.hasL[3], hU3 = c.hasU[3];
    double yn0[SPL], yn1[SPL];
    w_next<SPL>(w.y0, yn0); w_next<SPL>(w.y1, yn1);
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        const int s = lane * SPL + j;
        if (s > N) continue;
        const double v = w.v[j], cs = w.cs[j], sn = w.sn[j];
        const bool ing = s >= c.gk_lo && s <= c.gk_hi;
        const double h0 = ing ? df * 2.0 * c.W[0] : 0.0, h1 = ing ? df * 2.0 * c.W[1] : 0.0, h2 = ing ? df * 2.0 * c.W[2] : 0.0;
        const double *pv = priv + s;
        // sigma of a bounded variable as wb_terms forms it: zL * rsL, then fma(zU, rsU, .)
        auto sigma = [](bool hL, bool hU, double zL, double zU, double rsL, double rsU) {
            double sg = 0.0;
            if (FULL || hL) sg = zL * rsL;
            if (FULL || hU) sg = fma(zU, rsU, sg);
            return sg;
        };
        const double sg0 = sigma(hL0, hU0, w.zLx[j], w.zUx[j], pv[V_RL0 * NSTG], pv[V_RU0 * NSTG]);
        const double sg1 = sigma(hL1, hU1, w.zLy[j], w.zUy[j], pv[V_RL1 * NSTG], pv[V_RU1 * NSTG]);
        const double sgv = sigma(hL2, hU2, w.zLv[j], w.zUv[j], pv[V_RL2 * NSTG], pv[V_RU2 * NSTG]);
        const double sgw = sigma(hL3, hU3, w.zLw[j], w.zUw[j], pv[V_RL3 * NSTG], pv[V_RU3 * NSTG]);
        double gv, hvv;
        vcost(c, df, v, &gv, &hvv);
        const double hww = df * 2.0 * c.Ww;
        const double curv = KMUL(KMUL(T, v), fma(yn0[j], cs, KMUL(yn1[j], sn)));
        const double *q = coop + s;
#pragma unroll
        for (int k = 0; k < (KMPC_NCAND > 1 ? KMPC_NCAND - 1 : 1); ++k) {
            if (k >= cand.n) break;
            const double dk = cand.delta[k];
            double *qk = cand.coop[k] + s;
            qk[C_A13 * NSTG] = q[C_A13 * NSTG]; qk[C_A23 * NSTG] = q[C_A23 * NSTG]; qk[C_B11 * NSTG] = q[C_B11 * NSTG]; qk[C_B21 * NSTG] = q[C_B21 * NSTG];
            qk[C_E0 * NSTG] = q[C_E0 * NSTG]; qk[C_E1 * NSTG] = q[C_E1 * NSTG]; qk[C_E2 * NSTG] = q[C_E2 * NSTG];
            qk[C_Q0 * NSTG] = q[C_Q0 * NSTG]; qk[C_Q1 * NSTG] = q[C_Q1 * NSTG]; qk[C_Q2 * NSTG] = q[C_Q2 * NSTG];
            qk[C_QV * NSTG] = q[C_QV * NSTG]; qk[C_QW * NSTG] = q[C_QW * NSTG]; qk[C_HTV * NSTG] = q[C_HTV * NSTG];
            // w_assemble: Q00 = h0 + sg0 + delta; Q22 = h2 + delta (+= curvature for s < N); dv = hvv + (sgv + delta); terminal stage: unit Quu
            qk[C_Q00 * NSTG] = KADD(KADD(h0, sg0), dk); qk[C_Q11 * NSTG] = KADD(KADD(h1, sg1), dk);
            qk[C_Q22 * NSTG] = s < N ? KADD(KADD(h2, dk), curv) : KADD(h2, dk);
            qk[C_DV * NSTG] = s < N ? KADD(hvv, KADD(sgv, dk)) : 1.0; qk[C_DW * NSTG] = s < N ? KADD(hww, KADD(sgw, dk)) : 1.0;
            if (s == 0) { cand.sc[k]->d0[0] = sc->d0[0]; cand.sc[k]->d0[1] = sc->d0[1]; cand.sc[k]->d0[2] = sc->d0[2]; }
        }
    }
}

// two consecutive stages (an even one and the next) of one field in one shared-memory access.  The serial warp is bound by
// instruction issue, and a 64-bit shared-memory access costs it ~4 issue cycles (ncu, r02d): half as many accesses.
#ifdef __CUDA_ARCH__
KMPC_W void w_ld2(const double *p, double &lo, double &hi) { const double2 v = *reinterpret_cast<const double2 *>(p); lo = v.x; hi = v.y; }
KMPC_W void w_st2(double *p, double lo, double hi) { *reinterpret_cast<double2 *>(p) = make_double2(lo, hi); }
#else
KMPC_W void w_ld2(const double *p, double &lo, double &hi) { lo = p[0]; hi = p[1]; }
KMPC_W void w_st2(double *p, double lo, double hi) { p[0] = lo; p[1] = hi; }
#endif
// ---- phase 1b, RICCATI: the two serial recursions of ONE instance, executed by ONE lane of the block's Riccati warp ----
// Backward sweep (K, k_ff, P, p of every stage; false = some Q_uu not positive definite = wrong inertia), then the
// forward substitution dx+ = A dx + B du + e, du = K dx + k_ff.  Same algebra as riccati_step (kmpc_core.cuh) for the
// unicycle stage A = I + a13 e1 e3^T + a23 e2 e3^T, B = [b11 0; b21 0; 0 T].  The lane is alone on its dependency chain,
// so the loop is software-pipelined: iteration s runs the MATRIX part of stage s (P, K: the long chain through the
// Q_uu inverse) together with the independent VECTOR part of stage s+1 (p, k_ff: a short chain), and no sign flips are
// left on either chain.
struct WRicCarry { double P00, P10, P11, P20, P21, P22, K00, K01, K02, K10, K11, K12, m00, m01, m11, a13, a23, b11, b21; };
// vector part of stage k (q = coop + k): uses the cost-to-go P of stage k+1 (in cy) and p of stage k+1 (p0..p2, updated)
KMPC_W void w_ric_vec(const WRicCarry &cy, const double *q, double *qs, const int NSTG, const double T, double &p0, double &p1, double &p2) {
    const double e0 = q[C_E0 * NSTG], e1 = q[C_E1 * NSTG], e2 = q[C_E2 * NSTG];
    const double q0 = q[C_Q0 * NSTG], q1 = q[C_Q1 * NSTG], q2 = q[C_Q2 * NSTG], qv = q[C_QV * NSTG], qw = q[C_QW * NSTG];
    const double Pe0 = fma(cy.P00, e0, fma(cy.P10, e1, cy.P20 * e2)) + p0, Pe1 = fma(cy.P10, e0, fma(cy.P11, e1, cy.P21 * e2)) + p1,
                 Pe2 = fma(cy.P20, e0, fma(cy.P21, e1, cy.P22 * e2)) + p2;
    const double qu0 = fma(cy.b11, Pe0, fma(cy.b21, Pe1, qv)), qu1 = fma(T, Pe2, qw);
    p0 = fma(cy.K00, qu0, fma(cy.K10, qu1, q0 + Pe0));
    p1 = fma(cy.K01, qu0, fma(cy.K11, qu1, q1 + Pe1));
    p2 = fma(cy.K02, qu0, fma(cy.K12, qu1, fma(cy.a13, Pe0, fma(cy.a23, Pe1, q2 + Pe2))));
    const double kf0 = fma(cy.m00, qu0, cy.m01 * qu1), kf1 = fma(cy.m01, qu0, cy.m11 * qu1);
    qs[C_KF0 * NSTG] = kf0; qs[C_KF1 * NSTG] = kf1; qs[C_PV0 * NSTG] = p0; qs[C_PV1 * NSTG] = p1; qs[C_PV2 * NSTG] = p2;
}
// matrix part of stage k: (P of stage k+1 in P..) -> K, M = -Quu^-1, P of stage k; leaves what the vector part needs in cy
// qs: where K, P go.  (The blocks are used as assembled: the candidates of other delta_w have their own loop.  An added shift of 0.0
// and the x-y coupling 0.0 of the rows-free problem are not folded by the compiler -- x + 0.0 is not x for x = -0.0 -- and cost six
// FP64 issue slots per stage on the one warp the whole block waits for.)
template <bool OBS>
KMPC_W bool w_ric_mat(WRicCarry &cy, double *q, const int NSTG, const double T, const double TT, double &P00, double &P10,
                      double &P11, double &P20, double &P21, double &P22, double *qs) {
    const double a13 = q[C_A13 * NSTG], a23 = q[C_A23 * NSTG], b11 = q[C_B11 * NSTG], b21 = q[C_B21 * NSTG];
    const double Q00 = q[C_Q00 * NSTG], Q11 = q[C_Q11 * NSTG], Q22 = q[C_Q22 * NSTG];
    const double dv = q[C_DV * NSTG], dw = q[C_DW * NSTG], htv = q[C_HTV * NSTG];
    cy.P00 = P00; cy.P10 = P10; cy.P11 = P11; cy.P20 = P20; cy.P21 = P21; cy.P22 = P22;
    cy.a13 = a13; cy.a23 = a23; cy.b11 = b11; cy.b21 = b21;
    // P A (third column), symmetric Qxx = A^T P A + Q
    const double PA02 = fma(P00, a13, fma(P10, a23, P20)), PA12 = fma(P10, a13, fma(P11, a23, P21)), PA22 = fma(P20, a13, fma(P21, a23, P22));
    const double X00 = P00 + Q00, X10 = OBS ? P10 + q[C_Q01 * NSTG] : P10, X11 = P11 + Q11, X20 = PA02, X21 = PA12;
    const double X22 = fma(a13, PA02, fma(a23, PA12, PA22)) + Q22;
    // Qux = B^T P A (+ W_v,theta)
    const double U00 = fma(b11, P00, b21 * P10), U01 = fma(b11, P10, b21 * P11), U02 = fma(b11, PA02, fma(b21, PA12, htv));
    const double U10 = T * P20, U11 = T * P21, U12 = T * PA22;
    // Quu = B^T P B + diag = [d1 qb; qb qc];  positive definite <=> d1 > 0 and det > 0
    const double d1 = fma(b11, U00, fma(b21, U01, dv)), qb = fma(b11, U10, b21 * U11), qc = fma(TT, P22, dw);
    const double ndet = fma(qb, qb, -(d1 * qc));  // -det
    // M = -Quu^-1 = [m00 m01; m01 m11]
    const double rn = KRCPF(ndet), m00 = qc * rn, m01 = qb * -rn, m11 = d1 * rn;
    // K = -Quu^-1 Qux
    const double K00 = fma(m00, U00, m01 * U10), K01 = fma(m00, U01, m01 * U11), K02 = fma(m00, U02, m01 * U12);
    const double K10 = fma(m01, U00, m11 * U10), K11 = fma(m01, U01, m11 * U11), K12 = fma(m01, U02, m11 * U12);
    // P <- Qxx + Qux^T K  (lower triangle; symmetric in exact arithmetic)
    P00 = fma(U00, K00, fma(U10, K10, X00)); P10 = fma(U01, K00, fma(U11, K10, X10)); P11 = fma(U01, K01, fma(U11, K11, X11));
    P20 = fma(U02, K00, fma(U12, K10, X20)); P21 = fma(U02, K01, fma(U12, K11, X21)); P22 = fma(U02, K02, fma(U12, K12, X22));
    cy.K00 = K00; cy.K01 = K01; cy.K02 = K02; cy.K10 = K10; cy.K11 = K11; cy.K12 = K12; cy.m00 = m00; cy.m01 = m01; cy.m11 = m11;
    // (plain unconditional stores: the whole stage stays in one basic block, so the compiler interleaves the vector part of
    //  the stage behind with this chain)
    qs[C_K00 * NSTG] = K00; qs[C_K01 * NSTG] = K01; qs[C_K02 * NSTG] = K02;
    qs[C_K10 * NSTG] = K10; qs[C_K11 * NSTG] = K11; qs[C_K12 * NSTG] = K12;
    qs[C_P00 * NSTG] = P00; qs[C_P10 * NSTG] = P10; qs[C_P11 * NSTG] = P11;
    qs[C_P20 * NSTG] = P20; qs[C_P21 * NSTG] = P21; qs[C_P22 * NSTG] = P22;
    return d1 > 0.0 && ndet < 0.0;
}
// inertia test only: the matrix recursion of one instance with delta_w raised by dshift (a speculative candidate of IPOPT's
// perturbation sequence); reads the assembled blocks, stores nothing.  Runs on a second warp next to the solving sweep.
template <bool OBS>
KMPC_WN inline bool w_serial_candidate(const Cfg &c, const double *coop, const int NSTG, const double dshift) {
    const int N = c.N;
    const double T = c.T, TT = T * T;
    double P00 = 0, P10 = 0, P11 = 0, P20 = 0, P21 = 0, P22 = 0;
    bool pd = true;
#pragma unroll 1
    for (int s = N; s >= 0; --s) {
        const double *q = coop + s;
        const double a13 = q[C_A13 * NSTG], a23 = q[C_A23 * NSTG], b11 = q[C_B11 * NSTG], b21 = q[C_B21 * NSTG];
        const double du = s < N ? dshift : 0.0;   // the terminal stage is a pass-through with unit Quu
        // delta_w sits on the diagonal and, with obstacle rows, in every condensed slack block Ds n n^T
        const double Q00 = fma(dshift, OBS ? 1.0 + q[C_S00 * NSTG] : 1.0, q[C_Q00 * NSTG]), Q11 = fma(dshift, OBS ? 1.0 + q[C_S11 * NSTG] : 1.0, q[C_Q11 * NSTG]);
        const double Q22 = q[C_Q22 * NSTG] + dshift, Q01 = OBS ? fma(dshift, q[C_S01 * NSTG], q[C_Q01 * NSTG]) : 0.0;
        const double dv = q[C_DV * NSTG] + du, dw = q[C_DW * NSTG] + du, htv = q[C_HTV * NSTG];
        const double PA02 = fma(P00, a13, fma(P10, a23, P20)), PA12 = fma(P10, a13, fma(P11, a23, P21)), PA22 = fma(P20, a13, fma(P21, a23, P22));
        const double X00 = P00 + Q00, X10 = P10 + Q01, X11 = P11 + Q11, X20 = PA02, X21 = PA12;
        const double X22 = fma(a13, PA02, fma(a23, PA12, PA22)) + Q22;
        const double U00 = fma(b11, P00, b21 * P10), U01 = fma(b11, P10, b21 * P11), U02 = fma(b11, PA02, fma(b21, PA12, htv));
        const double U10 = T * P20, U11 = T * P21, U12 = T * PA22;
        const double d1 = fma(b11, U00, fma(b21, U01, dv)), qb = fma(b11, U10, b21 * U11), qc = fma(TT, P22, dw);
        const double ndet = fma(qb, qb, -(d1 * qc));
        pd = pd && d1 > 0.0 && ndet < 0.0;
        const double rn = KRCPF(ndet), m00 = qc * rn, m01 = qb * -rn, m11 = d1 * rn;
        const double K00 = fma(m00, U00, m01 * U10), K01 = fma(m00, U01, m01 * U11), K02 = fma(m00, U02, m01 * U12);
        const double K10 = fma(m01, U00, m11 * U10), K11 = fma(m01, U01, m11 * U11), K12 = fma(m01, U02, m11 * U12);
        P00 = fma(U00, K00, fma(U10, K10, X00)); P10 = fma(U01, K00, fma(U11, K10, X10)); P11 = fma(U01, K01, fma(U11, K11, X11));
        P20 = fma(U02, K00, fma(U12, K10, X20)); P21 = fma(U02, K01, fma(U12, K11, X21)); P22 = fma(U02, K02, fma(U12, K12, X22));
    }
    return pd;
}
struct WFwdIn { double K00, K01, K02, K10, K11, K12, kf0, kf1, a13, a23, b11, b21, e0, e1, e2; };
KMPC_W void w_fwd_load2(WFwdIn &a, WFwdIn &b, const double *q, const int NSTG) {   // q: an even stage
    w_ld2(q + C_K00 * NSTG, a.K00, b.K00); w_ld2(q + C_K01 * NSTG, a.K01, b.K01); w_ld2(q + C_K02 * NSTG, a.K02, b.K02);
    w_ld2(q + C_K10 * NSTG, a.K10, b.K10); w_ld2(q + C_K11 * NSTG, a.K11, b.K11); w_ld2(q + C_K12 * NSTG, a.K12, b.K12);
    w_ld2(q + C_KF0 * NSTG, a.kf0, b.kf0); w_ld2(q + C_KF1 * NSTG, a.kf1, b.kf1);
    w_ld2(q + C_A13 * NSTG, a.a13, b.a13); w_ld2(q + C_A23 * NSTG, a.a23, b.a23); w_ld2(q + C_B11 * NSTG, a.b11, b.b11); w_ld2(q + C_B21 * NSTG, a.b21, b.b21);
    w_ld2(q + C_E0 * NSTG, a.e0, b.e0); w_ld2(q + C_E1 * NSTG, a.e1, b.e1); w_ld2(q + C_E2 * NSTG, a.e2, b.e2);
}
// one stage of the roll-out without the stores: (dx, du) of the stage come back in o[0..4], x0..x2 become dx of the next stage
KMPC_W void w_fwd_step(const WFwdIn &f, const double T, double &x0, double &x1, double &x2, double (&o)[5]) {
    const double du0 = fma(f.K00, x0, fma(f.K01, x1, fma(f.K02, x2, f.kf0)));
    const double du1 = fma(f.K10, x0, fma(f.K11, x1, fma(f.K12, x2, f.kf1)));
    o[0] = x0; o[1] = x1; o[2] = x2; o[3] = du0; o[4] = du1;
    const double n0 = fma(f.b11, du0, fma(f.a13, x2, x0 + f.e0));
    const double n1 = fma(f.b21, du0, fma(f.a23, x2, x1 + f.e1));
    const double n2 = fma(T, du1, x2 + f.e2);
    x0 = n0; x1 = n1; x2 = n2;
}
KMPC_W void w_fwd_load(WFwdIn &r, const double *q, const int NSTG) {
    r.K00 = q[C_K00 * NSTG]; r.K01 = q[C_K01 * NSTG]; r.K02 = q[C_K02 * NSTG];
    r.K10 = q[C_K10 * NSTG]; r.K11 = q[C_K11 * NSTG]; r.K12 = q[C_K12 * NSTG];
    r.kf0 = q[C_KF0 * NSTG]; r.kf1 = q[C_KF1 * NSTG];
    r.a13 = q[C_A13 * NSTG]; r.a23 = q[C_A23 * NSTG]; r.b11 = q[C_B11 * NSTG]; r.b21 = q[C_B21 * NSTG];
    r.e0 = q[C_E0 * NSTG]; r.e1 = q[C_E1 * NSTG]; r.e2 = q[C_E2 * NSTG];
}
// one stage of the forward roll-out: stores (dx, du) of the stage over its K, returns dx of the next stage in x0..x2
KMPC_W void w_fwd_stage(const WFwdIn &f, double *q, const int NSTG, const double T, double &x0, double &x1, double &x2) {
    const double du0 = fma(f.K00, x0, fma(f.K01, x1, fma(f.K02, x2, f.kf0)));
    const double du1 = fma(f.K10, x0, fma(f.K11, x1, fma(f.K12, x2, f.kf1)));
    q[C_DX0 * NSTG] = x0; q[C_DX1 * NSTG] = x1; q[C_DX2 * NSTG] = x2; q[C_DU0 * NSTG] = du0; q[C_DU1 * NSTG] = du1;
    // (fused form: one FMA behind du on the dependency chain of the recursion instead of a multiply and two additions)
    const double n0 = fma(f.b11, du0, fma(f.a13, x2, x0 + f.e0));
    const double n1 = fma(f.b21, du0, fma(f.a23, x2, x1 + f.e1));
    const double n2 = fma(T, du1, x2 + f.e2);
    x0 = n0; x1 = n1; x2 = n2;
}
// forward roll-out dx+ = A dx + B du + e, du = K dx + k_ff over all stages
KMPC_WN inline void w_serial_fwd(const Cfg &c, double *coop, const int NSTG, const double *d0) {
    const int N = c.N;
    const double T = c.T;
    double x0 = d0[0], x1 = d0[1], x2 = d0[2];
    // two stages (an even one and the next) per trip: their operands arrive and their results leave in 128-bit accesses.  Every
    // address is the loop base plus a constant, so the loads move freely above the stores.
    int s = 0;
#pragma unroll 1
    for (; s + 1 <= N; s += 2) {
        double *q = coop + s;
        WFwdIn fa, fb;
        double oa[5], ob[5];
        w_fwd_load2(fa, fb, q, NSTG);
        w_fwd_step(fa, T, x0, x1, x2, oa);
        w_fwd_step(fb, T, x0, x1, x2, ob);
        w_st2(q + C_DX0 * NSTG, oa[0], ob[0]); w_st2(q + C_DX1 * NSTG, oa[1], ob[1]); w_st2(q + C_DX2 * NSTG, oa[2], ob[2]);
        w_st2(q + C_DU0 * NSTG, oa[3], ob[3]); w_st2(q + C_DU1 * NSTG, oa[4], ob[4]);
    }
    if (s == N) {   // N even: the terminal stage is left
        WFwdIn fa;
        w_fwd_load(fa, coop + s, NSTG);
        w_fwd_stage(fa, coop + s, NSTG, T, x0, x1, x2);
    }
}
// the solving sweep of one instance: factors, vector part and roll-out of the system as assembled (the speculative inertia
// candidates run w_serial_candidate on another warp)
template <bool OBS>
KMPC_WN inline bool w_serial(const Cfg &c, double *coop, const int NSTG, const double *d0) {
    const int N = c.N;
    const double T = c.T, TT = T * T;
    double P00 = 0, P10 = 0, P11 = 0, P20 = 0, P21 = 0, P22 = 0, p0 = 0, p1 = 0, p2 = 0;
    WRicCarry cy;
    double *const st = coop;      // loads and stores through ONE base: the compiler can tell the fields apart and keeps hoisting
                                  // the next stage's loads above this stage's stores (a second base pointer cost 15 %)
    bool pd = w_ric_mat<OBS>(cy, coop + N, NSTG, T, TT, P00, P10, P11, P20, P21, P22, st + N);
    // two stages per trip of the loop with the carry structs swapping roles (no register copies between iterations)
    // (tried: the blocks of an even-odd stage pair loaded, and their K, P stored, with 128-bit accesses as the roll-out below does --
    //  2.6 % faster for the sweep alone, 3 % SLOWER inside the kernel, whose 128-register budget the longer live ranges do not fit)
    WRicCarry cz;
    int s = N - 1;
#pragma unroll 1
    for (; s >= 1; s -= 2) {
        pd = w_ric_mat<OBS>(cz, coop + s, NSTG, T, TT, P00, P10, P11, P20, P21, P22, st + s) && pd;
        w_ric_vec(cy, coop + s + 1, st + s + 1, NSTG, T, p0, p1, p2);
        pd = w_ric_mat<OBS>(cy, coop + s - 1, NSTG, T, TT, P00, P10, P11, P20, P21, P22, st + s - 1) && pd;
        w_ric_vec(cz, coop + s, st + s, NSTG, T, p0, p1, p2);
    }
    if (s == 0) {  // (only when the pair loop did not end on stage 0)
        pd = w_ric_mat<OBS>(cz, coop, NSTG, T, TT, P00, P10, P11, P20, P21, P22, st) && pd;
        w_ric_vec(cy, coop + 1, st + 1, NSTG, T, p0, p1, p2);
        cy = cz;
    }
    w_ric_vec(cy, coop, coop, NSTG, T, p0, p1, p2);
    if (!pd) return false;
    w_serial_fwd(c, coop, NSTG, d0);
    return true;
}

// ---- phase 2, STEP: multiplier step dy = -(P dx + p), step-size limits, directional derivative (all stages at once) ----
template <int SPL, int NST, bool FULL, bool OBS>
KMPC_WN inline void w_step(const Cfg &c, const WScal *sc, const WState<SPL> &w, const double *coop, const double *priv,
                           const double *ob, WStep<SPL> &d, double *alpha_pr, double *alpha_du, double *gBD, double *ymax) {
    constexpr int NSTG = WLay<SPL, NST>::NSTG;
    const int N = c.N, lane = w_lane(), O = OBS ? c.O : 0;
    const bool lsq = sc->t.mode == M_LSQ, soc = sc->t.mode == M_SOC;
    const double mu = sc->t.mu, df = sc->t.df, tau = sc->t.tau, delta = sc->t.delta;
    const double gl0 = sc->gl[0], gl1 = sc->gl[1], gl2 = sc->gl[2];
    const bool hL0 = c.hasL[0], hU0 = c.hasU[0], hL1 = c.hasL[1], hU1 = c.hasU[1], hL2 = c.hasL[2], hU2 = c.hasU[2], hL3 = c.hasL[3], hU3 = c.hasU[3];
    double rpr = 0.0, rdu = 0.0, gbd = 0.0, ym = 0.0;
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
        if (OBS && s >= 1) {  // slack steps ds = n^T dx - bd, multiplier steps dyd = Ds ds - bs of the obstacle rows
            const WCen cen = w_cen(c, ob, O, NSTG, s);
            for (int o = 0; o < O; ++o) {
                const double *po = ob + o * NSTG + s;
                const double vL = po[B_VL * O * NSTG];
                const WObsT ot = w_obs_terms(c, w.x0[j], w.x1[j], w_cx(cen, o), w_cy(cen, o), w_cr(cen, o), po[B_S * O * NSTG], po[B_YD * O * NSTG], vL, mu,
                                             delta, lsq, soc, soc ? po[B_DSOC * O * NSTG] : 0.0);
                const double ds = fma(ot.nx, d0, ot.ny * d1) - ot.bd;
                ym = maxabs_nan(ym, fma(ot.Ds, ds, -ot.bs));
                if (!lsq) {
                    rpr = kmax(-ds * ot.rsl, rpr);
                    const double dvl = fma(ot.rsl, fma(-vL, ds, mu), -vL);
                    rdu = kmax(-dvl * KRCPF(vL), rdu);
                    gbd += (K_KAPPA_D * mu - mu * ot.rsl) * ds;
                }
            }
        }
        if (lsq) continue;
        const double *pv = priv + s;
        const double x0 = w.x0[j], x1 = w.x1[j], x2 = w.x2[j];
        const double rl0 = pv[V_RL0 * NSTG], ru0 = pv[V_RU0 * NSTG], rl1 = pv[V_RL1 * NSTG], ru1 = pv[V_RU1 * NSTG];
        wb_ftb<FULL>(d0, hL0, hU0, w.zLx[j], w.zUx[j], rl0, ru0, mu, rpr, rdu);
        wb_ftb<FULL>(d1, hL1, hU1, w.zLy[j], w.zUy[j], rl1, ru1, mu, rpr, rdu);
        const bool ing = s >= c.gk_lo && s <= c.gk_hi;
        gbd += ((ing ? df * 2.0 * c.W[0] * (x0 - gl0) : 0.0) + wb_rb<FULL>(hL0, hU0, rl0, ru0, mu)) * d0;
        gbd += ((ing ? df * 2.0 * c.W[1] * (x1 - gl1) : 0.0) + wb_rb<FULL>(hL1, hU1, rl1, ru1, mu)) * d1;
        gbd += (ing ? df * 2.0 * c.W[2] * (x2 - gl2) : 0.0) * d2;
        if (s < N) {
            const double v = w.v[j], om = w.om[j];
            const double rl2 = pv[V_RL2 * NSTG], ru2 = pv[V_RU2 * NSTG], rl3 = pv[V_RL3 * NSTG], ru3 = pv[V_RU3 * NSTG];
            wb_ftb<FULL>(du0, hL2, hU2, w.zLv[j], w.zUv[j], rl2, ru2, mu, rpr, rdu);
            wb_ftb<FULL>(du1, hL3, hU3, w.zLw[j], w.zUw[j], rl3, ru3, mu, rpr, rdu);
            double gv, hv;
            vcost(c, df, v, &gv, &hv);
            gbd += (gv + wb_rb<FULL>(hL2, hU2, rl2, ru2, mu)) * du0;
            gbd += (df * 2.0 * c.Ww * om + wb_rb<FULL>(hL3, hU3, rl3, ru3, mu)) * du1;
        }
    }
    *alpha_pr = w_ftb_alpha(w_max_nn(rpr), tau); *alpha_du = w_ftb_alpha(w_max_nn(rdu), tau);
    *gBD = w_sum(gbd); *ymax = w_max_nn(ym);
}

// kept Newton step <-> global private area
template <int SPL, int NST>
KMPC_W void w_step_store(const WStep<SPL> &d, double *gp) {
    constexpr int NSTG = WLay<SPL, NST>::NSTG;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        if (w_lane() * SPL + j >= NSTG) continue;
        double *p = gp + w_lane() * SPL + j;
        p[G_DX0 * NSTG] = d.dx0[j]; p[G_DX1 * NSTG] = d.dx1[j]; p[G_DX2 * NSTG] = d.dx2[j]; p[G_DU0 * NSTG] = d.du0[j];
        p[G_DU1 * NSTG] = d.du1[j]; p[G_DY0 * NSTG] = d.dy0[j]; p[G_DY1 * NSTG] = d.dy1[j]; p[G_DY2 * NSTG] = d.dy2[j];
    }
}
template <int SPL, int NST>
KMPC_W void w_step_load(WStep<SPL> &d, const double *gp) {
    constexpr int NSTG = WLay<SPL, NST>::NSTG;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        if (w_lane() * SPL + j >= NSTG) { d.dx0[j] = d.dx1[j] = d.dx2[j] = d.du0[j] = d.du1[j] = d.dy0[j] = d.dy1[j] = d.dy2[j] = 0.0; continue; }
        const double *p = gp + w_lane() * SPL + j;
        d.dx0[j] = p[G_DX0 * NSTG]; d.dx1[j] = p[G_DX1 * NSTG]; d.dx2[j] = p[G_DX2 * NSTG]; d.du0[j] = p[G_DU0 * NSTG];
        d.du1[j] = p[G_DU1 * NSTG]; d.dy0[j] = p[G_DY0 * NSTG]; d.dy1[j] = p[G_DY1 * NSTG]; d.dy2[j] = p[G_DY2 * NSTG];
    }
}

// kept search direction <-> the instance's own coop area.  Between the step phase and the trial evaluations the direction (8 doubles
// per stage) does not stay in registers: at 128 registers per thread the compiler spilled it to local memory (8 STL.64 + ~10 LDL.64
// per trip, and 106 KB of stack per block do not fit the L1 left beside 185 KB of shared memory -- ncu r02e: a third of all
// long-scoreboard stalls sat on those reloads).  dx, du already are in the coop area (the roll-out left them in C_DX0.. C_DU1); dy
// goes to the fields of p, which the step phase has consumed (C_PV0..2).  The reduction scratch of w_trial (the first 7 x 33
// doubles of the area) does not reach either.
#ifndef KMPC_STEP_SMEM
#define KMPC_STEP_SMEM 1
#endif
template <int SPL, int NST>
KMPC_W void w_step_park(const WStep<SPL> &d, double *coop) {
    constexpr int NSTG = WLay<SPL, NST>::NSTG;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        if (w_lane() * SPL + j >= NSTG) continue;
        double *p = coop + w_lane() * SPL + j;
        p[C_DX0 * NSTG] = d.dx0[j]; p[C_DX1 * NSTG] = d.dx1[j]; p[C_DX2 * NSTG] = d.dx2[j]; p[C_DU0 * NSTG] = d.du0[j];
        p[C_DU1 * NSTG] = d.du1[j]; p[C_PV0 * NSTG] = d.dy0[j]; p[C_PV1 * NSTG] = d.dy1[j]; p[C_PV2 * NSTG] = d.dy2[j];
    }
}
template <int SPL, int NST>
KMPC_W void w_step_fetch(WStep<SPL> &d, const double *coop) {
    constexpr int NSTG = WLay<SPL, NST>::NSTG;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        if (w_lane() * SPL + j >= NSTG) { d.dx0[j] = d.dx1[j] = d.dx2[j] = d.du0[j] = d.du1[j] = d.dy0[j] = d.dy1[j] = d.dy2[j] = 0.0; continue; }
        const double *p = coop + w_lane() * SPL + j;
        d.dx0[j] = p[C_DX0 * NSTG]; d.dx1[j] = p[C_DX1 * NSTG]; d.dx2[j] = p[C_DX2 * NSTG]; d.du0[j] = p[C_DU0 * NSTG];
        d.du1[j] = p[C_DU1 * NSTG]; d.dy0[j] = p[C_PV0 * NSTG]; d.dy1[j] = p[C_PV1 * NSTG]; d.dy2[j] = p[C_PV2 * NSTG];
    }
}

// The bound multipliers of a trial point (8 doubles per stage) are needed again only if the point is accepted: they wait in fields of
// the coop area that are dead during the trial phase (k_ff, the Q_uu diagonal, W_v,theta, the first P entries -- read by the sweep,
// its candidates and the step phase only) instead of 16 registers of the most register-starved function of the kernel.
#ifndef KMPC_TRIZ_SMEM
#define KMPC_TRIZ_SMEM 1
#endif
#if KMPC_TRIZ_SMEM
#define TRIZ(reg, field, val) (scr + s)[(field) * NSTG] = (val)
#else
#define TRIZ(reg, field, val) (reg) = (val)
#endif
// accepted trial point -> current iterate
template <int SPL, int NST>
KMPC_W void w_accept(const Cfg &c, WState<SPL> &cur, const WState<SPL> &tri, const double *coop) {
#if KMPC_TRIZ_SMEM
    constexpr int NSTG = WLay<SPL, NST>::NSTG;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        const int s = w_lane() * SPL + j;
        cur.x0[j] = tri.x0[j]; cur.x1[j] = tri.x1[j]; cur.x2[j] = tri.x2[j]; cur.v[j] = tri.v[j]; cur.om[j] = tri.om[j];
        cur.y0[j] = tri.y0[j]; cur.y1[j] = tri.y1[j]; cur.y2[j] = tri.y2[j]; cur.cs[j] = tri.cs[j]; cur.sn[j] = tri.sn[j];
        const bool in = s <= c.N;
        const double *p = coop + (in ? s : 0);
        cur.zLx[j] = in ? p[C_QV * NSTG] : 0.0; cur.zUx[j] = in ? p[C_QW * NSTG] : 0.0; cur.zLy[j] = in ? p[C_DV * NSTG] : 0.0; cur.zUy[j] = in ? p[C_DW * NSTG] : 0.0;
        cur.zLv[j] = in ? p[C_HTV * NSTG] : 0.0; cur.zUv[j] = in ? p[C_P00 * NSTG] : 0.0; cur.zLw[j] = in ? p[C_P10 * NSTG] : 0.0; cur.zUw[j] = in ? p[C_P11 * NSTG] : 0.0;
    }
#else
    cur = tri;
#endif
}

// ---- phase 3, TRIAL: trial point + speculative multiplier update + residual norms (all stages at once) ----
template <int SPL, int NST, bool FULL, bool OBS>
KMPC_WN inline bool w_trial(const Cfg &c, const WScal *sc, const WState<SPL> &w, const WStep<SPL> &d, double alpha, double ay,
                            double adu, bool clamp, WState<SPL> &n, const double *priv, double *gp, double *scr, double *ob, Stats *out) {
    constexpr int NSTG = WLay<SPL, NST>::NSTG;
    const int N = c.N, lane = w_lane(), O = OBS ? c.O : 0;
    const double mu = sc->t.mu, df = sc->t.df, T = c.T, delta = sc->t.delta;
    const bool lsq = sc->t.mode == M_LSQ, soc = sc->t.mode == M_SOC;
    const double xc0 = sc->xc[0], xc1 = sc->xc[1], xc2 = sc->xc[2], gl0 = sc->gl[0], gl1 = sc->gl[1], gl2 = sc->gl[2];
    const bool hL0 = c.hasL[0], hU0 = c.hasU[0], hL1 = c.hasL[1], hU1 = c.hasU[1], hL2 = c.hasL[2], hU2 = c.hasU[2], hL3 = c.hasL[3], hU3 = c.hasU[3];
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
#if !KMPC_TRIZ_SMEM
        n.zLx[j] = n.zUx[j] = n.zLy[j] = n.zUy[j] = n.zLv[j] = n.zUv[j] = n.zLw[j] = n.zUw[j] = 0.0;
#endif
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
        const double *pv = priv + s;
        double *pg = gp + s;
        pg[G_CT0 * NSTG] = c0; pg[G_CT1 * NSTG] = c1; pg[G_CT2 * NSTG] = c2;
        st.theta += fabs(c0) + fabs(c1) + fabs(c2);
        st.pinf = maxabs_nan(maxabs_nan(maxabs_nan(st.pinf, c0), c1), c2);
        st.sumy += fabs(n.y0[j]) + fabs(n.y1[j]) + fabs(n.y2[j]);
        st.wmax = kmax(fabs(x0), kmax(fabs(x1), kmax(fabs(x2), st.wmax)));
        double r0 = n.y0[j], r1 = n.y1[j], r2 = n.y2[j];
        if (s >= c.gk_lo && s <= c.gk_hi) {
            const double e0 = x0 - gl0, e1 = x1 - gl1, e2 = x2 - gl2;
            st.f += c.W[0] * e0 * e0; st.f += c.W[1] * e1 * e1; st.f += c.W[2] * e2 * e2;
            r0 += df * 2.0 * c.W[0] * e0; r1 += df * 2.0 * c.W[1] * e1; r2 += df * 2.0 * c.W[2] * e2;
        }
        double prod = 1.0, zLn, zUn;
        valid &= wb_trial<FULL>(d.dx0[j], x0, c.lb[0], c.ub[0], hL0, hU0, w.zLx[j], w.zUx[j], pv[V_RL0 * NSTG], pv[V_RU0 * NSTG], mu, adu,
                                clamp, zLn, zUn, prod, st);
        TRIZ(n.zLx[j], C_QV, zLn); TRIZ(n.zUx[j], C_QW, zUn); r0 += zUn - zLn;
        valid &= wb_trial<FULL>(d.dx1[j], x1, c.lb[1], c.ub[1], hL1, hU1, w.zLy[j], w.zUy[j], pv[V_RL1 * NSTG], pv[V_RU1 * NSTG], mu, adu,
                                clamp, zLn, zUn, prod, st);
        TRIZ(n.zLy[j], C_DV, zLn); TRIZ(n.zUy[j], C_DW, zUn); r1 += zUn - zLn;
        if (OBS && s >= 1) {  // obstacle rows at the trial point
            const WCen cen = w_cen(c, ob, O, NSTG, s);
            for (int o = 0; o < O; ++o) {
                double *po = ob + o * NSTG + s;
                const double so = po[B_S * O * NSTG], ydo = po[B_YD * O * NSTG], vL = po[B_VL * O * NSTG];
                const double cx = w_cx(cen, o), cy = w_cy(cen, o);
                const WObsT ot = w_obs_terms(c, w.x0[j], w.x1[j], cx, cy, w_cr(cen, o), so, ydo, vL, mu, delta, lsq, soc, soc ? po[B_DSOC * O * NSTG] : 0.0);
                const WObsV tv = w_obs_vals(c, ot, so, ydo, vL, fma(ot.nx, d.dx0[j], ot.ny * d.dx1[j]), mu, alpha, ay, adu, clamp);
                const double ex = x0 - cx, ey = x1 - cy, rr = sqrt(ex * ex + ey * ey), ir = KRCPF(rr);
                const double dm = (rr - w_cr(cen, o)) - tv.s;
                po[B_DM * O * NSTG] = dm;
                st.theta += fabs(dm); st.pinf = maxabs_nan(st.pinf, dm);
                valid &= tv.sln > 0;
                prod *= tv.sln; st.damp += tv.sln;
                if (!(prod < 1e250 && prod > 1e-250)) { st.bar += log(prod); prod = 1.0; }   // keep the product in range
                r0 += ex * ir * tv.yd; r1 += ey * ir * tv.yd;
                st.dinf = maxabs_nan(st.dinf, -tv.yd - tv.z);
                const double p = tv.sln * tv.z;
                st.mn = kmin(p, st.mn); st.mx = kmax(p, st.mx); st.sumz += fabs(tv.z); st.sumy += fabs(tv.yd);
            }
        }
        if (s < N) {
            const double v = n.v[j], om = n.om[j], cs = n.cs[j], sn = n.sn[j];
            st.wmax = kmax(fabs(v), kmax(fabs(om), st.wmax));
            const double a13 = -T * v * sn, a23 = T * v * cs;
            r0 -= yn0[j]; r1 -= yn1[j]; r2 -= a13 * yn0[j] + a23 * yn1[j] + yn2[j];
            double gv, hv;
            vcost(c, df, v, &gv, &hv);
            double rv = gv - (T * cs * yn0[j] + T * sn * yn1[j]), rw = df * 2.0 * c.Ww * om - T * yn2[j];
            if (c.cost_mode == 0) { const double vm = v < 0 ? v : 0.0, vp = v > 0 ? v : 0.0; st.f += c.Wvn * vm * vm + c.Wvp * vp * vp; }
            else st.f += c.Wvn * (v < 0 ? v : 0.0);
            st.f += c.Ww * om * om;
            valid &= wb_trial<FULL>(d.du0[j], v, c.lb[2], c.ub[2], hL2, hU2, w.zLv[j], w.zUv[j], pv[V_RL2 * NSTG], pv[V_RU2 * NSTG], mu, adu,
                                    clamp, zLn, zUn, prod, st);
            TRIZ(n.zLv[j], C_HTV, zLn); TRIZ(n.zUv[j], C_P00, zUn); rv += zUn - zLn;
            valid &= wb_trial<FULL>(d.du1[j], om, c.lb[3], c.ub[3], hL3, hU3, w.zLw[j], w.zUw[j], pv[V_RL3 * NSTG], pv[V_RU3 * NSTG], mu, adu,
                                    clamp, zLn, zUn, prod, st);
            TRIZ(n.zLw[j], C_P10, zLn); TRIZ(n.zUw[j], C_P11, zUn); rw += zUn - zLn;
            st.dinf = maxabs_nan(maxabs_nan(st.dinf, rv), rw);
        } else {
            TRIZ(n.zLv[j], C_HTV, 0.0); TRIZ(n.zUv[j], C_P00, 0.0); TRIZ(n.zLw[j], C_P10, 0.0); TRIZ(n.zUw[j], C_P11, 0.0);
        }
        st.dinf = maxabs_nan(maxabs_nan(maxabs_nan(st.dinf, r0), r1), r2);
        st.bar += log(prod);
    }
    // Sums: every lane drops its partials into the (now dead) coop area, six lanes add up one column each (33-double
    // rows: conflict-free); max / min: two CREDUX each.  One barrier pair instead of eleven shuffle butterflies.
    Stats g;
    {
        constexpr int NS = (FULL && !OBS) ? 5 : 6;
        scr[0 * 33 + lane] = st.f; scr[1 * 33 + lane] = st.bar; scr[2 * 33 + lane] = st.theta; scr[3 * 33 + lane] = st.sumy;
        scr[4 * 33 + lane] = st.sumz;
        if (NS == 6) scr[5 * 33 + lane] = st.damp;
        w_sync();
        if (lane < NS) {
            const double *r = scr + lane * 33;
            double a0 = r[0], a1 = r[1], a2 = r[2], a3 = r[3];
#pragma unroll
            for (int i = 4; i < 32; i += 4) { a0 += r[i]; a1 += r[i + 1]; a2 += r[i + 2]; a3 += r[i + 3]; }
            scr[6 * 33 + lane] = (a0 + a1) + (a2 + a3);
        }
        w_sync();
        g.f = scr[6 * 33 + 0] * df; g.bar = scr[6 * 33 + 1]; g.theta = scr[6 * 33 + 2]; g.sumy = scr[6 * 33 + 3];
        g.sumz = scr[6 * 33 + 4]; g.damp = NS == 6 ? scr[6 * 33 + 5] : 0.0;
    }
    g.dinf = w_max_nn(st.dinf); g.pinf = w_max_nn(st.pinf); g.mn = w_min_nn(st.mn); g.mx = w_max_nn(st.mx); g.wmax = w_max_nn(st.wmax);
    if (c.nb == 0) g.mn = 0.0;
    *out = g;
    const double phi = g.f - mu * g.bar + K_KAPPA_D * mu * g.damp;
    return w_all(valid) && isfinite(phi) && isfinite(g.theta);
}

// accepted trial point: the slack / multiplier values of the obstacle rows are recomputed (same formulas, same inputs as
// in the trial pass, hence the same bits) and become current.  Must run before `cur` is overwritten.
template <int SPL, int NST>
KMPC_WN inline void w_obs_commit(const Cfg &c, const WState<SPL> &w, const WStep<SPL> &d, double mu, double delta, double alpha,
                                 double ay, double adu, bool clamp, bool lsq, bool soc, double *ob) {
    constexpr int NSTG = WLay<SPL, NST>::NSTG;
    const int N = c.N, lane = w_lane(), O = c.O;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        const int s = lane * SPL + j;
        if (s < 1 || s > N) continue;
        const WCen cen = w_cen(c, ob, O, NSTG, s);
        for (int o = 0; o < O; ++o) {
            double *po = ob + o * NSTG + s;
            const double so = po[B_S * O * NSTG], ydo = po[B_YD * O * NSTG], vL = po[B_VL * O * NSTG];
            const WObsT ot = w_obs_terms(c, w.x0[j], w.x1[j], w_cx(cen, o), w_cy(cen, o), w_cr(cen, o), so, ydo, vL, mu, delta, lsq, soc,
                                         soc ? po[B_DSOC * O * NSTG] : 0.0);
            const WObsV tv = w_obs_vals(c, ot, so, ydo, vL, fma(ot.nx, d.dx0[j], ot.ny * d.dx1[j]), mu, alpha, ay, adu, clamp);
            po[B_S * O * NSTG] = tv.s; po[B_YD * O * NSTG] = tv.yd; po[B_VL * O * NSTG] = tv.z;
        }
    }
}

// c_soc <- al * base + c(trial); base = c(current) for the first correction, else the previous c_soc
template <int SPL, int NST, bool OBS>
KMPC_WN inline void w_soc_rhs(const Cfg &c, const WScal *sc, const WState<SPL> &w, double al, bool first, double *gp, double *ob) {
    constexpr int NSTG = WLay<SPL, NST>::NSTG;
    const int N = c.N, lane = w_lane(), O = OBS ? c.O : 0;
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
        double *pv = gp + s;
        double b0, b1, b2;
        if (first) { b0 = w.x0[j] - (s == 0 ? sc->xc[0] : pp0[j]); b1 = w.x1[j] - (s == 0 ? sc->xc[1] : pp1[j]); b2 = w.x2[j] - (s == 0 ? sc->xc[2] : pp2[j]); }
        else { b0 = pv[G_CS0 * NSTG]; b1 = pv[G_CS1 * NSTG]; b2 = pv[G_CS2 * NSTG]; }
        pv[G_CS0 * NSTG] = al * b0 + pv[G_CT0 * NSTG]; pv[G_CS1 * NSTG] = al * b1 + pv[G_CT1 * NSTG]; pv[G_CS2 * NSTG] = al * b2 + pv[G_CT2 * NSTG];
        if (OBS && s >= 1) {
            const WCen cen = w_cen(c, ob, O, NSTG, s);
            for (int o = 0; o < O; ++o) {
                double *po = ob + o * NSTG + s;
                double base;
                if (first) { const double ex = w.x0[j] - w_cx(cen, o), ey = w.x1[j] - w_cy(cen, o); base = (sqrt(ex * ex + ey * ey) - w_cr(cen, o)) - po[B_S * O * NSTG]; }
                else base = po[B_DSOC * O * NSTG];
                po[B_DSOC * O * NSTG] = al * base + po[B_DM * O * NSTG];
            }
        }
    }
    w_sync();
}

// Optional phase timers (tuning builds only, -DKMPC_PHASE_TIMING): cycles per phase summed over all owner warps.
#if defined(KMPC_PHASE_TIMING) && defined(__CUDACC__)
#define KMPC_NPHASE 18
__device__ unsigned long long g_phase_cycles[KMPC_NPHASE];
#define PT_DECL long long pt_acc[KMPC_NPHASE] = {0}; long long pt_last = clock64();
#define PT(i) { const long long pt_now = clock64(); pt_acc[i] += pt_now - pt_last; pt_last = pt_now; }
#define PT_COUNT(i) pt_acc[i] += 1;
#define PT_SERIAL_BEGIN const long long pt_s0 = clock64();
#define PT_SERIAL_END { w_sync(); pt_acc[17] += clock64() - pt_s0; pt_acc[16] += 1; }   /* whole serial window of the block, per trip */
#define PT_FLUSH if (lane == 0) { for (int i = 0; i < KMPC_NPHASE; ++i) atomicAdd(&g_phase_cycles[i], (unsigned long long)pt_acc[i]); }
#else
#define PT_DECL
#define PT(i)
#define PT_COUNT(i)
#define PT_SERIAL_BEGIN
#define PT_SERIAL_END
#define PT_FLUSH
#endif

// rejected trial points re-evaluated within one trip before the instance goes round the block loop again (bounds what the
// other warps of the block can be made to wait for)
#ifndef KMPC_TAIL
#define KMPC_TAIL 1   /* tail mode (full-solve inertia candidates in borrowed instance slots); 0 compiles it out (tuning builds) */
#endif
#ifndef KMPC_HOLD_TRIPS
#define KMPC_HOLD_TRIPS 256   /* an instance older than this many trips gets its block to itself (w_block_holds); 0 switches that off */
#endif
#ifndef KMPC_INLINE_BACKTRACKS
#define KMPC_INLINE_BACKTRACKS 24
#endif

#if defined(KMPC_SCHED_TRACE) && defined(__CUDACC__)
// tuning builds only: per instance the times (globaltimer, ns) at which a warp took it and finished it, its trips and its SM
__device__ unsigned long long *g_sched;
__device__ __forceinline__ unsigned long long kmpc_gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned kmpc_smid() { unsigned r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
#define SCHED_START(b) if (lane == 0 && g_sched) g_sched[4 * (size_t)(b)] = kmpc_gtime();
#define SCHED_END(b, trips) if (g_sched) { g_sched[4 * (size_t)(b) + 1] = kmpc_gtime(); g_sched[4 * (size_t)(b) + 2] = (unsigned long long)(trips); g_sched[4 * (size_t)(b) + 3] = ((unsigned long long)kmpc_smid() << 32) | (blockIdx.x << 8) | wid; }
#else
#define SCHED_START(b)
#define SCHED_END(b, trips)
#endif

// The regular line search of this instance has failed (trial_decide returned ST_RESTORATION) and the point is not "almost feasible":
// IPOPT's restoration phase is due.  It is rare and serial in nature, so it does not run here: the instance is handed over to the
// finisher (finish_instance, kmpc_core.cuh) -- iterate, multipliers, obstacle-row state, problem data, filter and solver context go
// to a column of the hand-over workspace in the thread solver's layout.  False: no column left (the caller reports Restoration_Failed).
template <int SPL, int NST, bool OBS>
KMPC_WN inline bool w_hand_over(const Cfg &c, WScal *sc, const WState<SPL> &w, const double *ob, const FiltSplit &filt, const IO &io, int b) {
    constexpr int NSTG = WLay<SPL, NST>::NSTG;
    const int N = c.N, lane = w_lane(), O = OBS ? c.O : 0;
    const Rows &L = c.L;
    int slot = -1;
    if (lane == 0 && io.resto_ws) slot = w_take_slot(io.resto_count);
    slot = w_bcast_i(slot, 0);
    if (slot < 0 || slot >= io.resto_cap) return false;
    double *base = io.resto_ws + (size_t)slot * io.resto_rows;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        const int s = lane * SPL + j;
        if (s > N) continue;
        double *ps = base + L.rState[0] + NSTATE * s;
        ps[F_X0] = w.x0[j]; ps[F_X1] = w.x1[j]; ps[F_X2] = w.x2[j]; ps[F_V] = w.v[j]; ps[F_OM] = w.om[j];
        ps[F_Y0] = w.y0[j]; ps[F_Y1] = w.y1[j]; ps[F_Y2] = w.y2[j];
        ps[F_ZLX] = w.zLx[j]; ps[F_ZUX] = w.zUx[j]; ps[F_ZLY] = w.zLy[j]; ps[F_ZUY] = w.zUy[j];
        ps[F_ZLV] = w.zLv[j]; ps[F_ZUV] = w.zUv[j]; ps[F_ZLW] = w.zLw[j]; ps[F_ZUW] = w.zUw[j];
        ps[F_CS] = w.cs[j]; ps[F_SN] = w.sn[j];
        if (OBS && s >= 1) {
            const WCen cen = w_cen(c, ob, O, NSTG, s);
            for (int o = 0; o < O; ++o) {
                const double *po = ob + o * NSTG + s;
                double *pr = base + L.rState[0] + L.sObs + 3 * ((s - 1) * O + o);
                pr[0] = po[B_S * O * NSTG]; pr[1] = po[B_YD * O * NSTG]; pr[2] = po[B_VL * O * NSTG];
                if (c.obs_sw) { base[L.rSc + CEN_ROW(o, s, 0)] = w_cx(cen, o); base[L.rSc + CEN_ROW(o, s, 1)] = w_cy(cen, o); }
            }
        }
    }
    if (OBS) {
        const WCen cen = w_cen(c, ob, O, NSTG, 1);
        for (int o = lane; o < O; o += 32) {
            base[L.rSc + RAD_ROW(o)] = w_cr(cen, o);
            if (!c.obs_sw) { base[L.rSc + CEN_ROW(o, 1, 0)] = w_cx(cen, o); base[L.rSc + CEN_ROW(o, 1, 1)] = w_cy(cen, o); }
        }
    }
    for (int i = lane; i < 2 * sc->t.fn; i += 32) base[L.rFilt + i] = filt.at(i >> 1, i & 1);
    if (lane == 0) {
        for (int j = 0; j < 3; ++j) { base[L.rSc + j] = sc->xc[j]; base[L.rSc + 3 + j] = sc->gl[j]; }
        sc->t.cur = 0; sc->t.inst = b;
        ctx_store(sc->t, L, base, 1);
        io.resto_list[slot] = b;
    }
    w_sync();
    return true;
}

// next instance of the queue that is to be solved (masked-out instances only get their status / iteration records)
KMPC_WN inline int w_fetch_active(const Cfg &c, const IO &io, int *queue) {
    for (;;) {
        const int q = w_fetch(queue);
        if (q >= c.B) return q;
        const int b = io.order ? io.order[q] : q;   // likely-long instances first (kmpc_order_kernel)
        if (!io.active || io.active[b]) return b;
        if (w_lane() == 0) { if (io.status) io.status[b] = KMPC_STATUS_SKIPPED; if (io.iters) io.iters[b] = 0; }
    }
}

// A block that carries a very old instance stops taking new ones: the trips of a block are as long as its busiest phase, an
// instance in a full block of 16 advances at ~17 us per trip but at ~8 us when it is alone, and the one instance in 10^4..10^5
// that needs many hundred trips otherwise sets the end of the whole launch (65,536 x N = 30: a 770-trip instance = 13.5 ms of
// a 12.6 ms batch).  The other blocks absorb the work of the 15 slots that drain, < 1 % of the grid.  Scheduling only: which
// warp solves an instance never changes its arithmetic.  (Reads the neighbours' contexts in shared memory; a stale value costs
// or saves one fetch, nothing else -- the caller never lets a hold empty the block, see the two-pass fetch in w_worker.)
KMPC_WN inline bool w_block_holds(const WScal *scal0, int W) {
    if (KMPC_HOLD_TRIPS <= 0) return false;
    const int l = w_lane();
    bool old = false;
    if (l < W) { const Ctx &o = scal0[l].t; old = o.mode != M_DONE && o.trips > KMPC_HOLD_TRIPS; }
    return w_any(old);
}

// Inertia prediction.  IPOPT factorises every new iteration's KKT matrix with delta_w = 0 first and walks its perturbation sequence
// (0, d1 = max(delta_min, delta_last / 3) or 1e-4, d2 = 8 d1 or 100 d1, ...) until the inertia is right.  A wrong inertia costs this
// solver a whole trip (the instance sits out the rest of the block's trip), and it clusters: after an iteration that needed a
// perturbation the first factorisation of the next one fails in 87 % of the cases (after an unperturbed one: 1.6 %; 65,536 x N = 30),
// and which element then works follows from the last one -- after d1 mostly d2 (a third of the last value is too small), after d2 or
// later mostly d1.  So such an iteration is ASSEMBLED AND SOLVED with the predicted element right away, and the candidate lanes
// test the other elements of the sequence, delta_w = 0 included (diagonal shifts of the assembled blocks, as before).  The decision
// rule is IPOPT's: the first element of the sequence with the right inertia is the one used -- if that is not the predicted one, the
// instance repeats the sweep with it (what a failed first factorisation costs anyway).  Same iterates, 5 % fewer trips.
#ifndef KMPC_PREDICT_INERTIA
#define KMPC_PREDICT_INERTIA 1
#endif
#ifndef KMPC_MERGE_TOP
#define KMPC_MERGE_TOP 1   /* the top-of-trip barrier merged into the one behind the assemble phase (kernels without the tail mode) */
#endif
// element i of the sequence that starts at delta_w = 0 (i = 0: no perturbation)
KMPC_W double w_inertia_seq(int i, double delta_last) {
    double d = 0.0;
    for (int k = 0; k < i; ++k) d = inertia_next_delta(d, delta_last);
    return d;
}

// ---- persistent worker: one warp pulls instances from a queue and solves each start to finish; the warps of a block
// walk through the phases of a trip in step (block barriers) so that warp 0 can run every instance's serial recursions.
// smem: WLay<SPL, NST>::bytes(warps per block, O) bytes of block-shared scratch.
// TAIL: the tail mode is compiled in.  Its mere presence costs the common path 3 % (65,536 x N = 30: 12.95 vs 12.60 ms; neither
// keeping its state in shared memory nor moving its code out of line changed that), so batches of many waves run the kernel
// without it and only batches in which the drained-queue phase matters -- small ones, the slices of a batch sharded over
// several GPUs -- run the kernel with it (launch_warp_kernel).
template <int SPL, int NST, bool FULL, bool OBS, bool TAIL = true>
KMPC_WN inline void w_worker(const Cfg &c, const IO &io, double *smem, int *queue, unsigned long long *trips_total) {
    typedef WLay<SPL, NST, OBS> LY;
    const int N = c.N, lane = w_lane(), wid = w_warp(), W = w_warps();
    double *coop = smem + (size_t)wid * LY::COOP;
    double *priv = smem + (size_t)W * LY::COOP + (size_t)wid * LY::PRIV;
    double *gp = io.wscratch + ((size_t)w_block() * W + wid) * LY::GPRIV;
    const int OBD = LY::obs_doubles(OBS ? c.O : 0, c.obs_sw);
    double *ob = smem + (size_t)W * (LY::COOP + LY::PRIV) + (size_t)wid * OBD;
    WScal *scal0 = (WScal *)(smem + (size_t)W * (LY::COOP + LY::PRIV + OBD));
    WScal *sc = scal0 + wid;
    // the filter (touched by lane 0 only; 512 entries at most as in the oracle, a handful as a rule): the first few entries in shared
    // memory, the others in the global scratch slot (all of it there cost 5 % of the batch: a dependent L2 round trip per entry and trial)
    const FiltSplit filt{sc->fnear, gp + LY::GFILT};
    unsigned *hmask = (unsigned *)(scal0 + W);   // which warps hold an instance, this trip / next trip
    Ctx &t = sc->t;
    WState<SPL> cur;
#if !KMPC_STEP_SMEM
    WStep<SPL> act;
#endif
    bool have = false;
    int b = -1;
    if (lane == 0) { t.mode = M_DONE; sc->flag = 0; sc->ok = 0; sc->tinfo = 0; sc->ncand = 0; sc->pred = 0; sc->pstat = 0; for (int k = 0; k < KMPC_NCAND; ++k) { sc->dshift[k] = NAN; sc->pdc[k] = 0; } }
    if (wid == 0 && lane == 0) { hmask[0] = 0; hmask[1] = 0; }
    w_block_sync();
    PT_DECL
    bool drained = false;  // the queue has no more instances for this warp
    // the block's Riccati warp: co-resident blocks (b and b + gridDim/2 under round-robin placement) pick different warp
    // slots, hence different schedulers, so that two serial phases that coincide do not share one FP64 issue port
    const int swid = w_serial_warp(W);
    // the warps that test the speculative inertia candidates: (KMPC_NCAND - 1) * W lanes, i.e. one warp for blocks of up to 10
    // instances, two for 16; candidate-lane g = cj * 32 + lane serves instance g % W with candidate 1 + g / W
    const int ncw = W >= 2 ? ((KMPC_NCAND - 1) * W + 31) / 32 < W - 1 ? ((KMPC_NCAND - 1) * W + 31) / 32 : W - 1 : 0;
    const int cj = (wid - swid - 1 + W) % W;          // 0 .. ncw - 1: this warp is a candidate warp
    const bool is_cand = W >= 2 && wid != swid && cj < ncw;
    const bool can_predict = W >= 2 && (KMPC_NCAND - 1) * W <= 32 * ncw && KMPC_NCAND >= 3;   // every other element of the sequence has a candidate lane
    bool fetch_all = false;   // (merged top barrier only)
#pragma unroll 1
    for (;;) {
        // warp 0 is busy in the serial window below, so it takes its next instance here; the other warps take theirs
        // in that window (the global-memory round trip then costs the block nothing)
        // (second pass: nothing is live, so a hold -- possibly read stale -- is void and every warp may fetch)
        int nlive = 0;
        // Without the tail mode nothing needs the number of live warps before the stage blocks are assembled, so the count rides on the
        // barrier behind the assemble phase (one block barrier less per trip; a warp that is done with its trial phase goes straight on
        // to assemble -- between the barrier behind the serial window and the next one every warp touches its own areas only).
        constexpr bool top_barrier = (TAIL && KMPC_TAIL && !OBS && KMPC_NCAND > 1) || !KMPC_MERGE_TOP;
        if (top_barrier) {
#pragma unroll 1
            for (int pass = 0; pass < 2 && !nlive; ++pass) {
                if (!have && !drained && (pass || wid == swid || is_cand) && (pass || !w_block_holds(scal0, W))) {
                    b = w_fetch_active(c, io, queue);
                    if (b < c.B) { SCHED_START(b) w_init<SPL, NST, OBS>(c, sc, io, b, cur, ob); have = true; } else drained = true;
                }
                PT(0)
                nlive = w_block_warps_with(have);   // warps that hold an instance this trip
            }
            if (!nlive) break;
        } else {
            // (fetch_all: the last count found nothing live, so a hold -- possibly read stale -- is void and every warp may fetch)
            if (!have && !drained && (fetch_all || wid == swid || is_cand) && (fetch_all || !w_block_holds(scal0, W))) {
                b = w_fetch_active(c, io, queue);
                if (b < c.B) { SCHED_START(b) w_init<SPL, NST, OBS>(c, sc, io, b, cur, ob); have = true; } else drained = true;
            }
            PT(0)
        }
        PT(1)
        int status = 100;
        // ---- tail mode: once at most a quarter of the block's instance slots are taken (the queue is drained, or the batch is
        // small) every live instance borrows KMPC_NCAND - 1 free slots.  Into those it assembles the SAME Newton system with the
        // next perturbations delta_w of IPOPT's inertia-correction sequence, the Riccati warp solves them side by side with the
        // base system on lanes that would idle anyway, and a factorisation with the wrong inertia costs no retry trip: the
        // instance continues at once with the first candidate that has the right one.  Same arithmetic as the retry, same bits.
        // (All of its state lives in shared memory -- sc->tinfo, sc->ncand, hmask[1] -- so that the common path carries no
        //  register for it: with them in registers the whole kernel ran 3 % slower.)
        constexpr int NCF = KMPC_NCAND > 1 ? KMPC_NCAND - 1 : 1;
        if (TAIL && KMPC_TAIL && !OBS && KMPC_NCAND > 1) {
            if (W - nlive >= NCF * nlive) {          // uniform over the block: two more barriers, next to idle warps
                if (lane == 0 && have) w_smem_or(hmask, 1u << wid);
                w_block_sync();
                const unsigned hm = *(volatile unsigned *)hmask;
                if (wid == 0 && lane == 0) hmask[1] = hm;   // kept for the rest of the trip (read after the barrier below)
                w_block_sync();
                if (wid == 0 && lane == 0) hmask[0] = 0;    // (the next word is built after the next trip's first barrier)
                // live: 1 + number of the first free slot it borrows (in the order of the free slots), shifted by 8; free: bit 0 = lent out
                if (lane == 0) sc->tinfo = have ? (w_popc(hm & ((1u << wid) - 1u)) * NCF + 1) << 8 : (w_popc(~hm & ((1u << wid) - 1u)) < nlive * NCF ? 1 : 0);
            } else if (lane == 0) sc->tinfo = 0;
            w_sync();
        }
        // ---- phase 1a: assemble the stage blocks ----
        const int mode = t.mode;
        const bool do_sweep = have && mode != M_TRIAL;
        {
            const int tinfo = (TAIL && KMPC_TAIL && !OBS && KMPC_NCAND > 1) ? sc->tinfo : 0;
            const bool full_cands = (tinfo >> 8) != 0 && do_sweep && mode == M_NEWTON;
            if (KMPC_PREDICT_INERTIA && full_cands && sc->pred) {   // borrowed slots solve the whole sequence side by side: no prediction needed
                w_sync();
                if (lane == 0) { t.delta = 0.0; t.alpha_min = 0.0; sc->pred = 0; }
                w_sync();
            }
            int ncand = 0;       // full candidates of this trip (fewer than NCF when the sequence runs past delta_w_max)
            if (do_sweep) w_assemble<SPL, NST, FULL, OBS>(c, sc, cur, priv, gp, coop, ob);   // (ONE call site: two inlined copies need not round alike)
            if (full_cands) {
                const unsigned hm = ((volatile unsigned *)hmask)[1];
                WCands cand;     // (uniform over the warp: every lane forms the sequence itself)
                double dk = t.delta;
#pragma unroll
                for (int k = 0; k < NCF; ++k) {
                    dk = inertia_next_delta(dk, t.delta_last);
                    if (!(dk <= K_DW_MAX)) break;
                    const int slot = w_nth_clear(hm, W, (tinfo >> 8) - 1 + k);
                    cand.coop[k] = smem + (size_t)slot * LY::COOP; cand.delta[k] = dk; cand.sc[k] = scal0 + slot;
                    ncand = k + 1;
                }
                cand.n = ncand;
                w_sync();   // (the base blocks and d0 are read back below)
                if (!OBS) w_assemble_cands<SPL, NST, FULL>(c, sc, cur, priv, coop, cand);
            }
            if (lane == 0) {
                if (!(tinfo & 1)) sc->flag = do_sweep ? 1 : 0;   // (a lent slot's flag belongs to its borrower)
                if (tinfo >> 8) {
                    const unsigned hm = ((volatile unsigned *)hmask)[1];
                    for (int k = 0; k < NCF; ++k) {
                        WScal *sk = scal0 + w_nth_clear(hm, W, (tinfo >> 8) - 1 + k);
                        sk->flag = k < ncand ? 1 : 0;
                        for (int q = 1; q < KMPC_NCAND; ++q) sk->dshift[q] = NAN;
                    }
                }
                if (do_sweep) {
                    t.trips++;
                    sc->ncand = ncand;
                    // speculative inertia candidates (Newton systems; delta_w is a diagonal shift of the assembled blocks plus, with
                    // obstacle rows, delta_w * sum n n^T): the next perturbations IPOPT would try if this factorisation has the wrong inertia
                    double dk = t.delta;
                    sc->dshift[0] = 0.0;
                    if (KMPC_PREDICT_INERTIA && sc->pred) {
                        // predicted sweep: the candidates are the OTHER elements of the sequence, in its order (delta_w = 0 first)
                        int k = 1;
                        for (int i = 0; i < KMPC_NCAND; ++i) {
                            if (i == sc->pred) continue;
                            const double di = w_inertia_seq(i, t.delta_last);
                            sc->dshift[k] = di <= K_DW_MAX ? di - t.delta : NAN;
                            sc->pdc[k] = 0;
                            ++k;
                        }
                    } else
                    for (int k = 1; k < KMPC_NCAND; ++k) {
                        dk = inertia_next_delta(dk, t.delta_last);
                        sc->dshift[k] = (!full_cands && W >= 2 && mode == M_NEWTON && dk <= K_DW_MAX && k * W <= 32 * ncw) ? dk - t.delta : NAN;
                        sc->pdc[k] = 0;
                    }
                }
            }
        }
        PT(2)
        if (top_barrier) w_block_sync();
        else {
            nlive = w_block_warps_with(have);
            if (!nlive) { if (fetch_all) break; fetch_all = true; continue; }   // (uniform over the block)
            fetch_all = false;
        }
        PT(3)
        // ---- phase 1b: the serial recursions of all the block's instances, one lane each ----
        bool fresh = false;  // an instance taken in this window joins the next trip
        if (wid == swid) {
            PT_SERIAL_BEGIN
            if (lane < W) {
                WScal *so = scal0 + lane;
                if (so->flag) so->ok = w_serial<OBS>(c, smem + (size_t)lane * LY::COOP, LY::NSTG, so->d0) ? 1 : 0;
            }
            PT_SERIAL_END
        } else if (is_cand && KMPC_NCAND > 1) {
            // speculative inertia candidates: candidate-lane g = (candidate - 1) * W + instance
            const int g = cj * 32 + lane, inst = g % W, cand = 1 + g / W;
            WScal *so = scal0 + inst;
            const double ds = cand < KMPC_NCAND ? so->dshift[cand] : NAN;
            if (so->flag && ds == ds) so->pdc[cand] = w_serial_candidate<OBS>(c, smem + (size_t)inst * LY::COOP, LY::NSTG, ds) ? 1 : 0;
        } else if (!have && !drained && !w_block_holds(scal0, W)) {
            b = w_fetch_active(c, io, queue);
            if (b < c.B) { SCHED_START(b) w_init<SPL, NST, OBS>(c, sc, io, b, cur, ob); fresh = true; } else drained = true;
        }
        PT(4)
        w_block_sync();
        PT(5)
        // ---- phase 2: search direction, step sizes, line-search set-up (or inertia correction) ----
        bool go_trial = have && !do_sweep;
        if (do_sweep) {
            // the system that was solved: the base one, or -- tail mode -- the first full candidate with the right inertia
            const double *solved = coop;
            if (KMPC_PREDICT_INERTIA && sc->pred) {
                // IPOPT's rule over the tested elements: the first one of the sequence with the right inertia is the one to use
                if (lane == 0) {
                    const int pj = sc->pred;
                    int first = -1, k = 1;
                    for (int i = 0; i < KMPC_NCAND && first < 0; ++i) {
                        const bool works = i == pj ? sc->ok != 0 : (sc->dshift[k] == sc->dshift[k] && sc->pdc[k] != 0);
                        if (i != pj) ++k;
                        if (works) first = i;
                    }
                    if (first != pj) {
                        sc->ok = 0;
                        if (first >= 0) { t.delta = w_inertia_seq(first, t.delta_last); t.alpha_min = (double)first; sc->pstat = R_RETRY; }
                        else { t.delta = w_inertia_seq(KMPC_NCAND - 1, t.delta_last); t.alpha_min = (double)KMPC_NCAND; sc->pstat = inertia_update(t); }
                    }
                }
                w_sync();
            }
            bool ok = sc->ok != 0;
            const int ncand = (TAIL && KMPC_TAIL && !OBS && KMPC_NCAND > 1 && !ok) ? sc->ncand : 0;
            for (int k = 0; k < ncand; ++k) {
                const int slot = w_nth_clear(((volatile unsigned *)hmask)[1], W, (sc->tinfo >> 8) - 1 + k);
                if (scal0[slot].ok) {
                    ok = true; solved = smem + (size_t)slot * LY::COOP;
                    w_sync();
                    if (lane == 0) { for (int q = 0; q <= k; ++q) t.delta = inertia_next_delta(t.delta, t.delta_last); t.alpha_min += (double)(k + 1); }   // IPOPT's sequence up to the perturbation that works
                    w_sync();
                    break;
                }
            }
            if (!ok) {
                PT_COUNT(10)
                if (lane == 0) {
                    if (KMPC_PREDICT_INERTIA && sc->pred) { sc->status = sc->pstat; sc->pred = 0; }   // (judged above)
                    else {
                        // wrong inertia: raise delta_w (IPOPT's sequence) past the candidates already known to fail, sweep again next trip
                        // (t.alpha_min -- a field the warp solver has no other use for -- counts the elements of the sequence tried)
                        int st = mode != M_NEWTON ? sweep_failure_status(t) : inertia_update(t);
                        t.alpha_min += 1.0;
                        for (int k = 0; k < ncand && st == R_RETRY; ++k) { st = inertia_update(t); t.alpha_min += 1.0; }
                        for (int k = 1; k < KMPC_NCAND && st == R_RETRY && sc->dshift[k] == sc->dshift[k] && !sc->pdc[k]; ++k) { st = inertia_update(t); t.alpha_min += 1.0; }
                        sc->status = st;
                    }
                }
                w_sync();
                status = sc->status;
            } else {
                double apr, adu, gbd, ym;
#if KMPC_STEP_SMEM
                WStep<SPL> act;
#endif
                w_step<SPL, NST, FULL, OBS>(c, sc, cur, solved, priv, ob, act, &apr, &adu, &gbd, &ym);
                if (lane == 0) {
                    if (mode == M_NEWTON) t.pw_t = t.alpha_min;   // which element of the perturbation sequence this iteration needed (0: none)
                    sc->pred = 0;
                    rollout_logic(t, apr, adu, gbd, ym);
                }
                w_sync();
                if (t.sel == 0) w_step_store<SPL, NST>(act, gp);
#if KMPC_STEP_SMEM
                w_step_park<SPL, NST>(act, coop);   // (every lane parks and later fetches its own stages only: no warp barrier needed)
#endif
                go_trial = true;
            }
        } else if (have) {
            if (lane == 0) trial_setup(t);
            w_sync();
#if KMPC_STEP_SMEM
            { WStep<SPL> act; w_step_load<SPL, NST>(act, gp); w_step_park<SPL, NST>(act, coop); }
#else
            w_step_load<SPL, NST>(act, gp);
#endif
        }
        PT(6)
        // ---- phase 3: trial point + acceptance logic ----
        // A rejected trial point is followed at once by the next, shorter one: back-tracking needs no new factorisation, and an
        // instance with a difficult line search (hundreds of rejected points) would otherwise pay a whole block trip for each.
        for (int nbt = 0; go_trial; ++nbt) {
            Stats ts;
            WState<SPL> tri;
#if KMPC_STEP_SMEM
            WStep<SPL> act;
            w_step_fetch<SPL, NST>(act, coop);
#endif
            // step sizes / barrier parameters of THIS trial, read before lane 0 moves the context on (begin_iteration)
            const double ta_pr = t.a_pr, ta_y = t.a_y, ta_du = t.a_du, ta_mu = t.mu, ta_delta = t.delta;
            const bool tclamp = t.tu == TU_STEP, tlsq = t.mode == M_LSQ, tsoc = t.mode == M_SOC;
            const bool evok = w_trial<SPL, NST, FULL, OBS>(c, sc, cur, act, ta_pr, ta_y, ta_du, tclamp, tri, priv, gp, coop, ob, &ts);
            PT(7)
            if (lane == 0) {
                bool aug; double ath, aph;
                int r = trial_decide(t, filt, ts, evok, &aug, &ath, &aph);
                if (aug && !filter_add(t, filt, ath, aph)) r = ST_INTERNAL;   // filter full: ends the instance loudly (Internal_Error)
                sc->r = r;
            }
            w_sync();
            const int r = sc->r;
            if (r == R_SOC1 || r == R_SOC2) { PT_COUNT(11) w_soc_rhs<SPL, NST, OBS>(c, sc, cur, t.alpha_soc, r == R_SOC1, gp, ob); }
            else if (r == R_ACCEPT) {
                PT_COUNT(12)
                if (mode == M_SOC) { PT_COUNT(14) }
                if (OBS) w_obs_commit<SPL, NST>(c, cur, act, ta_mu, ta_delta, ta_pr, ta_y, ta_du, tclamp, tlsq, tsoc, ob);
                w_accept<SPL, NST>(c, cur, tri, coop);
                if (lane == 0) {
                    t.c = ts; sc->status = begin_iteration(c, t);
                    t.alpha_min = 0.0;
                    if (KMPC_PREDICT_INERTIA && can_predict && sc->status == 100 && t.pw_t > 0.0) {
                        const int pj = t.pw_t == 1.0 ? 2 : 1;
                        const double dj = w_inertia_seq(pj, t.delta_last);
                        if (dj <= K_DW_MAX) { t.delta = dj; t.alpha_min = (double)pj; sc->pred = pj; }
                    }
                }
                w_sync();
                status = sc->status;
            } else if (r != R_BACKTRACK) status = r;
            else {
                PT_COUNT(13)
                // (after a failed second-order correction `act` holds the corrected step: the original one is re-read next trip)
                if (nbt < KMPC_INLINE_BACKTRACKS && mode != M_SOC) { if (lane == 0) trial_setup(t); w_sync(); continue; }
            }
            break;
        }
        PT(8)
        if (have && status == ST_RESTORATION) {
            // IPOPT's BacktrackingLineSearch: restoration phase, unless the point is almost feasible ("Restoration phase called, but
            // point is almost feasible": Restoration_Failed).  The phase itself runs in the finisher kernel.
            if (!(t.c.theta <= 1e-2 * c.tol) && w_hand_over<SPL, NST, OBS>(c, sc, cur, ob, filt, io, b)) {
                if (lane == 0) { w_count_trips(trips_total, t.trips); SCHED_END(b, t.trips) t.mode = M_DONE; }
                have = false;
                status = 100;
            }
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
#ifdef KMPC_ITERS_OUT_TRIPS   /* tuning build for scripts/fit_order_prior.py: report trips (what an instance costs) as "iterations" */
                if (io.iters) io.iters[b] = t.trips;
#else
                if (io.iters) io.iters[b] = t.iter;
#endif
                if (io.cost_out) io.cost_out[b] = t.trips;
                w_count_trips(trips_total, t.trips);
                SCHED_END(b, t.trips)
                t.mode = M_DONE;
            }
            have = false;
        }
        if (fresh) have = true;
        PT(9)
    }
    PT_FLUSH
}

}  // namespace kmpc
