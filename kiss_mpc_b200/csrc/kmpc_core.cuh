// kmpc_core.cuh -- per-instance primal-dual interior-point solver for the unicycle MPC NLP, written for ONE CUDA
// thread per problem instance with all per-instance state in a structure-of-arrays HBM workspace (row r of slot s
// lives at ws[r * S + s], so every load/store of a warp is one contiguous 256-byte segment).
//
// What it replaces: the arithmetic behind mpc/optimizer.py:354 (ca.nlpsol "ipopt") + :375-391 (the solve call):
// IPOPT's filter line-search interior point method (Waechter & Biegler 2006; IPOPT 3.14 defaults + the options of
// optimizer.py:344-352) applied to the NLP that optimizer.py:79-317 builds (README.md:15-81 form by default).
// The linear algebra is NOT IPOPT's (MUMPS LDL^T on the 246x246 sparse KKT): the block-tridiagonal KKT system is
// solved by a Riccati recursion over the 3-state/2-control stages, with the inertia test "every Q_uu is positive
// definite" standing in for MUMPS' inertia count (equivalent because the dynamics Jacobian has full row rank).
//
// Control flow is a state machine ("trip" = [backward Riccati sweep] -> [forward roll-out] -> [trial-point
// evaluation + speculative iterate/dual update] -> [scalar filter logic]) so that the 32 instances of a warp execute
// the same heavy code every trip no matter whether a lane is in a Newton step, an inertia-correction retry, a
// second-order correction, a back-tracking trial or the initial least-squares multiplier estimate.
//
// The code is plain C++ in KMPC_HD functions: the CUDA kernel in kmpc.cu wraps it; tests/host_emul compiles the very
// same source with g++ to check the algorithm against the oracle on the GPU-less build machine (test harness only --
// the product library contains only the CUDA path).
#pragma once
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <float.h>

#ifdef __CUDACC__
#define KMPC_HD __host__ __device__ __forceinline__
#define KMPC_HDN __host__ __device__
#else
#define KMPC_HD inline
#define KMPC_HDN
#endif

namespace kmpc {

// ---- IPOPT 3.14 defaults (option names in brackets) ----
#define K_BOUND_RELAX 1e-8        /* bound_relax_factor */
#define K_SCALING_MAX_GRAD 100.0  /* nlp_scaling_max_gradient */
#define K_SCALING_MIN 1e-8        /* nlp_scaling_min_value */
#define K_BOUND_PUSH 0.01         /* bound_push */
#define K_BOUND_FRAC 0.01         /* bound_frac */
#define K_YINIT_MAX 1e3           /* constr_mult_init_max */
#define K_MU_INIT 0.1             /* mu_init */
#define K_TAU_MIN 0.99            /* tau_min */
#define K_KAPPA_EPS 10.0          /* barrier_tol_factor */
#define K_MU_LIN 0.2              /* mu_linear_decrease_factor */
#define K_MU_SUPER 1.5            /* mu_superlinear_decrease_power */
#define K_S_MAX 100.0             /* s_max */
#define K_KAPPA_SIGMA 1e10        /* kappa_sigma */
#define K_KAPPA_D 1e-5            /* kappa_d */
#define K_GAMMA_THETA 1e-5
#define K_GAMMA_PHI 1e-8
#define K_ETA_PHI 1e-8
#define K_S_THETA 1.1
#define K_S_PHI 2.3
#define K_DELTA_LS 1.0
#define K_THETA_MAX_FACT 1e4
#define K_THETA_MIN_FACT 1e-4
#define K_ALPHA_MIN_FRAC 0.05
#define K_ALPHA_RED 0.5
#define K_MAX_SOC 4
#define K_KAPPA_SOC 0.99
#define K_OBJ_MAX_INC 5.0
#define K_DW_INIT 1e-4            /* first_hessian_perturbation */
#define K_DW_MIN 1e-20
#define K_DW_MAX 1e20
#define K_DW_INC_FIRST 100.0
#define K_DW_INC 8.0
#define K_DW_DEC (1.0 / 3.0)
#define K_DUAL_INF_TOL 1.0
#define K_CONSTR_VIOL_TOL 1e-4
#define K_COMPL_INF_TOL 1e-4
#define K_DIVERGING 1e20
#define K_FILTER_CAP 512         /* filter entries kept per instance, as in the oracle (IPOPT's filter is unbounded; a full filter ends the instance with Internal_Error) */
#define KMPC_NCTX 40  /* rows reserved for the per-instance solver context in the workspace */

enum { ST_SUCCESS = 0, ST_MAXITER = -1, ST_RESTORATION = -2, ST_STEP_ERROR = -3, ST_DIVERGING = 4, ST_INVALID = -13, ST_INTERNAL = -199 };
enum { M_FETCH = 0, M_LSQ = 1, M_NEWTON = 2, M_SOC = 3, M_TRIAL = 4, M_DONE = 5 };
enum { TU_INIT = 0, TU_STEP = 1 };

// Workspace records.  All per-instance data of one stage k that a pass touches together is one RECORD of consecutive
// rows, so a pass walks one pointer per record type and prefetches the next record while it computes on this one.
enum { F_X0 = 0, F_X1, F_X2, F_V, F_OM, F_Y0, F_Y1, F_Y2, F_ZLX, F_ZUX, F_ZLY, F_ZUY, F_ZLV, F_ZUV, F_ZLW, F_ZUW, F_CS, F_SN, NSTATE };
enum { A_K00 = 0, A_K01, A_K02, A_K10, A_K11, A_K12, A_KF0, A_KF1, A_P00, A_P10, A_P11, A_P20, A_P21, A_P22, A_PV0, A_PV1, A_PV2, NFACT };
enum { D_X0 = 0, D_X1, D_X2, D_U0, D_U1, D_Y0, D_Y1, D_Y2, NSTEP };

// Row map of the per-slot workspace (all offsets in rows of S doubles).
struct Rows {
    int N, O;
    int sObs, state_rows;  // one state buffer: (N+1) STATE records, then N*O obstacle records [s, yd, vL]
    int rState[2];
    int rFact;             // (N+1) FACT records
    int dObs, step_rows;   // one step buffer: (N+1) STEP records, then N*O records [ds, dyd]
    int rStep[2];
    int rCsoc, rDsoc, rFilt, rSc, rCtx, total;
};

KMPC_HD Rows make_rows(int N, int O, int stagewise = 0) {
    Rows L;
    L.N = N; L.O = O;
    const int NO = N * O;
    L.sObs = NSTATE * (N + 1);
    L.state_rows = L.sObs + 3 * NO;
    int r = 0;
    L.rState[0] = r; r += L.state_rows;
    L.rState[1] = r; r += L.state_rows;
    L.rFact = r; r += NFACT * (N + 1);
    L.dObs = NSTEP * (N + 1);
    L.step_rows = L.dObs + 2 * NO;
    L.rStep[0] = r; r += L.step_rows;
    L.rStep[1] = r; r += L.step_rows;
    L.rCsoc = r; r += 3 * (N + 1);
    L.rDsoc = r; r += NO;
    L.rFilt = r; r += 2 * K_FILTER_CAP;
    L.rSc = r; r += 6 + 2 * O * (stagewise ? N : 1) + O;   // x_cur, goal, circle centres (per obstacle, or per obstacle and stage), radii
    L.rCtx = r; r += KMPC_NCTX;
    L.total = r;
    return L;
}

// Per-solve constants (kernel parameter, by value).
struct Cfg {
    int N, O, cost_mode, gk_lo, gk_hi, max_iter, layout, B;
    int obs_sw;            // circle centres given per obstacle AND stage (dynamic_obstacle.py:47-56) instead of per obstacle
    int hasL[4], hasU[4];  // x, y, v, omega
    int nb, m;             // number of bound sides incl. obstacle slacks; number of constraint rows
    double r_mnb, r_nb;    // 1 / (m + nb), 1 / nb (0 if nb == 0): the averaging factors of IPOPT's s_d, s_c
    double mu_floor;       // min(tol, compl_inf_tol) / (barrier_tol_factor + 1): the smallest barrier parameter (mu_min of IPOPT's monotone update)
    double T, W[3], Wvn, Wvp, Ww;
    double lb[4], ub[4];   // relaxed bounds
    double tol, obs_radius, dL;
    Rows L;
};

// I/O pointers (device memory owned by the caller)
struct IO {
    const double *x_cur, *goal, *X0, *U0, *obs;
    const double *orad;     // per-slot obstacle radii [B][O] / [O][B] (optimizer.py:231-250: one radius per obstacle class); NULL: Cfg::obs_radius
    double *X_out, *U_out, *obj;
    int32_t *status, *iters;
    int32_t *cost_out;      // optional: trips the instance took (what it cost; the closed loops order the next step's queue by it)
    double *wscratch;       // warp solver: global scratch, WLay::GPRIV doubles per resident warp (owned by the handle)
    const int32_t *active;  // optional per-instance mask (closed loop: agents that reached their goal are not solved again)
    const int32_t *order;   // warp solver: instance handed out at queue position q (NULL: q itself); a permutation of 0..B-1
    // warp solver -> finisher hand-over of the instances whose line search failed (restoration phase, kmpc_resto.cuh): a workspace
    // column per instance (resto_rows doubles each, contiguous), the instance each column holds, the number of columns in use
    double *resto_ws;
    int32_t *resto_list;
    int *resto_count;
    int resto_cap, resto_rows;
};
#define KMPC_STATUS_SKIPPED 1000  /* status_log value of an agent that was not solved in a closed-loop step */

struct Stats {  // residual norms + merit ingredients of one point
    double f, bar, damp, theta, dinf, pinf, mn, mx, sumy, sumz, wmax;
};

struct Ctx {  // per-thread solver state (registers / local memory)
    int mode, inst, iter, cur, nsteps, soc_count, fn, trips, sel, tu;
    double mu, tau, delta, delta_last, df, theta_max, theta_min;
    double alpha, alpha_test, alpha_min, alpha_du0, alpha_soc, gBD, theta_soc_old, theta_trial;
    double a_pr, a_y, a_du;  // step sizes of the pending trial: primal, equality multipliers, bound multipliers
    double pw_g, pw_t;       // pw_g: alpha_min of the current line search once a trial point has been rejected (0: not computed yet); pw_t: warp solver, inertia prediction
    Stats c;  // current iterate
};

KMPC_HD double dmin0(double v) { return v < 0 ? 1.0 : (v == 0 ? 0.5 : 0.0); }
KMPC_HD double dmax0(double v) { return v > 0 ? 1.0 : (v == 0 ? 0.5 : 0.0); }
KMPC_HD bool cmp_le(double lhs, double rhs, double bas) { return lhs - rhs <= 10.0 * DBL_EPSILON * fabs(bas); }
KMPC_HD void sincos_(double a, double *s, double *c) {
#ifdef __CUDA_ARCH__
    sincos(a, s, c);
#else
    *s = sin(a); *c = cos(a);
#endif
}

#define RW(r) wsp[(size_t)(r) * S]

// I/O index helpers (kmpc.h layouts)
KMPC_HD size_t io_vec3(const Cfg &c, int b, int j) { return c.layout ? (size_t)j * c.B + b : (size_t)b * 3 + j; }
KMPC_HD size_t io_X(const Cfg &c, int b, int j, int k) {
    return c.layout ? ((size_t)j * (c.N + 1) + k) * c.B + b : ((size_t)b * 3 + j) * (c.N + 1) + k;
}
KMPC_HD size_t io_U(const Cfg &c, int b, int j, int k) {
    return c.layout ? ((size_t)j * c.N + k) * c.B + b : ((size_t)b * 2 + j) * c.N + k;
}
KMPC_HD size_t io_obs(const Cfg &c, int b, int o, int j) {
    return c.layout ? ((size_t)o * 2 + j) * c.B + b : ((size_t)b * c.O + o) * 2 + j;
}
// stage-wise centres: column t of obstacle o's track is paired with X_{t+1} (dynamic_obstacle.py:47-56)
KMPC_HD size_t io_obs_sw(const Cfg &c, int b, int o, int t, int j) {
    return c.layout ? (((size_t)o * c.N + t) * 2 + j) * c.B + b : (((size_t)b * c.O + o) * c.N + t) * 2 + j;
}
// row (within the scalar rows of the thread solver's workspace) of coordinate j of obstacle o's centre at stage k >= 1
#define CEN_ROW(o, k, j) (6 + 2 * (c.obs_sw ? (o) * c.N + ((k) - 1) : (o)) + (j))
// row of obstacle o's radius (behind the centres)
#define RAD_ROW(o) (6 + 2 * c.O * (c.obs_sw ? c.N : 1) + (o))
KMPC_HD size_t io_orad(const Cfg &c, int b, int o) { return c.layout ? (size_t)o * c.B + b : (size_t)b * c.O + o; }

// cost gradient of v (scaled) and its second derivative; optimizer.py:91-96 (literal) / README.md:23-24
KMPC_HD void vcost(const Cfg &c, double df, double v, double *g, double *h) {
    if (c.cost_mode == 0) {
        // W_v- min(0,v)^2 + W_v+ max(0,v)^2 with CasADi's derivative convention at the kink (both halves weigh 1/2 at v == 0)
        const double w = v < 0 ? c.Wvn : c.Wvp;
        *g = df * (2.0 * w * v);
        *h = v == 0 ? df * (0.5 * c.Wvn + 0.5 * c.Wvp) : df * (2.0 * w);
    } else {
        *g = df * c.Wvn * dmin0(v);
        *h = 0.0;
    }
}

// correctly rounded reciprocal (one MUFU + Newton steps on the device instead of a full division)
// additions / multiplications the compiler must not contract into an FMA: expressions that two different functions have to
// evaluate to the same bits (kmpc_warp.cuh: w_assemble and w_assemble_cands)
#ifdef __CUDA_ARCH__
#define KADD(a, b) __dadd_rn((a), (b))
#define KMUL(a, b) __dmul_rn((a), (b))
#else
#define KADD(a, b) ((a) + (b))
#define KMUL(a, b) ((a) * (b))
#endif
#ifdef __CUDA_ARCH__
#define KRCP(x) __drcp_rn(x)
// reciprocal for the Riccati pivots and slacks: MUFU.RCP64H seed (~20 bits) + one cubic step (<= 1-2 ulp), branch-free.  Operands
// are positive and far from the denormal/overflow range whenever the result is used (the pivot test discards the rest).
__device__ __forceinline__ double krcp_fast(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    // one cubic step r (1 + e + e^2), e = 1 - d r (|e| ~ 2^-20 from the seed -> 2^-60 after): three dependent FMAs
    const double e = fma(-d, r, 1.0);
    return fma(r, fma(e, e, e), r);
}
#define KRCPF(x) krcp_fast(x)
#else
#define KRCP(x) (1.0 / (x))
#define KRCPF(x) (1.0 / (x))
#endif

// barrier contributions of one bounded variable at the current iterate:
//   sigma = zL/sl + zU/su,  rb = -mu/sl + mu/su (+- kappa_d mu for one-sided bounds)
KMPC_HD void bound_terms(double val, double lb, double ub, int hL, int hU, double zL, double zU, double mu, double *sigma,
                         double *rb) {
    double sg = 0.0, r = 0.0;
    if (hL) { const double rs = KRCPF(val - lb); sg += zL * rs; r -= mu * rs; if (!hU) r += K_KAPPA_D * mu; }
    if (hU) { const double rs = KRCPF(ub - val); sg += zU * rs; r += mu * rs; if (!hL) r -= K_KAPPA_D * mu; }
    *sigma = sg; *rb = r;
}

// fraction-to-the-boundary of one bounded variable for the primal step d, and of its multipliers for the
// induced dual steps; accumulates the directional derivative of the barrier term.
KMPC_HD void bound_ftb(double val, double d, double lb, double ub, int hL, int hU, double zL, double zU, double mu,
                       double tau, double *apr, double *adu) {
    if (hL) {
        const double sl = val - lb, rs = KRCPF(sl);
        if (d < 0) *apr = fmin(*apr, -tau * sl / d);
        const double dz = mu * rs - zL - zL * rs * d;
        if (dz < 0) *adu = fmin(*adu, -tau * zL / dz);
    }
    if (hU) {
        const double su = ub - val, rs = KRCPF(su);
        if (d > 0) *apr = fmin(*apr, tau * su / d);
        const double dz = mu * rs - zU + zU * rs * d;
        if (dz < 0) *adu = fmin(*adu, -tau * zU / dz);
    }
}

// trial-point treatment of one bounded variable: new multipliers (with the kappa_sigma safeguard), barrier product,
// damping, complementarity stats.  Returns false if the trial value is not strictly inside its bounds.
KMPC_HD bool bound_trial(double val, double d, double vt, double lb, double ub, int hL, int hU, double zL, double zU,
                         double mu, double adu, bool clamp, double *zLn, double *zUn, double *prod, double *damp,
                         Stats *st) {
    bool ok = true;
    *zLn = 0.0; *zUn = 0.0;
    if (hL) {
        const double sl = val - lb, sn = vt - lb, rs = KRCPF(sl);
        if (!(sn > 0)) ok = false;
        *prod *= sn;
        if (!hU) *damp += sn;
        double z = zL + adu * (mu * rs - zL - zL * rs * d);
        if (clamp) { const double mr = mu * KRCPF(sn); z = fmax(fmin(z, K_KAPPA_SIGMA * mr), mr * (1.0 / K_KAPPA_SIGMA)); }
        *zLn = z;
        double p = sn * z;
        st->mn = fmin(st->mn, p); st->mx = fmax(st->mx, p); st->sumz += fabs(z);
    }
    if (hU) {
        const double su = ub - val, sn = ub - vt, rs = KRCPF(su);
        if (!(sn > 0)) ok = false;
        *prod *= sn;
        if (!hL) *damp += sn;
        double z = zU + adu * (mu * rs - zU + zU * rs * d);
        if (clamp) { const double mr = mu * KRCPF(sn); z = fmax(fmin(z, K_KAPPA_SIGMA * mr), mr * (1.0 / K_KAPPA_SIGMA)); }
        *zUn = z;
        double p = sn * z;
        st->mn = fmin(st->mn, p); st->mx = fmax(st->mx, p); st->sumz += fabs(z);
    }
    return ok;
}

// compare-and-select min / max (3 instructions on the device; fmin / fmax expand to ~9 with their NaN canonicalisation).
// A NaN in the FIRST argument is ignored (as fmin / fmax would), so candidates go first, accumulators second.
KMPC_HD double kmin(double cand, double acc) { return cand < acc ? cand : acc; }
KMPC_HD double kmax(double cand, double acc) { return cand > acc ? cand : acc; }
// fmax / fmin with their exact NaN semantics (a NaN operand is dropped) in 4 instructions instead of ~9
KMPC_HD double kfmax(double a, double b) { return (a > b || b != b) ? a : b; }
KMPC_HD double kfmin(double a, double b) { return (a < b || b != b) ? a : b; }
KMPC_HD double maxabs_nan(double m, double v) { double t = fabs(v); return (t > m || t != t) ? t : m; }

// ------------------------------------------------------------------------------------------------
// One step of the backward Riccati recursion for the unicycle stage
//   A = I + a13 e1 e3^T + a23 e2 e3^T,   B = [b11 0; b21 0; 0 T],
// Q_uu (2x2) is inverted through its determinant: d1 > 0 and det > 0 <=> Q_uu positive definite (the inertia test).
//   in : (P, p) of stage k+1, stage blocks Q (xx, with Q01 only for obstacle rows), q, (qv, qw), (dv, dw) = diag of
//        W_uu + Sigma_u + delta, htv = W_v,theta, e = bc_{k+1}
//   out: (P, p) of stage k (overwritten), feedback K (2x3), feed-forward kf.  Returns false on a non-positive pivot.
// ------------------------------------------------------------------------------------------------
struct RicK { double K00, K01, K02, K10, K11, K12, kf0, kf1; };
KMPC_HD bool riccati_step(double &P00, double &P10, double &P11, double &P20, double &P21, double &P22, double &p0, double &p1,
                          double &p2, double a13, double a23, double b11, double b21, double T, double Q00, double Q01, double Q11,
                          double Q22, double q0, double q1, double q2, double qv, double qw, double dv, double dw, double htv,
                          double e0, double e1, double e2, RicK &o) {
    // P A (third column), symmetric Qxx = A^T P A + Q
    const double PA02 = fma(P00, a13, fma(P10, a23, P20)), PA12 = fma(P10, a13, fma(P11, a23, P21)), PA22 = fma(P20, a13, fma(P21, a23, P22));
    const double X00 = P00 + Q00, X10 = P10 + Q01, X11 = P11 + Q11, X20 = PA02, X21 = PA12;
    const double X22 = fma(a13, PA02, fma(a23, PA12, PA22)) + Q22;
    // Qux = B^T P A (+ W_v,theta)
    const double U00 = fma(b11, P00, b21 * P10), U01 = fma(b11, P10, b21 * P11), U02 = fma(b11, PA02, fma(b21, PA12, htv));
    const double U10 = T * P20, U11 = T * P21, U12 = T * PA22;
    // Quu = B^T P B + diag = [d1 qb; qb qc];  positive definite <=> d1 > 0 and det > 0
    const double d1 = fma(b11, U00, fma(b21, U01, dv)), qb = fma(b11, U10, b21 * U11), qc = fma(T * T, P22, dw);
    const double det = fma(d1, qc, -(qb * qb));
    const bool pd = d1 > 0.0 && det > 0.0;  // branch-free: the caller discards the outputs when Quu is not positive definite
    const double r = KRCPF(det), i00 = qc * r, i01 = -(qb * r), i11 = d1 * r;
    // K = -Quu^-1 Qux
    const double K00 = -fma(i00, U00, i01 * U10), K01 = -fma(i00, U01, i01 * U11), K02 = -fma(i00, U02, i01 * U12);
    const double K10 = -fma(i01, U00, i11 * U10), K11 = -fma(i01, U01, i11 * U11), K12 = -fma(i01, U02, i11 * U12);
    // vector part
    const double Pe0 = fma(P00, e0, fma(P10, e1, fma(P20, e2, p0))), Pe1 = fma(P10, e0, fma(P11, e1, fma(P21, e2, p1))),
                 Pe2 = fma(P20, e0, fma(P21, e1, fma(P22, e2, p2)));
    const double qu0 = fma(b11, Pe0, fma(b21, Pe1, qv)), qu1 = fma(T, Pe2, qw);
    p0 = fma(K00, qu0, fma(K10, qu1, q0 + Pe0));
    p1 = fma(K01, qu0, fma(K11, qu1, q1 + Pe1));
    p2 = fma(K02, qu0, fma(K12, qu1, fma(a13, Pe0, fma(a23, Pe1, q2 + Pe2))));
    // P <- Qxx + Qux^T K  (lower triangle; symmetric in exact arithmetic)
    P00 = fma(U00, K00, fma(U10, K10, X00)); P10 = fma(U01, K00, fma(U11, K10, X10)); P11 = fma(U01, K01, fma(U11, K11, X11));
    P20 = fma(U02, K00, fma(U12, K10, X20)); P21 = fma(U02, K01, fma(U12, K11, X21)); P22 = fma(U02, K02, fma(U12, K12, X22));
    o.K00 = K00; o.K01 = K01; o.K02 = K02; o.K10 = K10; o.K11 = K11; o.K12 = K12;
    o.kf0 = -fma(i00, qu0, i01 * qu1); o.kf1 = -fma(i01, qu0, i11 * qu1);
    return pd;
}

// ------------------------------------------------------------------------------------------------
// record access helpers.  ST(p, f): field f of the record p points at.
// ------------------------------------------------------------------------------------------------
#define FD(p, f) (p)[(size_t)(f) * S]

template <int NR>
KMPC_HD void rec_load(double (&r)[NR], const double *p, size_t S) {
#pragma unroll
    for (int j = 0; j < NR; ++j) r[j] = FD(p, j);
}
template <int NR>
KMPC_HD void rec_copy(double (&d)[NR], const double (&s)[NR]) {
#pragma unroll
    for (int j = 0; j < NR; ++j) d[j] = s[j];
}

KMPC_HD double push_in(double v, double lb, double ub, int hL, int hU) {
    if (hL && hU) {
        const double pl = fmin(K_BOUND_PUSH * fmax(1.0, fabs(lb)), K_BOUND_FRAC * (ub - lb));
        const double pu = fmin(K_BOUND_PUSH * fmax(1.0, fabs(ub)), K_BOUND_FRAC * (ub - lb));
        return fmin(fmax(v, lb + pl), ub - pu);
    }
    if (hL) return fmax(v, lb + K_BOUND_PUSH * fmax(1.0, fabs(lb)));
    if (hU) return fmin(v, ub - K_BOUND_PUSH * fmax(1.0, fabs(ub)));
    return v;
}

// ------------------------------------------------------------------------------------------------
// INIT pass: starting point (optimizer.py:375-385; cold start agent.py:59-60), objective scaling, push into the
// interior, z = 1, y = 0, slacks.  Writes state buffer 0.
// ------------------------------------------------------------------------------------------------
KMPC_HDN inline void pass_init(const Cfg &c, Ctx &t, double *wsp, size_t S, const IO &io) {
    const int N = c.N, O = c.O, b = t.inst;
    const Rows &L = c.L;
    double *sc = wsp + (size_t)L.rSc * S;
    double xc[3], gl[3];
    for (int j = 0; j < 3; ++j) {
        xc[j] = io.x_cur[io_vec3(c, b, j)]; gl[j] = io.goal[io_vec3(c, b, j)];
        FD(sc, j) = xc[j]; FD(sc, 3 + j) = gl[j];
    }
    if (c.obs_sw) {
        for (int o = 0; o < O; ++o) for (int t = 0; t < N; ++t) for (int j = 0; j < 2; ++j)
            FD(sc, CEN_ROW(o, t + 1, j)) = io.obs[io_obs_sw(c, b, o, t, j)];
    } else {
        for (int o = 0; o < O; ++o) for (int j = 0; j < 2; ++j) FD(sc, 6 + 2 * o + j) = io.obs[io_obs(c, b, o, j)];
    }
    for (int o = 0; o < O; ++o) FD(sc, RAD_ROW(o)) = io.orad ? io.orad[io_orad(c, b, o)] : c.obs_radius;
    double gm = 0.0;
    const double dLpush = c.dL + K_BOUND_PUSH * fmax(1.0, fabs(c.dL));
    double *ps = wsp + (size_t)L.rState[0] * S;
    double *po = wsp + (size_t)(L.rState[0] + L.sObs) * S;
#pragma unroll 1
    for (int k = 0; k <= N; ++k, ps += (size_t)NSTATE * S) {
        double x[3], u[2] = {0.0, 0.0};
        for (int j = 0; j < 3; ++j) x[j] = io.X0 ? io.X0[io_X(c, b, j, k)] : xc[j];
        if (k < N) for (int j = 0; j < 2; ++j) u[j] = io.U0 ? io.U0[io_U(c, b, j, k)] : 0.0;
        if (k >= c.gk_lo && k <= c.gk_hi)
            for (int j = 0; j < 3; ++j) gm = maxabs_nan(gm, 2.0 * c.W[j] * (x[j] - gl[j]));
        x[0] = push_in(x[0], c.lb[0], c.ub[0], c.hasL[0], c.hasU[0]);
        x[1] = push_in(x[1], c.lb[1], c.ub[1], c.hasL[1], c.hasU[1]);
        FD(ps, F_X0) = x[0]; FD(ps, F_X1) = x[1]; FD(ps, F_X2) = x[2];
        FD(ps, F_Y0) = 0.0; FD(ps, F_Y1) = 0.0; FD(ps, F_Y2) = 0.0;
        FD(ps, F_ZLX) = c.hasL[0] ? 1.0 : 0.0; FD(ps, F_ZUX) = c.hasU[0] ? 1.0 : 0.0;
        FD(ps, F_ZLY) = c.hasL[1] ? 1.0 : 0.0; FD(ps, F_ZUY) = c.hasU[1] ? 1.0 : 0.0;
        double cs = 1.0, sn = 0.0;
        const bool hasu = k < N;
        if (hasu) {
            double gv, hv;
            vcost(c, 1.0, u[0], &gv, &hv);
            gm = maxabs_nan(gm, gv); gm = maxabs_nan(gm, 2.0 * c.Ww * u[1]);
            u[0] = push_in(u[0], c.lb[2], c.ub[2], c.hasL[2], c.hasU[2]);
            u[1] = push_in(u[1], c.lb[3], c.ub[3], c.hasL[3], c.hasU[3]);
            sincos_(x[2], &sn, &cs);
        }
        FD(ps, F_V) = u[0]; FD(ps, F_OM) = u[1];
        FD(ps, F_ZLV) = (hasu && c.hasL[2]) ? 1.0 : 0.0; FD(ps, F_ZUV) = (hasu && c.hasU[2]) ? 1.0 : 0.0;
        FD(ps, F_ZLW) = (hasu && c.hasL[3]) ? 1.0 : 0.0; FD(ps, F_ZUW) = (hasu && c.hasU[3]) ? 1.0 : 0.0;
        FD(ps, F_CS) = cs; FD(ps, F_SN) = sn;
        if (k >= 1)
            for (int o = 0; o < O; ++o, po += (size_t)3 * S) {
                const double ex = x[0] - FD(sc, CEN_ROW(o, k, 0)), ey = x[1] - FD(sc, CEN_ROW(o, k, 1));
                const double d = sqrt(ex * ex + ey * ey) - FD(sc, RAD_ROW(o));
                FD(po, 0) = fmax(d, dLpush); FD(po, 1) = 0.0; FD(po, 2) = 1.0;
            }
    }
    t.df = gm > K_SCALING_MAX_GRAD ? fmax(K_SCALING_MAX_GRAD / gm, K_SCALING_MIN) : 1.0;
    t.cur = 0; t.iter = 0; t.mu = K_MU_INIT; t.tau = fmax(K_TAU_MIN, 1.0 - K_MU_INIT);
    t.delta = 0.0; t.delta_last = 0.0; t.theta_max = -1.0; t.theta_min = -1.0; t.fn = 0;
    t.nsteps = 0; t.soc_count = 0; t.trips = 0; t.sel = 0; t.tu = TU_INIT;
    t.alpha = t.alpha_test = t.alpha_min = t.alpha_du0 = t.alpha_soc = t.gBD = t.theta_soc_old = t.theta_trial = 0.0;
    t.a_pr = t.a_y = t.a_du = 0.0; t.pw_g = t.pw_t = 0.0;
    t.c.f = t.c.bar = t.c.damp = t.c.theta = t.c.dinf = t.c.pinf = t.c.mn = t.c.mx = t.c.sumy = t.c.sumz = t.c.wmax = 0.0;
    t.mode = M_LSQ;
}

// per-(stage, obstacle) quantities shared by the sweep and the roll-out
struct ObsT { double nx, ny, rr, Ds, bd, bs; };
KMPC_HD ObsT obs_terms(const Cfg &c, double px, double py, double cx, double cy, double rad, double s, double yd, double vL,
                       double mu, double delta, bool lsq, bool soc, double dsoc) {
    ObsT r;
    const double ex = px - cx, ey = py - cy;
    r.rr = sqrt(ex * ex + ey * ey);
    r.nx = ex / r.rr; r.ny = ey / r.rr;
    if (lsq) { r.Ds = 1.0; r.bd = 0.0; r.bs = -yd - vL; }
    else {
        const double sl = s - c.dL;
        r.Ds = vL / sl + delta;
        r.bs = yd + mu / sl - K_KAPPA_D * mu;
        r.bd = soc ? -dsoc : -((r.rr - rad) - s);
    }
    return r;
}

// ------------------------------------------------------------------------------------------------
// SWEEP: backward Riccati recursion over the stages of the primal-dual system
//   [W + Sigma + delta I   J^T] [dx ]   [bx]
//   [J                      0 ] [dy ] = [bc]        (slacks of the obstacle rows condensed into the x-x blocks)
// kind M_LSQ: W = 0, Sigma = I, rhs = (grad f - zL + zU, 0)  (least-squares multiplier estimate)
// kind M_NEWTON: rhs = -(grad of the barrier Lagrangian, c);  kind M_SOC: same matrix, bc = -c_soc.
// Stores the feedback gains K, k_ff and the cost-to-go (P, p) as FACT records.  Returns false when some Q_uu is not
// positive definite (wrong inertia -> the caller raises delta, IPOPT's inertia correction).
// The STATE record of stage k-1 is loaded while stage k is being processed (register double buffer).
// ------------------------------------------------------------------------------------------------
template <bool OBS>
KMPC_HDN inline bool pass_sweep(const Cfg &c, const Ctx &t, double *wsp, size_t S) {
    const int N = c.N, O = OBS ? c.O : 0;
    const Rows &L = c.L;
    const bool lsq = t.mode == M_LSQ, soc = t.mode == M_SOC;
    const double mu = t.mu, delta = t.delta, df = t.df, T = c.T;
    const double *__restrict__ sc = wsp + (size_t)L.rSc * S;
    const double g0 = FD(sc, 3), g1 = FD(sc, 4), g2 = FD(sc, 5);
    const double *__restrict__ ps = wsp + ((size_t)L.rState[t.cur] + (size_t)NSTATE * N) * S;
    const double *__restrict__ po = wsp + ((size_t)L.rState[t.cur] + L.sObs + (size_t)3 * N * O) * S;  // one past the last obstacle record
    const double *__restrict__ pcs = wsp + ((size_t)L.rCsoc + 3 * (N + 1)) * S;  // record k+1 once stage k is reached
    const double *__restrict__ pds = wsp + ((size_t)L.rDsoc + (size_t)N * O) * S;
    double *__restrict__ pf = wsp + ((size_t)L.rFact + (size_t)NFACT * N) * S;
    double P00 = 0, P10 = 0, P11 = 0, P20 = 0, P21 = 0, P22 = 0, p0 = 0, p1 = 0, p2 = 0;
    double xn0 = 0, xn1 = 0, xn2 = 0, yn0 = 0, yn1 = 0, yn2 = 0;
    double a[NSTATE], nx[NSTATE];
    rec_load(a, ps, S);
    bool ok = true;
#pragma unroll 1
    for (int k = N; k >= 0; --k) {
        if (k > 0) rec_load(nx, ps - (size_t)NSTATE * S, S);
        const double x0 = a[F_X0], x1 = a[F_X1], x2 = a[F_X2], y0 = a[F_Y0], y1 = a[F_Y1], y2 = a[F_Y2];
        const bool ing = k >= c.gk_lo && k <= c.gk_hi;
        double gx0 = 0, gx1 = 0, gx2 = 0, h0 = 0, h1 = 0, h2 = 0;
        if (ing) {
            gx0 = df * 2.0 * c.W[0] * (x0 - g0); gx1 = df * 2.0 * c.W[1] * (x1 - g1); gx2 = df * 2.0 * c.W[2] * (x2 - g2);
            h0 = df * 2.0 * c.W[0]; h1 = df * 2.0 * c.W[1]; h2 = df * 2.0 * c.W[2];
        }
        // q = -bx (x part), Q = W_xx + Sigma_x + delta
        double q0, q1, q2, Q00, Q01 = 0.0, Q11, Q22;
        if (lsq) {
            q0 = -(gx0 - a[F_ZLX] + a[F_ZUX]); q1 = -(gx1 - a[F_ZLY] + a[F_ZUY]); q2 = -gx2;
            Q00 = 1.0; Q11 = 1.0; Q22 = 1.0;
        } else {
            double sg0, rb0, sg1, rb1;
            bound_terms(x0, c.lb[0], c.ub[0], c.hasL[0], c.hasU[0], a[F_ZLX], a[F_ZUX], mu, &sg0, &rb0);
            bound_terms(x1, c.lb[1], c.ub[1], c.hasL[1], c.hasU[1], a[F_ZLY], a[F_ZUY], mu, &sg1, &rb1);
            q0 = gx0 + y0 + rb0; q1 = gx1 + y1 + rb1; q2 = gx2 + y2;
            Q00 = h0 + sg0 + delta; Q11 = h1 + sg1 + delta; Q22 = h2 + delta;
        }
        if (OBS && k >= 1) {
#pragma unroll 1
            for (int o = O - 1; o >= 0; --o) {
                po -= (size_t)3 * S; pds -= S;
                const double yd = FD(po, 1);
                ObsT ot = obs_terms(c, x0, x1, FD(sc, CEN_ROW(o, k, 0)), FD(sc, CEN_ROW(o, k, 1)), FD(sc, RAD_ROW(o)), FD(po, 0), yd, FD(po, 2), mu, delta, lsq,
                                    soc, soc ? FD(pds, 0) : 0.0);
                if (!lsq) {
                    const double h = yd / ot.rr;
                    Q00 += h * (1.0 - ot.nx * ot.nx); Q01 += h * (-ot.nx * ot.ny); Q11 += h * (1.0 - ot.ny * ot.ny);
                    q0 += ot.nx * yd; q1 += ot.ny * yd;
                }
                Q00 += ot.Ds * ot.nx * ot.nx; Q01 += ot.Ds * ot.nx * ot.ny; Q11 += ot.Ds * ot.ny * ot.ny;
                const double tt = ot.Ds * ot.bd + ot.bs;
                q0 -= ot.nx * tt; q1 -= ot.ny * tt;
            }
        }
        if (k == N) {
            P00 = Q00; P10 = Q01; P11 = Q11; P20 = 0.0; P21 = 0.0; P22 = Q22;
            p0 = q0; p1 = q1; p2 = q2;
        } else {
            const double v = a[F_V], om = a[F_OM], cs = a[F_CS], sn = a[F_SN];
            const double a13 = -T * v * sn, a23 = T * v * cs, b11 = T * cs, b21 = T * sn;
            double gv, hvv, qv, qw, Dv, Dw, htv = 0.0;
            vcost(c, df, v, &gv, &hvv);
            const double gw = df * 2.0 * c.Ww * om;
            double hww = df * 2.0 * c.Ww;
            if (lsq) {
                qv = -(gv - a[F_ZLV] + a[F_ZUV]); qw = -(gw - a[F_ZLW] + a[F_ZUW]);
                Dv = 1.0; Dw = 1.0; hvv = 0.0; hww = 0.0;
            } else {
                double sgv, rbv, sgw, rbw;
                bound_terms(v, c.lb[2], c.ub[2], c.hasL[2], c.hasU[2], a[F_ZLV], a[F_ZUV], mu, &sgv, &rbv);
                bound_terms(om, c.lb[3], c.ub[3], c.hasL[3], c.hasU[3], a[F_ZLW], a[F_ZUW], mu, &sgw, &rbw);
                // J^T y of dynamics row k+1 (multiplier yn)
                q0 -= yn0; q1 -= yn1; q2 -= a13 * yn0 + a23 * yn1 + yn2;
                qv = gv - (b11 * yn0 + b21 * yn1) + rbv;
                qw = gw - T * yn2 + rbw;
                Dv = sgv + delta; Dw = sgw + delta;
                // curvature of the dynamics in the Lagrangian (the only indefinite terms)
                Q22 += T * v * (yn0 * cs + yn1 * sn);
                htv = T * (yn0 * sn - yn1 * cs);
            }
            // e = bc_{k+1}
            double e0, e1, e2;
            if (lsq) { e0 = e1 = e2 = 0.0; }
            else if (soc) { e0 = -FD(pcs, 0); e1 = -FD(pcs, 1); e2 = -FD(pcs, 2); }
            else { e0 = -(xn0 - (x0 + T * v * cs)); e1 = -(xn1 - (x1 + T * v * sn)); e2 = -(xn2 - (x2 + T * om)); }
            RicK rk;
            if (!riccati_step(P00, P10, P11, P20, P21, P22, p0, p1, p2, a13, a23, b11, b21, T, Q00, Q01, Q11, Q22, q0, q1, q2, qv, qw,
                              hvv + Dv, hww + Dw, htv, e0, e1, e2, rk)) { ok = false; break; }
            const double K00 = rk.K00, K01 = rk.K01, K02 = rk.K02, K10 = rk.K10, K11 = rk.K11, K12 = rk.K12, kf0 = rk.kf0, kf1 = rk.kf1;
            FD(pf, A_K00) = K00; FD(pf, A_K01) = K01; FD(pf, A_K02) = K02;
            FD(pf, A_K10) = K10; FD(pf, A_K11) = K11; FD(pf, A_K12) = K12;
            FD(pf, A_KF0) = kf0; FD(pf, A_KF1) = kf1;
        }
        FD(pf, A_P00) = P00; FD(pf, A_P10) = P10; FD(pf, A_P11) = P11;
        FD(pf, A_P20) = P20; FD(pf, A_P21) = P21; FD(pf, A_P22) = P22;
        FD(pf, A_PV0) = p0; FD(pf, A_PV1) = p1; FD(pf, A_PV2) = p2;
        xn0 = x0; xn1 = x1; xn2 = x2; yn0 = y0; yn1 = y1; yn2 = y2;
        rec_copy(a, nx);
        ps -= (size_t)NSTATE * S; pf -= (size_t)NFACT * S; pcs -= (size_t)3 * S;
    }
    return ok;
}

// ------------------------------------------------------------------------------------------------
// ROLL-OUT: forward substitution dx_0 = bc_0, du = K dx + k_ff, dx+ = A dx + B du + e, dy = -(P dx + p);
// fraction-to-the-boundary step sizes for the primal step and for the bound multipliers, and the
// directional derivative of the barrier objective.  Writes STEP records of step buffer `sel`.
// ------------------------------------------------------------------------------------------------
template <bool OBS>
KMPC_HDN inline void pass_rollout(const Cfg &c, const Ctx &t, double *wsp, size_t S, int sel, double *alpha_pr,
                                  double *alpha_du, double *gBD, double *ymax) {
    const int N = c.N, O = OBS ? c.O : 0;
    const Rows &L = c.L;
    const bool lsq = t.mode == M_LSQ, soc = t.mode == M_SOC;
    const double mu = t.mu, delta = t.delta, df = t.df, T = c.T, tau = t.tau;
    const double *__restrict__ sc = wsp + (size_t)L.rSc * S;
    const double g0 = FD(sc, 3), g1 = FD(sc, 4), g2 = FD(sc, 5);
    const double *__restrict__ ps = wsp + (size_t)L.rState[t.cur] * S;
    const double *__restrict__ po = wsp + ((size_t)L.rState[t.cur] + L.sObs) * S;
    const double *__restrict__ pf = wsp + (size_t)L.rFact * S;
    const double *__restrict__ pcs = wsp + (size_t)L.rCsoc * S;
    const double *__restrict__ pds = wsp + (size_t)L.rDsoc * S;
    double *__restrict__ pd = wsp + (size_t)L.rStep[sel] * S;
    double *__restrict__ pdo = wsp + ((size_t)L.rStep[sel] + L.dObs) * S;
    double apr = 1.0, adu = 1.0, gbd = 0.0, ym = 0.0;
    double a[NSTATE], an[NSTATE], f[NFACT], fn[NFACT];
    rec_load(a, ps, S); rec_load(f, pf, S);
    double d0, d1, d2;
    if (lsq) { d0 = d1 = d2 = 0.0; }
    else if (soc) { d0 = -FD(pcs, 0); d1 = -FD(pcs, 1); d2 = -FD(pcs, 2); }
    else { d0 = -(a[F_X0] - FD(sc, 0)); d1 = -(a[F_X1] - FD(sc, 1)); d2 = -(a[F_X2] - FD(sc, 2)); }
#pragma unroll 1
    for (int k = 0; k <= N; ++k) {
        if (k < N) { rec_load(an, ps + (size_t)NSTATE * S, S); rec_load(fn, pf + (size_t)NFACT * S, S); }
        const double x0 = a[F_X0], x1 = a[F_X1], x2 = a[F_X2];
        const double dy0 = -(f[A_P00] * d0 + f[A_P10] * d1 + f[A_P20] * d2 + f[A_PV0]);
        const double dy1 = -(f[A_P10] * d0 + f[A_P11] * d1 + f[A_P21] * d2 + f[A_PV1]);
        const double dy2 = -(f[A_P20] * d0 + f[A_P21] * d1 + f[A_P22] * d2 + f[A_PV2]);
        FD(pd, D_Y0) = dy0; FD(pd, D_Y1) = dy1; FD(pd, D_Y2) = dy2;
        FD(pd, D_X0) = d0; FD(pd, D_X1) = d1; FD(pd, D_X2) = d2;
        ym = maxabs_nan(maxabs_nan(maxabs_nan(ym, dy0), dy1), dy2);
        if (!lsq) {
            double sg, rb;
            bound_ftb(x0, d0, c.lb[0], c.ub[0], c.hasL[0], c.hasU[0], a[F_ZLX], a[F_ZUX], mu, tau, &apr, &adu);
            bound_ftb(x1, d1, c.lb[1], c.ub[1], c.hasL[1], c.hasU[1], a[F_ZLY], a[F_ZUY], mu, tau, &apr, &adu);
            const bool ing = k >= c.gk_lo && k <= c.gk_hi;
            bound_terms(x0, c.lb[0], c.ub[0], c.hasL[0], c.hasU[0], a[F_ZLX], a[F_ZUX], mu, &sg, &rb);
            gbd += ((ing ? df * 2.0 * c.W[0] * (x0 - g0) : 0.0) + rb) * d0;
            bound_terms(x1, c.lb[1], c.ub[1], c.hasL[1], c.hasU[1], a[F_ZLY], a[F_ZUY], mu, &sg, &rb);
            gbd += ((ing ? df * 2.0 * c.W[1] * (x1 - g1) : 0.0) + rb) * d1;
            gbd += (ing ? df * 2.0 * c.W[2] * (x2 - g2) : 0.0) * d2;
        }
        if (OBS && k >= 1) {
#pragma unroll 1
            for (int o = 0; o < O; ++o, po += (size_t)3 * S, pdo += (size_t)2 * S, pds += S) {
                const double s = FD(po, 0), yd = FD(po, 1), vL = FD(po, 2);
                ObsT ot = obs_terms(c, x0, x1, FD(sc, CEN_ROW(o, k, 0)), FD(sc, CEN_ROW(o, k, 1)), FD(sc, RAD_ROW(o)), s, yd, vL, mu, delta, lsq, soc,
                                    soc ? FD(pds, 0) : 0.0);
                const double ds = ot.nx * d0 + ot.ny * d1 - ot.bd;
                const double dyd = ot.Ds * ds - ot.bs;
                FD(pdo, 0) = ds; FD(pdo, 1) = dyd;
                ym = maxabs_nan(ym, dyd);
                if (!lsq) {
                    const double sl = s - c.dL;
                    if (ds < 0) apr = fmin(apr, -tau * sl / ds);
                    const double dv = mu / sl - vL - vL / sl * ds;
                    if (dv < 0) adu = fmin(adu, -tau * vL / dv);
                    gbd += (-mu / sl + K_KAPPA_D * mu) * ds;
                }
            }
        }
        if (k < N) {
            const double v = a[F_V], om = a[F_OM], cs = a[F_CS], sn = a[F_SN];
            const double du0 = f[A_K00] * d0 + f[A_K01] * d1 + f[A_K02] * d2 + f[A_KF0];
            const double du1 = f[A_K10] * d0 + f[A_K11] * d1 + f[A_K12] * d2 + f[A_KF1];
            FD(pd, D_U0) = du0; FD(pd, D_U1) = du1;
            double e0, e1, e2;
            if (lsq) { e0 = e1 = e2 = 0.0; }
            else {
                bound_ftb(v, du0, c.lb[2], c.ub[2], c.hasL[2], c.hasU[2], a[F_ZLV], a[F_ZUV], mu, tau, &apr, &adu);
                bound_ftb(om, du1, c.lb[3], c.ub[3], c.hasL[3], c.hasU[3], a[F_ZLW], a[F_ZUW], mu, tau, &apr, &adu);
                double gv, hv, sg, rb;
                vcost(c, df, v, &gv, &hv);
                bound_terms(v, c.lb[2], c.ub[2], c.hasL[2], c.hasU[2], a[F_ZLV], a[F_ZUV], mu, &sg, &rb);
                gbd += (gv + rb) * du0;
                bound_terms(om, c.lb[3], c.ub[3], c.hasL[3], c.hasU[3], a[F_ZLW], a[F_ZUW], mu, &sg, &rb);
                gbd += (df * 2.0 * c.Ww * om + rb) * du1;
                if (soc) { e0 = -FD(pcs, 3); e1 = -FD(pcs, 4); e2 = -FD(pcs, 5); }
                else { e0 = -(an[F_X0] - (x0 + T * v * cs)); e1 = -(an[F_X1] - (x1 + T * v * sn)); e2 = -(an[F_X2] - (x2 + T * om)); }
            }
            const double a13 = -T * v * sn, a23 = T * v * cs;
            const double n0 = d0 + a13 * d2 + T * cs * du0 + e0;
            const double n1 = d1 + a23 * d2 + T * sn * du0 + e1;
            const double n2 = d2 + T * du1 + e2;
            d0 = n0; d1 = n1; d2 = n2;
        } else { FD(pd, D_U0) = 0.0; FD(pd, D_U1) = 0.0; }
        rec_copy(a, an); rec_copy(f, fn);
        ps += (size_t)NSTATE * S; pf += (size_t)NFACT * S; pd += (size_t)NSTEP * S; pcs += (size_t)3 * S;
    }
    *alpha_pr = apr; *alpha_du = adu; *gBD = gbd; *ymax = ym;
}

// ------------------------------------------------------------------------------------------------
// TRIAL + speculative UPDATE: evaluates the trial point w + alpha d (barrier objective ingredients, constraint
// violation) and, in the same sweep over the stages, the iterate that would result from accepting it (multiplier
// updates with their own step sizes, kappa_sigma safeguard) together with its optimality-error norms.  Everything
// is written to the OTHER state buffer; accepting the trial point just flips t.cur.
//   tu = TU_INIT : alpha = 0, y' = ay * dy (least-squares estimate), multipliers untouched (no safeguard)
// ------------------------------------------------------------------------------------------------
template <bool OBS>
KMPC_HDN inline bool pass_trial(const Cfg &c, const Ctx &t, double *wsp, size_t S, int sel, int tu, double alpha, double ay,
                                double adu, Stats *out) {
    const int N = c.N, O = OBS ? c.O : 0;
    const Rows &L = c.L;
    const double mu = t.mu, df = t.df, T = c.T;
    const bool clamp = tu == TU_STEP;
    const double *__restrict__ sc = wsp + (size_t)L.rSc * S;
    const double g0 = FD(sc, 3), g1 = FD(sc, 4), g2 = FD(sc, 5);
    const double *__restrict__ ps = wsp + (size_t)L.rState[t.cur] * S;
    const double *__restrict__ po = wsp + ((size_t)L.rState[t.cur] + L.sObs) * S;
    const double *__restrict__ pd = wsp + (size_t)L.rStep[sel] * S;
    const double *__restrict__ pdo = wsp + ((size_t)L.rStep[sel] + L.dObs) * S;
    double *__restrict__ pn = wsp + (size_t)L.rState[t.cur ^ 1] * S;
    double *__restrict__ pno = wsp + ((size_t)L.rState[t.cur ^ 1] + L.sObs) * S;
    Stats st;
    st.f = 0; st.bar = 0; st.damp = 0; st.theta = 0; st.dinf = 0; st.pinf = 0; st.mn = INFINITY; st.mx = 0; st.sumy = 0;
    st.sumz = 0; st.wmax = 0;
    bool valid = true;
    double xp0 = FD(sc, 0), xp1 = FD(sc, 1), xp2 = FD(sc, 2);  // predicted state (k = 0: x_cur)
    double a[NSTATE - 2], an[NSTATE - 2], d[NSTEP], dn[NSTEP];  // CS/SN of the old point are not needed
    rec_load(a, ps, S); rec_load(d, pd, S);
    double yk0 = a[F_Y0] + ay * d[D_Y0], yk1 = a[F_Y1] + ay * d[D_Y1], yk2 = a[F_Y2] + ay * d[D_Y2];
#pragma unroll 1
    for (int k = 0; k <= N; ++k) {
        if (k < N) { rec_load(an, ps + (size_t)NSTATE * S, S); rec_load(dn, pd + (size_t)NSTEP * S, S); }
        const double x0 = a[F_X0] + alpha * d[D_X0], x1 = a[F_X1] + alpha * d[D_X1], x2 = a[F_X2] + alpha * d[D_X2];
        FD(pn, F_X0) = x0; FD(pn, F_X1) = x1; FD(pn, F_X2) = x2;
        FD(pn, F_Y0) = yk0; FD(pn, F_Y1) = yk1; FD(pn, F_Y2) = yk2;
        const double c0 = x0 - xp0, c1 = x1 - xp1, c2 = x2 - xp2;
        st.theta += fabs(c0) + fabs(c1) + fabs(c2);
        st.pinf = maxabs_nan(maxabs_nan(maxabs_nan(st.pinf, c0), c1), c2);
        st.sumy += fabs(yk0) + fabs(yk1) + fabs(yk2);
        st.wmax = fmax(st.wmax, fmax(fabs(x0), fmax(fabs(x1), fabs(x2))));
        double r0 = yk0, r1 = yk1, r2 = yk2;  // dual residual of x_k
        if (k >= c.gk_lo && k <= c.gk_hi) {
            const double e0 = x0 - g0, e1 = x1 - g1, e2 = x2 - g2;
            st.f += c.W[0] * e0 * e0; st.f += c.W[1] * e1 * e1; st.f += c.W[2] * e2 * e2;
            r0 += df * 2.0 * c.W[0] * e0; r1 += df * 2.0 * c.W[1] * e1; r2 += df * 2.0 * c.W[2] * e2;
        }
        double prod = 1.0;
        {
            double zLn, zUn;
            valid &= bound_trial(a[F_X0], d[D_X0], x0, c.lb[0], c.ub[0], c.hasL[0], c.hasU[0], a[F_ZLX], a[F_ZUX], mu, adu, clamp,
                                 &zLn, &zUn, &prod, &st.damp, &st);
            FD(pn, F_ZLX) = zLn; FD(pn, F_ZUX) = zUn;
            r0 += zUn - zLn;
            valid &= bound_trial(a[F_X1], d[D_X1], x1, c.lb[1], c.ub[1], c.hasL[1], c.hasU[1], a[F_ZLY], a[F_ZUY], mu, adu, clamp,
                                 &zLn, &zUn, &prod, &st.damp, &st);
            FD(pn, F_ZLY) = zLn; FD(pn, F_ZUY) = zUn;
            r1 += zUn - zLn;
        }
        if (OBS && k >= 1) {
#pragma unroll 1
            for (int o = 0; o < O; ++o, po += (size_t)3 * S, pdo += (size_t)2 * S, pno += (size_t)3 * S) {
                const double so = FD(po, 0), ds = FD(pdo, 0), vL = FD(po, 2);
                const double s = so + alpha * ds;
                const double ex = x0 - FD(sc, CEN_ROW(o, k, 0)), ey = x1 - FD(sc, CEN_ROW(o, k, 1));
                const double rr = sqrt(ex * ex + ey * ey), nx = ex / rr, ny = ey / rr;
                const double dm = (rr - FD(sc, RAD_ROW(o))) - s;
                st.theta += fabs(dm); st.pinf = maxabs_nan(st.pinf, dm);
                const double slo = so - c.dL, sln = s - c.dL;
                if (!(sln > 0)) valid = false;
                prod *= sln; st.damp += sln;
                // many obstacle rows: take the logarithm before the running product of slacks leaves the double range
                if (!(prod < 1e250 && prod > 1e-250)) { st.bar += log(prod); prod = 1.0; }
                const double yd = FD(po, 1) + ay * FD(pdo, 1);
                double z = vL + adu * (mu / slo - vL - vL / slo * ds);
                if (clamp) z = fmax(fmin(z, K_KAPPA_SIGMA * mu / sln), mu / (K_KAPPA_SIGMA * sln));
                FD(pno, 0) = s; FD(pno, 1) = yd; FD(pno, 2) = z;
                r0 += nx * yd; r1 += ny * yd;
                st.dinf = maxabs_nan(st.dinf, -yd - z);
                const double p = sln * z;
                st.mn = fmin(st.mn, p); st.mx = fmax(st.mx, p); st.sumz += fabs(z); st.sumy += fabs(yd);
            }
        }
        if (k < N) {
            const double v = a[F_V] + alpha * d[D_U0], om = a[F_OM] + alpha * d[D_U1];
            FD(pn, F_V) = v; FD(pn, F_OM) = om;
            double sn, cs;
            sincos_(x2, &sn, &cs);
            FD(pn, F_CS) = cs; FD(pn, F_SN) = sn;
            st.wmax = fmax(st.wmax, fmax(fabs(v), fabs(om)));
            // multiplier of dynamics row k+1 at the updated point
            const double yn0 = an[F_Y0] + ay * dn[D_Y0], yn1 = an[F_Y1] + ay * dn[D_Y1], yn2 = an[F_Y2] + ay * dn[D_Y2];
            const double a13 = -T * v * sn, a23 = T * v * cs;
            r0 -= yn0; r1 -= yn1; r2 -= a13 * yn0 + a23 * yn1 + yn2;
            double gv, hv;
            vcost(c, df, v, &gv, &hv);
            double rv = gv - (T * cs * yn0 + T * sn * yn1), rw = df * 2.0 * c.Ww * om - T * yn2;
            if (c.cost_mode == 0) { const double vm = fmin(v, 0.0), vp = fmax(v, 0.0); st.f += c.Wvn * vm * vm + c.Wvp * vp * vp; }
            else st.f += c.Wvn * fmin(v, 0.0);
            st.f += c.Ww * om * om;
            double zLn, zUn;
            valid &= bound_trial(a[F_V], d[D_U0], v, c.lb[2], c.ub[2], c.hasL[2], c.hasU[2], a[F_ZLV], a[F_ZUV], mu, adu, clamp,
                                 &zLn, &zUn, &prod, &st.damp, &st);
            FD(pn, F_ZLV) = zLn; FD(pn, F_ZUV) = zUn;
            rv += zUn - zLn;
            valid &= bound_trial(a[F_OM], d[D_U1], om, c.lb[3], c.ub[3], c.hasL[3], c.hasU[3], a[F_ZLW], a[F_ZUW], mu, adu, clamp,
                                 &zLn, &zUn, &prod, &st.damp, &st);
            FD(pn, F_ZLW) = zLn; FD(pn, F_ZUW) = zUn;
            rw += zUn - zLn;
            st.dinf = maxabs_nan(maxabs_nan(st.dinf, rv), rw);
            xp0 = x0 + T * v * cs; xp1 = x1 + T * v * sn; xp2 = x2 + T * om;
            yk0 = yn0; yk1 = yn1; yk2 = yn2;
        } else {
            FD(pn, F_V) = 0.0; FD(pn, F_OM) = 0.0; FD(pn, F_CS) = 1.0; FD(pn, F_SN) = 0.0;
            FD(pn, F_ZLV) = 0.0; FD(pn, F_ZUV) = 0.0; FD(pn, F_ZLW) = 0.0; FD(pn, F_ZUW) = 0.0;
        }
        st.dinf = maxabs_nan(maxabs_nan(maxabs_nan(st.dinf, r0), r1), r2);
        st.bar += log(prod);
        rec_copy(a, an); rec_copy(d, dn);
        ps += (size_t)NSTATE * S; pd += (size_t)NSTEP * S; pn += (size_t)NSTATE * S;
    }
    st.f *= df;
    if (c.nb == 0) st.mn = 0.0;
    *out = st;
    const double phi = st.f - mu * st.bar + K_KAPPA_D * mu * st.damp;
    return valid && isfinite(phi) && isfinite(st.theta);
}

// SOC right-hand side: c_soc <- a * base + c(trial), base = c(current) for the first correction, else the previous c_soc
KMPC_HDN inline void pass_soc_rhs(const Cfg &c, const Ctx &t, double *wsp, size_t S, double al, bool first) {
    const int N = c.N, O = c.O;
    const Rows &L = c.L;
    const double T = c.T;
    const double *sc = wsp + (size_t)L.rSc * S;
    const double *ps = wsp + (size_t)L.rState[t.cur] * S, *pt = wsp + (size_t)L.rState[t.cur ^ 1] * S;
    const double *po = wsp + ((size_t)L.rState[t.cur] + L.sObs) * S, *pto = wsp + ((size_t)L.rState[t.cur ^ 1] + L.sObs) * S;
    double *pcs = wsp + (size_t)L.rCsoc * S, *pds = wsp + (size_t)L.rDsoc * S;
    double cp0 = FD(sc, 0), cp1 = FD(sc, 1), cp2 = FD(sc, 2);  // predicted (current point)
    double tp0 = cp0, tp1 = cp1, tp2 = cp2;                    // predicted (trial point)
#pragma unroll 1
    for (int k = 0; k <= N; ++k) {
        const double x0 = FD(ps, F_X0), x1 = FD(ps, F_X1), x2 = FD(ps, F_X2);
        const double t0 = FD(pt, F_X0), t1 = FD(pt, F_X1), t2 = FD(pt, F_X2);
        double b0, b1, b2;
        if (first) { b0 = x0 - cp0; b1 = x1 - cp1; b2 = x2 - cp2; }
        else { b0 = FD(pcs, 0); b1 = FD(pcs, 1); b2 = FD(pcs, 2); }
        FD(pcs, 0) = al * b0 + (t0 - tp0); FD(pcs, 1) = al * b1 + (t1 - tp1); FD(pcs, 2) = al * b2 + (t2 - tp2);
        if (O > 0 && k >= 1)
            for (int o = 0; o < O; ++o, po += (size_t)3 * S, pto += (size_t)3 * S, pds += S) {
                const double cx = FD(sc, CEN_ROW(o, k, 0)), cy = FD(sc, CEN_ROW(o, k, 1));
                double base;
                if (first) { const double ex = x0 - cx, ey = x1 - cy; base = (sqrt(ex * ex + ey * ey) - FD(sc, RAD_ROW(o))) - FD(po, 0); }
                else base = FD(pds, 0);
                const double ex = t0 - cx, ey = t1 - cy;
                FD(pds, 0) = al * base + ((sqrt(ex * ex + ey * ey) - FD(sc, RAD_ROW(o))) - FD(pto, 0));
            }
        if (k < N) {
            cp0 = x0 + T * FD(ps, F_V) * FD(ps, F_CS); cp1 = x1 + T * FD(ps, F_V) * FD(ps, F_SN); cp2 = x2 + T * FD(ps, F_OM);
            tp0 = t0 + T * FD(pt, F_V) * FD(pt, F_CS); tp1 = t1 + T * FD(pt, F_V) * FD(pt, F_SN); tp2 = t2 + T * FD(pt, F_OM);
        }
        ps += (size_t)NSTATE * S; pt += (size_t)NSTATE * S; pcs += (size_t)3 * S;
    }
}

// OUTPUT pass: returned matrices (optimizer.py:392-400) + objective / status / iteration count
KMPC_HDN inline void pass_output(const Cfg &c, const Ctx &t, double *wsp, size_t S, const IO &io, int status) {
    const int N = c.N, b = t.inst;
    const Rows &L = c.L;
    const double *ps = wsp + (size_t)L.rState[t.cur] * S;
#pragma unroll 1
    for (int k = 0; k <= N; ++k, ps += (size_t)NSTATE * S) {
        io.X_out[io_X(c, b, 0, k)] = FD(ps, F_X0); io.X_out[io_X(c, b, 1, k)] = FD(ps, F_X1); io.X_out[io_X(c, b, 2, k)] = FD(ps, F_X2);
        if (k < N) { io.U_out[io_U(c, b, 0, k)] = FD(ps, F_V); io.U_out[io_U(c, b, 1, k)] = FD(ps, F_OM); }
    }
    if (io.obj) io.obj[b] = t.c.f / t.df;
    if (io.status) io.status[b] = status;
    if (io.iters) io.iters[b] = t.iter;
}

// ---- scalar logic -------------------------------------------------------------------------------
KMPC_HD double cfg_mu_floor(double tol) { return kfmin(tol, K_COMPL_INF_TOL) / (K_KAPPA_EPS + 1.0); }   // (an IEEE division wherever it is evaluated)
KMPC_HD double compl_inf(const Cfg &c, const Stats &s, double mu) {
    return c.nb ? kfmax(fabs(s.mx - mu), fabs(s.mn - mu)) : 0.0;
}
KMPC_HD double opt_error(const Cfg &c, const Stats &s, double mu) {
    // s_d = max(s_max, (|y|_1 + |z|_1) / (m + n_b)) / s_max, s_c likewise; both are >= 1 and almost always exactly 1
    // (tried: the two rare divisions behind a branch in an out-of-line function -- 0.7 % slower than these selects)
    const double sd = kfmax(K_S_MAX, (s.sumy + s.sumz) * c.r_mnb) * (1.0 / K_S_MAX);
    const double sc = c.nb ? kfmax(K_S_MAX, s.sumz * c.r_nb) * (1.0 / K_S_MAX) : 1.0;
    const double di = sd > 1.0 ? s.dinf / sd : s.dinf, ci = compl_inf(c, s, mu);
    return kfmax(di, kfmax(s.pinf, sc > 1.0 ? ci / sc : ci));
}
KMPC_HD double phi_of(const Stats &s, double mu) { return s.f - mu * s.bar + K_KAPPA_D * mu * s.damp; }

// the filter: entry i = (theta, phi).  Two stores: FiltStrided -- entry i at filt[(2 i) * FS], filt[(2 i + 1) * FS] (a band of the
// thread solver's workspace); FiltSplit -- the first KMPC_FILTER_NEAR entries in a small fast array (the warp solver's shared
// memory: nearly every instance keeps fewer), the others at the same index of a far array (its global scratch slot).
#ifndef KMPC_FILTER_NEAR
#define KMPC_FILTER_NEAR 6
#endif
struct FiltStrided { double *p; size_t FS; KMPC_HD double &at(int i, int k) const { return p[(size_t)(2 * i + k) * FS]; } };
struct FiltSplit { double *nearp; double *farp; KMPC_HD double &at(int i, int k) const { return i < KMPC_FILTER_NEAR ? nearp[2 * i + k] : farp[2 * i + k]; } };
template <class F>
KMPC_HD bool filter_ok(const Ctx &t, const F &filt, double theta, double phi) {
    for (int i = 0; i < t.fn; ++i)
        if (!(theta <= filt.at(i, 0) || phi <= filt.at(i, 1))) return false;
    return true;
}
// false: the filter is full (the entry is NOT recorded; the caller ends the instance with ST_INTERNAL rather than go on with a
// filter that has forgotten an entry)
template <class F>
KMPC_HD bool filter_add(Ctx &t, const F &filt, double theta, double phi) {
    int m = 0;
    for (int i = 0; i < t.fn; ++i) {
        const double th = filt.at(i, 0), ph = filt.at(i, 1);
        if (!(th >= theta && ph >= phi)) { filt.at(m, 0) = th; filt.at(m, 1) = ph; ++m; }
    }
    t.fn = m;
    if (t.fn >= K_FILTER_CAP) return false;
    filt.at(t.fn, 0) = theta; filt.at(t.fn, 1) = phi; t.fn++;
    return true;
}
KMPC_HD bool filter_add(Ctx &t, double *filt, size_t FS, double theta, double phi) { return filter_add(t, FiltStrided{filt, FS}, theta, phi); }

// FilterLSAcceptor::CheckAcceptabilityOfTrialPoint
// switching condition  gBD < 0  and  a (-gBD)^s_phi > delta theta^s_theta  (theta, gBD of the current iterate).
// On the device the comparison is screened in single precision in the log2 domain (three MUFU.LG2); only a near-tie
// (|log2 ratio| < 2^-6, far above the float error of ~1e-5) or an out-of-range operand takes the double-precision pow path.
KMPC_HD bool is_ftype(const Ctx &t, double a) {
    if (!(t.gBD < 0)) return false;
#ifdef __CUDA_ARCH__
    const float L = __log2f((float)a) + (float)K_S_PHI * __log2f((float)(-t.gBD)) - (float)K_S_THETA * __log2f((float)t.c.theta)
                    - __log2f((float)K_DELTA_LS);
    if (fabsf(L) > 0.015625f && fabsf(L) < 1e30f) return L > 0.0f;
#endif
    return a * pow(-t.gBD, K_S_PHI) > K_DELTA_LS * pow(t.c.theta, K_S_THETA);
}
// smallest step size of the back-tracking line search before IPOPT would switch to the restoration phase
// (FilterLSAcceptor::CalculateAlphaMin); evaluated only when a trial point has been rejected
KMPC_HD double alpha_min_of(const Ctx &t) {
    double amin = K_GAMMA_THETA;
    if (t.gBD < 0) {
        amin = fmin(K_GAMMA_THETA, K_GAMMA_PHI * t.c.theta / (-t.gBD));
        if (t.c.theta <= t.theta_min) amin = fmin(amin, K_DELTA_LS * pow(t.c.theta, K_S_THETA) / pow(-t.gBD, K_S_PHI));
    }
    return amin * K_ALPHA_MIN_FRAC;
}
KMPC_HD bool armijo(const Ctx &t, double a, double tphi, double cphi) { return cmp_le(tphi - cphi, K_ETA_PHI * a * t.gBD, cphi); }
// ftype: 0 / 1 = the switching condition at alpha_test as evaluated here (the caller needs it again for the filter augmentation of an
// accepted point), -1 = not evaluated
template <class F>
KMPC_HD bool acceptable(const Ctx &t, const F &filt, const Stats &tri, int *ftype) {
    const double cphi = phi_of(t.c, t.mu), tphi = phi_of(tri, t.mu), cth = t.c.theta;
    bool acc;
    *ftype = -1;
    if (tri.theta > t.theta_max) return false;
    const bool ft = is_ftype(t, t.alpha_test);
    *ftype = ft ? 1 : 0;
    if (ft && cth <= t.theta_min) acc = armijo(t, t.alpha_test, tphi, cphi);
    else {
        acc = true;
        if (tphi > cphi) {
            // obj_max_inc test: log10(tphi - cphi) > obj_max_inc + max(1, log10 |cphi|), i.e. an increase by more than a factor 10^5 of
            // max(10, |cphi|).  Screened without the logarithms: 10 % on either side of that threshold the outcome cannot depend on
            // their rounding (a few ulp), and only inside the band are they evaluated -- same decisions, two log10() less per trial.
            const double d = tphi - cphi, sc = fabs(cphi) > 10.0 ? fabs(cphi) : 10.0;
            if (d >= 1.1e5 * sc) acc = false;
            else if (d > 0.9e5 * sc) {
                const double bas = fabs(cphi) > 10.0 ? log10(fabs(cphi)) : 1.0;
                if (log10(d) > K_OBJ_MAX_INC + bas) acc = false;
            }
        }
        if (acc) acc = cmp_le(tri.theta, (1.0 - K_GAMMA_THETA) * cth, cth) || cmp_le(tphi - cphi, -K_GAMMA_PHI * cth, cphi);
    }
    if (acc) acc = filter_ok(t, filt, tri.theta, tphi);
    return acc;
}

// Top of IPOPT's main loop at a (new) current iterate: termination tests, monotone barrier update.
// Returns a status < 100 to finish the instance, 100 to continue with a Newton step.
// opt_error(c, s, mu) with the parts that do not depend on mu (the scaled dual infeasibility, the scale of the complementarity) taken
// from the caller: begin_iteration evaluates the error for mu = 0 and then for one or more barrier parameters on the same point
struct OptErrParts { double di, sc; };
KMPC_HD OptErrParts opt_error_parts(const Cfg &c, const Stats &s) {
    const double sd = kfmax(K_S_MAX, (s.sumy + s.sumz) * c.r_mnb) * (1.0 / K_S_MAX);
    OptErrParts p;
    p.sc = c.nb ? kfmax(K_S_MAX, s.sumz * c.r_nb) * (1.0 / K_S_MAX) : 1.0;
    p.di = s.dinf;
    if (__builtin_expect(sd > 1.0, 0)) p.di = s.dinf / sd;   // (s_d is 1 unless the multipliers are huge: keep the division off the common path)
    return p;
}
KMPC_HD double opt_error_with(const Cfg &c, const Stats &s, double mu, const OptErrParts &p) {
    double ci = compl_inf(c, s, mu);
    if (__builtin_expect(p.sc > 1.0, 0)) ci = ci / p.sc;
    return kfmax(p.di, kfmax(s.pinf, ci));
}
KMPC_HD int begin_iteration(const Cfg &c, Ctx &t) {
    const OptErrParts ep = opt_error_parts(c, t.c);
    const double E0 = opt_error_with(c, t.c, 0.0, ep);
    // (IPOPT checks every evaluated quantity for non-finite numbers; the max-norms carry a NaN / inf through, the scaled max may lose it)
    if (!isfinite(E0) || !isfinite(t.c.pinf) || !isfinite(t.c.dinf)) return ST_INVALID;
    if (E0 <= c.tol && t.c.dinf / t.df <= K_DUAL_INF_TOL && t.c.pinf <= K_CONSTR_VIOL_TOL &&
        compl_inf(c, t.c, 0.0) / t.df <= K_COMPL_INF_TOL)
        return ST_SUCCESS;
    if (t.iter >= c.max_iter) return ST_MAXITER;
    if (t.c.wmax > K_DIVERGING) return ST_DIVERGING;
    bool done = false;
    while (!done && opt_error_with(c, t.c, t.mu, ep) <= K_KAPPA_EPS * t.mu) {
        const double nm = kfmax(kfmin(K_MU_LIN * t.mu, t.mu * sqrt(t.mu)), c.mu_floor);  // mu^1.5 (mu_superlinear_decrease_power); the floor is formed once, on the host (cfg_mu_floor)
        const bool changed = nm != t.mu;
        t.mu = nm; t.tau = kfmax(K_TAU_MIN, 1.0 - t.mu);
        if (changed) t.fn = 0; else done = true;
    }
    t.delta = 0.0;
    t.mode = M_NEWTON;
    return 100;
}

// ------------------------------------------------------------------------------------------------
// The state machine of one instance, cut into the three phases the kernels run (kmpc.cu):
//   phase_sweep   -> 100: factorisation ok, go on to the roll-out; 101: wrong inertia, delta raised, sweep again;
//                    otherwise a final status
//   phase_rollout -> search direction, step sizes, line-search set-up
//   phase_trial   -> trial point + acceptance logic; 100: continue (t.mode tells which phase is next), else final status
// ------------------------------------------------------------------------------------------------
// inertia correction (IPOPT PDPerturbationHandler): raise delta_w; 101 = factorise again, else a final status
// the perturbation IPOPT tries after `delta` has failed (delta_last = the last one that worked, 0 if none yet)
KMPC_HD double inertia_next_delta(double delta, double delta_last) {
    if (delta == 0.0) return delta_last == 0.0 ? K_DW_INIT : fmax(K_DW_MIN, delta_last * K_DW_DEC);
    return (delta_last == 0.0 || 1e5 * delta_last < delta) ? K_DW_INC_FIRST * delta : K_DW_INC * delta;
}
KMPC_HD int inertia_update(Ctx &t) {
    t.delta = inertia_next_delta(t.delta, t.delta_last);
    if (t.delta > K_DW_MAX) return ST_STEP_ERROR;
    return 101;
}

// status of a factorisation that failed outside a Newton step.  The least-squares system of the multiplier estimate ([I J^T; J 0],
// J of full row rank) always has the right inertia for finite data, so a failure there means non-finite numbers in the starting
// point: IPOPT keeps y = 0 and stops at its first convergence check with Invalid_Number_Detected -- the same status, one trip earlier.
KMPC_HD int sweep_failure_status(const Ctx &t) { return t.mode == M_LSQ ? (int)ST_INVALID : (int)ST_STEP_ERROR; }

template <bool OBS>
KMPC_HDN inline int phase_sweep(const Cfg &c, Ctx &t, double *wsp, size_t S) {
    t.trips++;
    const bool ok = pass_sweep<OBS>(c, t, wsp, S);
    if (ok) return 100;
    if (t.mode != M_NEWTON) return sweep_failure_status(t);
    return inertia_update(t);
}

// line-search set-up once the search direction and its step-size limits are known
KMPC_HD void rollout_logic(Ctx &t, double apr, double adu, double gbd, double ym) {
    t.sel = t.mode == M_SOC ? 1 : 0;
    t.tu = TU_STEP;
    if (t.mode == M_LSQ) {
        t.tu = TU_INIT; t.a_pr = 0.0; t.a_du = 0.0;
        t.a_y = (ym <= K_YINIT_MAX && isfinite(ym)) ? -1.0 : 0.0;
    } else if (t.mode == M_NEWTON) {
        if (t.delta > 0.0) t.delta_last = t.delta;
        t.gBD = gbd; t.pw_g = 0.0;   // (new search direction: alpha_min of the line search is not known yet)
        if (t.theta_max < 0) { t.theta_max = K_THETA_MAX_FACT * fmax(1.0, t.c.theta); t.theta_min = K_THETA_MIN_FACT * fmax(1.0, t.c.theta); }
        t.alpha = apr; t.alpha_test = apr; t.alpha_du0 = adu; t.nsteps = 0; t.soc_count = 0;
        t.a_pr = apr; t.a_y = apr; t.a_du = adu;
    } else {  // M_SOC
        t.alpha_soc = apr;
        t.a_pr = apr; t.a_y = apr; t.a_du = adu;
    }
}

template <bool OBS>
KMPC_HDN inline void phase_rollout(const Cfg &c, Ctx &t, double *wsp, size_t S) {
    double apr, adu, gbd, ym;
    pass_rollout<OBS>(c, t, wsp, S, t.mode == M_SOC ? 1 : 0, &apr, &adu, &gbd, &ym);
    rollout_logic(t, apr, adu, gbd, ym);
}

// back-tracking trial on the original step: no sweep / roll-out this trip
KMPC_HD void trial_setup(Ctx &t) {
    if (t.mode == M_TRIAL) {
        t.trips++;
        t.sel = 0; t.tu = TU_STEP; t.a_pr = t.alpha; t.a_y = t.alpha; t.a_du = t.alpha_du0;
        t.alpha_test = t.alpha;
    }
}

// Acceptance logic of one evaluated trial point.  Returns
//   R_SOC1 / R_SOC2  a (first / follow-up) second-order correction is due: caller builds c_soc with step t.alpha_soc
//   R_BACKTRACK      next trip evaluates a shorter step;   R_ACCEPT  caller makes the trial point current, then begin_iteration
//   a final status (ST_RESTORATION)
enum { R_CONTINUE = 100, R_RETRY = 101, R_ACCEPT = 102, R_SOC1 = 103, R_SOC2 = 104, R_BACKTRACK = 105 };
template <class F>
KMPC_HD int trial_decide(Ctx &t, const F &filt, const Stats &tri, bool evok, bool *augment, double *aug_theta, double *aug_phi) {
    *augment = false; *aug_theta = 0.0; *aug_phi = 0.0;
    if (t.tu == TU_INIT) return R_ACCEPT;
    bool accept = false;
    int soc_rhs = 0;
    int ftype = -1;
    if (evok) accept = acceptable(t, filt, tri, &ftype);
    if (!accept && evok) {
        if (t.mode == M_SOC) {
            t.soc_count++; t.theta_trial = tri.theta;
            if (t.soc_count < K_MAX_SOC && t.theta_trial <= K_KAPPA_SOC * t.theta_soc_old) soc_rhs = 2;
        } else if (t.nsteps == 0 && t.c.theta <= tri.theta) {
            // second-order correction from the first trial point (max_soc 4)
            t.alpha_soc = t.alpha; t.soc_count = 0;
            soc_rhs = 1;
        }
    }
    if (soc_rhs) {
        t.theta_soc_old = tri.theta; t.theta_trial = tri.theta;
        t.mode = M_SOC;
        return soc_rhs == 1 ? R_SOC1 : R_SOC2;
    }
    if (!accept) {
        // back-track on the original step (also after a failed correction)
        t.alpha *= K_ALPHA_RED; t.nsteps++;
        // (alpha_min depends on gBD, theta and theta_min of the current iterate only -- two pow() and two divisions -- and an instance with a
        //  difficult line search rejects hundreds of trial points: computed at the first rejection of a line search, then reused)
        double amin = t.pw_g;
        if (!(amin > 0.0)) { amin = alpha_min_of(t); t.pw_g = amin; }
        if (!(t.alpha > amin)) return ST_RESTORATION;  // IPOPT would enter the restoration phase here
        t.mode = M_TRIAL;
        return R_BACKTRACK;
    }
    // accepted: filter augmentation (FilterLSAcceptor::UpdateForNextIteration)
    const double cphi = phi_of(t.c, t.mu), tphi = phi_of(tri, t.mu);
    if (!(ftype < 0 ? is_ftype(t, t.alpha_test) : ftype != 0) || !armijo(t, t.alpha_test, tphi, cphi)) {
        *augment = true; *aug_theta = (1.0 - K_GAMMA_THETA) * t.c.theta; *aug_phi = cphi - K_GAMMA_PHI * t.c.theta;
    }
    t.iter++;
    return R_ACCEPT;
}

}  // namespace kmpc
#include "kmpc_resto.cuh"   // the feasibility restoration phase (uses everything above)
namespace kmpc {

template <bool OBS>
KMPC_HDN inline int phase_trial(const Cfg &c, Ctx &t, double *wsp, size_t S) {
    trial_setup(t);
    Stats tri;
    const bool evok = pass_trial<OBS>(c, t, wsp, S, t.sel, t.tu, t.a_pr, t.a_y, t.a_du, &tri);
    bool aug; double ath, aph;
    double *filt = wsp + (size_t)c.L.rFilt * S;
    const int r = trial_decide(t, FiltStrided{filt, S}, tri, evok, &aug, &ath, &aph);
    if (aug && !filter_add(t, filt, S, ath, aph)) return ST_INTERNAL;
    if (r == R_SOC1 || r == R_SOC2) { pass_soc_rhs(c, t, wsp, S, t.alpha_soc, r == R_SOC1); return 100; }
    if (r == R_BACKTRACK) return 100;
    // the step size fell below alpha_min: IPOPT's restoration phase (100: it handed a point back, the next trip is a Newton step there)
    if (r == ST_RESTORATION) return resto_enter<OBS>(c, t, make_resto_rows(c.L), wsp, S);
    if (r != R_ACCEPT) return r;
    // the trial buffer becomes the current iterate
    t.c = tri; t.cur ^= 1;
    return begin_iteration(c, t);
}

// One whole trip for one instance (used by the sequential test harness): sweep -> roll-out -> trial.
template <bool OBS>
KMPC_HDN inline int trip(const Cfg &c, Ctx &t, double *wsp, size_t S) {
    if (t.mode != M_TRIAL) {
        const int r = phase_sweep<OBS>(c, t, wsp, S);
        if (r == 101) return 100;
        if (r != 100) return r;
        phase_rollout<OBS>(c, t, wsp, S);
    }
    return phase_trial<OBS>(c, t, wsp, S);
}

// (finish_instance is defined at the end of the file: it needs ctx_load)

// ---- solver context <-> workspace (the kernels keep no state between launches) -------------------
enum { X_MODE = 0, X_ITER, X_CUR, X_NSTEPS, X_SOCC, X_FN, X_TRIPS, X_SEL, X_TU, X_MU, X_TAU, X_DELTA, X_DLAST, X_DF, X_THMAX,
       X_THMIN, X_ALPHA, X_ATEST, X_AMIN, X_ADU0, X_ASOC, X_GBD, X_THSOC, X_THTRI, X_APR, X_AY, X_ADU, X_CF, X_CBAR, X_CDAMP,
       X_CTHETA, X_CDINF, X_CPINF, X_CMN, X_CMX, X_CSUMY, X_CSUMZ, X_CWMAX, X_PWG, X_PWT, X_COUNT };
static_assert(X_COUNT <= KMPC_NCTX, "context rows");

KMPC_HD void ctx_store(const Ctx &t, const Rows &L, double *wsp, size_t S) {
    double *p = wsp + (size_t)L.rCtx * S;
    FD(p, X_MODE) = t.mode; FD(p, X_ITER) = t.iter; FD(p, X_CUR) = t.cur; FD(p, X_NSTEPS) = t.nsteps; FD(p, X_SOCC) = t.soc_count;
    FD(p, X_FN) = t.fn; FD(p, X_TRIPS) = t.trips; FD(p, X_SEL) = t.sel; FD(p, X_TU) = t.tu;
    FD(p, X_MU) = t.mu; FD(p, X_TAU) = t.tau; FD(p, X_DELTA) = t.delta; FD(p, X_DLAST) = t.delta_last; FD(p, X_DF) = t.df;
    FD(p, X_THMAX) = t.theta_max; FD(p, X_THMIN) = t.theta_min; FD(p, X_ALPHA) = t.alpha; FD(p, X_ATEST) = t.alpha_test;
    FD(p, X_AMIN) = t.alpha_min; FD(p, X_ADU0) = t.alpha_du0; FD(p, X_ASOC) = t.alpha_soc; FD(p, X_GBD) = t.gBD;
    FD(p, X_THSOC) = t.theta_soc_old; FD(p, X_THTRI) = t.theta_trial; FD(p, X_APR) = t.a_pr; FD(p, X_AY) = t.a_y; FD(p, X_ADU) = t.a_du;
    FD(p, X_CF) = t.c.f; FD(p, X_CBAR) = t.c.bar; FD(p, X_CDAMP) = t.c.damp; FD(p, X_CTHETA) = t.c.theta; FD(p, X_CDINF) = t.c.dinf;
    FD(p, X_CPINF) = t.c.pinf; FD(p, X_CMN) = t.c.mn; FD(p, X_CMX) = t.c.mx; FD(p, X_CSUMY) = t.c.sumy; FD(p, X_CSUMZ) = t.c.sumz;
    FD(p, X_CWMAX) = t.c.wmax; FD(p, X_PWG) = t.pw_g; FD(p, X_PWT) = t.pw_t;
}
KMPC_HD void ctx_load(Ctx &t, const Rows &L, const double *wsp, size_t S) {
    const double *p = wsp + (size_t)L.rCtx * S;
    t.mode = (int)FD(p, X_MODE); t.iter = (int)FD(p, X_ITER); t.cur = (int)FD(p, X_CUR); t.nsteps = (int)FD(p, X_NSTEPS);
    t.soc_count = (int)FD(p, X_SOCC); t.fn = (int)FD(p, X_FN); t.trips = (int)FD(p, X_TRIPS); t.sel = (int)FD(p, X_SEL); t.tu = (int)FD(p, X_TU);
    t.mu = FD(p, X_MU); t.tau = FD(p, X_TAU); t.delta = FD(p, X_DELTA); t.delta_last = FD(p, X_DLAST); t.df = FD(p, X_DF);
    t.theta_max = FD(p, X_THMAX); t.theta_min = FD(p, X_THMIN); t.alpha = FD(p, X_ALPHA); t.alpha_test = FD(p, X_ATEST);
    t.alpha_min = FD(p, X_AMIN); t.alpha_du0 = FD(p, X_ADU0); t.alpha_soc = FD(p, X_ASOC); t.gBD = FD(p, X_GBD);
    t.theta_soc_old = FD(p, X_THSOC); t.theta_trial = FD(p, X_THTRI); t.a_pr = FD(p, X_APR); t.a_y = FD(p, X_AY); t.a_du = FD(p, X_ADU);
    t.c.f = FD(p, X_CF); t.c.bar = FD(p, X_CBAR); t.c.damp = FD(p, X_CDAMP); t.c.theta = FD(p, X_CTHETA); t.c.dinf = FD(p, X_CDINF);
    t.c.pinf = FD(p, X_CPINF); t.c.mn = FD(p, X_CMN); t.c.mx = FD(p, X_CMX); t.c.sumy = FD(p, X_CSUMY); t.c.sumz = FD(p, X_CSUMZ);
    t.c.wmax = FD(p, X_CWMAX); t.pw_g = FD(p, X_PWG); t.pw_t = FD(p, X_PWT);
}

// Finisher of the warp solver (kmpc_finish_kernel): column i of the hand-over workspace holds an instance whose regular line search
// failed (its iterate, multipliers, filter and solver context as the warp solver left them, w_hand_over in kmpc_warp.cuh).  One thread
// runs IPOPT's restoration phase on it and, if that hands a point back, the rest of the regular algorithm (the thread solver's trips),
// then writes the instance's outputs.
template <bool OBS>
KMPC_HDN inline void finish_instance(const Cfg &c, const IO &io, int i, double *wsp /* the instance's column (HBM, or a copy of it in shared memory) */) {
    Ctx t;
    ctx_load(t, c.L, wsp, 1);
    t.inst = io.resto_list[i];
    int r = resto_enter<OBS>(c, t, make_resto_rows(c.L), wsp, 1);
    while (r == 100) r = trip<OBS>(c, t, wsp, 1);
    pass_output(c, t, wsp, 1, io, r);
}

}  // namespace kmpc
