/*
 * kmpc_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, NOT PRODUCT CODE)
 *
 * A plain-C restatement of the numerical path the reference runs per MPC step:
 *   mpc/optimizer.py:319-400  MotionPlanner.solve  ->  ca.nlpsol("solver","ipopt",...)  (optimizer.py:354, :375-391)
 * i.e. the unicycle-MPC NLP (optimizer.py:57-60 weights, :79-110 cost, :111-156 bounds, :158-196 dynamics,
 * :198-258 obstacle rows in the *intended* README.md:78-81 form) solved by IPOPT's filter line-search
 * primal-dual interior-point method with the options the reference sets (optimizer.py:344-352) and IPOPT
 * 3.14 defaults for everything else.
 *
 * The arithmetic of the reference lives in a third-party dependency that is NOT in /root/reference and NOT
 * installable here: casadi==3.7.1 (requirements.txt:1), which bundles IPOPT 3.14.x + MUMPS.  This file
 * restates IPOPT's published algorithm (Waechter & Biegler, Math. Program. 106(1), 2006, and the IPOPT 3.14
 * option documentation).  The reference holds no tests, golden vectors or fixtures for this path, and
 * CasADi cannot be run in this image:   **PARITY UNPINNED**  (see DESIGN.md).
 *
 * What is restated (IPOPT names in brackets):
 *   - bound relaxation [bound_relax_factor 1e-8], gradient-based objective scaling [nlp_scaling_max_gradient 100]
 *   - starting point push [bound_push/bound_frac 0.01], z=1 [bound_mult_init_val], least-squares equality
 *     multipliers, discarded above 1e3 [constr_mult_init_max], mu0 = 0.1 [mu_init]
 *   - primal-dual step from the augmented system with inertia correction
 *     [first_hessian_perturbation 1e-4, perturb_inc_fact_first 100, perturb_inc_fact 8, perturb_dec_fact 1/3]
 *   - fraction-to-the-boundary [tau_min 0.99], filter line search with switching/Armijo conditions,
 *     second-order correction [max_soc 4, kappa_soc 0.99], alpha_min rule, kappa_sigma dual reset 1e10
 *   - monotone barrier update [barrier_tol_factor 10, mu_linear_decrease_factor 0.2,
 *     mu_superlinear_decrease_power 1.5, mu_allow_fast_monotone_decrease yes]
 *   - termination on the scaled optimality error E_0 <= tol [tol 1e-8, s_max 100] plus the unscaled
 *     dual_inf_tol 1 / constr_viol_tol 1e-4 / compl_inf_tol 1e-4 tests; max_iter 2000.
 *   - the feasibility restoration phase [MinC_1NrmRestorationPhase: resto_penalty_parameter 1000, resto_proximity_weight 1,
 *     required_infeasibility_reduction 0.9, bound_mult_reset_threshold 1000, constr_mult_reset_threshold 0,
 *     resto.theta_max_fact 1e8]: entered when the line search runs below alpha_min; "almost feasible" entry -> Restoration_Failed
 *     (-2); converged to a stationary point of the infeasibility that is not feasible -> Infeasible_Problem_Detected (2).
 *     Restated from the published algorithm (Waechter & Biegler 2006, section 3.3) and IPOPT 3.14's documented behaviour; like the
 *     rest of this file it cannot be held against IPOPT itself here.
 * What is NOT restated (documented deviations):
 *   - watchdog, soft restoration phase, iterative refinement of the linear solve, tiny-step detection, slack safeguards at
 *     machine precision, the recursive restoration inside the restoration phase.
 *
 * Linear algebra: linsolve=0 assembles the full IPOPT augmented system (x, s, y_c, y_d) densely and factors it
 * with a Bunch-Kaufman LDL^T, reading the inertia off D exactly as IPOPT does with MUMPS.  linsolve=1 solves the
 * same system stage-wise (condensed slacks + Riccati recursion) -- used only for the timed CPU baseline and to
 * cross-check the dense path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this file.
 */
#include <alloca.h>
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define KMO_VERSION 200

/* ---- IPOPT 3.14 defaults (+ the four options optimizer.py:344-352 sets) ---- */
#define BOUND_RELAX 1e-8
#define NLP_LOWER_INF (-1e19)
#define NLP_UPPER_INF (1e19)
#define SCALING_MAX_GRAD 100.0
#define SCALING_MIN_VALUE 1e-8
#define BOUND_PUSH 0.01
#define BOUND_FRAC 0.01
#define CONSTR_MULT_INIT_MAX 1e3
#define MU_INIT 0.1
#define TAU_MIN 0.99
#define KAPPA_EPS 10.0
#define MU_LIN 0.2
#define MU_SUPER 1.5
#define S_MAX 100.0
#define KAPPA_SIGMA 1e10
#define KAPPA_D 1e-5
#define GAMMA_THETA 1e-5
#define GAMMA_PHI 1e-8
#define ETA_PHI 1e-8
#define S_THETA 1.1
#define S_PHI 2.3
#define DELTA_LS 1.0
#define THETA_MAX_FACT 1e4
#define THETA_MIN_FACT 1e-4
#define ALPHA_MIN_FRAC 0.05
#define ALPHA_RED 0.5
#define MAX_SOC 4
#define KAPPA_SOC 0.99
#define OBJ_MAX_INC 5.0
#define DELTA_W_INIT 1e-4
#define DELTA_W_MIN 1e-20
#define DELTA_W_MAX 1e20
#define DELTA_W_INC_FIRST 100.0
#define DELTA_W_INC 8.0
#define DELTA_W_DEC (1.0 / 3.0)
#define DUAL_INF_TOL 1.0
#define CONSTR_VIOL_TOL 1e-4
#define COMPL_INF_TOL 1e-4
#define DIVERGING_TOL 1e20
#define FILTER_CAP 512
#define RESTO_RHO 1000.0          /* resto_penalty_parameter */
#define RESTO_KAPPA 0.9           /* required_infeasibility_reduction */
#define RESTO_THETA_MAX_FACT 1e8  /* resto.theta_max_fact */
#define BOUND_MULT_RESET 1000.0   /* bound_mult_reset_threshold */

/* IPOPT ApplicationReturnStatus numbering */
#define ST_SUCCESS 0
#define ST_INFEASIBLE 2
#define ST_MAXITER (-1)
#define ST_RESTORATION (-2)
#define ST_STEP_ERROR (-3)
#define ST_DIVERGING 4
#define ST_INVALID_NUMBER (-13)

typedef struct {
    int32_t N;         /* horizon */
    int32_t O;         /* obstacles per instance */
    int32_t cost_mode; /* 0 = README squared penalties, 1 = code-literal 300*fmin(v,0) (optimizer.py:91-96) */
    int32_t goal_k_lo; /* goal cost over k = goal_k_lo .. goal_k_hi (README: 1..N, code: 1..N-1, optimizer.py:80) */
    int32_t goal_k_hi;
    int32_t max_iter;
    int32_t linsolve; /* 0 dense Bunch-Kaufman on the full augmented system, 1 stage-wise Riccati */
    int32_t obs_stagewise; /* 0: one centre per obstacle, obs[O][2] (optimizer.py:217-221); 1: a centre per obstacle and stage,
                            * obs[O][N][2], column t paired with X_{t+1} (dynamic_obstacle.py:47-56 over optimizer.py:201-215) */
    double T;
    double W[3];      /* optimizer.py:57 */
    double Wv_neg;    /* optimizer.py:59 */
    double Wv_pos;    /* README.md:24 */
    double Ww;        /* optimizer.py:60 */
    double lo[5];     /* x, y, theta, v, omega lower bounds (<= -1e19: none) */
    double hi[5];
    double obs_radius;
    double inflation; /* lower bound on the obstacle distance rows (optimizer.py:254-258) */
    double tol;
} kmo_config;

typedef struct {
    int32_t n_factor;    /* KKT factorisations (incl. inertia retries) */
    int32_t n_trials;    /* line-search trial evaluations */
    int32_t n_soc;       /* second-order-correction solves */
    int32_t max_filter;  /* largest filter size seen */
    double mu;           /* final barrier parameter */
    double err;          /* final scaled E_0 */
    double obj_scaling;  /* df */
    double max_delta_w;  /* largest Hessian perturbation used */
    int32_t n_resto;     /* restoration phases entered */
    int32_t reserved;
} kmo_diag;

/* ------------------------------------------------------------------ */
/* per-thread workspace                                                */
/* ------------------------------------------------------------------ */
typedef struct {
    int N, O, n, ns, mc, md, nk;
    /* problem data */
    double xcur[3], goal[3], df;
    double *obs; /* [O][2], or [O][N][2] when the centres move with the stage */
    double *orad; /* [O] radius subtracted in every row of obstacle o (optimizer.py:231-250: one radius per obstacle class) */
    int obs_sw;
    double *lb, *ub;
    int *hasL, *hasU;
    double dL; /* relaxed slack lower bound */
    /* iterate */
    double *w, *s, *yc, *yd, *zL, *zU, *vL;
    /* evaluation scratch */
    double *g, *c, *dms, *cs, *sn, *nrm /*[ns][2]*/, *dist;
    /* stage Hessian of the Lagrangian */
    double *Wxx /*[N+1][6]: xx xy yy xt yt tt*/, *Wtv, *Wvv, *Www;
    /* linear system */
    double *Dx, *Ds, *bx, *bs, *bc, *bd;
    double *dx, *ds, *dyc, *dyd, *dzL, *dzU, *dvL;
    double *dx2, *ds2, *dyc2, *dyd2, *csoc, *dsoc;
    double *wt, *st, *ct, *dmst;
    /* dense solver */
    double *K, *L;
    int *perm, *pivtype;
    double *rhs;
    int fact_valid;
    /* riccati solver */
    double *Kg /*[N][6]*/, *Quui /*[N][3]*/, *Pm /*[N+1][6]*/, *pv, *Sigc /*[ns]*/;
    int linsolve;
    /* restoration phase: -Dc / -Dd on the diagonal of the constraint blocks (eliminated n, p), proximity Hessian eta * DR^2 instead of the objective's */
    double *Dc, *Dd, *DR2;
    double eta;
    int resto;
} work_t;

static void *xcalloc(size_t n, size_t sz) { void *p = calloc(n ? n : 1, sz); if (!p) abort(); return p; }

static work_t *work_new(const kmo_config *cf) {
    work_t *w = (work_t *)xcalloc(1, sizeof(work_t));
    int N = cf->N, O = cf->O;
    w->N = N; w->O = O; w->n = 5 * N + 3; w->ns = N * O; w->mc = 3 * (N + 1); w->md = N * O;
    w->nk = w->n + w->ns + w->mc + w->md;
    w->linsolve = cf->linsolve;
    int n = w->n, ns = w->ns, mc = w->mc;
#define D(name, cnt) w->name = (double *)xcalloc((cnt), sizeof(double))
    w->obs_sw = cf->obs_stagewise != 0; D(obs, 2 * O * (w->obs_sw ? N : 1)); D(orad, O); D(lb, n); D(ub, n);
    w->hasL = (int *)xcalloc(n, sizeof(int)); w->hasU = (int *)xcalloc(n, sizeof(int));
    D(w, n); D(s, ns); D(yc, mc); D(yd, ns); D(zL, n); D(zU, n); D(vL, ns);
    D(g, n); D(c, mc); D(dms, ns); D(cs, N + 1); D(sn, N + 1); D(nrm, 2 * ns); D(dist, ns);
    D(Wxx, 6 * (N + 1)); D(Wtv, N + 1); D(Wvv, N + 1); D(Www, N + 1);
    D(Dx, n); D(Ds, ns); D(bx, n); D(bs, ns); D(bc, mc); D(bd, ns);
    D(dx, n); D(ds, ns); D(dyc, mc); D(dyd, ns); D(dzL, n); D(dzU, n); D(dvL, ns);
    D(dx2, n); D(ds2, ns); D(dyc2, mc); D(dyd2, ns); D(csoc, mc); D(dsoc, ns);
    D(wt, n); D(st, ns); D(ct, mc); D(dmst, ns);
    if (cf->linsolve == 0) {
        D(K, (size_t)w->nk * w->nk); D(L, (size_t)w->nk * w->nk); D(rhs, w->nk);
        w->perm = (int *)xcalloc(w->nk, sizeof(int)); w->pivtype = (int *)xcalloc(w->nk, sizeof(int));
    }
    D(Kg, 6 * N); D(Quui, 3 * N); D(Pm, 6 * (N + 1)); D(pv, 3 * (N + 1)); D(Sigc, ns);
    D(Dc, mc); D(Dd, ns); D(DR2, n);
#undef D
    return w;
}

static void work_free(work_t *w) {
    double **ptrs[] = {&w->obs, &w->orad, &w->lb, &w->ub, &w->w, &w->s, &w->yc, &w->yd, &w->zL, &w->zU, &w->vL, &w->g, &w->c,
        &w->dms, &w->cs, &w->sn, &w->nrm, &w->dist, &w->Wxx, &w->Wtv, &w->Wvv, &w->Www, &w->Dx, &w->Ds, &w->bx,
        &w->bs, &w->bc, &w->bd, &w->dx, &w->ds, &w->dyc, &w->dyd, &w->dzL, &w->dzU, &w->dvL, &w->dx2, &w->ds2,
        &w->dyc2, &w->dyd2, &w->csoc, &w->dsoc, &w->wt, &w->st, &w->ct, &w->dmst, &w->K, &w->L, &w->rhs, &w->Kg,
        &w->Quui, &w->Pm, &w->pv, &w->Sigc, &w->Dc, &w->Dd, &w->DR2};
    for (size_t i = 0; i < sizeof(ptrs) / sizeof(ptrs[0]); ++i) free(*ptrs[i]);
    free(w->hasL); free(w->hasU); free(w->perm); free(w->pivtype);
    free(w);
}

/* decision vector ordering (optimizer.py:74-77): vec(X) column-major then vec(U) column-major */
#define IX(k, j) (3 * (k) + (j))
#define IU(k, j) (3 * (N + 1) + 2 * (k) + (j))
/* obstacle rows (optimizer.py:229, (N x O) column-major): obstacle-major, k = 1..N */
#define IS(o, k) ((o) * N + ((k) - 1))

/* CasADi derivative convention for fmin/fmax at ties: a/(a+b) */
static double dmin0(double v) { return v < 0 ? 1.0 : (v == 0 ? 0.5 : 0.0); }
static double dmax0(double v) { return v > 0 ? 1.0 : (v == 0 ? 0.5 : 0.0); }

/* ---------------- problem functions (scaled objective) ---------------- */
/* cost: optimizer.py:79-110 (code-literal) / README.md:15-27 (README form) */
static double eval_f(const kmo_config *cf, const work_t *wk, const double *w) {
    int N = wk->N; double f = 0;
    for (int k = cf->goal_k_lo; k <= cf->goal_k_hi; ++k)
        for (int j = 0; j < 3; ++j) { double e = w[IX(k, j)] - wk->goal[j]; f += cf->W[j] * e * e; }
    for (int k = 0; k < N; ++k) {
        double v = w[IU(k, 0)], om = w[IU(k, 1)];
        if (cf->cost_mode == 0) { double vm = fmin(v, 0.0), vp = fmax(v, 0.0); f += cf->Wv_neg * vm * vm + cf->Wv_pos * vp * vp; }
        else f += cf->Wv_neg * fmin(v, 0.0);
        f += cf->Ww * om * om;
    }
    return wk->df * f;
}

static void eval_grad(const kmo_config *cf, const work_t *wk, const double *w, double df, double *g) {
    int N = wk->N;
    memset(g, 0, sizeof(double) * wk->n);
    for (int k = cf->goal_k_lo; k <= cf->goal_k_hi; ++k)
        for (int j = 0; j < 3; ++j) g[IX(k, j)] = df * 2.0 * cf->W[j] * (w[IX(k, j)] - wk->goal[j]);
    for (int k = 0; k < N; ++k) {
        double v = w[IU(k, 0)], om = w[IU(k, 1)];
        if (cf->cost_mode == 0) g[IU(k, 0)] = df * (2.0 * cf->Wv_neg * fmin(v, 0.0) * dmin0(v) + 2.0 * cf->Wv_pos * fmax(v, 0.0) * dmax0(v));
        else g[IU(k, 0)] = df * cf->Wv_neg * dmin0(v);
        g[IU(k, 1)] = df * 2.0 * cf->Ww * om;
    }
}

/* dynamics defects: optimizer.py:163-196 */
static void eval_c(const kmo_config *cf, const work_t *wk, const double *w, double *c) {
    int N = wk->N; double T = cf->T;
    for (int j = 0; j < 3; ++j) c[j] = w[IX(0, j)] - wk->xcur[j];
    for (int k = 0; k < N; ++k) {
        double th = w[IX(k, 2)], v = w[IU(k, 0)], om = w[IU(k, 1)];
        c[3 * (k + 1) + 0] = w[IX(k + 1, 0)] - (w[IX(k, 0)] + T * v * cos(th));
        c[3 * (k + 1) + 1] = w[IX(k + 1, 1)] - (w[IX(k, 1)] + T * v * sin(th));
        c[3 * (k + 1) + 2] = w[IX(k + 1, 2)] - (th + T * om);
    }
}

/* obstacle distances, intended form (README.md:78-81; optimizer.py:217-250 with the dangling-minus lines repaired) */
static void eval_d(const kmo_config *cf, const work_t *wk, const double *w, double *d, double *nrm, double *dist) {
    int N = wk->N, O = wk->O;
    for (int o = 0; o < O; ++o)
        for (int k = 1; k <= N; ++k) {
            const double *cc = wk->obs + (wk->obs_sw ? 2 * ((size_t)o * N + (k - 1)) : 2 * (size_t)o);
            double ex = w[IX(k, 0)] - cc[0], ey = w[IX(k, 1)] - cc[1];
            double r = sqrt(ex * ex + ey * ey);
            d[IS(o, k)] = r - wk->orad[o];
            if (nrm) { nrm[2 * IS(o, k)] = ex / r; nrm[2 * IS(o, k) + 1] = ey / r; dist[IS(o, k)] = r; }
        }
}

static void eval_trig(const work_t *wk, const double *w, double *cs, double *sn) {
    int N = wk->N;
    for (int k = 0; k < N; ++k) { cs[k] = cos(w[IX(k, 2)]); sn[k] = sin(w[IX(k, 2)]); }
}

/* out += Jc^T y  (x and u parts) */
static void add_JcT(const kmo_config *cf, const work_t *wk, const double *w, const double *cs, const double *sn,
                    const double *y, double *out) {
    int N = wk->N; double T = cf->T;
    for (int k = 0; k <= N; ++k) for (int j = 0; j < 3; ++j) out[IX(k, j)] += y[3 * k + j];
    for (int k = 0; k < N; ++k) {
        const double *yn = y + 3 * (k + 1); double v = w[IU(k, 0)];
        out[IX(k, 0)] -= yn[0]; out[IX(k, 1)] -= yn[1];
        out[IX(k, 2)] -= (-T * v * sn[k]) * yn[0] + (T * v * cs[k]) * yn[1] + yn[2];
        out[IU(k, 0)] -= T * cs[k] * yn[0] + T * sn[k] * yn[1];
        out[IU(k, 1)] -= T * yn[2];
    }
}

static void add_JdT(const work_t *wk, const double *nrm, const double *y, double *out) {
    int N = wk->N, O = wk->O;
    for (int o = 0; o < O; ++o) for (int k = 1; k <= N; ++k) {
        int i = IS(o, k); out[IX(k, 0)] += nrm[2 * i] * y[i]; out[IX(k, 1)] += nrm[2 * i + 1] * y[i];
    }
}

/* Hessian of the Lagrangian, stage blocks */
static void eval_hess(const kmo_config *cf, work_t *wk, const double *w, const double *yc, const double *yd,
                      const double *cs, const double *sn) {
    int N = wk->N, O = wk->O; double T = cf->T, df = wk->df;
    memset(wk->Wxx, 0, sizeof(double) * 6 * (N + 1));
    for (int k = 0; k <= N; ++k) {
        if (wk->resto) {   /* restoration objective: rho (n + p) + eta/2 |DR (x - x_ref)|^2 -> eta DR^2 on the diagonal */
            wk->Wxx[6 * k + 0] = wk->eta * wk->DR2[IX(k, 0)]; wk->Wxx[6 * k + 2] = wk->eta * wk->DR2[IX(k, 1)]; wk->Wxx[6 * k + 5] = wk->eta * wk->DR2[IX(k, 2)];
        } else if (k >= cf->goal_k_lo && k <= cf->goal_k_hi) {
            wk->Wxx[6 * k + 0] = df * 2.0 * cf->W[0]; wk->Wxx[6 * k + 2] = df * 2.0 * cf->W[1]; wk->Wxx[6 * k + 5] = df * 2.0 * cf->W[2];
        }
        wk->Wtv[k] = 0; wk->Wvv[k] = 0; wk->Www[k] = 0;
    }
    for (int k = 0; k < N; ++k) {
        double v = w[IU(k, 0)]; const double *yn = yc + 3 * (k + 1);
        wk->Wxx[6 * k + 5] += T * v * (yn[0] * cs[k] + yn[1] * sn[k]);
        wk->Wtv[k] = T * (yn[0] * sn[k] - yn[1] * cs[k]);
        if (wk->resto) { wk->Wvv[k] = wk->eta * wk->DR2[IU(k, 0)]; wk->Www[k] = wk->eta * wk->DR2[IU(k, 1)]; continue; }
        if (cf->cost_mode == 0) { double a = dmin0(v), b = dmax0(v); wk->Wvv[k] = df * (2.0 * cf->Wv_neg * a * a + 2.0 * cf->Wv_pos * b * b); }
        wk->Www[k] = df * 2.0 * cf->Ww;
    }
    for (int o = 0; o < O; ++o) for (int k = 1; k <= N; ++k) {
        int i = IS(o, k); double nx = wk->nrm[2 * i], ny = wk->nrm[2 * i + 1], h = yd[i] / wk->dist[i];
        wk->Wxx[6 * k + 0] += h * (1.0 - nx * nx); wk->Wxx[6 * k + 1] += h * (-nx * ny); wk->Wxx[6 * k + 2] += h * (1.0 - ny * ny);
    }
}

/* ------------------------------------------------------------------ */
/* dense symmetric indefinite factorisation (Bunch-Kaufman, full storage)
 * P A P^T = L D L^T ; perm[i] = original index at position i          */
/* ------------------------------------------------------------------ */
static void sym_swap(double *A, int n, int i, int j) {
    if (i == j) return;
    for (int c = 0; c < n; ++c) { double t = A[(size_t)i * n + c]; A[(size_t)i * n + c] = A[(size_t)j * n + c]; A[(size_t)j * n + c] = t; }
    for (int r = 0; r < n; ++r) { double t = A[(size_t)r * n + i]; A[(size_t)r * n + i] = A[(size_t)r * n + j]; A[(size_t)r * n + j] = t; }
}

/* returns 0 ok, 1 singular.  On exit A holds D on its (block) diagonal and the unit-lower L below it. */
static int bk_factor(double *A, int n, int *perm, int *pivtype, int *npos, int *nneg, int *nzero) {
    const double alpha = (1.0 + sqrt(17.0)) / 8.0;
    for (int i = 0; i < n; ++i) perm[i] = i;
    *npos = *nneg = *nzero = 0;
    int k = 0;
    while (k < n) {
        int kstep = 1, kp = k;
        double absakk = fabs(A[(size_t)k * n + k]);
        int imax = -1; double colmax = 0;
        for (int i = k + 1; i < n; ++i) { double t = fabs(A[(size_t)i * n + k]); if (t > colmax) { colmax = t; imax = i; } }
        if (fmax(absakk, colmax) == 0.0) { pivtype[k] = 1; (*nzero)++; k += 1; continue; }
        if (absakk >= alpha * colmax) kp = k;
        else {
            double rowmax = 0;
            for (int j = k; j < n; ++j) if (j != imax) { double t = fabs(A[(size_t)imax * n + j]); if (t > rowmax) rowmax = t; }
            if (absakk >= alpha * colmax * (colmax / rowmax)) kp = k;
            else if (fabs(A[(size_t)imax * n + imax]) >= alpha * rowmax) kp = imax;
            else { kp = imax; kstep = 2; }
        }
        int kk = k + kstep - 1;
        if (kp != kk) { sym_swap(A, n, kk, kp); int t = perm[kk]; perm[kk] = perm[kp]; perm[kp] = t; }
        if (kstep == 1) {
            double d = A[(size_t)k * n + k];
            pivtype[k] = 1;
            if (d > 0) (*npos)++; else if (d < 0) (*nneg)++; else (*nzero)++;
            double r = 1.0 / d;
            for (int i = k + 1; i < n; ++i) {
                double lik = A[(size_t)i * n + k] * r;
                if (lik != 0.0) for (int j = k + 1; j < n; ++j) A[(size_t)i * n + j] -= lik * A[(size_t)k * n + j];
            }
            for (int i = k + 1; i < n; ++i) { A[(size_t)i * n + k] *= r; A[(size_t)k * n + i] = 0; }
        } else {
            double a = A[(size_t)k * n + k], b = A[(size_t)(k + 1) * n + k], c = A[(size_t)(k + 1) * n + k + 1];
            double det = a * c - b * b;
            pivtype[k] = 2; pivtype[k + 1] = 0;
            if (det < 0) { (*npos)++; (*nneg)++; }
            else if (det > 0) { if (a > 0) (*npos) += 2; else (*nneg) += 2; }
            else (*nzero)++;
            /* W = A[i, k:k+2] D^{-1} */
            for (int i = k + 2; i < n; ++i) {
                double x0 = A[(size_t)i * n + k], x1 = A[(size_t)i * n + k + 1];
                double w0 = (c * x0 - b * x1) / det, w1 = (-b * x0 + a * x1) / det;
                if (w0 != 0.0 || w1 != 0.0)
                    for (int j = k + 2; j < n; ++j) A[(size_t)i * n + j] -= w0 * A[(size_t)k * n + j] + w1 * A[(size_t)(k + 1) * n + j];
                A[(size_t)i * n + k] = w0; A[(size_t)i * n + k + 1] = w1;
            }
            for (int j = k + 2; j < n; ++j) { A[(size_t)k * n + j] = 0; A[(size_t)(k + 1) * n + j] = 0; }
        }
        k += kstep;
    }
    return *nzero ? 1 : 0;
}

static void bk_solve(const double *A, int n, const int *perm, const int *pivtype, double *b /* in/out */, double *tmp) {
    for (int i = 0; i < n; ++i) tmp[i] = b[perm[i]];
    /* forward: L y = Pb ; the 2x2 blocks have an identity L-diagonal block */
    for (int k = 0; k < n;) {
        int st = pivtype[k] == 2 ? 2 : 1;
        for (int i = k + st; i < n; ++i) {
            double acc = A[(size_t)i * n + k] * tmp[k];
            if (st == 2) acc += A[(size_t)i * n + k + 1] * tmp[k + 1];
            tmp[i] -= acc;
        }
        k += st;
    }
    for (int k = 0; k < n;) {
        if (pivtype[k] == 2) {
            double a = A[(size_t)k * n + k], bb = A[(size_t)(k + 1) * n + k], c = A[(size_t)(k + 1) * n + k + 1];
            double det = a * c - bb * bb, x0 = tmp[k], x1 = tmp[k + 1];
            tmp[k] = (c * x0 - bb * x1) / det; tmp[k + 1] = (-bb * x0 + a * x1) / det; k += 2;
        } else { tmp[k] /= A[(size_t)k * n + k]; k += 1; }
    }
    /* backward: L^T x = y : process pivot blocks from last to first */
    {
        int k = n - 1;
        while (k >= 0) {
            int st = 1, k0 = k;
            if (pivtype[k] == 0) { st = 2; k0 = k - 1; }
            for (int c = k0; c < k0 + st; ++c) {
                double acc = 0;
                for (int i = k0 + st; i < n; ++i) acc += A[(size_t)i * n + c] * tmp[i];
                tmp[c] -= acc;
            }
            k = k0 - 1;
        }
    }
    for (int i = 0; i < n; ++i) b[perm[i]] = tmp[i];
}

/* ------------------------------------------------------------------ */
/* the linear system of one IPM step:
 *  [W+Dx   0    Jc^T  Jd^T] [dx ]   [bx]
 *  [ 0     Ds    0    -I  ] [ds ] = [bs]
 *  [ Jc    0     0     0  ] [dyc]   [bc]
 *  [ Jd   -I     0     0  ] [dyd]   [bd]
 * useW = 0 replaces W by zero (least-squares multiplier system).          */
/* ------------------------------------------------------------------ */
static int kkt_factor_dense(const kmo_config *cf, work_t *wk, const double *w, int useW) {
    int N = wk->N, O = wk->O, n = wk->n, ns = wk->ns, mc = wk->mc, nk = wk->nk; double T = cf->T;
    double *K = wk->K;
    memset(K, 0, sizeof(double) * (size_t)nk * nk);
#define KS(i, j, v) do { K[(size_t)(i) * nk + (j)] += (v); if ((i) != (j)) K[(size_t)(j) * nk + (i)] += (v); } while (0)
    for (int i = 0; i < n; ++i) KS(i, i, wk->Dx[i]);
    for (int i = 0; i < ns; ++i) KS(n + i, n + i, wk->Ds[i]);
    if (useW) {
        for (int k = 0; k <= N; ++k) {
            const double *h = wk->Wxx + 6 * k;
            KS(IX(k, 0), IX(k, 0), h[0]); KS(IX(k, 1), IX(k, 0), h[1]); KS(IX(k, 1), IX(k, 1), h[2]);
            KS(IX(k, 2), IX(k, 0), h[3]); KS(IX(k, 2), IX(k, 1), h[4]); KS(IX(k, 2), IX(k, 2), h[5]);
        }
        for (int k = 0; k < N; ++k) { KS(IU(k, 0), IX(k, 2), wk->Wtv[k]); KS(IU(k, 0), IU(k, 0), wk->Wvv[k]); KS(IU(k, 1), IU(k, 1), wk->Www[k]); }
    }
    int rc = n + ns, rd = n + ns + mc;
    for (int j = 0; j < 3; ++j) KS(rc + j, IX(0, j), 1.0);
    for (int k = 0; k < N; ++k) {
        int r = rc + 3 * (k + 1); double v = w[IU(k, 0)];
        for (int j = 0; j < 3; ++j) { KS(r + j, IX(k + 1, j), 1.0); KS(r + j, IX(k, j), -1.0); }
        KS(r + 0, IX(k, 2), T * v * wk->sn[k]); KS(r + 1, IX(k, 2), -T * v * wk->cs[k]);
        KS(r + 0, IU(k, 0), -T * wk->cs[k]); KS(r + 1, IU(k, 0), -T * wk->sn[k]); KS(r + 2, IU(k, 1), -T);
    }
    for (int o = 0; o < O; ++o) for (int k = 1; k <= N; ++k) {
        int i = IS(o, k);
        KS(rd + i, IX(k, 0), wk->nrm[2 * i]); KS(rd + i, IX(k, 1), wk->nrm[2 * i + 1]); KS(rd + i, n + i, -1.0);
    }
    if (wk->resto) {
        for (int i = 0; i < mc; ++i) KS(rc + i, rc + i, -wk->Dc[i]);
        for (int i = 0; i < ns; ++i) KS(rd + i, rd + i, -wk->Dd[i]);
    }
#undef KS
    int npos, nneg, nzero;
    int sing = bk_factor(K, nk, wk->perm, wk->pivtype, &npos, &nneg, &nzero);
    wk->fact_valid = 1;
    if (sing) return 0;
    return (npos == n + ns && nneg == mc + wk->md) ? 1 : 0;
}

static void kkt_solve_dense(work_t *wk, const double *bx, const double *bs, const double *bc, const double *bd,
                            double *dx, double *ds, double *dyc, double *dyd) {
    int n = wk->n, ns = wk->ns, mc = wk->mc, nk = wk->nk;
    memcpy(wk->rhs, bx, sizeof(double) * n); memcpy(wk->rhs + n, bs, sizeof(double) * ns);
    memcpy(wk->rhs + n + ns, bc, sizeof(double) * mc); memcpy(wk->rhs + n + ns + mc, bd, sizeof(double) * ns);
    bk_solve(wk->K, nk, wk->perm, wk->pivtype, wk->rhs, wk->L);
    memcpy(dx, wk->rhs, sizeof(double) * n); memcpy(ds, wk->rhs + n, sizeof(double) * ns);
    memcpy(dyc, wk->rhs + n + ns, sizeof(double) * mc); memcpy(dyd, wk->rhs + n + ns + mc, sizeof(double) * ns);
}

/* ---- stage-wise solver: condensed slacks + Riccati recursion (same system as above) ---- */
static void sym3_get(const double *h, double M[3][3]) {
    M[0][0] = h[0]; M[1][0] = M[0][1] = h[1]; M[1][1] = h[2]; M[2][0] = M[0][2] = h[3]; M[2][1] = M[1][2] = h[4]; M[2][2] = h[5];
}

static int kkt_factor_riccati(const kmo_config *cf, work_t *wk, const double *w, int useW) {
    int N = wk->N, O = wk->O; double T = cf->T;
    double P[3][3];
    /* terminal */
    for (int k = N; k >= 0; --k) {
        double Q[3][3] = {{0}};
        if (useW) sym3_get(wk->Wxx + 6 * k, Q);
        for (int j = 0; j < 3; ++j) Q[j][j] += wk->Dx[IX(k, j)];
        if (k >= 1) for (int o = 0; o < O; ++o) {
            int i = IS(o, k); double nx = wk->nrm[2 * i], ny = wk->nrm[2 * i + 1], sg = wk->Ds[i];
            Q[0][0] += sg * nx * nx; Q[0][1] += sg * nx * ny; Q[1][0] += sg * nx * ny; Q[1][1] += sg * ny * ny;
        }
        if (k == N) { memcpy(P, Q, sizeof(P)); }
        else {
            double v = w[IU(k, 0)], a13 = -T * v * wk->sn[k], a23 = T * v * wk->cs[k];
            double b11 = T * wk->cs[k], b21 = T * wk->sn[k], b32 = T;
            /* PA = P*A, A = I + e1 e3^T a13 + e2 e3^T a23 */
            double PA[3][3], PB[3][2];
            for (int i = 0; i < 3; ++i) { PA[i][0] = P[i][0]; PA[i][1] = P[i][1]; PA[i][2] = P[i][0] * a13 + P[i][1] * a23 + P[i][2];
                PB[i][0] = P[i][0] * b11 + P[i][1] * b21; PB[i][1] = P[i][2] * b32; }
            double Qxx[3][3], Qux[2][3], Quu[2][2];
            for (int j = 0; j < 3; ++j) {
                Qxx[0][j] = PA[0][j]; Qxx[1][j] = PA[1][j]; Qxx[2][j] = a13 * PA[0][j] + a23 * PA[1][j] + PA[2][j];
                Qux[0][j] = b11 * PA[0][j] + b21 * PA[1][j]; Qux[1][j] = b32 * PA[2][j];
            }
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Qxx[i][j] += Q[i][j];
            double Svt = useW ? wk->Wtv[k] : 0.0;
            Qux[0][2] += Svt;
            Quu[0][0] = b11 * PB[0][0] + b21 * PB[1][0] + (useW ? wk->Wvv[k] : 0.0) + wk->Dx[IU(k, 0)];
            Quu[0][1] = b11 * PB[0][1] + b21 * PB[1][1];
            Quu[1][1] = b32 * PB[2][1] + (useW ? wk->Www[k] : 0.0) + wk->Dx[IU(k, 1)];
            /* inertia: reduced Hessian PD <=> every Quu PD */
            double a = Quu[0][0], b = Quu[0][1], c = Quu[1][1];
            if (!(a > 0.0)) return 0;
            double sch = c - b * b / a;
            if (!(sch > 0.0)) return 0;
            double det = a * c - b * b;
            double i00 = c / det, i01 = -b / det, i11 = a / det;
            wk->Quui[3 * k] = i00; wk->Quui[3 * k + 1] = i01; wk->Quui[3 * k + 2] = i11;
            double *Kg = wk->Kg + 6 * k; /* K = -Quu^{-1} Qux (2x3) */
            for (int j = 0; j < 3; ++j) { Kg[j] = -(i00 * Qux[0][j] + i01 * Qux[1][j]); Kg[3 + j] = -(i01 * Qux[0][j] + i11 * Qux[1][j]); }
            double Pn[3][3];
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Pn[i][j] = Qxx[i][j] + Qux[0][i] * Kg[j] + Qux[1][i] * Kg[3 + j];
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) P[i][j] = 0.5 * (Pn[i][j] + Pn[j][i]);
        }
        double *pm = wk->Pm + 6 * k;
        pm[0] = P[0][0]; pm[1] = P[1][0]; pm[2] = P[1][1]; pm[3] = P[2][0]; pm[4] = P[2][1]; pm[5] = P[2][2];
    }
    return 1;
}

static void kkt_solve_riccati(const kmo_config *cf, work_t *wk, const double *w, int useW,
                              const double *bx, const double *bs, const double *bc, const double *bd,
                              double *dx, double *ds, double *dyc, double *dyd) {
    int N = wk->N, O = wk->O; double T = cf->T;
    double p[3] = {0, 0, 0}, P[3][3];
    double *kff = (double *)alloca(sizeof(double) * 2 * (N > 0 ? N : 1));
    /* backward vector pass */
    for (int k = N; k >= 0; --k) {
        double q[3];
        for (int j = 0; j < 3; ++j) q[j] = -bx[IX(k, j)];
        if (k >= 1) for (int o = 0; o < O; ++o) {
            int i = IS(o, k); double t = wk->Ds[i] * bd[i] + bs[i];
            q[0] -= wk->nrm[2 * i] * t; q[1] -= wk->nrm[2 * i + 1] * t;
        }
        if (k == N) { for (int j = 0; j < 3; ++j) p[j] = q[j]; }
        else {
            sym3_get(wk->Pm + 6 * (k + 1), P);
            const double *e = bc + 3 * (k + 1);
            double Pe[3];
            for (int i = 0; i < 3; ++i) Pe[i] = P[i][0] * e[0] + P[i][1] * e[1] + P[i][2] * e[2] + p[i];
            double v = w[IU(k, 0)], a13 = -T * v * wk->sn[k], a23 = T * v * wk->cs[k];
            double b11 = T * wk->cs[k], b21 = T * wk->sn[k], b32 = T;
            double qu[2] = {-bx[IU(k, 0)] + b11 * Pe[0] + b21 * Pe[1], -bx[IU(k, 1)] + b32 * Pe[2]};
            double qx[3] = {q[0] + Pe[0], q[1] + Pe[1], q[2] + a13 * Pe[0] + a23 * Pe[1] + Pe[2]};
            const double *qi = wk->Quui + 3 * k;
            kff[2 * k] = -(qi[0] * qu[0] + qi[1] * qu[1]); kff[2 * k + 1] = -(qi[1] * qu[0] + qi[2] * qu[1]);
            /* p_k = qx + Qux^T kff = qx - K^T Quu kff ... use p = qx + K^T qu (equivalent: Qux^T kff = K^T qu) */
            const double *Kg = wk->Kg + 6 * k;
            for (int j = 0; j < 3; ++j) p[j] = qx[j] + Kg[j] * qu[0] + Kg[3 + j] * qu[1];
        }
        for (int j = 0; j < 3; ++j) wk->pv[3 * k + j] = p[j];
    }
    /* forward */
    for (int j = 0; j < 3; ++j) dx[IX(0, j)] = bc[j];
    for (int k = 0; k < N; ++k) {
        const double *Kg = wk->Kg + 6 * k; const double *xk = dx + IX(k, 0);
        double du0 = Kg[0] * xk[0] + Kg[1] * xk[1] + Kg[2] * xk[2] + kff[2 * k];
        double du1 = Kg[3] * xk[0] + Kg[4] * xk[1] + Kg[5] * xk[2] + kff[2 * k + 1];
        dx[IU(k, 0)] = du0; dx[IU(k, 1)] = du1;
        double v = w[IU(k, 0)], a13 = -T * v * wk->sn[k], a23 = T * v * wk->cs[k];
        const double *e = bc + 3 * (k + 1);
        dx[IX(k + 1, 0)] = xk[0] + a13 * xk[2] + T * wk->cs[k] * du0 + e[0];
        dx[IX(k + 1, 1)] = xk[1] + a23 * xk[2] + T * wk->sn[k] * du0 + e[1];
        dx[IX(k + 1, 2)] = xk[2] + T * du1 + e[2];
    }
    for (int k = 0; k <= N; ++k) {
        sym3_get(wk->Pm + 6 * k, P); const double *xk = dx + IX(k, 0);
        for (int i = 0; i < 3; ++i) dyc[3 * k + i] = -(P[i][0] * xk[0] + P[i][1] * xk[1] + P[i][2] * xk[2] + wk->pv[3 * k + i]);
    }
    for (int o = 0; o < O; ++o) for (int k = 1; k <= N; ++k) {
        int i = IS(o, k);
        ds[i] = wk->nrm[2 * i] * dx[IX(k, 0)] + wk->nrm[2 * i + 1] * dx[IX(k, 1)] - bd[i];
        dyd[i] = wk->Ds[i] * ds[i] - bs[i];
    }
    (void)useW;
}

static int kkt_factor(const kmo_config *cf, work_t *wk, const double *w, int useW) {
    return wk->linsolve == 0 ? kkt_factor_dense(cf, wk, w, useW) : kkt_factor_riccati(cf, wk, w, useW);
}
static void kkt_solve(const kmo_config *cf, work_t *wk, const double *w, int useW, const double *bx, const double *bs,
                      const double *bc, const double *bd, double *dx, double *ds, double *dyc, double *dyd) {
    if (wk->linsolve == 0) kkt_solve_dense(wk, bx, bs, bc, bd, dx, ds, dyc, dyd);
    else kkt_solve_riccati(cf, wk, w, useW, bx, bs, bc, bd, dx, ds, dyc, dyd);
}

/* ------------------------------------------------------------------ */
/* IPM helpers                                                         */
/* ------------------------------------------------------------------ */
static double vmaxabs(const double *a, int n) { double m = 0; for (int i = 0; i < n; ++i) { double t = fabs(a[i]); if (t > m || t != t) m = t; } return m; }
static double vsumabs(const double *a, int n) { double m = 0; for (int i = 0; i < n; ++i) m += fabs(a[i]); return m; }

/* IPOPT Compare_le(lhs, rhs, BasVal): lhs - rhs <= 10 eps |BasVal| */
static int compare_le(double lhs, double rhs, double bas) { return lhs - rhs <= 10.0 * DBL_EPSILON * fabs(bas); }

typedef struct { double theta[FILTER_CAP], phi[FILTER_CAP]; int n; } filter_t;
static int filter_ok(const filter_t *F, double theta, double phi) {
    for (int i = 0; i < F->n; ++i) if (!(theta <= F->theta[i] || phi <= F->phi[i])) return 0;
    return 1;
}
static void filter_add(filter_t *F, double theta, double phi) {
    int m = 0; /* drop entries dominated by the new one */
    for (int i = 0; i < F->n; ++i) if (!(F->theta[i] >= theta && F->phi[i] >= phi)) { F->theta[m] = F->theta[i]; F->phi[m] = F->phi[i]; ++m; }
    F->n = m;
    if (F->n < FILTER_CAP) { F->theta[F->n] = theta; F->phi[F->n] = phi; F->n++; }
}

typedef struct {
    double theta, phi, f;       /* at a point */
} merit_t;

/* barrier objective and constraint violation at (w, s) */
static int eval_merit(const kmo_config *cf, work_t *wk, const double *w, const double *s, double mu, double *c, double *dms,
                      merit_t *m) {
    int n = wk->n, ns = wk->ns;
    double f = eval_f(cf, wk, w), bar = 0, damp = 0;
    for (int i = 0; i < n; ++i) {
        if (wk->hasL[i]) { double sl = w[i] - wk->lb[i]; if (!(sl > 0)) return 0; bar += log(sl); if (!wk->hasU[i]) damp += sl; }
        if (wk->hasU[i]) { double su = wk->ub[i] - w[i]; if (!(su > 0)) return 0; bar += log(su); if (!wk->hasL[i]) damp += su; }
    }
    for (int i = 0; i < ns; ++i) { double sl = s[i] - wk->dL; if (!(sl > 0)) return 0; bar += log(sl); damp += sl; }
    eval_c(cf, wk, w, c);
    double theta = vsumabs(c, wk->mc);
    if (ns) { eval_d(cf, wk, w, dms, NULL, NULL); for (int i = 0; i < ns; ++i) { dms[i] -= s[i]; theta += fabs(dms[i]); } }
    m->f = f; m->phi = f - mu * bar + KAPPA_D * mu * damp; m->theta = theta;
    if (!isfinite(m->phi) || !isfinite(m->theta)) return 0;
    return 1;
}

typedef struct {
    double dual_inf, primal_inf, min_sz, max_sz, sum_y, sum_z;
    int nb;
} errs_t;

/* residual norms at the current iterate; also leaves g, c, dms, trig, normals evaluated in wk */
static void eval_errors(const kmo_config *cf, work_t *wk, errs_t *e, double *rx /* n */, double *rs /* ns */) {
    int n = wk->n, ns = wk->ns;
    eval_trig(wk, wk->w, wk->cs, wk->sn);
    eval_grad(cf, wk, wk->w, wk->df, wk->g);
    eval_c(cf, wk, wk->w, wk->c);
    if (ns) { eval_d(cf, wk, wk->w, wk->dms, wk->nrm, wk->dist); for (int i = 0; i < ns; ++i) wk->dms[i] -= wk->s[i]; }
    memcpy(rx, wk->g, sizeof(double) * n);
    add_JcT(cf, wk, wk->w, wk->cs, wk->sn, wk->yc, rx);
    if (ns) add_JdT(wk, wk->nrm, wk->yd, rx);
    double mn = INFINITY, mx = 0, sz = 0; int nb = 0;
    for (int i = 0; i < n; ++i) {
        if (wk->hasL[i]) { rx[i] -= wk->zL[i]; double p = (wk->w[i] - wk->lb[i]) * wk->zL[i]; mn = fmin(mn, p); mx = fmax(mx, p); sz += fabs(wk->zL[i]); nb++; }
        if (wk->hasU[i]) { rx[i] += wk->zU[i]; double p = (wk->ub[i] - wk->w[i]) * wk->zU[i]; mn = fmin(mn, p); mx = fmax(mx, p); sz += fabs(wk->zU[i]); nb++; }
    }
    for (int i = 0; i < ns; ++i) {
        rs[i] = -wk->yd[i] - wk->vL[i];
        double p = (wk->s[i] - wk->dL) * wk->vL[i]; mn = fmin(mn, p); mx = fmax(mx, p); sz += fabs(wk->vL[i]); nb++;
    }
    /* (fmax would drop a NaN; the invalid-number test of the main loop needs it) */
    { double a = vmaxabs(rx, n), b = ns ? vmaxabs(rs, ns) : 0.0; e->dual_inf = (a != a || b != b) ? NAN : fmax(a, b); }
    { double a = vmaxabs(wk->c, wk->mc), b = ns ? vmaxabs(wk->dms, ns) : 0.0; e->primal_inf = (a != a || b != b) ? NAN : fmax(a, b); }
    e->min_sz = nb ? mn : 0.0; e->max_sz = mx; e->sum_z = sz; e->nb = nb;
    e->sum_y = vsumabs(wk->yc, wk->mc) + (ns ? vsumabs(wk->yd, ns) : 0.0);
}

static double compl_inf(const errs_t *e, double mu) { return e->nb ? fmax(fabs(e->max_sz - mu), fabs(e->min_sz - mu)) : 0.0; }

static double opt_error(const work_t *wk, const errs_t *e, double mu) {
    int m = wk->mc + wk->md;
    double sd = fmax(S_MAX, (e->sum_y + e->sum_z) / (double)(m + e->nb)) / S_MAX;
    double sc = e->nb ? fmax(S_MAX, e->sum_z / (double)e->nb) / S_MAX : 1.0;
    return fmax(e->dual_inf / sd, fmax(e->primal_inf, compl_inf(e, mu) / sc));
}

/* fraction to the boundary for the primal variables */
static double ftb_primal(const work_t *wk, const double *dx, const double *ds, double tau) {
    double a = 1.0; int n = wk->n, ns = wk->ns;
    for (int i = 0; i < n; ++i) {
        if (wk->hasL[i] && dx[i] < 0) a = fmin(a, -tau * (wk->w[i] - wk->lb[i]) / dx[i]);
        if (wk->hasU[i] && dx[i] > 0) a = fmin(a, tau * (wk->ub[i] - wk->w[i]) / dx[i]);
    }
    for (int i = 0; i < ns; ++i) if (ds[i] < 0) a = fmin(a, -tau * (wk->s[i] - wk->dL) / ds[i]);
    return a;
}

/* bound-multiplier steps from the primal step + fraction to the boundary for them */
static double dual_steps(work_t *wk, const double *dx, const double *ds, double mu, double tau) {
    double a = 1.0; int n = wk->n, ns = wk->ns;
    for (int i = 0; i < n; ++i) {
        wk->dzL[i] = wk->dzU[i] = 0;
        if (wk->hasL[i]) { double sl = wk->w[i] - wk->lb[i]; wk->dzL[i] = mu / sl - wk->zL[i] - wk->zL[i] / sl * dx[i]; if (wk->dzL[i] < 0) a = fmin(a, -tau * wk->zL[i] / wk->dzL[i]); }
        if (wk->hasU[i]) { double su = wk->ub[i] - wk->w[i]; wk->dzU[i] = mu / su - wk->zU[i] + wk->zU[i] / su * dx[i]; if (wk->dzU[i] < 0) a = fmin(a, -tau * wk->zU[i] / wk->dzU[i]); }
    }
    for (int i = 0; i < ns; ++i) { double sl = wk->s[i] - wk->dL; wk->dvL[i] = mu / sl - wk->vL[i] - wk->vL[i] / sl * ds[i]; if (wk->dvL[i] < 0) a = fmin(a, -tau * wk->vL[i] / wk->dvL[i]); }
    return a;
}

/* right-hand sides of the step system at the current iterate (needs eval_errors to have run) */
static void build_rhs(const kmo_config *cf, work_t *wk, double mu, double delta_w, const double *cvec, const double *dvec) {
    int n = wk->n, ns = wk->ns;
    memcpy(wk->bx, wk->g, sizeof(double) * n);
    add_JcT(cf, wk, wk->w, wk->cs, wk->sn, wk->yc, wk->bx);
    if (ns) add_JdT(wk, wk->nrm, wk->yd, wk->bx);
    for (int i = 0; i < n; ++i) {
        double sig = 0, r = wk->bx[i];
        if (wk->hasL[i]) { double sl = wk->w[i] - wk->lb[i]; sig += wk->zL[i] / sl; r -= mu / sl; if (!wk->hasU[i]) r += KAPPA_D * mu; }
        if (wk->hasU[i]) { double su = wk->ub[i] - wk->w[i]; sig += wk->zU[i] / su; r += mu / su; if (!wk->hasL[i]) r -= KAPPA_D * mu; }
        wk->Dx[i] = sig + delta_w; wk->bx[i] = -r;
    }
    for (int i = 0; i < ns; ++i) {
        double sl = wk->s[i] - wk->dL;
        wk->Ds[i] = wk->vL[i] / sl + delta_w;
        wk->bs[i] = -(-wk->yd[i] - mu / sl + KAPPA_D * mu);
        wk->bd[i] = -dvec[i];
    }
    for (int i = 0; i < wk->mc; ++i) wk->bc[i] = -cvec[i];
}

/* ------------------------------------------------------------------ */
/* feasibility restoration phase (IPOPT MinC_1NrmRestorationPhase / RestoIpoptNLP / RestoIterateInitializer /
 * RestoFilterConvergenceCheck; Waechter & Biegler 2006, section 3.3):
 *     min  rho (sum n_c + p_c + n_d + p_d) + eta/2 |D_R (x - x_ref)|^2      rho = 1000, eta = sqrt(mu), D_R = diag 1/max(1,|x_ref|)
 *     s.t. c(x) + n_c - p_c = 0,   d(x) - s + n_d - p_d = 0,   x and s within their bounds,   n, p >= 0
 * solved by the same interior-point algorithm (own filter, own barrier parameter starting at max(mu, |c|_inf, |d - s|_inf)); n and p
 * are eliminated from the step system, which leaves -(n/z_n + p/z_p) on the diagonal of the constraint blocks.  It returns to the
 * regular algorithm as soon as an iterate reduces the ORIGINAL constraint violation to 0.9 of its value at entry and is acceptable
 * to the original filter and current point; if instead it converges on its own problem the original NLP is locally infeasible.
 * Always uses the dense augmented system.  Returns 0 (x, s, bound multipliers replaced; y = 0) or a final status.
 * ------------------------------------------------------------------ */
typedef struct {
    double *nc, *pc, *znc, *zpc, *nd, *pd, *znd, *zpd, *xref;
    double *dnc, *dpc, *dnd, *dpd, *nct, *pct, *ndt, *pdt;
    double *rc, *rd, *gR, *rnc, *rpc, *rnd, *rpd;
} resto_t;

static double resto_f(const work_t *wk, const resto_t *R, const double *w, const double *nc, const double *pc, const double *nd, const double *pd) {
    double f = 0, q = 0;
    for (int i = 0; i < wk->mc; ++i) f += nc[i] + pc[i];
    for (int i = 0; i < wk->ns; ++i) f += nd[i] + pd[i];
    for (int i = 0; i < wk->n; ++i) { double e = w[i] - R->xref[i]; q += wk->DR2[i] * e * e; }
    return RESTO_RHO * f + 0.5 * wk->eta * q;
}

/* barrier objective and constraint violation of the restoration problem; leaves c + n - p in ct, d - s + n - p in dt */
static int resto_merit(const kmo_config *cf, work_t *wk, const resto_t *R, const double *w, const double *s, const double *nc, const double *pc,
                       const double *nd, const double *pd, double mu, double *ct, double *dt, merit_t *m) {
    int n = wk->n, ns = wk->ns, mc = wk->mc;
    double bar = 0, damp = 0;
    for (int i = 0; i < n; ++i) {
        if (wk->hasL[i]) { double sl = w[i] - wk->lb[i]; if (!(sl > 0)) return 0; bar += log(sl); if (!wk->hasU[i]) damp += sl; }
        if (wk->hasU[i]) { double su = wk->ub[i] - w[i]; if (!(su > 0)) return 0; bar += log(su); if (!wk->hasL[i]) damp += su; }
    }
    for (int i = 0; i < ns; ++i) { double sl = s[i] - wk->dL; if (!(sl > 0)) return 0; bar += log(sl); damp += sl; }
    for (int i = 0; i < mc; ++i) { if (!(nc[i] > 0) || !(pc[i] > 0)) return 0; bar += log(nc[i]) + log(pc[i]); damp += nc[i] + pc[i]; }
    for (int i = 0; i < ns; ++i) { if (!(nd[i] > 0) || !(pd[i] > 0)) return 0; bar += log(nd[i]) + log(pd[i]); damp += nd[i] + pd[i]; }
    eval_c(cf, wk, w, ct);
    double theta = 0;
    for (int i = 0; i < mc; ++i) { ct[i] += nc[i] - pc[i]; theta += fabs(ct[i]); }
    if (ns) { eval_d(cf, wk, w, dt, NULL, NULL); for (int i = 0; i < ns; ++i) { dt[i] += -s[i] + nd[i] - pd[i]; theta += fabs(dt[i]); } }
    m->f = resto_f(wk, R, w, nc, pc, nd, pd); m->phi = m->f - mu * bar + KAPPA_D * mu * damp; m->theta = theta;
    return isfinite(m->phi) && isfinite(m->theta);
}

static int resto_acceptable(const filter_t *F, const merit_t *cur, const merit_t *tri, double gBD, double alpha_test, double theta_max, double theta_min) {
    if (tri->theta > theta_max) return 0;
    int acc;
    int ftype = gBD < 0 && alpha_test * pow(-gBD, S_PHI) > DELTA_LS * pow(cur->theta, S_THETA);
    if (ftype && cur->theta <= theta_min) acc = compare_le(tri->phi - cur->phi, ETA_PHI * alpha_test * gBD, cur->phi);
    else {
        acc = 1;
        if (tri->phi > cur->phi) { double bas = fabs(cur->phi) > 10.0 ? log10(fabs(cur->phi)) : 1.0; if (log10(tri->phi - cur->phi) > OBJ_MAX_INC + bas) acc = 0; }
        if (acc) acc = compare_le(tri->theta, (1.0 - GAMMA_THETA) * cur->theta, cur->theta) || compare_le(tri->phi - cur->phi, -GAMMA_PHI * cur->theta, cur->phi);
    }
    if (acc) acc = filter_ok(F, tri->theta, tri->phi);
    return acc;
}

static int restoration(const kmo_config *cf, work_t *wk, double mu_orig, const filter_t *F_orig, double theta_ref, double phi_ref,
                       int *iter, kmo_diag *dg) {
    const int n = wk->n, ns = wk->ns, mc = wk->mc;
    const double rho = RESTO_RHO;
    int st = -100;
    if (!wk->K) {   /* the restoration phase always factors the dense augmented system */
        wk->K = (double *)xcalloc((size_t)wk->nk * wk->nk, sizeof(double)); wk->L = (double *)xcalloc((size_t)wk->nk * wk->nk, sizeof(double));
        wk->rhs = (double *)xcalloc(wk->nk, sizeof(double)); wk->perm = (int *)xcalloc(wk->nk, sizeof(int)); wk->pivtype = (int *)xcalloc(wk->nk, sizeof(int));
    }
    resto_t R;
    const size_t tot = (size_t)16 * mc + 16 * (ns ? ns : 1) + 2 * n + 16;
    double *pool = (double *)xcalloc(tot, sizeof(double)), *q = pool;
#define TAKE(name, cnt) R.name = q; q += (cnt)
    TAKE(nc, mc); TAKE(pc, mc); TAKE(znc, mc); TAKE(zpc, mc); TAKE(dnc, mc); TAKE(dpc, mc); TAKE(nct, mc); TAKE(pct, mc); TAKE(rc, mc); TAKE(rnc, mc); TAKE(rpc, mc);
    TAKE(nd, ns); TAKE(pd, ns); TAKE(znd, ns); TAKE(zpd, ns); TAKE(dnd, ns); TAKE(dpd, ns); TAKE(ndt, ns); TAKE(pdt, ns); TAKE(rd, ns); TAKE(rnd, ns); TAKE(rpd, ns);
    TAKE(xref, n); TAKE(gR, n);
#undef TAKE
    filter_t *F = (filter_t *)xcalloc(1, sizeof(filter_t));
    const int saved_linsolve = wk->linsolve;
    wk->linsolve = 0; wk->resto = 1;

    /* RestoIterateInitializer: x, s kept; mu = max(mu, |c|_inf, |d - s|_inf); n, p from the complementarity-consistent formula */
    eval_trig(wk, wk->w, wk->cs, wk->sn);
    eval_c(cf, wk, wk->w, wk->c);
    if (ns) { eval_d(cf, wk, wk->w, wk->dms, wk->nrm, wk->dist); for (int i = 0; i < ns; ++i) wk->dms[i] -= wk->s[i]; }
    double mu = fmax(mu_orig, fmax(vmaxabs(wk->c, mc), ns ? vmaxabs(wk->dms, ns) : 0.0)), tau = fmax(TAU_MIN, 1.0 - mu);
    wk->eta = sqrt(mu_orig);
    for (int i = 0; i < n; ++i) { R.xref[i] = wk->w[i]; double dr = 1.0 / fmax(1.0, fabs(wk->w[i])); wk->DR2[i] = dr * dr; }
    for (int i = 0; i < mc + ns; ++i) {
        double cv = i < mc ? wk->c[i] : wk->dms[i - mc];
        double a = (mu - rho * cv) / (2.0 * rho), nn = a + sqrt(a * a + mu * cv / (2.0 * rho)), pp = cv + nn;
        if (i < mc) { R.nc[i] = nn; R.pc[i] = pp; R.znc[i] = mu / nn; R.zpc[i] = mu / pp; }
        else { R.nd[i - mc] = nn; R.pd[i - mc] = pp; R.znd[i - mc] = mu / nn; R.zpd[i - mc] = mu / pp; }
    }
    for (int i = 0; i < n; ++i) { if (wk->hasL[i]) wk->zL[i] = fmin(rho, wk->zL[i]); if (wk->hasU[i]) wk->zU[i] = fmin(rho, wk->zU[i]); }
    for (int i = 0; i < ns; ++i) wk->vL[i] = fmin(rho, wk->vL[i]);
    memset(wk->yc, 0, sizeof(double) * mc); if (ns) memset(wk->yd, 0, sizeof(double) * ns);

    double theta_max = -1, theta_min = -1, delta_last = 0.0;
    int first = 1;
    double *rx = wk->dx2, *rs = wk->ds2;
    for (;;) {
        /* ---- residuals of the restoration problem at the current iterate ---- */
        eval_trig(wk, wk->w, wk->cs, wk->sn);
        eval_c(cf, wk, wk->w, wk->c);
        if (ns) { eval_d(cf, wk, wk->w, wk->dms, wk->nrm, wk->dist); for (int i = 0; i < ns; ++i) wk->dms[i] -= wk->s[i]; }
        for (int i = 0; i < n; ++i) R.gR[i] = wk->eta * wk->DR2[i] * (wk->w[i] - R.xref[i]);
        memcpy(rx, R.gR, sizeof(double) * n);
        add_JcT(cf, wk, wk->w, wk->cs, wk->sn, wk->yc, rx);
        if (ns) add_JdT(wk, wk->nrm, wk->yd, rx);
        double mn = INFINITY, mx = 0, sz = 0, dinf = 0, pinf = 0; int nb = 0;
#define CP(slack, z) do { double p_ = (slack) * (z); mn = fmin(mn, p_); mx = fmax(mx, p_); sz += fabs(z); nb++; } while (0)
        for (int i = 0; i < n; ++i) {
            if (wk->hasL[i]) { rx[i] -= wk->zL[i]; CP(wk->w[i] - wk->lb[i], wk->zL[i]); }
            if (wk->hasU[i]) { rx[i] += wk->zU[i]; CP(wk->ub[i] - wk->w[i], wk->zU[i]); }
        }
        for (int i = 0; i < ns; ++i) { rs[i] = -wk->yd[i] - wk->vL[i]; CP(wk->s[i] - wk->dL, wk->vL[i]); }
        for (int i = 0; i < mc; ++i) {
            R.rc[i] = wk->c[i] + R.nc[i] - R.pc[i];
            dinf = fmax(dinf, fmax(fabs(rho + wk->yc[i] - R.znc[i]), fabs(rho - wk->yc[i] - R.zpc[i])));
            CP(R.nc[i], R.znc[i]); CP(R.pc[i], R.zpc[i]);
        }
        for (int i = 0; i < ns; ++i) {
            R.rd[i] = wk->dms[i] + R.nd[i] - R.pd[i];
            dinf = fmax(dinf, fmax(fabs(rho + wk->yd[i] - R.znd[i]), fabs(rho - wk->yd[i] - R.zpd[i])));
            CP(R.nd[i], R.znd[i]); CP(R.pd[i], R.zpd[i]);
        }
#undef CP
        { double a = vmaxabs(rx, n), b = ns ? vmaxabs(rs, ns) : 0.0; dinf = (a != a || b != b) ? NAN : fmax(dinf, fmax(a, b)); }
        { double a = vmaxabs(R.rc, mc), b = ns ? vmaxabs(R.rd, ns) : 0.0; pinf = (a != a || b != b) ? NAN : fmax(a, b); }
        const double sum_y = vsumabs(wk->yc, mc) + (ns ? vsumabs(wk->yd, ns) : 0.0);
        const double sd = fmax(S_MAX, (sum_y + sz) / (double)(mc + ns + nb)) / S_MAX, sc = fmax(S_MAX, sz / (double)nb) / S_MAX;
#define RCOMPL(m_) fmax(fabs(mx - (m_)), fabs(mn - (m_)))
#define RERR(m_) fmax(dinf / sd, fmax(pinf, RCOMPL(m_) / sc))
        const double E0 = RERR(0.0);
        if (!isfinite(E0) || !isfinite(dinf) || !isfinite(pinf)) { st = ST_INVALID_NUMBER; break; }

        /* ---- RestoFilterConvergenceCheck (not in the first iteration: the start is the point the line search failed at) ---- */
        if (!first) {
            const double theta_o = vsumabs(wk->c, mc) + (ns ? vsumabs(wk->dms, ns) : 0.0);
            if (theta_o <= RESTO_KAPPA * theta_ref) {
                merit_t mo;
                if (eval_merit(cf, wk, wk->w, wk->s, mu_orig, wk->ct, wk->dmst, &mo) && filter_ok(F_orig, mo.theta, mo.phi) &&
                    (compare_le(mo.theta, (1.0 - GAMMA_THETA) * theta_ref, theta_ref) || compare_le(mo.phi - phi_ref, -GAMMA_PHI * theta_ref, phi_ref))) { st = 0; break; }
            }
            if (E0 <= cf->tol && dinf <= DUAL_INF_TOL && pinf <= CONSTR_VIOL_TOL && RCOMPL(0.0) <= COMPL_INF_TOL) {
                const double po = fmax(vmaxabs(wk->c, mc), ns ? vmaxabs(wk->dms, ns) : 0.0);
                st = po <= 1e2 * cf->tol ? ST_RESTORATION : ST_INFEASIBLE;   /* converged to a feasible point the filter rejects / local infeasibility */
                break;
            }
        }
        first = 0;
        if (*iter >= cf->max_iter) { st = ST_MAXITER; break; }
        if (vmaxabs(wk->w, n) > DIVERGING_TOL) { st = ST_DIVERGING; break; }

        /* monotone barrier update */
        {
            int done = 0;
            while (!done && RERR(mu) <= KAPPA_EPS * mu) {
                double nm = fmax(fmin(MU_LIN * mu, pow(mu, MU_SUPER)), fmin(cf->tol, COMPL_INF_TOL) / (KAPPA_EPS + 1.0));
                int changed = nm != mu;
                mu = nm; tau = fmax(TAU_MIN, 1.0 - mu);
                if (changed) F->n = 0; else done = 1;
            }
        }
#undef RERR
#undef RCOMPL

        /* ---- search direction: n, p eliminated (their blocks leave -Dc, -Dd on the constraint diagonals) ---- */
        eval_hess(cf, wk, wk->w, wk->yc, wk->yd, wk->cs, wk->sn);
        for (int i = 0; i < mc; ++i) {
            R.rnc[i] = rho + wk->yc[i] - mu / R.nc[i] + KAPPA_D * mu; R.rpc[i] = rho - wk->yc[i] - mu / R.pc[i] + KAPPA_D * mu;
            wk->Dc[i] = R.nc[i] / R.znc[i] + R.pc[i] / R.zpc[i];
        }
        for (int i = 0; i < ns; ++i) {
            R.rnd[i] = rho + wk->yd[i] - mu / R.nd[i] + KAPPA_D * mu; R.rpd[i] = rho - wk->yd[i] - mu / R.pd[i] + KAPPA_D * mu;
            wk->Dd[i] = R.nd[i] / R.znd[i] + R.pd[i] / R.zpd[i];
        }
        double delta = 0.0; int ok = 0;
        for (;;) {
            memcpy(wk->g, R.gR, sizeof(double) * n);            /* build_rhs works on wk->g */
            build_rhs(cf, wk, mu, delta, R.rc, R.rd);
            for (int i = 0; i < mc; ++i) wk->bc[i] += R.rnc[i] * R.nc[i] / R.znc[i] - R.rpc[i] * R.pc[i] / R.zpc[i];
            for (int i = 0; i < ns; ++i) wk->bd[i] += R.rnd[i] * R.nd[i] / R.znd[i] - R.rpd[i] * R.pd[i] / R.zpd[i];
            ok = kkt_factor_dense(cf, wk, wk->w, 1); if (dg) dg->n_factor++;
            if (ok) break;
            if (delta == 0.0) delta = delta_last == 0.0 ? DELTA_W_INIT : fmax(DELTA_W_MIN, delta_last * DELTA_W_DEC);
            else delta = (delta_last == 0.0 || 1e5 * delta_last < delta) ? DELTA_W_INC_FIRST * delta : DELTA_W_INC * delta;
            if (delta > DELTA_W_MAX) break;
        }
        if (!ok) { st = ST_STEP_ERROR; break; }
        if (delta > 0.0) delta_last = delta;
        kkt_solve_dense(wk, wk->bx, wk->bs, wk->bc, wk->bd, wk->dx, wk->ds, wk->dyc, wk->dyd);
#define ELIM(dyc_, dyd_) do { \
        for (int i = 0; i < mc; ++i) { R.dnc[i] = -(R.rnc[i] + (dyc_)[i]) * R.nc[i] / R.znc[i]; R.dpc[i] = -(R.rpc[i] - (dyc_)[i]) * R.pc[i] / R.zpc[i]; } \
        for (int i = 0; i < ns; ++i) { R.dnd[i] = -(R.rnd[i] + (dyd_)[i]) * R.nd[i] / R.znd[i]; R.dpd[i] = -(R.rpd[i] - (dyd_)[i]) * R.pd[i] / R.zpd[i]; } } while (0)
#define FTB_NP(a_) do { \
        for (int i = 0; i < mc; ++i) { if (R.dnc[i] < 0) (a_) = fmin((a_), -tau * R.nc[i] / R.dnc[i]); if (R.dpc[i] < 0) (a_) = fmin((a_), -tau * R.pc[i] / R.dpc[i]); } \
        for (int i = 0; i < ns; ++i) { if (R.dnd[i] < 0) (a_) = fmin((a_), -tau * R.nd[i] / R.dnd[i]); if (R.dpd[i] < 0) (a_) = fmin((a_), -tau * R.pd[i] / R.dpd[i]); } } while (0)
        ELIM(wk->dyc, wk->dyd);

        /* ---- line search on the restoration problem ---- */
        merit_t cur;
        if (!resto_merit(cf, wk, &R, wk->w, wk->s, R.nc, R.pc, R.nd, R.pd, mu, wk->ct, wk->dmst, &cur)) { st = ST_INVALID_NUMBER; break; }
        double gBD = 0;
        for (int i = 0; i < n; ++i) {
            double gp = R.gR[i];
            if (wk->hasL[i]) { gp -= mu / (wk->w[i] - wk->lb[i]); if (!wk->hasU[i]) gp += KAPPA_D * mu; }
            if (wk->hasU[i]) { gp += mu / (wk->ub[i] - wk->w[i]); if (!wk->hasL[i]) gp -= KAPPA_D * mu; }
            gBD += gp * wk->dx[i];
        }
        for (int i = 0; i < ns; ++i) gBD += (-mu / (wk->s[i] - wk->dL) + KAPPA_D * mu) * wk->ds[i];
        for (int i = 0; i < mc; ++i) gBD += (rho - mu / R.nc[i] + KAPPA_D * mu) * R.dnc[i] + (rho - mu / R.pc[i] + KAPPA_D * mu) * R.dpc[i];
        for (int i = 0; i < ns; ++i) gBD += (rho - mu / R.nd[i] + KAPPA_D * mu) * R.dnd[i] + (rho - mu / R.pd[i] + KAPPA_D * mu) * R.dpd[i];
        if (theta_max < 0) { theta_max = RESTO_THETA_MAX_FACT * fmax(1.0, cur.theta); theta_min = THETA_MIN_FACT * fmax(1.0, cur.theta); }
        double alpha_min = GAMMA_THETA;
        if (gBD < 0) {
            alpha_min = fmin(GAMMA_THETA, GAMMA_PHI * cur.theta / (-gBD));
            if (cur.theta <= theta_min) alpha_min = fmin(alpha_min, DELTA_LS * pow(cur.theta, S_THETA) / pow(-gBD, S_PHI));
        }
        alpha_min *= ALPHA_MIN_FRAC;
        double alpha_max = ftb_primal(wk, wk->dx, wk->ds, tau);
        FTB_NP(alpha_max);
        double alpha = alpha_max, alpha_test = alpha_max;
        /* the step finally taken (the corrected one after an accepted second-order correction) */
        double *sdx = wk->dx, *sds = wk->ds, *sdyc = wk->dyc, *sdyd = wk->dyd;
        int accept = 0, nsteps = 0;
        merit_t tri;
#define TRIAL(a_, dx_, ds_) do { \
        for (int i = 0; i < n; ++i) wk->wt[i] = wk->w[i] + (a_) * (dx_)[i]; \
        for (int i = 0; i < ns; ++i) wk->st[i] = wk->s[i] + (a_) * (ds_)[i]; \
        for (int i = 0; i < mc; ++i) { R.nct[i] = R.nc[i] + (a_) * R.dnc[i]; R.pct[i] = R.pc[i] + (a_) * R.dpc[i]; } \
        for (int i = 0; i < ns; ++i) { R.ndt[i] = R.nd[i] + (a_) * R.dnd[i]; R.pdt[i] = R.pd[i] + (a_) * R.dpd[i]; } } while (0)
        while (alpha > alpha_min || nsteps == 0) {
            TRIAL(alpha, wk->dx, wk->ds);
            int evok = resto_merit(cf, wk, &R, wk->wt, wk->st, R.nct, R.pct, R.ndt, R.pdt, mu, wk->ct, wk->dmst, &tri); if (dg) dg->n_trials++;
            alpha_test = alpha;
            if (evok) accept = resto_acceptable(F, &cur, &tri, gBD, alpha_test, theta_max, theta_min);
            if (accept) break;
            if (evok && alpha == alpha_max && cur.theta <= tri.theta) {   /* second-order correction */
                double theta_soc_old = 0, theta_trial = tri.theta, alpha_soc = alpha;
                memcpy(wk->csoc, R.rc, sizeof(double) * mc); if (ns) memcpy(wk->dsoc, R.rd, sizeof(double) * ns);
                int count = 0;
                while (count < MAX_SOC && !accept && (count == 0 || theta_trial <= KAPPA_SOC * theta_soc_old)) {
                    theta_soc_old = theta_trial;
                    for (int i = 0; i < mc; ++i) { wk->csoc[i] = alpha_soc * wk->csoc[i] + wk->ct[i]; wk->bc[i] = -wk->csoc[i] + R.rnc[i] * R.nc[i] / R.znc[i] - R.rpc[i] * R.pc[i] / R.zpc[i]; }
                    for (int i = 0; i < ns; ++i) { wk->dsoc[i] = alpha_soc * wk->dsoc[i] + wk->dmst[i]; wk->bd[i] = -wk->dsoc[i] + R.rnd[i] * R.nd[i] / R.znd[i] - R.rpd[i] * R.pd[i] / R.zpd[i]; }
                    kkt_solve_dense(wk, wk->bx, wk->bs, wk->bc, wk->bd, wk->dx2, wk->ds2, wk->dyc2, wk->dyd2); if (dg) dg->n_soc++;
                    ELIM(wk->dyc2, wk->dyd2);
                    alpha_soc = ftb_primal(wk, wk->dx2, wk->ds2, tau);
                    FTB_NP(alpha_soc);
                    TRIAL(alpha_soc, wk->dx2, wk->ds2);
                    merit_t ts; int e2 = resto_merit(cf, wk, &R, wk->wt, wk->st, R.nct, R.pct, R.ndt, R.pdt, mu, wk->ct, wk->dmst, &ts); if (dg) dg->n_trials++;
                    if (!e2) break;
                    if (resto_acceptable(F, &cur, &ts, gBD, alpha_test, theta_max, theta_min)) { accept = 1; tri = ts; alpha = alpha_soc; sdx = wk->dx2; sds = wk->ds2; sdyc = wk->dyc2; sdyd = wk->dyd2; }
                    else { count++; theta_trial = ts.theta; }
                }
                if (accept) break;
                ELIM(wk->dyc, wk->dyd);   /* back to the original step */
            }
            alpha *= ALPHA_RED; nsteps++;
        }
        if (!accept) { st = ST_RESTORATION; break; }   /* the restoration phase's own line search failed: Restoration_Failed */
        {
            int ftype = gBD < 0 && alpha_test * pow(-gBD, S_PHI) > DELTA_LS * pow(cur.theta, S_THETA);
            if (!ftype || !compare_le(tri.phi - cur.phi, ETA_PHI * alpha_test * gBD, cur.phi)) filter_add(F, (1.0 - GAMMA_THETA) * cur.theta, cur.phi - GAMMA_PHI * cur.theta);
        }
        /* ---- accept: primal with alpha, y with alpha, every bound multiplier with its own fraction-to-the-boundary step ---- */
        double adu = dual_steps(wk, sdx, sds, mu, tau);
        for (int i = 0; i < mc; ++i) {   /* dz = mu/slack - z - z/slack d */
            double a = mu / R.nc[i] - R.znc[i] - R.znc[i] / R.nc[i] * R.dnc[i], b = mu / R.pc[i] - R.zpc[i] - R.zpc[i] / R.pc[i] * R.dpc[i];
            if (a < 0) adu = fmin(adu, -tau * R.znc[i] / a);
            if (b < 0) adu = fmin(adu, -tau * R.zpc[i] / b);
        }
        for (int i = 0; i < ns; ++i) {
            double a = mu / R.nd[i] - R.znd[i] - R.znd[i] / R.nd[i] * R.dnd[i], b = mu / R.pd[i] - R.zpd[i] - R.zpd[i] / R.pd[i] * R.dpd[i];
            if (a < 0) adu = fmin(adu, -tau * R.znd[i] / a);
            if (b < 0) adu = fmin(adu, -tau * R.zpd[i] / b);
        }
#define ZUP(z_, sl_old, d_, sl_new) do { double dz_ = mu / (sl_old) - (z_) - (z_) / (sl_old) * (d_), zn_ = (z_) + adu * dz_; (z_) = fmax(fmin(zn_, KAPPA_SIGMA * mu / (sl_new)), mu / (KAPPA_SIGMA * (sl_new))); } while (0)
        for (int i = 0; i < mc; ++i) {
            double nn = R.nc[i] + alpha * R.dnc[i], pp = R.pc[i] + alpha * R.dpc[i];
            ZUP(R.znc[i], R.nc[i], R.dnc[i], nn); ZUP(R.zpc[i], R.pc[i], R.dpc[i], pp);
            R.nc[i] = nn; R.pc[i] = pp;
        }
        for (int i = 0; i < ns; ++i) {
            double nn = R.nd[i] + alpha * R.dnd[i], pp = R.pd[i] + alpha * R.dpd[i];
            ZUP(R.znd[i], R.nd[i], R.dnd[i], nn); ZUP(R.zpd[i], R.pd[i], R.dpd[i], pp);
            R.nd[i] = nn; R.pd[i] = pp;
        }
#undef ZUP
        for (int i = 0; i < n; ++i) wk->w[i] += alpha * sdx[i];
        for (int i = 0; i < ns; ++i) wk->s[i] += alpha * sds[i];
        for (int i = 0; i < mc; ++i) wk->yc[i] += alpha * sdyc[i];
        for (int i = 0; i < ns; ++i) wk->yd[i] += alpha * sdyd[i];
        for (int i = 0; i < n; ++i) {
            if (wk->hasL[i]) { double sl = wk->w[i] - wk->lb[i], z = wk->zL[i] + adu * wk->dzL[i]; wk->zL[i] = fmax(fmin(z, KAPPA_SIGMA * mu / sl), mu / (KAPPA_SIGMA * sl)); }
            if (wk->hasU[i]) { double su = wk->ub[i] - wk->w[i], z = wk->zU[i] + adu * wk->dzU[i]; wk->zU[i] = fmax(fmin(z, KAPPA_SIGMA * mu / su), mu / (KAPPA_SIGMA * su)); }
        }
        for (int i = 0; i < ns; ++i) { double sl = wk->s[i] - wk->dL, z = wk->vL[i] + adu * wk->dvL[i]; wk->vL[i] = fmax(fmin(z, KAPPA_SIGMA * mu / sl), mu / (KAPPA_SIGMA * sl)); }
        (*iter)++;
#undef ELIM
#undef FTB_NP
#undef TRIAL
    }
    if (st == 0) {
        /* back to the regular algorithm: bound multipliers kept unless one exceeds bound_mult_reset_threshold (then all 1); equality
         * multipliers zero (constr_mult_reset_threshold 0) */
        double zm = fmax(vmaxabs(wk->zL, n), fmax(vmaxabs(wk->zU, n), ns ? vmaxabs(wk->vL, ns) : 0.0));
        if (zm > BOUND_MULT_RESET) {
            for (int i = 0; i < n; ++i) { wk->zL[i] = wk->hasL[i] ? 1.0 : 0.0; wk->zU[i] = wk->hasU[i] ? 1.0 : 0.0; }
            for (int i = 0; i < ns; ++i) wk->vL[i] = 1.0;
        }
        memset(wk->yc, 0, sizeof(double) * mc); if (ns) memset(wk->yd, 0, sizeof(double) * ns);
    }
    wk->resto = 0; wk->linsolve = saved_linsolve;
    free(F); free(pool);
    return st;
}

/* ------------------------------------------------------------------ */
/* one NLP                                                             */
/* ------------------------------------------------------------------ */
typedef struct {
    int cap, len;
    double *rows; /* [cap][8]: mu, alpha_pr, alpha_du, delta_w, theta, phi, E0, f */
} trace_t;

static void solve_one(const kmo_config *cf, work_t *wk, const double *xcur, const double *goal, const double *X0,
                      const double *U0, const double *obs, const double *orad, double *Xout, double *Uout, double *duals_out,
                      double *obj, int32_t *status, int32_t *iters, kmo_diag *dg, trace_t *tr) {
    int N = wk->N, O = wk->O, n = wk->n, ns = wk->ns, mc = wk->mc;
    memcpy(wk->xcur, xcur, 3 * sizeof(double)); memcpy(wk->goal, goal, 3 * sizeof(double));
    if (O) memcpy(wk->obs, obs, sizeof(double) * 2 * O * (wk->obs_sw ? N : 1));
    for (int o = 0; o < O; ++o) wk->orad[o] = orad ? orad[o] : cf->obs_radius;
    kmo_diag dloc; memset(&dloc, 0, sizeof dloc);

    /* bounds, relaxed (optimizer.py:111-156 + IPOPT bound_relax_factor) */
    for (int i = 0; i < n; ++i) {
        int t = i < 3 * (N + 1) ? i % 3 : 3 + (i - 3 * (N + 1)) % 2;
        double lo = cf->lo[t], hi = cf->hi[t];
        wk->hasL[i] = lo > NLP_LOWER_INF; wk->hasU[i] = hi < NLP_UPPER_INF;
        wk->lb[i] = wk->hasL[i] ? lo - BOUND_RELAX * fmax(1.0, fabs(lo)) : -INFINITY;
        wk->ub[i] = wk->hasU[i] ? hi + BOUND_RELAX * fmax(1.0, fabs(hi)) : INFINITY;
    }
    wk->dL = cf->inflation - BOUND_RELAX * fmax(1.0, fabs(cf->inflation));

    /* starting point (optimizer.py:375-385; agent.py:59-60 for the cold start) */
    for (int k = 0; k <= N; ++k) for (int j = 0; j < 3; ++j) wk->w[IX(k, j)] = X0 ? X0[j * (N + 1) + k] : xcur[j];
    for (int k = 0; k < N; ++k) for (int j = 0; j < 2; ++j) wk->w[IU(k, j)] = U0 ? U0[j * N + k] : 0.0;

    /* gradient-based objective scaling at the user's starting point */
    eval_grad(cf, wk, wk->w, 1.0, wk->g);
    { double gm = vmaxabs(wk->g, n); wk->df = gm > SCALING_MAX_GRAD ? fmax(SCALING_MAX_GRAD / gm, SCALING_MIN_VALUE) : 1.0; }
    dloc.obj_scaling = wk->df;

    /* push into the interior */
    for (int i = 0; i < n; ++i) {
        if (wk->hasL[i] && wk->hasU[i]) {
            double pl = fmin(BOUND_PUSH * fmax(1.0, fabs(wk->lb[i])), BOUND_FRAC * (wk->ub[i] - wk->lb[i]));
            double pu = fmin(BOUND_PUSH * fmax(1.0, fabs(wk->ub[i])), BOUND_FRAC * (wk->ub[i] - wk->lb[i]));
            wk->w[i] = fmin(fmax(wk->w[i], wk->lb[i] + pl), wk->ub[i] - pu);
        } else if (wk->hasL[i]) wk->w[i] = fmax(wk->w[i], wk->lb[i] + BOUND_PUSH * fmax(1.0, fabs(wk->lb[i])));
        else if (wk->hasU[i]) wk->w[i] = fmin(wk->w[i], wk->ub[i] - BOUND_PUSH * fmax(1.0, fabs(wk->ub[i])));
        wk->zL[i] = wk->hasL[i] ? 1.0 : 0.0; wk->zU[i] = wk->hasU[i] ? 1.0 : 0.0;
    }
    if (ns) {
        eval_d(cf, wk, wk->w, wk->s, NULL, NULL);
        for (int i = 0; i < ns; ++i) { wk->s[i] = fmax(wk->s[i], wk->dL + BOUND_PUSH * fmax(1.0, fabs(wk->dL))); wk->vL[i] = 1.0; }
    }
    memset(wk->yc, 0, sizeof(double) * mc);
    if (ns) memset(wk->yd, 0, sizeof(double) * ns);

    int st = ST_SUCCESS, iter = 0;
    double mu = MU_INIT, tau = fmax(TAU_MIN, 1.0 - MU_INIT);
    double delta_last = 0.0;
    errs_t er; double *rx = wk->dx2, *rs = wk->ds2; /* scratch until the step solve */

    /* least-squares multiplier estimate */
    {
        eval_errors(cf, wk, &er, rx, rs); /* with y = 0: rx = grad f - zL + zU ; rs = -vL */
        for (int i = 0; i < n; ++i) wk->Dx[i] = 1.0;
        for (int i = 0; i < ns; ++i) wk->Ds[i] = 1.0;
        memset(wk->bc, 0, sizeof(double) * mc); if (ns) memset(wk->bd, 0, sizeof(double) * ns);
        int ok = kkt_factor(cf, wk, wk->w, 0); dloc.n_factor++;
        if (ok) {
            kkt_solve(cf, wk, wk->w, 0, rx, rs, wk->bc, wk->bd, wk->dx, wk->ds, wk->dyc, wk->dyd);
            double ym = fmax(vmaxabs(wk->dyc, mc), ns ? vmaxabs(wk->dyd, ns) : 0.0);
            if (ym <= CONSTR_MULT_INIT_MAX && isfinite(ym)) {
                for (int i = 0; i < mc; ++i) wk->yc[i] = -wk->dyc[i];
                for (int i = 0; i < ns; ++i) wk->yd[i] = -wk->dyd[i];
            }
        }
    }

    filter_t *F = (filter_t *)xcalloc(1, sizeof(filter_t));
    double theta_max = -1, theta_min = -1;
    double E0 = 0;

    for (;;) {
        eval_errors(cf, wk, &er, rx, rs);
        E0 = opt_error(wk, &er, 0.0);
        if (tr && tr->len < tr->cap) {
            double *r = tr->rows + 8 * tr->len; r[0] = mu; r[6] = E0; r[7] = eval_f(cf, wk, wk->w) / wk->df;
            r[4] = vsumabs(wk->c, mc) + (ns ? vsumabs(wk->dms, ns) : 0.0);
        }
        /* IPOPT checks every evaluated quantity for non-finite numbers; the max-norms carry a NaN / inf through, the scaled max may lose it */
        if (!isfinite(E0) || !isfinite(er.dual_inf) || !isfinite(er.primal_inf)) { st = ST_INVALID_NUMBER; break; }
        /* convergence: scaled E_0 <= tol and the unscaled safeguards (objective scaling undone) */
        if (E0 <= cf->tol && er.dual_inf / wk->df <= DUAL_INF_TOL && er.primal_inf <= CONSTR_VIOL_TOL &&
            compl_inf(&er, 0.0) / wk->df <= COMPL_INF_TOL) { st = ST_SUCCESS; break; }
        if (iter >= cf->max_iter) { st = ST_MAXITER; break; }
        if (vmaxabs(wk->w, n) > DIVERGING_TOL) { st = ST_DIVERGING; break; }

        /* monotone barrier update */
        {
            int done = 0;
            while (!done && opt_error(wk, &er, mu) <= KAPPA_EPS * mu) {
                double nm = fmax(fmin(MU_LIN * mu, pow(mu, MU_SUPER)), fmin(cf->tol, COMPL_INF_TOL) / (KAPPA_EPS + 1.0));
                int changed = nm != mu;
                mu = nm; tau = fmax(TAU_MIN, 1.0 - mu);
                if (changed) F->n = 0; else done = 1;
            }
        }

        /* search direction with inertia correction */
        eval_hess(cf, wk, wk->w, wk->yc, wk->yd, wk->cs, wk->sn);
        double delta = 0.0; int ok = 0;
        for (;;) {
            build_rhs(cf, wk, mu, delta, wk->c, wk->dms);
            ok = kkt_factor(cf, wk, wk->w, 1); dloc.n_factor++;
            if (ok) break;
            if (delta == 0.0) delta = delta_last == 0.0 ? DELTA_W_INIT : fmax(DELTA_W_MIN, delta_last * DELTA_W_DEC);
            else delta = (delta_last == 0.0 || 1e5 * delta_last < delta) ? DELTA_W_INC_FIRST * delta : DELTA_W_INC * delta;
            if (delta > DELTA_W_MAX) break;
        }
        if (!ok) { st = ST_STEP_ERROR; break; }
        if (delta > 0.0) { delta_last = delta; if (delta > dloc.max_delta_w) dloc.max_delta_w = delta; }
        kkt_solve(cf, wk, wk->w, 1, wk->bx, wk->bs, wk->bc, wk->bd, wk->dx, wk->ds, wk->dyc, wk->dyd);

        /* line search */
        merit_t cur;
        if (!eval_merit(cf, wk, wk->w, wk->s, mu, wk->ct, wk->dmst, &cur)) { st = ST_INVALID_NUMBER; break; }
        double gBD = 0;
        for (int i = 0; i < n; ++i) {
            double gp = wk->g[i];
            if (wk->hasL[i]) { gp -= mu / (wk->w[i] - wk->lb[i]); if (!wk->hasU[i]) gp += KAPPA_D * mu; }
            if (wk->hasU[i]) { gp += mu / (wk->ub[i] - wk->w[i]); if (!wk->hasL[i]) gp -= KAPPA_D * mu; }
            gBD += gp * wk->dx[i];
        }
        for (int i = 0; i < ns; ++i) gBD += (-mu / (wk->s[i] - wk->dL) + KAPPA_D * mu) * wk->ds[i];
        if (theta_max < 0) { theta_max = THETA_MAX_FACT * fmax(1.0, cur.theta); theta_min = THETA_MIN_FACT * fmax(1.0, cur.theta); }
        double alpha_min = GAMMA_THETA;
        if (gBD < 0) {
            alpha_min = fmin(GAMMA_THETA, GAMMA_PHI * cur.theta / (-gBD));
            if (cur.theta <= theta_min) alpha_min = fmin(alpha_min, DELTA_LS * pow(cur.theta, S_THETA) / pow(-gBD, S_PHI));
        }
        alpha_min *= ALPHA_MIN_FRAC;
        double alpha_max = ftb_primal(wk, wk->dx, wk->ds, tau);
        double alpha = alpha_max, alpha_test = alpha_max;
        const double *sdx = wk->dx, *sds = wk->ds, *sdyc = wk->dyc, *sdyd = wk->dyd; /* the step finally taken */
        int accept = 0, nsteps = 0;
        merit_t tri;
#define IS_FTYPE(a) (gBD < 0 && (a) * pow(-gBD, S_PHI) > DELTA_LS * pow(cur.theta, S_THETA))
#define ARMIJO(a, t) compare_le((t).phi - cur.phi, ETA_PHI * (a) * gBD, cur.phi)
        while (alpha > alpha_min || nsteps == 0) {
            for (int i = 0; i < n; ++i) wk->wt[i] = wk->w[i] + alpha * wk->dx[i];
            for (int i = 0; i < ns; ++i) wk->st[i] = wk->s[i] + alpha * wk->ds[i];
            int evok = eval_merit(cf, wk, wk->wt, wk->st, mu, wk->ct, wk->dmst, &tri); dloc.n_trials++;
            alpha_test = alpha;
            if (evok) {
                /* FilterLSAcceptor::CheckAcceptabilityOfTrialPoint */
                int acc;
                if (tri.theta > theta_max) acc = 0;
                else {
                    if (IS_FTYPE(alpha_test) && cur.theta <= theta_min) acc = ARMIJO(alpha_test, tri);
                    else {
                        acc = 1;
                        if (tri.phi > cur.phi) { double bas = fabs(cur.phi) > 10.0 ? log10(fabs(cur.phi)) : 1.0; if (log10(tri.phi - cur.phi) > OBJ_MAX_INC + bas) acc = 0; }
                        if (acc) acc = compare_le(tri.theta, (1.0 - GAMMA_THETA) * cur.theta, cur.theta) || compare_le(tri.phi - cur.phi, -GAMMA_PHI * cur.theta, cur.phi);
                    }
                    if (acc) acc = filter_ok(F, tri.theta, tri.phi);
                }
                accept = acc;
            }
            if (accept) break;
            /* second-order correction on the first trial point if the violation got worse */
            if (evok && alpha == alpha_max && cur.theta <= tri.theta) {
                double theta_soc_old = 0, theta_trial = tri.theta, alpha_soc = alpha;
                memcpy(wk->csoc, wk->c, sizeof(double) * mc); if (ns) memcpy(wk->dsoc, wk->dms, sizeof(double) * ns);
                int count = 0;
                while (count < MAX_SOC && !accept && (count == 0 || theta_trial <= KAPPA_SOC * theta_soc_old)) {
                    theta_soc_old = theta_trial;
                    for (int i = 0; i < mc; ++i) wk->csoc[i] = alpha_soc * wk->csoc[i] + wk->ct[i];
                    for (int i = 0; i < ns; ++i) wk->dsoc[i] = alpha_soc * wk->dsoc[i] + wk->dmst[i];
                    for (int i = 0; i < mc; ++i) wk->bc[i] = -wk->csoc[i];
                    for (int i = 0; i < ns; ++i) wk->bd[i] = -wk->dsoc[i];
                    kkt_solve(cf, wk, wk->w, 1, wk->bx, wk->bs, wk->bc, wk->bd, wk->dx2, wk->ds2, wk->dyc2, wk->dyd2); dloc.n_soc++;
                    alpha_soc = ftb_primal(wk, wk->dx2, wk->ds2, tau);
                    for (int i = 0; i < n; ++i) wk->wt[i] = wk->w[i] + alpha_soc * wk->dx2[i];
                    for (int i = 0; i < ns; ++i) wk->st[i] = wk->s[i] + alpha_soc * wk->ds2[i];
                    merit_t ts; int e2 = eval_merit(cf, wk, wk->wt, wk->st, mu, wk->ct, wk->dmst, &ts); dloc.n_trials++;
                    if (!e2) break;
                    int acc;
                    if (ts.theta > theta_max) acc = 0;
                    else {
                        if (IS_FTYPE(alpha_test) && cur.theta <= theta_min) acc = ARMIJO(alpha_test, ts);
                        else {
                            acc = 1;
                            if (ts.phi > cur.phi) { double bas = fabs(cur.phi) > 10.0 ? log10(fabs(cur.phi)) : 1.0; if (log10(ts.phi - cur.phi) > OBJ_MAX_INC + bas) acc = 0; }
                            if (acc) acc = compare_le(ts.theta, (1.0 - GAMMA_THETA) * cur.theta, cur.theta) || compare_le(ts.phi - cur.phi, -GAMMA_PHI * cur.theta, cur.phi);
                        }
                        if (acc) acc = filter_ok(F, ts.theta, ts.phi);
                    }
                    if (acc) { accept = 1; tri = ts; alpha = alpha_soc; sdx = wk->dx2; sds = wk->ds2; sdyc = wk->dyc2; sdyd = wk->dyd2; }
                    else { count++; theta_trial = ts.theta; }
                }
                if (accept) break;
            }
            alpha *= ALPHA_RED; nsteps++;
        }
        if (!accept) {
            /* BacktrackingLineSearch: the step size fell below alpha_min -> restoration phase, unless the point is almost feasible */
            if (cur.theta <= 1e-2 * cf->tol) { st = ST_RESTORATION; break; }   /* "Restoration phase called, but point is almost feasible" */
            filter_add(F, (1.0 - GAMMA_THETA) * cur.theta, cur.phi - GAMMA_PHI * cur.theta);   /* PrepareRestoPhaseStart */
            if (F->n > dloc.max_filter) dloc.max_filter = F->n;
            const int rs_ = restoration(cf, wk, mu, F, cur.theta, cur.phi, &iter, &dloc);
            dloc.n_resto++;
            if (rs_ != 0) { st = rs_; break; }
            iter++;
            continue;
        }

        /* filter augmentation (FilterLSAcceptor::UpdateForNextIteration) */
        if (!IS_FTYPE(alpha_test) || !ARMIJO(alpha_test, tri)) {
            filter_add(F, (1.0 - GAMMA_THETA) * cur.theta, cur.phi - GAMMA_PHI * cur.theta);
            if (F->n > dloc.max_filter) dloc.max_filter = F->n;
        }
#undef IS_FTYPE
#undef ARMIJO
        /* accept the trial point; duals: y with the primal step size, z with its own fraction-to-the-boundary */
        double alpha_du = dual_steps(wk, sdx, sds, mu, tau);
        for (int i = 0; i < n; ++i) wk->w[i] += alpha * sdx[i];
        for (int i = 0; i < ns; ++i) wk->s[i] += alpha * sds[i];
        for (int i = 0; i < mc; ++i) wk->yc[i] += alpha * sdyc[i];
        for (int i = 0; i < ns; ++i) wk->yd[i] += alpha * sdyd[i];
        for (int i = 0; i < n; ++i) {
            if (wk->hasL[i]) { double sl = wk->w[i] - wk->lb[i], z = wk->zL[i] + alpha_du * wk->dzL[i]; wk->zL[i] = fmax(fmin(z, KAPPA_SIGMA * mu / sl), mu / (KAPPA_SIGMA * sl)); }
            if (wk->hasU[i]) { double su = wk->ub[i] - wk->w[i], z = wk->zU[i] + alpha_du * wk->dzU[i]; wk->zU[i] = fmax(fmin(z, KAPPA_SIGMA * mu / su), mu / (KAPPA_SIGMA * su)); }
        }
        for (int i = 0; i < ns; ++i) { double sl = wk->s[i] - wk->dL, z = wk->vL[i] + alpha_du * wk->dvL[i]; wk->vL[i] = fmax(fmin(z, KAPPA_SIGMA * mu / sl), mu / (KAPPA_SIGMA * sl)); }
        if (tr && tr->len < tr->cap) { double *r = tr->rows + 8 * tr->len; r[1] = alpha; r[2] = alpha_du; r[3] = delta; r[5] = cur.phi; tr->len++; }
        iter++;
    }
    if (tr && tr->len < tr->cap) tr->len++;

    for (int k = 0; k <= N; ++k) for (int j = 0; j < 3; ++j) Xout[j * (N + 1) + k] = wk->w[IX(k, j)];
    for (int k = 0; k < N; ++k) for (int j = 0; j < 2; ++j) Uout[j * N + k] = wk->w[IU(k, j)];
    if (duals_out) { /* [yc(mc) | zL(n) | zU(n) | s(ns) | yd(ns) | vL(ns)] of the SCALED problem */
        double *p = duals_out;
        memcpy(p, wk->yc, sizeof(double) * mc); p += mc; memcpy(p, wk->zL, sizeof(double) * n); p += n;
        memcpy(p, wk->zU, sizeof(double) * n); p += n;
        if (ns) { memcpy(p, wk->s, sizeof(double) * ns); p += ns; memcpy(p, wk->yd, sizeof(double) * ns); p += ns; memcpy(p, wk->vL, sizeof(double) * ns); }
    }
    *obj = eval_f(cf, wk, wk->w) / wk->df; *status = st; *iters = iter;
    dloc.mu = mu; dloc.err = E0;
    if (dg) *dg = dloc;
    free(F);
}

/* ------------------------------------------------------------------ */
/* exported entry points (ctypes)                                      */
/* ------------------------------------------------------------------ */
int kmo_version(void) { return KMO_VERSION; }
int kmo_duals_len(const kmo_config *cf) { int N = cf->N, O = cf->O; return 3 * (N + 1) + 2 * (5 * N + 3) + 3 * N * O; }

/* Layouts (row-major, per-instance contiguous = numpy C order):
 *   x_cur[B][3], goal[B][3], X0[B][3][N+1] or NULL, U0[B][2][N] or NULL, obs[B][O][2] (obs_stagewise: [B][O][N][2]) or NULL,
 *   obs_rad[B][O] or NULL (NULL: cf->obs_radius for every obstacle),
 *   X_out[B][3][N+1], U_out[B][2][N], obj[B], status[B], iters[B], duals[B][kmo_duals_len] or NULL, diag[B] or NULL */
int kmo_solve(const kmo_config *cf, int B, const double *x_cur, const double *goal, const double *X0, const double *U0,
              const double *obs, const double *obs_rad, double *X_out, double *U_out, double *obj, int32_t *status, int32_t *iters,
              double *duals, kmo_diag *diag, int nthreads) {
    if (!cf || cf->N < 1 || cf->O < 0 || B < 0) return -1;
    if (cf->O > 0 && !obs) return -1;
    int N = cf->N, O = cf->O, dl = kmo_duals_len(cf);
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
    {
        work_t *wk = work_new(cf);
#pragma omp for schedule(dynamic, 1)
        for (int b = 0; b < B; ++b) {
            solve_one(cf, wk, x_cur + 3 * (size_t)b, goal + 3 * (size_t)b, X0 ? X0 + (size_t)b * 3 * (N + 1) : NULL,
                      U0 ? U0 + (size_t)b * 2 * N : NULL, O ? obs + (size_t)b * 2 * O * (cf->obs_stagewise ? N : 1) : NULL,
                      (O && obs_rad) ? obs_rad + (size_t)b * O : NULL, X_out + (size_t)b * 3 * (N + 1), U_out + (size_t)b * 2 * N, duals ? duals + (size_t)b * dl : NULL,
                      obj + b, status + b, iters + b, diag ? diag + b : NULL, NULL);
        }
        work_free(wk);
    }
    return 0;
}

/* single instance with a per-iteration trace: rows[cap][8] = mu, alpha_pr, alpha_du, delta_w, theta, phi, E0, f */
int kmo_solve_trace(const kmo_config *cf, const double *x_cur, const double *goal, const double *X0, const double *U0,
                    const double *obs, const double *obs_rad, double *X_out, double *U_out, double *obj, int32_t *status, int32_t *iters,
                    double *rows, int cap, int32_t *len) {
    if (!cf || cf->N < 1) return -1;
    work_t *wk = work_new(cf);
    trace_t tr = {cap, 0, rows};
    solve_one(cf, wk, x_cur, goal, X0, U0, obs, obs_rad, X_out, U_out, NULL, obj, status, iters, NULL, &tr);
    *len = tr.len;
    work_free(wk);
    return 0;
}
