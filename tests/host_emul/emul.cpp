// TEST HARNESS ONLY -- never built into or loaded by the product package.
// Compiles kiss_mpc_b200/csrc/kmpc_core.cuh (the per-thread solver the CUDA kernel runs) with g++ so that the
// algorithm can be checked against the oracle on a machine without a GPU.  One "slot" is executed at a time; the
// structure-of-arrays workspace indexing (row * S + slot) is exercised with S > 1.
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <vector>
#include "../../include/kmpc.h"
#include "../../kiss_mpc_b200/csrc/kmpc_core.cuh"
#include "../../kiss_mpc_b200/csrc/kmpc_warp.cuh"

using namespace kmpc;

// ---- fibre emulator of a block of warps (see simt.h) ----
namespace kmpc {
thread_local Simt *g_simt = nullptr;
static void simt_entry() {
    Simt &s = *g_simt;
    s.fn();
    // this lane is finished: hand over to the next unfinished lane, the last one returns to the caller
    const int me = s.cur;
    if (me < 32 * s.nw - 1) { s.cur = me + 1; setcontext(&s.ctx[me + 1]); }
    setcontext(&s.main);
}
void simt_run(const std::function<void()> &fn, int warps) {
    Simt *sp = new Simt;
    Simt &s = *sp;
    g_simt = sp;
    s.fn = fn;
    s.nw = warps;
    for (int i = 0; i < 32 * warps; ++i) {
        s.parity[i] = 0; s.bparity[i] = 0;
        s.stack[i].resize(1 << 20);
        getcontext(&s.ctx[i]);
        s.ctx[i].uc_stack.ss_sp = s.stack[i].data();
        s.ctx[i].uc_stack.ss_size = s.stack[i].size();
        s.ctx[i].uc_link = nullptr;
        makecontext(&s.ctx[i], (void (*)())simt_entry, 0);
    }
    s.cur = 0;
    swapcontext(&s.main, &s.ctx[0]);
    g_simt = nullptr;
    delete sp;
}
}  // namespace kmpc

static Cfg make_cfg(const kmpc_config *cf, int B, int O, int stagewise, double obs_radius, double inflation) {
    Cfg c;
    memset(&c, 0, sizeof c);
    c.N = cf->N; c.O = O; c.cost_mode = cf->cost_mode; c.gk_lo = cf->goal_k_lo; c.gk_hi = cf->goal_k_hi;
    c.max_iter = cf->max_iter; c.layout = cf->layout; c.B = B;
    for (int i = 0; i < 4; ++i) {
        c.hasL[i] = cf->lo[i] > -KMPC_NO_BOUND; c.hasU[i] = cf->hi[i] < KMPC_NO_BOUND;
        c.lb[i] = c.hasL[i] ? cf->lo[i] - K_BOUND_RELAX * fmax(1.0, fabs(cf->lo[i])) : -INFINITY;
        c.ub[i] = c.hasU[i] ? cf->hi[i] + K_BOUND_RELAX * fmax(1.0, fabs(cf->hi[i])) : INFINITY;
    }
    c.T = cf->T; c.W[0] = cf->W[0]; c.W[1] = cf->W[1]; c.W[2] = cf->W[2];
    c.Wvn = cf->Wv_neg; c.Wvp = cf->Wv_pos; c.Ww = cf->Ww; c.tol = cf->tol;
    c.obs_radius = obs_radius; c.dL = inflation - K_BOUND_RELAX * fmax(1.0, fabs(inflation));
    c.obs_sw = (stagewise && O > 0) ? 1 : 0;
    c.L = make_rows(cf->N, O, c.obs_sw);
    c.nb = (cf->N + 1) * (c.hasL[0] + c.hasU[0] + c.hasL[1] + c.hasU[1]) + cf->N * (c.hasL[2] + c.hasU[2] + c.hasL[3] + c.hasU[3]) + cf->N * O;
    c.m = 3 * (cf->N + 1) + cf->N * O;
    c.r_mnb = 1.0 / (double)(c.m + c.nb); c.r_nb = c.nb ? 1.0 / (double)c.nb : 0.0; c.mu_floor = cfg_mu_floor(c.tol);
    return c;
}

extern "C" int emul_solve(const kmpc_config *cf, int B, const double *x_cur, const double *goal, const double *X0,
                          const double *U0, const double *obs, const double *obs_rad, int O, int stagewise, double obs_radius, double inflation, double *X_out,
                          double *U_out, double *obj, int32_t *status, int32_t *iters, int32_t *trips) {
    Cfg c = make_cfg(cf, B, O, stagewise, obs_radius, inflation);
    IO io;
    memset(&io, 0, sizeof io);
    io.x_cur = x_cur; io.goal = goal; io.X0 = X0; io.U0 = U0; io.obs = obs; io.orad = O > 0 ? obs_rad : NULL;
    io.X_out = X_out; io.U_out = U_out; io.obj = obj; io.status = status; io.iters = iters; io.active = NULL; io.wscratch = NULL; io.order = NULL;
    // Mirrors the launch structure of kmpc.cu on the host: per trip, sweep(LA[p]) -> rollout(LT[p]) -> trial(LT[p]), with the
    // solver context stored in / reloaded from the workspace between the phases exactly as the kernels do.
    const size_t S = (size_t)((B + 31) / 32 * 32);
    const size_t rows_total = (size_t)make_resto_rows(c.L).total;   // solver rows + the rows of the restoration phase
    double *ws = (double *)malloc(sizeof(double) * S * rows_total);
    for (size_t i = 0; i < S * rows_total; ++i) ws[i] = NAN;  // poison: reads of never-written rows show up
    std::vector<int> LA[2], LT[2];
    for (int b = 0; b < B; ++b) {
        Ctx t; memset(&t, 0, sizeof t); t.inst = b;
        pass_init(c, t, ws + b, S, io);
        ctx_store(t, c.L, ws + b, S);
        LA[0].push_back(b);
    }
    for (int tr = 0; !(LA[tr & 1].empty() && LT[tr & 1].empty()); ++tr) {
        const int p = tr & 1;
        std::vector<int> outcome(LA[p].size());
#pragma omp parallel for schedule(dynamic, 8)
        for (size_t i = 0; i < LA[p].size(); ++i) {
            const int b = LA[p][i]; double *wsp = ws + b;
            Ctx t; ctx_load(t, c.L, wsp, S); t.inst = b;
            const int r = O > 0 ? phase_sweep<true>(c, t, wsp, S) : phase_sweep<false>(c, t, wsp, S);
            double *pw = wsp + (size_t)c.L.rCtx * S;
            pw[(size_t)X_TRIPS * S] = t.trips; pw[(size_t)X_DELTA * S] = t.delta;
            if (r != 100 && r != 101) { pass_output(c, t, wsp, S, io, r); if (trips) trips[b] = t.trips; }
            outcome[i] = r;
        }
        for (size_t i = 0; i < LA[p].size(); ++i) {
            if (outcome[i] == 100) LT[p].push_back(LA[p][i]);
            else if (outcome[i] == 101) LA[1 - p].push_back(LA[p][i]);
        }
#pragma omp parallel for schedule(dynamic, 8)
        for (size_t i = 0; i < LT[p].size(); ++i) {
            const int b = LT[p][i]; double *wsp = ws + b;
            Ctx t; ctx_load(t, c.L, wsp, S);
            if (t.mode == M_TRIAL) continue;
            t.inst = b;
            if (O > 0) phase_rollout<true>(c, t, wsp, S); else phase_rollout<false>(c, t, wsp, S);
            ctx_store(t, c.L, wsp, S);
        }
        std::vector<int> out2(LT[p].size());
#pragma omp parallel for schedule(dynamic, 8)
        for (size_t i = 0; i < LT[p].size(); ++i) {
            const int b = LT[p][i]; double *wsp = ws + b;
            Ctx t; ctx_load(t, c.L, wsp, S); t.inst = b;
            const int r = O > 0 ? phase_trial<true>(c, t, wsp, S) : phase_trial<false>(c, t, wsp, S);
            if (r == 100) { ctx_store(t, c.L, wsp, S); out2[i] = t.mode == M_TRIAL ? 1 : 2; }
            else { pass_output(c, t, wsp, S, io, r); if (trips) trips[b] = t.trips; out2[i] = 0; }
        }
        for (size_t i = 0; i < LT[p].size(); ++i) {
            if (out2[i] == 1) LT[1 - p].push_back(LT[p][i]);
            else if (out2[i] == 2) LA[1 - p].push_back(LT[p][i]);
        }
        LA[p].clear(); LT[p].clear();
    }
    free(ws);
    return 0;
}

// warp-per-instance solver (kmpc_warp.cuh) on the fibre emulator; N + 1 <= 64.
// warps <= 1: every instance on its own one-warp block.  warps > 1: blocks of `warps` warps, block j working through the
// instances [j * chunk, (j + 1) * chunk) from its own queue -- the block-level machinery of the kernel (the Riccati warp
// serving the other warps' instances, the speculative inertia candidates, refills in the serial window, the tail mode with
// borrowed instance slots) runs exactly as on the device, only one fibre at a time.
extern "C" int emul_solve_warp(const kmpc_config *cf, int B, const double *x_cur, const double *goal, const double *X0,
                               const double *U0, const double *obs, const double *obs_rad, int O, int stagewise, double obs_radius, double inflation, double *X_out,
                               double *U_out, double *obj, int32_t *status, int32_t *iters, int32_t *trips, int warps, int chunk) {
    if (cf->N + 1 > 64) return -1;
    if (warps > SIMT_MAX_WARPS) return -1;
    Cfg c = make_cfg(cf, B, O, stagewise, obs_radius, inflation);
    IO io;
    memset(&io, 0, sizeof io);
    io.x_cur = x_cur; io.goal = goal; io.X0 = X0; io.U0 = U0; io.obs = obs; io.orad = O > 0 ? obs_rad : NULL;
    io.X_out = X_out; io.U_out = U_out; io.obj = obj; io.status = status; io.iters = iters; io.active = NULL; io.wscratch = NULL; io.order = NULL;
    const int spl = cf->N + 1 <= 32 ? 1 : 2;
    const int W = warps > 1 ? warps : 1;
    if (W == 1) chunk = 1;
    if (chunk < 1) chunk = B;
    const int nblk = (B + chunk - 1) / chunk;
    unsigned long long tr_total = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int j = 0; j < nblk; ++j) {
        const int lo = j * chunk, hi = lo + chunk < B ? lo + chunk : B;
        std::vector<double> smem((spl == 1 ? WLay<1, 32>::bytes(W, O, c.obs_sw) : cf->N + 1 <= 52 ? WLay<2, 52>::bytes(W, O, c.obs_sw) : WLay<2, 64>::bytes(W, O, c.obs_sw)) / sizeof(double) + 8, NAN);
        unsigned long long tr = 0;
        int queue = lo;          // this emulated block is handed exactly the instances lo .. hi - 1
        Cfg cb = c; cb.B = hi;
        std::vector<double> gscr((size_t)(G_NF * 64 + 2 * K_FILTER_CAP) * W, NAN);   // the block's global scratch slots (>= WLay::GPRIV each)
        IO iob = io; iob.wscratch = gscr.data();
        // hand-over workspace of the restoration phase: a column per instance of this block, run by finish_instance below
        const int rrows = make_resto_rows(c.L).total;
        std::vector<double> rws((size_t)rrows * (hi - lo), NAN);
        std::vector<int32_t> rlist(hi - lo, -1);
        int rcount = 0;
        iob.resto_ws = rws.data(); iob.resto_list = rlist.data(); iob.resto_count = &rcount; iob.resto_cap = hi - lo; iob.resto_rows = rrows;
        bool full = true;
        for (int i = 0; i < 4; ++i) full = full && c.hasL[i] && c.hasU[i];
        const int nst = spl == 1 ? 32 : cf->N + 1 <= 52 ? 52 : 64;
#define EMUL_RUN(SPL, NST) do { \
            if (O > 0) { if (full) w_worker<SPL, NST, true, true>(cb, iob, smem.data(), &queue, &tr); else w_worker<SPL, NST, false, true>(cb, iob, smem.data(), &queue, &tr); } \
            else { if (full) w_worker<SPL, NST, true, false>(cb, iob, smem.data(), &queue, &tr); else w_worker<SPL, NST, false, false>(cb, iob, smem.data(), &queue, &tr); } } while (0)
        simt_run([&]() {
            if (nst == 32) EMUL_RUN(1, 32); else if (nst == 52) EMUL_RUN(2, 52); else EMUL_RUN(2, 64);
        }, W);
#undef EMUL_RUN
        for (int i = 0; i < rcount; ++i) { double *col = iob.resto_ws + (size_t)i * iob.resto_rows; if (O > 0) finish_instance<true>(cb, iob, i, col); else finish_instance<false>(cb, iob, i, col); }   // kmpc_finish_kernel
        if (W == 1 && trips) trips[lo] = (int)tr;
#pragma omp atomic
        tr_total += tr;
    }
    if (W > 1 && trips) trips[0] = (int)tr_total;
    return 0;
}
