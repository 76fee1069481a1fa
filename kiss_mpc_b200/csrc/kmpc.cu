// kmpc.cu -- CUDA kernels (sm_100a) and the C ABI of include/kmpc.h.
//
// Replaces the numerical core of mpc/optimizer.py:319-400 (MotionPlanner.solve -> CasADi/IPOPT) for B instances at
// once.  One persistent CUDA thread per problem instance; per-instance state lives in a structure-of-arrays HBM
// workspace (see kmpc_core.cuh); finished lanes pull the next instance from a global work counter, so a warp keeps
// all 32 lanes busy although iteration counts differ by 5x between instances.  No tensor cores: the stage blocks are
// 3x3 / 2x2 / 2x3 and the work is FP64 FMA + HBM streaming.  No CPU fallback exists in this library.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/kmpc.h"
#include "kmpc_core.cuh"

using namespace kmpc;

#define KMPC_TPB 128

// ------------------------------------------------------------------------------------------------
// The solver kernel: grid of resident threads; thread "slot" owns workspace column `slot`.
// counter starts at the number of launched threads: the first instance of a slot is b = slot (coalesced I/O for the
// batch-minor layout), later ones come from atomicAdd.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(KMPC_TPB, 2)
kmpc_ipm_kernel(const Cfg c, const IO io, double *__restrict__ ws, const size_t S, int *__restrict__ counter,
                unsigned long long *__restrict__ trips_total) {
    const size_t slot = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    double *wsp = ws + slot;
    Ctx t;
    t.mode = M_FETCH; t.trips = 0; t.inst = -1;
    bool first = true;
    for (;;) {
        if (t.mode == M_FETCH) {
            const int b = first ? (int)slot : atomicAdd(counter, 1);
            first = false;
            if (b < c.B) { t.inst = b; pass_init(c, t, wsp, S, io); }
            else t.mode = M_DONE;
        }
        __syncwarp();
        if (__all_sync(0xffffffffu, t.mode == M_DONE)) break;
        if (t.mode != M_DONE) {
            const int r = trip(c, t, wsp, S);
            if (r != 100) { pass_output(c, t, wsp, S, io, r); t.mode = M_FETCH; }
        }
    }
    if (trips_total) {
        unsigned long long v = (unsigned long long)t.trips;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(trips_total, v);
    }
}

// Batched EgoAgent.step hand-off (agent.py:139-155, :70-72): applied control = U[:,0]; next current state = X[:,1].
__global__ void kmpc_handoff_kernel(int B, int N, int layout, const double *__restrict__ X, const double *__restrict__ U,
                                    double *__restrict__ x_cur, double *__restrict__ applied) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    for (int j = 0; j < 3; ++j) {
        const size_t src = layout ? ((size_t)j * (N + 1) + 1) * B + b : ((size_t)b * 3 + j) * (N + 1) + 1;
        const size_t dst = layout ? (size_t)j * B + b : (size_t)b * 3 + j;
        x_cur[dst] = X[src];
    }
    if (applied)
        for (int j = 0; j < 2; ++j) {
            const size_t src = layout ? ((size_t)j * N) * B + b : ((size_t)b * 2 + j) * N;
            const size_t dst = layout ? (size_t)j * B + b : (size_t)b * 2 + j;
            applied[dst] = U[src];
        }
}

// FP64 FMA throughput micro-benchmark: 8 independent DFMA chains per thread.
__global__ void kmpc_dfma_kernel(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct kmpc_handle {
    kmpc_config cfg;
    Rows rows;
    int device, sm_count, blocks, slots;
    double *ws;
    int *counter;
    unsigned long long *trips;
    cudaEvent_t ev0, ev1;
    int timing;
    double last_ms;
    long long launches, last_trips;
    // staging for kmpc_solve_host
    double *d_in, *d_out, *h_in, *h_out;
    int32_t *d_iout, *h_iout;
    size_t in_doubles, out_doubles;
    cudaStream_t stream;
    char err[256];
};

static char g_err[256] = "";

static int fail(kmpc_handle *h, int code, const char *fmt, const char *detail) {
    char *dst = h ? h->err : g_err;
    snprintf(dst, 256, fmt, detail ? detail : "");
    return code;
}

#define CU(call)                                                                      \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) return fail(h, KMPC_E_CUDA, #call ": %s", cudaGetErrorString(e_)); \
    } while (0)

static int check_cfg(const kmpc_config *cfg) {
    if (!cfg) return 0;
    if (cfg->N < 1 || cfg->N > 4096 || cfg->O_max < 0 || cfg->O_max > 256 || cfg->B_max < 1) return 0;
    if (cfg->cost_mode != 0 && cfg->cost_mode != 1) return 0;
    if (cfg->layout != 0 && cfg->layout != 1) return 0;
    if (cfg->goal_k_lo < 0 || cfg->goal_k_hi > cfg->N) return 0;
    if (!(cfg->T > 0) || !(cfg->tol > 0) || cfg->max_iter < 0) return 0;
    for (int i = 0; i < 4; ++i) if (!(cfg->lo[i] < cfg->hi[i])) return 0;
    return 1;
}

static int slots_for(const kmpc_config *cfg, int sm_count, int blocks_per_sm) {
    int want = (cfg->B_max + KMPC_TPB - 1) / KMPC_TPB;
    int cap = sm_count * blocks_per_sm;
    int blocks = want < cap ? want : cap;
    return blocks * KMPC_TPB;
}

extern "C" int kmpc_version(void) { return KMPC_VERSION; }

extern "C" size_t kmpc_workspace_bytes(const kmpc_config *cfg) {
    if (!check_cfg(cfg)) return 0;
    Rows r = make_rows(cfg->N, cfg->O_max);
    // upper bound without querying a device: 148 SMs x 2 blocks
    int slots = slots_for(cfg, 148, 2);
    return (size_t)r.total * slots * sizeof(double);
}

extern "C" const char *kmpc_last_error(const kmpc_handle *h) { return h ? h->err : g_err; }

extern "C" void kmpc_destroy(kmpc_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->ws) cudaFree(h->ws);
    if (h->counter) cudaFree(h->counter);
    if (h->trips) cudaFree(h->trips);
    if (h->d_in) cudaFree(h->d_in);
    if (h->d_out) cudaFree(h->d_out);
    if (h->d_iout) cudaFree(h->d_iout);
    if (h->h_in) cudaFreeHost(h->h_in);
    if (h->h_out) cudaFreeHost(h->h_out);
    if (h->h_iout) cudaFreeHost(h->h_iout);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    free(h);
}

extern "C" int kmpc_create(const kmpc_config *cfg, kmpc_handle **out) {
    kmpc_handle *h = NULL;
    if (!out) return fail(NULL, KMPC_E_BADARG, "kmpc_create: out is NULL%s", "");
    *out = NULL;
    if (!check_cfg(cfg)) return fail(NULL, KMPC_E_BADARG, "kmpc_create: invalid configuration%s", "");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail(NULL, KMPC_E_NODEVICE, "kmpc_create: no CUDA device (this library has no CPU path)%s", "");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(NULL, KMPC_E_BADARG, "kmpc_create: bad device ordinal%s", "");
    h = (kmpc_handle *)calloc(1, sizeof(kmpc_handle));
    if (!h) return fail(NULL, KMPC_E_NOMEM, "kmpc_create: out of host memory%s", "");
    h->cfg = *cfg;
    h->device = cfg->device;
    h->rows = make_rows(cfg->N, cfg->O_max);
    cudaError_t e = cudaSetDevice(h->device);
    int bps = 0;
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->device);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kmpc_ipm_kernel, KMPC_TPB, 0);
    if (e == cudaSuccess && bps < 1) bps = 1;
    if (e == cudaSuccess) {
        h->slots = slots_for(cfg, h->sm_count, bps);
        h->blocks = h->slots / KMPC_TPB;
        e = cudaMalloc(&h->ws, (size_t)h->rows.total * h->slots * sizeof(double));
    }
    if (e == cudaSuccess) e = cudaMalloc(&h->counter, sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&h->trips, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev1);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof g_err, "kmpc_create: %s", cudaGetErrorString(e));
        kmpc_destroy(h);
        return e == cudaErrorMemoryAllocation ? KMPC_E_NOMEM : KMPC_E_CUDA;
    }
    *out = h;
    return 0;
}

static void relax_bounds(const kmpc_config *cfg, Cfg *c) {
    for (int i = 0; i < 4; ++i) {
        c->hasL[i] = cfg->lo[i] > -KMPC_NO_BOUND;
        c->hasU[i] = cfg->hi[i] < KMPC_NO_BOUND;
        c->lb[i] = c->hasL[i] ? cfg->lo[i] - K_BOUND_RELAX * fmax(1.0, fabs(cfg->lo[i])) : -INFINITY;
        c->ub[i] = c->hasU[i] ? cfg->hi[i] + K_BOUND_RELAX * fmax(1.0, fabs(cfg->hi[i])) : INFINITY;
    }
}

extern "C" int kmpc_solve(kmpc_handle *h, int B, const double *x_cur, const double *goal, const double *X0, const double *U0,
                          const double *obs_centers, int O, double obs_radius, double inflation, double *X_out, double *U_out,
                          double *obj_out, int32_t *status_out, int32_t *iters_out, void *cuda_stream) {
    if (!h) return fail(NULL, KMPC_E_BADARG, "kmpc_solve: NULL handle%s", "");
    if (B < 0 || B > h->cfg.B_max) return fail(h, KMPC_E_BADARG, "kmpc_solve: B outside [0, B_max]%s", "");
    if (O < 0 || O > h->cfg.O_max || (O > 0 && !obs_centers)) return fail(h, KMPC_E_BADARG, "kmpc_solve: bad obstacle arguments%s", "");
    if ((X0 == NULL) != (U0 == NULL)) return fail(h, KMPC_E_BADARG, "kmpc_solve: X0 and U0 must both be given or both be NULL%s", "");
    if (B == 0) return 0;
    if (!x_cur || !goal || !X_out || !U_out) return fail(h, KMPC_E_BADARG, "kmpc_solve: NULL required pointer%s", "");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    CU(cudaSetDevice(h->device));
    Cfg c;
    memset(&c, 0, sizeof c);
    const kmpc_config *cf = &h->cfg;
    c.N = cf->N; c.O = O; c.cost_mode = cf->cost_mode; c.gk_lo = cf->goal_k_lo; c.gk_hi = cf->goal_k_hi;
    c.max_iter = cf->max_iter; c.layout = cf->layout; c.B = B;
    relax_bounds(cf, &c);
    c.T = cf->T; c.W[0] = cf->W[0]; c.W[1] = cf->W[1]; c.W[2] = cf->W[2];
    c.Wvn = cf->Wv_neg; c.Wvp = cf->Wv_pos; c.Ww = cf->Ww; c.tol = cf->tol;
    c.obs_radius = obs_radius; c.dL = inflation - K_BOUND_RELAX * fmax(1.0, fabs(inflation));
    c.L = make_rows(cf->N, O);
    c.nb = (cf->N + 1) * (c.hasL[0] + c.hasU[0] + c.hasL[1] + c.hasU[1]) + cf->N * (c.hasL[2] + c.hasU[2] + c.hasL[3] + c.hasU[3]) + cf->N * O;
    c.m = 3 * (cf->N + 1) + cf->N * O;
    IO io;
    io.x_cur = x_cur; io.goal = goal; io.X0 = X0; io.U0 = U0; io.obs = obs_centers;
    io.X_out = X_out; io.U_out = U_out; io.obj = obj_out; io.status = status_out; io.iters = iters_out;
    int blocks = (B + KMPC_TPB - 1) / KMPC_TPB;
    if (blocks > h->blocks) blocks = h->blocks;
    const int launched = blocks * KMPC_TPB;
    CU(cudaMemcpyAsync(h->counter, &launched, sizeof(int), cudaMemcpyHostToDevice, st));
    if (h->timing) { CU(cudaMemsetAsync(h->trips, 0, sizeof(unsigned long long), st)); CU(cudaEventRecord(h->ev0, st)); }
    kmpc_ipm_kernel<<<blocks, KMPC_TPB, 0, st>>>(c, io, h->ws, (size_t)h->slots, h->counter, h->timing ? h->trips : NULL);
    CU(cudaGetLastError());
    h->launches++;
    if (h->timing) {
        CU(cudaEventRecord(h->ev1, st));
        CU(cudaEventSynchronize(h->ev1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        h->last_ms = ms;
        unsigned long long tr = 0;
        CU(cudaMemcpy(&tr, h->trips, sizeof tr, cudaMemcpyDeviceToHost));
        h->last_trips = (long long)tr;
    }
    return 0;
}

static int ensure_staging(kmpc_handle *h) {
    if (h->d_in) return 0;
    const kmpc_config *cf = &h->cfg;
    const size_t Bm = cf->B_max, N = cf->N, O = cf->O_max;
    h->in_doubles = Bm * (6 + 5 * N + 3 + 2 * O);
    h->out_doubles = Bm * (5 * N + 3 + 1);
    CU(cudaMalloc(&h->d_in, h->in_doubles * sizeof(double)));
    CU(cudaMalloc(&h->d_out, h->out_doubles * sizeof(double)));
    CU(cudaMalloc(&h->d_iout, Bm * 2 * sizeof(int32_t)));
    CU(cudaMallocHost(&h->h_in, h->in_doubles * sizeof(double)));
    CU(cudaMallocHost(&h->h_out, h->out_doubles * sizeof(double)));
    CU(cudaMallocHost(&h->h_iout, Bm * 2 * sizeof(int32_t)));
    return 0;
}

extern "C" int kmpc_solve_host(kmpc_handle *h, int B, const double *x_cur, const double *goal, const double *X0, const double *U0,
                               const double *obs_centers, int O, double obs_radius, double inflation, double *X_out, double *U_out,
                               double *obj_out, int32_t *status_out, int32_t *iters_out) {
    if (!h) return fail(NULL, KMPC_E_BADARG, "kmpc_solve_host: NULL handle%s", "");
    if (B < 0 || B > h->cfg.B_max) return fail(h, KMPC_E_BADARG, "kmpc_solve_host: B outside [0, B_max]%s", "");
    if (O < 0 || O > h->cfg.O_max || (O > 0 && !obs_centers)) return fail(h, KMPC_E_BADARG, "kmpc_solve_host: bad obstacle arguments%s", "");
    if ((X0 == NULL) != (U0 == NULL)) return fail(h, KMPC_E_BADARG, "kmpc_solve_host: X0 and U0 must both be given or both be NULL%s", "");
    if (B == 0) return 0;
    if (!x_cur || !goal || !X_out || !U_out) return fail(h, KMPC_E_BADARG, "kmpc_solve_host: NULL required pointer%s", "");
    CU(cudaSetDevice(h->device));
    int rc = ensure_staging(h);
    if (rc) return rc;
    const size_t N = h->cfg.N, b = B;
    const size_t nX = b * 3 * (N + 1), nU = b * 2 * N;
    // pack inputs into the pinned buffer: [x_cur | goal | X0 | U0 | obs]
    size_t o = 0;
    double *hx = h->h_in + o; memcpy(hx, x_cur, b * 3 * sizeof(double)); o += b * 3;
    double *hg = h->h_in + o; memcpy(hg, goal, b * 3 * sizeof(double)); o += b * 3;
    size_t oX = 0, oU = 0, oO = 0;
    if (X0) { oX = o; memcpy(h->h_in + o, X0, nX * sizeof(double)); o += nX; oU = o; memcpy(h->h_in + o, U0, nU * sizeof(double)); o += nU; }
    if (O) { oO = o; memcpy(h->h_in + o, obs_centers, b * 2 * O * sizeof(double)); o += b * 2 * O; }
    (void)hx; (void)hg;
    CU(cudaMemcpyAsync(h->d_in, h->h_in, o * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    double *dX = h->d_out, *dU = h->d_out + nX, *dobj = h->d_out + nX + nU;
    rc = kmpc_solve(h, B, h->d_in, h->d_in + b * 3, X0 ? h->d_in + oX : NULL, X0 ? h->d_in + oU : NULL, O ? h->d_in + oO : NULL, O,
                    obs_radius, inflation, dX, dU, dobj, h->d_iout, h->d_iout + b, h->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(h->h_out, h->d_out, (nX + nU + b) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(h->h_iout, h->d_iout, b * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    memcpy(X_out, h->h_out, nX * sizeof(double));
    memcpy(U_out, h->h_out + nX, nU * sizeof(double));
    if (obj_out) memcpy(obj_out, h->h_out + nX + nU, b * sizeof(double));
    if (status_out) memcpy(status_out, h->h_iout, b * sizeof(int32_t));
    if (iters_out) memcpy(iters_out, h->h_iout + b, b * sizeof(int32_t));
    return 0;
}

extern "C" int kmpc_agent_handoff(kmpc_handle *h, int B, const double *X, const double *U, double *x_cur, double *applied_out,
                                  void *cuda_stream) {
    if (!h) return fail(NULL, KMPC_E_BADARG, "kmpc_agent_handoff: NULL handle%s", "");
    if (B < 0 || !X || !U || !x_cur) return fail(h, KMPC_E_BADARG, "kmpc_agent_handoff: bad arguments%s", "");
    if (B == 0) return 0;
    CU(cudaSetDevice(h->device));
    kmpc_handoff_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)cuda_stream>>>(B, h->cfg.N, h->cfg.layout, X, U, x_cur, applied_out);
    CU(cudaGetLastError());
    h->launches++;
    return 0;
}

extern "C" int kmpc_set_timing(kmpc_handle *h, int enable) {
    if (!h) return KMPC_E_BADARG;
    h->timing = enable ? 1 : 0;
    return 0;
}

extern "C" int kmpc_get_stats(kmpc_handle *h, kmpc_stats *out) {
    if (!h || !out) return KMPC_E_BADARG;
    out->last_kernel_ms = h->last_ms; out->launches = h->launches; out->slots = h->slots; out->blocks = h->blocks;
    out->threads_per_block = KMPC_TPB; out->sm_count = h->sm_count; out->trips = h->last_trips;
    return 0;
}

extern "C" int kmpc_measure_fp64_peak(kmpc_handle *h, double *tflops_out) {
    if (!h || !tflops_out) return KMPC_E_BADARG;
    CU(cudaSetDevice(h->device));
    const int blocks = h->sm_count * 8, tpb = 256, iters = 4096;
    double *buf = NULL;
    CU(cudaMalloc(&buf, (size_t)blocks * tpb * sizeof(double)));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CU(cudaEventRecord(h->ev0, 0));
        kmpc_dfma_kernel<<<blocks, tpb>>>(buf, iters, 0.999999, 1e-9);
        CU(cudaEventRecord(h->ev1, 0));
        CU(cudaEventSynchronize(h->ev1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        if (rep > 0 && ms < best) best = ms;
    }
    h->launches += 5;
    cudaFree(buf);
    const double flops = (double)blocks * tpb * (double)iters * 16.0 * 8.0 * 2.0;
    *tflops_out = flops / (best * 1e-3) / 1e12;
    return 0;
}
