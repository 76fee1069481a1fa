// kmpc.cu -- CUDA kernels (sm_100a) and the C ABI of include/kmpc.h.
//
// Replaces the numerical core of mpc/optimizer.py:319-400 (MotionPlanner.solve -> CasADi/IPOPT) for B instances at
// once.  The solver proper is kmpc_warp_kernel (kmpc_warp.cuh): a persistent grid, one warp per problem instance, the
// iterate in registers, the block's serial Riccati recursions side by side on one warp, nothing but the problem data and
// the result in HBM.  The thread-per-instance kernels below (state in a structure-of-arrays HBM workspace, host-driven
// trip loop) are the correctness fall-back for horizons above 63 stages or more obstacle rows than shared memory holds.
// No tensor cores: the stage blocks are 3x3 / 2x2 / 2x3 with a data-dependent 2x2 inverse between them.
// No CPU fallback exists in this library.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cub/device/device_radix_sort.cuh>

#include "../../include/kmpc.h"
#include "kmpc_core.cuh"
#include "kmpc_warp.cuh"
#include "kmpc_order_prior.h"

using namespace kmpc;

#define KMPC_TPB 128

// ------------------------------------------------------------------------------------------------
// Solver kernels.  One CUDA thread advances one instance by one phase; instance b owns workspace column b.
// Between launches all solver state lives in the workspace (HBM, structure of arrays), so every launch works on a
// COMPACTED list of the instances that still need that phase: a warp always has 32 live lanes although iteration
// counts differ by 5x between instances and although some instances need extra sweeps (inertia correction, second-
// order correction) or extra trial points (back-tracking).
//   lists: LA[p] = instances that need a sweep in trip parity p, LT[p] = instances that need a trial point.
//   trip t (p = t & 1):  sweep(LA[p]) -> ok: LT[p], wrong inertia: LA[1-p]
//                        rollout(LT[p]);  trial(LT[p]) -> Newton/SOC next: LA[1-p], back-track next: LT[1-p], done: outputs
// Survivors are appended per block with one atomicAdd (order inside a block chunk is preserved, so the columns a
// warp touches stay sorted and mostly adjacent -> coalesced 32-byte sectors).
// ------------------------------------------------------------------------------------------------
struct Lists {
    int *LA[2], *LT[2];
    int *cnt;  // [0,1] = |LA[0]|,|LA[1]|   [2,3] = |LT[0]|,|LT[1]|
    unsigned long long *trips;
};

// block-wide ordered append: every thread of the block calls it; returns the list position or -1
__device__ __forceinline__ int block_append(bool flag, int *counter) {
    __shared__ int s_cnt[KMPC_TPB / 32];
    __shared__ int s_base;
    const unsigned m = __ballot_sync(0xffffffffu, flag);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) s_cnt[w] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int i = 0; i < KMPC_TPB / 32; ++i) { const int v = s_cnt[i]; s_cnt[i] = tot; tot += v; }
        s_base = tot ? atomicAdd(counter, tot) : 0;
    }
    __syncthreads();
    const int pos = s_base + s_cnt[w] + __popc(m & ((1u << lane) - 1u));
    __syncthreads();  // s_cnt/s_base are reused by the next call
    return flag ? pos : -1;
}

__global__ void __launch_bounds__(KMPC_TPB)
kmpc_init_kernel(const Cfg c, const IO io, double *__restrict__ ws, const size_t S, int *__restrict__ LA0) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= c.B) return;
    double *wsp = ws + b;
    Ctx t;
    t.inst = b;
    pass_init(c, t, wsp, S, io);
    ctx_store(t, c.L, wsp, S);
    LA0[b] = b;
}

template <bool OBS>
__global__ void __launch_bounds__(KMPC_TPB, 3)
kmpc_sweep_kernel(const Cfg c, const IO io, double *__restrict__ ws, const size_t S, const Lists ls, const int p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = ls.cnt[p];
    if (blockIdx.x * blockDim.x >= n) return;  // whole block idle
    int r = -1000, b = -1;
    if (i < n) {
        b = ls.LA[p][i];
        double *wsp = ws + b;
        Ctx t;
        ctx_load(t, c.L, wsp, S);
        t.inst = b;
        r = phase_sweep<OBS>(c, t, wsp, S);
        const double *px = wsp + (size_t)c.L.rCtx * S;
        double *pw = wsp + (size_t)c.L.rCtx * S;
        (void)px;
        pw[(size_t)X_TRIPS * S] = t.trips; pw[(size_t)X_DELTA * S] = t.delta;
        if (r != 100 && r != 101) {
            pass_output(c, t, wsp, S, io, r);
            if (ls.trips) atomicAdd(ls.trips, (unsigned long long)t.trips);
        }
    }
    const int pos_t = block_append(r == 100, ls.cnt + 2 + p);
    if (pos_t >= 0) ls.LT[p][pos_t] = b;
    const int pos_a = block_append(r == 101, ls.cnt + (1 - p));
    if (pos_a >= 0) ls.LA[1 - p][pos_a] = b;
}

template <bool OBS>
__global__ void __launch_bounds__(KMPC_TPB, 4)
kmpc_rollout_kernel(const Cfg c, double *__restrict__ ws, const size_t S, const Lists ls, const int p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = ls.cnt[2 + p];
    if (i >= n) return;
    const int b = ls.LT[p][i];
    double *wsp = ws + b;
    Ctx t;
    ctx_load(t, c.L, wsp, S);
    if (t.mode == M_TRIAL) return;  // back-tracking instance: keeps its step, only a new trial point
    t.inst = b;
    phase_rollout<OBS>(c, t, wsp, S);
    ctx_store(t, c.L, wsp, S);
}

template <bool OBS>
__global__ void __launch_bounds__(KMPC_TPB, 3)
kmpc_trial_kernel(const Cfg c, const IO io, double *__restrict__ ws, const size_t S, const Lists ls, const int p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = ls.cnt[2 + p];
    if (blockIdx.x * blockDim.x >= n) return;
    int r = -1000, b = -1, mode = -1;
    if (i < n) {
        b = ls.LT[p][i];
        double *wsp = ws + b;
        Ctx t;
        ctx_load(t, c.L, wsp, S);
        t.inst = b;
        r = phase_trial<OBS>(c, t, wsp, S);
        mode = t.mode;
        if (r == 100) ctx_store(t, c.L, wsp, S);
        else {
            pass_output(c, t, wsp, S, io, r);
            if (ls.trips) atomicAdd(ls.trips, (unsigned long long)t.trips);
        }
    }
    const int pos_t = block_append(r == 100 && mode == M_TRIAL, ls.cnt + 2 + (1 - p));
    if (pos_t >= 0) ls.LT[1 - p][pos_t] = b;
    const int pos_a = block_append(r == 100 && mode != M_TRIAL, ls.cnt + (1 - p));
    if (pos_a >= 0) ls.LA[1 - p][pos_a] = b;
}

// ------------------------------------------------------------------------------------------------
// Warp-per-instance solver (kmpc_warp.cuh): a persistent grid of warps, each pulling the next instance from a global
// queue and solving it start to finish with the whole iterate in registers.  Used for problems without obstacle rows
// and N + 1 <= 32 * SPL.
// ------------------------------------------------------------------------------------------------
#define KMPC_LOOKAHEAD 4 /* trips enqueued before the host looks at the active-instance count of an older trip */

struct kmpc_handle {
    kmpc_config cfg;
    Rows rows;
    int device, sm_count, cols;  // cols = workspace columns (B_max rounded up to a multiple of 32)
    double *ws;
    int *lists;  // 4 x cols ints
    int *cnt;    // 4 counters of the thread solver's lists (cnt[0] doubles as the warp solver's queue head) + cnt[4]: restoration hand-over count
    unsigned long long *trips;
    int *h_cnt;  // pinned: KMPC_LOOKAHEAD x 4 counters
    cudaEvent_t ev0, ev1, evq[KMPC_LOOKAHEAD];
    int timing;
    double last_ms;
    long long launches, last_trips;
    int last_host_trips;
    // staging for kmpc_solve_host
    double *d_in, *d_out, *h_in, *h_out;
    int32_t *d_iout, *h_iout;
    size_t in_doubles, out_doubles;
    double *wscratch;      // warp solver: global scratch slots (WLay::GPRIV doubles per resident warp)
    size_t wscratch_doubles;
    int host_B;  // batch of the last kmpc_solve_host (addresses kmpc_host_result hands out)
    // queue order of the warp solver (likely-long instances first): keys / instance indices before and after the sort
    int order_mode;
    unsigned *okey;        // 2 x cols
    int32_t *oval;         // 2 x cols; the second half is the order the kernel reads
    void *osort_tmp;
    size_t osort_bytes;
    int32_t *cost_buf;     // closed loops: trips every agent's solve took this step / last step (2 x cols); the next step's queue is ordered by it
    const int32_t *order_hint;   // set by the closed loops for steps >= 1: the per-instance cost of the previous step (device pointer)
    double *env_obs;       // kmpc_environment_loop: the circles each agent kept this step: cols x O_max x 2 centres, then their N-column tracks
    double *env_rad;       //   ... their per-slot radii, cols x O_max
    int32_t *env_idx;      //   ... which dynamic candidate sits in every dynamic slot
    int max_smem;          // opt-in shared memory per block of the device
    struct { const void *fn; size_t smem; int bpsm; } kattr[16];   // per kernel instantiation: shared-memory opt-in done, resident blocks per SM
    int n_kattr;
    size_t fin_smem[2];    // shared-memory opt-in of the two finisher kernels done (bytes)
    // restoration-phase hand-over (kmpc_finish_kernel): workspace columns, the instance in each, the number in use
    double *resto_ws;
    int32_t *resto_list;
    int *resto_count;
    int resto_cap, resto_rows;
    int last_path;         // 1: the last solve ran the warp kernel, 0: the thread-per-instance fall-back
    cudaStream_t stream;
    char err[256];
};


// launch shape per stage-slot count: warps (= instances) per block, resident blocks per SM the register budget is cut for
#ifndef KMPC_WPB1
#define KMPC_WPB1 16   /* one block of 16 instances per SM: 13.6 ms vs 14.5 ms for 2 x 8 once the queue tail was gone */
#endif
#ifndef KMPC_MINB1
#define KMPC_MINB1 1
#endif
#ifndef KMPC_WPB_SMALL
#define KMPC_WPB_SMALL 4   /* instances per block of the low-latency launch shape for batches of at most one wave of such blocks */
#endif
#ifndef KMPC_WPB2
#define KMPC_WPB2 12   /* N <= 51: 12 warps at 170 registers (53.1 ms at N = 50) vs 8 at 254 (53.8 ms) */
#endif
#ifndef KMPC_MINB2
#define KMPC_MINB2 1
#endif
#ifndef KMPC_WPBO   /* obstacle-row kernels, N <= 31 */
#define KMPC_WPBO 12
#endif
#ifndef KMPC_MINBO
#define KMPC_MINBO 1
#endif
#ifndef KMPC_WPB3
#define KMPC_WPB3 9
#endif
#ifndef KMPC_MINB3
#define KMPC_MINB3 1
#endif
// FULL: every bound of x, y, v, omega exists (the default problem), so the per-side tests are compiled away.
// OBS: obstacle-distance rows present (their per-row state lives in shared memory; the block shrinks to what fits).
// TAIL: with the tail mode of w_worker (full-solve inertia candidates in borrowed instance slots once a block runs dry).
template <int SPL, int NST, bool FULL, bool OBS, int WPB, int MINB, bool TAIL = false>
__global__ void __launch_bounds__(32 * WPB, MINB)
kmpc_warp_kernel(const Cfg c, const IO io, int *__restrict__ queue, unsigned long long *__restrict__ trips_total) {
    extern __shared__ double s_dyn[];  // WLay<SPL, NST>::bytes(warps per block, O)
    w_worker<SPL, NST, FULL, OBS, TAIL>(c, io, s_dyn, queue, trips_total);
}

// Finisher of the warp solver: one block (one working thread) per instance that was handed over because its regular line search failed --
// IPOPT's restoration phase, then the rest of the regular algorithm (kmpc_resto.cuh, finish_instance).  Launched after every warp-solver
// launch with a fixed grid; the number of columns in use is read on the device (no host round trip), normally zero.
template <bool OBS>
__global__ void __launch_bounds__(32)
kmpc_finish_kernel(const Cfg c, const IO io, int in_smem) {
    extern __shared__ double s_fin[];
    const int i = blockIdx.x;   // one instance per block: the whole column of an instance fits the block's shared memory
    const int n = *io.resto_count < io.resto_cap ? *io.resto_count : io.resto_cap;
    if (i >= n) return;
    double *col = io.resto_ws + (size_t)i * io.resto_rows;
    if (in_smem) {
        // The phase is a long serial computation of ONE thread on ~5-15 k doubles of state: from HBM / L2 every dependent access costs
        // hundreds of cycles (2 ms per interior-point iteration at O = 10), from shared memory a tenth of that.
        const int rows = make_resto_rows(c.L).total;
        for (int r = threadIdx.x; r < rows; r += 32) s_fin[r] = col[r];
        __syncwarp();
        col = s_fin;
    }
    if (threadIdx.x == 0) finish_instance<OBS>(c, io, i, col);
}

#ifndef KMPC_TAIL_WAVES
#define KMPC_TAIL_WAVES 0   /* batches of at most this many waves of resident instances run the kernel with the tail mode; 0: never.
   r02b: 16.  Since the inertia prediction (1.05 instead of 2.36 retry trips per solve) the mode saves less than its two extra block
   barriers and the candidate assembly cost: B = 1 307 -> 266 us (N = 30), 137 -> 113 us (N = 7); 4,096 instances 4.59 -> 4.47 ms;
   the 1/2, 1/4, 1/8 slices of the headline batch 9.88 / 8.39 / 7.30 -> 9.69 / 8.25 / 7.24 ms (scripts/tail_onoff.py, strong_slices.py).
   The instantiation stays (KMPC_FORCE_TAIL=1 in the environment selects it, tests/test_parity_gpu.py keeps it bit-identical). */
#endif

// Queue order of the persistent kernel.  Iteration counts differ by more than 8x between instances and a long instance that is
// fetched late sets the end of the launch; how long an instance takes is largely a function of its geometry, so the instances
// are handed out in descending order of a prior (kmpc_order_prior.h, fitted by scripts/fit_order_prior.py): expected iteration
// count + 1 sd per bin of (|bearing of the goal from the start heading|, signed heading change, goal distance).  Scheduling
// only -- every instance is solved by the same arithmetic whatever its queue position.
// With obstacle rows the rare very long instances are the ones whose way is blocked: every circle closer than r_o + I + 0.4 m to
// the part of the straight start-goal segment the horizon can reach (v_max T N) adds 10 trips to the key (4,096 x 65,536
// instances, O = 10 and 4: tail 1.05-1.15 x ideal with the geometric prior alone, 1.015-1.03 x with this term).
__global__ void kmpc_order_key_kernel(int B, int layout, const double *__restrict__ x_cur, const double *__restrict__ goal,
                                      const double *__restrict__ obs, const double *__restrict__ orad, int O, int obs_sw, int N,
                                      double reach, double near_u, double near_add, unsigned *__restrict__ key, int32_t *__restrict__ val) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const size_t s0 = layout ? (size_t)b : (size_t)b * 3, st = layout ? (size_t)B : 1;
    const double dx = goal[s0] - x_cur[s0], dy = goal[s0 + st] - x_cur[s0 + st], th = x_cur[s0 + 2 * st], thg = goal[s0 + 2 * st];
    const double kPi = 3.14159265358979323846;
    double bs = atan2(dy, dx) - th;
    bs -= 2.0 * kPi * rint(bs / (2.0 * kPi));
    const double dt = (bs < 0 ? -1.0 : 1.0) * (thg - th);
    const double fa = floor(fabs(bs) / kPi * KMPC_PRIOR_NB), fd = floor((dt + 2.0 * kPi) / (4.0 * kPi) * KMPC_PRIOR_ND),
                 fr = floor(sqrt(dx * dx + dy * dy) / 6.0 * KMPC_PRIOR_NR);
    // (comparisons written so that a NaN input lands in bin 0)
    const int ia = fa >= 0 ? (fa < KMPC_PRIOR_NB ? (int)fa : KMPC_PRIOR_NB - 1) : 0;
    const int id = fd >= 0 ? (fd < KMPC_PRIOR_ND ? (int)fd : KMPC_PRIOR_ND - 1) : 0;
    const int ir = fr >= 0 ? (fr < KMPC_PRIOR_NR ? (int)fr : KMPC_PRIOR_NR - 1) : 0;
    float k = kmpc_order_prior[(ia * KMPC_PRIOR_ND + id) * KMPC_PRIOR_NR + ir];
    if (O > 0) {
        const double dist = sqrt(dx * dx + dy * dy), ux = dist > 0 ? dx / dist : 1.0, uy = dist > 0 ? dy / dist : 0.0;
        const double L = dist < reach ? dist : reach;
        int blocked = 0;
        for (int o = 0; o < O; ++o) {
            // (moving circles: where they are now, column 0 of the track)
            const size_t i0 = obs_sw ? (layout ? (((size_t)o * N) * 2) * B + b : (((size_t)b * O + o) * N) * 2)
                                     : (layout ? ((size_t)o * 2) * B + b : ((size_t)b * O + o) * 2);
            const double rx = obs[i0] - x_cur[s0], ry = obs[layout ? i0 + B : i0 + 1] - x_cur[s0 + st];
            double t = rx * ux + ry * uy;
            t = t < 0 ? 0 : (t > L ? L : t);
            const double ex = rx - t * ux, ey = ry - t * uy;
            const double near = orad ? orad[layout ? (size_t)o * B + b : (size_t)b * O + o] + near_add : near_u;
            blocked += (ex * ex + ey * ey < near * near) ? 1 : 0;
        }
        k += 10.0f * (float)blocked;
    }
    // 1/32-trip resolution, 16 bits: two radix passes instead of four
    key[b] = (unsigned)fminf(k * 32.0f, 65535.0f);
    val[b] = b;
}

// closed loops, step >= 1: an agent's solve costs about what its previous one did (same agent, one control interval later, warm-started
// from that solution), so the previous step's trip count is the key -- no model of the problem geometry involved.
__global__ void kmpc_order_hint_kernel(int B, const int32_t *__restrict__ cost, unsigned *__restrict__ key, int32_t *__restrict__ val) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int cst = cost[b];
    key[b] = cst < 0 ? 0u : (cst > 65535 ? 65535u : (unsigned)cst);
    val[b] = b;
}

// fills h->oval[cols..] with the queue order of this batch; returns NULL in *order when the natural order is kept
static cudaError_t queue_order(kmpc_handle *h, int B, int resident, const Cfg &c, const IO &io, int layout, cudaStream_t st, const int32_t **order) {
    *order = NULL;
    if (!h->order_mode || B <= resident) return cudaSuccess;   // a single wave: every instance starts at once
    const size_t S = (size_t)h->cols;
    cudaError_t e;
    if (!h->osort_tmp) {   // (the three buffers are committed to the handle together: a failed allocation leaves none behind)
        unsigned *k = NULL; int32_t *v = NULL; void *tmp = NULL; size_t nb = 0;
        if ((e = cudaMalloc(&k, 2 * S * sizeof(unsigned))) == cudaSuccess && (e = cudaMalloc(&v, 2 * S * sizeof(int32_t))) == cudaSuccess &&
            (e = cub::DeviceRadixSort::SortPairsDescending(NULL, nb, k, k + S, v, v + S, (int)S, 0, 16, st)) == cudaSuccess)
            e = cudaMalloc(&tmp, nb ? nb : 16);
        if (e != cudaSuccess) { cudaFree(k); cudaFree(v); cudaFree(tmp); return e; }
        h->okey = k; h->oval = v; h->osort_tmp = tmp; h->osort_bytes = nb;
    }
    const double inflation = c.dL + K_BOUND_RELAX * fmax(1.0, fabs(c.dL));   // (c.dL is the relaxed bound; close enough for a heuristic)
    if (h->order_hint) kmpc_order_hint_kernel<<<(B + 255) / 256, 256, 0, st>>>(B, h->order_hint, h->okey, h->oval);
    else kmpc_order_key_kernel<<<(B + 255) / 256, 256, 0, st>>>(B, layout, io.x_cur, io.goal, io.obs, io.orad, c.O, c.obs_sw, c.N, c.ub[2] * c.T * c.N,
                                                               c.obs_radius + inflation + 0.4, inflation + 0.4, h->okey, h->oval);
    size_t bytes = h->osort_bytes;
    if ((e = cub::DeviceRadixSort::SortPairsDescending(h->osort_tmp, bytes, h->okey, h->okey + S, h->oval, h->oval + S, B, 0, 16, st)) != cudaSuccess) return e;
    h->launches += 2;
    *order = h->oval + S;
    return cudaGetLastError();
}

// returns cudaErrorInvalidConfiguration if not even one instance fits into shared memory (caller falls back)
template <int SPL, int NST, bool FULL, bool OBS, int WPB, int MINB>
static cudaError_t launch_warp_kernel(kmpc_handle *h, int device, int sm_count, int B, const Cfg &c, const IO &io_in, int *queue,
                                      unsigned long long *trips, cudaStream_t st) {
    cudaError_t e = cudaSuccess;
    int max_smem = h->max_smem;
    if (max_smem <= 0) {
        e = cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
        if (e != cudaSuccess) return e;
    }
    int wpb = WPB;
    while (wpb > 0 && WLay<SPL, NST, OBS>::bytes(wpb, c.O, c.obs_sw) > (size_t)max_smem) --wpb;
    if (wpb < 1) return cudaErrorInvalidConfiguration;
    const size_t smem = WLay<SPL, NST, OBS>::bytes(wpb, c.O, c.obs_sw);
    // shared-memory opt-in + occupancy of a kernel instantiation: asked once per handle and shared-memory size, not once per solve
    // (three runtime calls, ~10 us of host time that a B = 1 solve of ~130 us would pay every time)
    auto prepare = [&](void (*k)(const Cfg, const IO, int *, unsigned long long *), int *bp) -> cudaError_t {
        for (int i = 0; i < h->n_kattr; ++i)
            if (h->kattr[i].fn == (const void *)k && h->kattr[i].smem == smem) { *bp = h->kattr[i].bpsm; return cudaSuccess; }
        cudaError_t r = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);   // (the device's limit: the attribute belongs to the function, not to this handle's problem size)
        if (r == cudaSuccess) r = cudaOccupancyMaxActiveBlocksPerMultiprocessor(bp, k, 32 * wpb, smem);
        if (r == cudaSuccess) {
            const int i = h->n_kattr < 16 ? h->n_kattr++ : 15;
            h->kattr[i].fn = (const void *)k; h->kattr[i].smem = smem; h->kattr[i].bpsm = *bp;
        }
        return r;
    };
    void (*kern)(const Cfg, const IO, int *, unsigned long long *) = kmpc_warp_kernel<SPL, NST, FULL, OBS, WPB, MINB, false>;
    int bpsm = 0;
    e = prepare(kern, &bpsm);
    if (e != cudaSuccess) return e;
    const bool no_tail = getenv("KMPC_NO_TAIL") != NULL, force_tail = getenv("KMPC_FORCE_TAIL") != NULL;
    if (!OBS && !no_tail && (force_tail || (long long)B <= (long long)KMPC_TAIL_WAVES * sm_count * (bpsm > 0 ? bpsm : 1) * wpb)) {
        // few waves: the phase in which the queue is drained is a large part of the launch -- the kernel with the tail mode
        kern = kmpc_warp_kernel<SPL, NST, FULL, OBS, WPB, MINB, !OBS>;
        e = prepare(kern, &bpsm);
        if (e != cudaSuccess) return e;
    }
    int grid = sm_count * (bpsm > 0 ? bpsm : 1);
    const int need = (B + wpb - 1) / wpb;
    if (grid > need) grid = need;
    // global scratch: one slot per resident warp (grown on demand; a solve on the handle is never concurrent with another)
    const size_t want = (size_t)grid * wpb * WLay<SPL, NST>::GPRIV;
    if (want > h->wscratch_doubles) {
        if (h->wscratch) { e = cudaStreamSynchronize(st); if (e != cudaSuccess) return e; cudaFree(h->wscratch); h->wscratch = NULL; h->wscratch_doubles = 0; }
        e = cudaMalloc(&h->wscratch, want * sizeof(double));
        if (e != cudaSuccess) return e;
        h->wscratch_doubles = want;
    }
    IO io = io_in;
    io.wscratch = h->wscratch;
    e = queue_order(h, B, grid * wpb, c, io, c.layout, st, &io.order);
    if (e != cudaSuccess) return e;
    // hand-over workspace of the restoration phase (allocated on the first solve, sized for the handle's largest problem)
    if (!h->resto_count) {
        const int rows = make_resto_rows(make_rows(h->cfg.N, h->cfg.O_max, 1)).total;
        long long cap = h->cfg.B_max < 1024 ? h->cfg.B_max : 1024;
        const long long budget = 256ll << 20;   // bytes
        if (cap * rows * 8ll > budget) cap = budget / (rows * 8ll);
        double *ws = NULL; int32_t *li = NULL;
        if (cap >= 1) {
            e = cudaMalloc(&ws, (size_t)cap * rows * sizeof(double));
            if (e == cudaSuccess) e = cudaMalloc(&li, (size_t)cap * sizeof(int32_t));
        }
        if (e != cudaSuccess) { cudaFree(ws); cudaFree(li); return e; }
        h->resto_ws = ws; h->resto_list = li; h->resto_count = h->cnt + 4; h->resto_cap = (int)(cap >= 1 ? cap : 0); h->resto_rows = rows;
    }
    io.resto_ws = h->resto_cap ? h->resto_ws : NULL; io.resto_list = h->resto_list; io.resto_count = h->resto_count;
    io.resto_cap = h->resto_cap; io.resto_rows = h->resto_rows;
    // (queue head and hand-over count were zeroed together by solve_impl: one memset of h->cnt[0..7])
    kern<<<grid, 32 * wpb, smem, st>>>(c, io, queue, trips);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (h->resto_cap) {
        const int cap = h->resto_cap < B ? h->resto_cap : B;
        void (*fin)(const Cfg, const IO, int) = OBS ? kmpc_finish_kernel<true> : kmpc_finish_kernel<false>;
        const size_t fbytes = (size_t)make_resto_rows(c.L).total * sizeof(double);
        const int in_smem = fbytes <= (size_t)max_smem ? 1 : 0;     // (else: the column stays in HBM, as large N x O needs)
        if (in_smem && h->fin_smem[OBS ? 1 : 0] < fbytes) {
            e = cudaFuncSetAttribute(fin, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
            if (e != cudaSuccess) return e;
            h->fin_smem[OBS ? 1 : 0] = (size_t)max_smem;
        }
        fin<<<cap, 32, in_smem ? fbytes : 0, st>>>(c, io, in_smem);
        h->launches++;
    }
    return cudaGetLastError();
}

#ifdef KMPC_TUNE_HEADLINE
template <int SPL, int NST, bool OBS, int WPB, int MINB>
static cudaError_t kmpc_tune_launch(bool full, kmpc_handle *h, int device, int sm_count, int B, const Cfg &c, const IO &io, int *queue,
                                    unsigned long long *trips, cudaStream_t st) {
    if constexpr (SPL == 1 && !OBS) { if (full) return launch_warp_kernel<SPL, NST, true, OBS, WPB, MINB>(h, device, sm_count, B, c, io, queue, trips, st); }
    return cudaErrorInvalidConfiguration;
}
#endif

// Batched EgoAgent.step hand-off (agent.py:139-155, :70-72): applied control = U[:,0]; next current state = X[:,1].
// With a mask: agents already at their goal keep their state (the reference stops stepping an agent whose final goal is
// reached, environment.py:31-33); after the hand-off the mask is refreshed with Agent.at_goal (agent.py:78-80):
// || (goal_xy - p_xy) - agent_radius ||_2 - goal_radius <= 0  -- the literal formula of geometry.py:44 (the radius is
// subtracted from both components); agent_radius = 0 gives the plain Euclidean goal distance.
__global__ void kmpc_handoff_kernel(int B, int N, int layout, const double *__restrict__ X, const double *__restrict__ U,
                                    double *__restrict__ x_cur, double *__restrict__ applied, const double *__restrict__ goal,
                                    int32_t *__restrict__ active, double goal_radius, double agent_radius) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const bool was_active = !active || active[b];
    double xn[3];
    for (int j = 0; j < 3; ++j) {
        const size_t src = layout ? ((size_t)j * (N + 1) + 1) * B + b : ((size_t)b * 3 + j) * (N + 1) + 1;
        const size_t dst = layout ? (size_t)j * B + b : (size_t)b * 3 + j;
        xn[j] = was_active ? X[src] : x_cur[dst];
        if (was_active) x_cur[dst] = xn[j];
    }
    if (applied)
        for (int j = 0; j < 2; ++j) {
            const size_t src = layout ? ((size_t)j * N) * B + b : ((size_t)b * 2 + j) * N;
            const size_t dst = layout ? (size_t)j * B + b : (size_t)b * 2 + j;
            applied[dst] = was_active ? U[src] : 0.0;
        }
    if (active && was_active && goal_radius > 0.0) {
        const double gx = goal[layout ? (size_t)b : (size_t)b * 3], gy = goal[layout ? (size_t)B + b : (size_t)b * 3 + 1];
        const double dx = (gx - xn[0]) - agent_radius, dy = (gy - xn[1]) - agent_radius;
        if (sqrt(dx * dx + dy * dy) - goal_radius <= 0.0) active[b] = 0;
    }
}

// Sensor filter of ROSEnvironment.step (environment.py:48-65): one thread per agent keeps the (at most O <= 32) nearest
// candidates within the sensor radius in a sorted list (insertion; M * O compares per agent, candidates are broadcast loads).
#define KMPC_SEL_MAX_O 32
// cc: candidate centres, `cstride` doubles apart (2: [M][2] centres; 3: [M][3] states x, y, heading); the O kept circles go to
// slots slot0 .. slot0 + O - 1 of an agent's Otot slots; rad_out: every slot of this class gets the radius of the nearest kept
// candidate -- the reference builds one radius per obstacle class from its first element (optimizer.py:231-250).
__global__ void kmpc_select_kernel(int B, int M, int layout, const double *__restrict__ x_cur, const double *__restrict__ cc, int cstride,
                                   const double *__restrict__ cr, double sensor_radius, int literal, int O, int Otot, int slot0, double pad_x,
                                   double pad_y, double *__restrict__ obs_out, int32_t *__restrict__ count_out, int32_t *__restrict__ index_out,
                                   double *__restrict__ rad_out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double px = x_cur[layout ? (size_t)b : (size_t)b * 3], py = x_cur[layout ? (size_t)B + b : (size_t)b * 3 + 1];
    double dist[KMPC_SEL_MAX_O];
    int idx[KMPC_SEL_MAX_O];
    int n = 0;
    for (int m = 0; m < M; ++m) {
        const double cx = cc[(size_t)cstride * m], cy = cc[(size_t)cstride * m + 1], r = cr[m];
        double d;
        if (literal) { const double ex = (px - cx) - r, ey = (py - cy) - r; d = sqrt(ex * ex + ey * ey); }   // geometry.py:44 as written
        else { const double ex = px - cx, ey = py - cy; d = sqrt(ex * ex + ey * ey) - r; }
        if (!(d <= sensor_radius)) continue;
        // position of d in the ascending list
        int pos = 0;
        while (pos < n && dist[pos] < d) ++pos;
        if (pos < n && dist[pos] == d) { idx[pos] = m; continue; }   // equal key: the later obstacle replaces the earlier one
        if (pos >= O) continue;                                        // farther than the O nearest kept so far
        const int last = n < O ? n : O - 1;
        for (int j = last; j > pos; --j) { dist[j] = dist[j - 1]; idx[j] = idx[j - 1]; }
        dist[pos] = d; idx[pos] = m;
        if (n < O) ++n;
    }
    const double rad = n > 0 ? cr[idx[0]] : 0.0;
    for (int o = 0; o < O; ++o) {
        const double ox = o < n ? cc[(size_t)cstride * idx[o]] : pad_x, oy = o < n ? cc[(size_t)cstride * idx[o] + 1] : pad_y;
        const int so = slot0 + o;
        if (layout) { obs_out[((size_t)so * 2) * B + b] = ox; obs_out[((size_t)so * 2 + 1) * B + b] = oy; }
        else { obs_out[((size_t)b * Otot + so) * 2] = ox; obs_out[((size_t)b * Otot + so) * 2 + 1] = oy; }
        if (index_out) index_out[(size_t)b * O + o] = o < n ? idx[o] : -1;
        if (rad_out) rad_out[layout ? (size_t)so * B + b : (size_t)b * Otot + so] = rad;
    }
    if (count_out) count_out[b] = n;
}

// a static circle as a track: its centre in every one of the N columns (slot o of Otot; the dynamic slots follow)
__global__ void kmpc_repeat_kernel(int B, int O, int Otot, int N, int layout, const double *__restrict__ cen, double *__restrict__ tracks) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * O) return;
    const int b = (int)(i / O), o = (int)(i % O);
    const double x = cen[layout ? ((size_t)o * 2) * B + b : ((size_t)b * Otot + o) * 2];
    const double y = cen[layout ? ((size_t)o * 2 + 1) * B + b : ((size_t)b * Otot + o) * 2 + 1];
    for (int t = 0; t < N; ++t) {
        const size_t ox = layout ? (((size_t)o * N + t) * 2) * B + b : (((size_t)b * Otot + o) * N + t) * 2;
        tracks[ox] = x; tracks[layout ? ox + B : ox + 1] = y;
    }
}

// Constant-velocity predictor of DynamicObstacle (dynamic_obstacle.py:20-37) for the obstacles each agent selected: one thread
// per (agent, slot) runs the N-column recursion  column 0 = current state, column t = column t-1 + [v cos(a) dt, v sin(a) dt,
// omega dt]  with a = deg2rad(heading) as the reference writes it (:24-25; literal == 0: the heading taken as radians) and
// stores the x, y of every column.  Slots without an obstacle (index < 0) get the padding point in every column.
__global__ void kmpc_predict_kernel(int B, int O, int Otot, int slot0, int M, int N, int layout, const int32_t *__restrict__ index, const double *__restrict__ state,
                                    const double *__restrict__ lin_vel, const double *__restrict__ ang_vel, double dt, int literal,
                                    double pad_x, double pad_y, double *__restrict__ tracks) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * O) return;
    const int b = (int)(i / O), o = (int)(i % O);
    const int m = index ? index[i] : o;
    const bool real = m >= 0 && m < M;
    double x = real ? state[3 * m] : pad_x, y = real ? state[3 * m + 1] : pad_y, th = real ? state[3 * m + 2] : 0.0;
    const double v = real ? lin_vel[m] : 0.0, w = real ? ang_vel[m] : 0.0;
    for (int t = 0; t < N; ++t) {
        const size_t ox = layout ? (((size_t)(slot0 + o) * N + t) * 2) * B + b : (((size_t)b * Otot + slot0 + o) * N + t) * 2;
        tracks[ox] = x; tracks[layout ? ox + B : ox + 1] = y;
        if (real) {
            const double a = literal ? th * 0.017453292519943295 : th;   // np.deg2rad = multiplication by pi / 180
            double sn, cs;
            sincos(a, &sn, &cs);
            x = x + v * cs * dt; y = y + v * sn * dt; th = th + w * dt;
        }
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

// The calling thread's current device is left as it was found (torch and other CUDA users of the thread keep theirs).
struct DeviceGuard {
    int prev;
    cudaError_t err;
    explicit DeviceGuard(int dev) : prev(-1) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        err = prev == dev ? cudaSuccess : cudaSetDevice(dev);
        if (prev == dev) prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define KMPC_ON_DEVICE(h)                 \
    DeviceGuard dev_guard_((h)->device);  \
    CU(dev_guard_.err)

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

static int cols_for(const kmpc_config *cfg) { return (cfg->B_max + 31) / 32 * 32; }

extern "C" int kmpc_version(void) { return KMPC_VERSION; }

extern "C" size_t kmpc_workspace_bytes(const kmpc_config *cfg) {
    if (!check_cfg(cfg)) return 0;
    Rows r = make_rows(cfg->N, cfg->O_max, 1);
    return (size_t)make_resto_rows(r).total * cols_for(cfg) * sizeof(double) + (size_t)4 * cols_for(cfg) * sizeof(int);
}

extern "C" const char *kmpc_last_error(const kmpc_handle *h) { return h ? h->err : g_err; }

extern "C" void kmpc_destroy(kmpc_handle *h) {
    if (!h) return;
    DeviceGuard dev_guard_(h->device);
    if (h->ws) cudaFree(h->ws);
    if (h->lists) cudaFree(h->lists);
    if (h->cnt) cudaFree(h->cnt);
    if (h->trips) cudaFree(h->trips);
    if (h->h_cnt) cudaFreeHost(h->h_cnt);
    if (h->wscratch) cudaFree(h->wscratch);
    if (h->okey) cudaFree(h->okey);
    if (h->env_obs) cudaFree(h->env_obs);
    if (h->env_rad) cudaFree(h->env_rad);
    if (h->resto_ws) cudaFree(h->resto_ws);
    if (h->resto_list) cudaFree(h->resto_list);
    if (h->env_idx) cudaFree(h->env_idx);
    if (h->oval) cudaFree(h->oval);
    if (h->osort_tmp) cudaFree(h->osort_tmp);
    if (h->cost_buf) cudaFree(h->cost_buf);
    if (h->d_in) cudaFree(h->d_in);
    if (h->d_out) cudaFree(h->d_out);
    if (h->d_iout) cudaFree(h->d_iout);
    if (h->h_in) cudaFreeHost(h->h_in);
    if (h->h_out) cudaFreeHost(h->h_out);
    if (h->h_iout) cudaFreeHost(h->h_iout);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    for (int i = 0; i < KMPC_LOOKAHEAD; ++i) if (h->evq[i]) cudaEventDestroy(h->evq[i]);
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
    { const char *om = getenv("KMPC_ORDER"); h->order_mode = om ? (atoi(om) != 0) : KMPC_ORDER_PRIOR; }   // measurement override
    h->device = cfg->device;
    h->rows = make_rows(cfg->N, cfg->O_max, 1);   // sized for stage-wise obstacle centres
    h->cols = cols_for(cfg);
    DeviceGuard dev_guard_(h->device);
    cudaError_t e = dev_guard_.err;
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&h->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);
    // (the thread solver's HBM workspace -- 1.2 GB at B_max = 65,536, N = 30 -- is allocated on its first use: the warp solver
    //  that serves N <= 63 keeps its state on chip and never touches it)
    if (e == cudaSuccess) e = cudaMalloc(&h->cnt, 8 * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&h->trips, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMallocHost(&h->h_cnt, KMPC_LOOKAHEAD * 4 * sizeof(int));
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev1);
    for (int i = 0; i < KMPC_LOOKAHEAD && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&h->evq[i], cudaEventDisableTiming);
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

static inline int nblocks(int n) { return (n + KMPC_TPB - 1) / KMPC_TPB; }

// Which solver serves a problem on this handle: the warp kernel whenever the horizon fits its stage slots (N + 1 <= 64) and the
// obstacle rows of one instance fit shared memory; otherwise the thread-per-instance fall-back.  Every entry point asks here,
// so what is decided before the launch (zero-copy result writes, the at-goal mask) is what runs.
static bool warp_path(const kmpc_handle *h, int O, int stagewise) {
    const int N1 = h->cfg.N + 1;
    if (N1 > 64 || getenv("KMPC_FORCE_THREAD") != NULL) return false;
    const int sw = (stagewise && O > 0) ? 1 : 0;
    const size_t one = N1 <= 32 ? WLay<1, 32>::bytes(1, O, sw) : N1 <= 52 ? WLay<2, 52>::bytes(1, O, sw) : WLay<2, 64>::bytes(1, O, sw);
    return one <= (size_t)h->max_smem;
}

struct SolveArgs {
    const double *x_cur, *goal, *X0, *U0, *obs, *orad;
    int O, stagewise;
    double obs_radius, inflation;
    double *X_out, *U_out, *obj;
    int32_t *status, *iters;
    const int32_t *active;
    int32_t *cost;   // optional: trips per instance (closed loops)
};

// One batch solve, asynchronous on `cuda_stream`.  Warp path: one persistent launch (+ the queue-order kernels).  Thread path:
// the trip loop is driven from the host, three launches per trip, KMPC_LOOKAHEAD trips in flight; the active-instance count
// of an older trip (async copy into pinned memory) sizes the grids and ends the loop.
static int solve_impl(kmpc_handle *h, int B, const SolveArgs &a, void *cuda_stream) {
    if (!h) return fail(NULL, KMPC_E_BADARG, "kmpc_solve: NULL handle%s", "");
    const int O = a.O;
    if (B < 0 || B > h->cfg.B_max) return fail(h, KMPC_E_BADARG, "kmpc_solve: B outside [0, B_max]%s", "");
    if (O < 0 || O > h->cfg.O_max || (O > 0 && !a.obs)) return fail(h, KMPC_E_BADARG, "kmpc_solve: bad obstacle arguments%s", "");
    if ((a.X0 == NULL) != (a.U0 == NULL)) return fail(h, KMPC_E_BADARG, "kmpc_solve: X0 and U0 must both be given or both be NULL%s", "");
    if (B == 0) return 0;
    if (!a.x_cur || !a.goal || !a.X_out || !a.U_out) return fail(h, KMPC_E_BADARG, "kmpc_solve: NULL required pointer%s", "");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    KMPC_ON_DEVICE(h);
    Cfg c;
    memset(&c, 0, sizeof c);
    const kmpc_config *cf = &h->cfg;
    c.N = cf->N; c.O = O; c.cost_mode = cf->cost_mode; c.gk_lo = cf->goal_k_lo; c.gk_hi = cf->goal_k_hi;
    c.max_iter = cf->max_iter; c.layout = cf->layout; c.B = B; c.obs_sw = (a.stagewise && O > 0) ? 1 : 0;
    relax_bounds(cf, &c);
    c.T = cf->T; c.W[0] = cf->W[0]; c.W[1] = cf->W[1]; c.W[2] = cf->W[2];
    c.Wvn = cf->Wv_neg; c.Wvp = cf->Wv_pos; c.Ww = cf->Ww; c.tol = cf->tol;
    c.obs_radius = a.obs_radius; c.dL = a.inflation - K_BOUND_RELAX * fmax(1.0, fabs(a.inflation));
    c.L = make_rows(cf->N, O, c.obs_sw);
    c.nb = (cf->N + 1) * (c.hasL[0] + c.hasU[0] + c.hasL[1] + c.hasU[1]) + cf->N * (c.hasL[2] + c.hasU[2] + c.hasL[3] + c.hasU[3]) + cf->N * O;
    c.m = 3 * (cf->N + 1) + cf->N * O;
    c.r_mnb = 1.0 / (double)(c.m + c.nb); c.r_nb = c.nb ? 1.0 / (double)c.nb : 0.0; c.mu_floor = cfg_mu_floor(c.tol);
    IO io;
    memset(&io, 0, sizeof io);
    io.x_cur = a.x_cur; io.goal = a.goal; io.X0 = a.X0; io.U0 = a.U0; io.obs = a.obs; io.orad = O > 0 ? a.orad : NULL;
    io.X_out = a.X_out; io.U_out = a.U_out; io.obj = a.obj; io.status = a.status; io.iters = a.iters; io.active = a.active; io.cost_out = a.cost;
    const size_t S = (size_t)h->cols;
    Lists ls;
    ls.LA[0] = h->lists; ls.LA[1] = h->lists + S; ls.LT[0] = h->lists + 2 * S; ls.LT[1] = h->lists + 3 * S;
    ls.cnt = h->cnt; ls.trips = h->timing ? h->trips : NULL;

    const bool use_warp = warp_path(h, O, c.obs_sw);
    if (a.active && !use_warp)
        return fail(h, KMPC_E_BADARG, "kmpc: the at-goal mask needs the warp solver (N <= 63, obstacle rows within shared memory, no KMPC_FORCE_THREAD)%s", "");
    if (h->timing) { CU(cudaMemsetAsync(h->trips, 0, sizeof(unsigned long long), st)); CU(cudaEventRecord(h->ev0, st)); }
    h->last_path = use_warp ? 1 : 0;
    if (use_warp) {
        // warp-per-instance path: one persistent launch, instances pulled from a queue, no workspace traffic
        CU(cudaMemsetAsync(h->cnt, 0, 8 * sizeof(int), st));   // queue head + restoration hand-over count (cnt[4])
        bool full = true;
        for (int i = 0; i < 4; ++i) full = full && c.hasL[i] && c.hasU[i];
        cudaError_t le;
        const int dv = h->device, sms = h->sm_count;
#ifdef KMPC_TUNE_HEADLINE   /* tuning builds only (scripts/variants.sh): nothing but the N <= 31, box-bounds, no-obstacle kernels -- compiles in a sixth of the time */
#define KMPC_LAUNCH(SPL, NST, OBS, WPB, MINB) kmpc_tune_launch<SPL, NST, OBS, WPB, MINB>(full, h, dv, sms, B, c, io, h->cnt, ls.trips, st)
#else
#define KMPC_LAUNCH(SPL, NST, OBS, WPB, MINB)                                                                          \
    (full ? launch_warp_kernel<SPL, NST, true, OBS, WPB, MINB>(h, dv, sms, B, c, io, h->cnt, ls.trips, st)                 \
          : launch_warp_kernel<SPL, NST, false, OBS, WPB, MINB>(h, dv, sms, B, c, io, h->cnt, ls.trips, st))
#endif
        // stage slots per field: 32 (N <= 31), 52 (N <= 51, e.g. the N = 50 configuration), 64 (N <= 63)
        // (a batch that fits one wave of 4-instance blocks runs those: 222 registers per thread instead of 128, nothing spilled, and a
        //  lone instance's trip is 12 % shorter -- B = 1: 321 -> 284 us at N = 30, 147 -> 124 us at N = 7)
        if (cf->N + 1 <= 32) le = O > 0 ? KMPC_LAUNCH(1, 32, true, KMPC_WPBO, KMPC_MINBO)
                                  : (B <= KMPC_WPB_SMALL * sms ? KMPC_LAUNCH(1, 32, false, KMPC_WPB_SMALL, 1) : KMPC_LAUNCH(1, 32, false, KMPC_WPB1, KMPC_MINB1));
        else if (cf->N + 1 <= 52) le = O > 0 ? KMPC_LAUNCH(2, 52, true, 6, 1) : KMPC_LAUNCH(2, 52, false, KMPC_WPB2, KMPC_MINB2);
        else le = O > 0 ? KMPC_LAUNCH(2, 64, true, 6, 1) : KMPC_LAUNCH(2, 64, false, KMPC_WPB3, KMPC_MINB3);
#undef KMPC_LAUNCH
        CU(le);
        CU(cudaGetLastError());
        h->launches++;
        h->last_host_trips = 0;
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
    if (!h->lists) {   // first thread-solver solve on this handle (both buffers are committed together)
        double *ws = NULL; int *li = NULL;
        cudaError_t e = cudaMalloc(&ws, (size_t)make_resto_rows(h->rows).total * h->cols * sizeof(double));   // (solver rows + the rows of the restoration phase)
        if (e == cudaSuccess) e = cudaMalloc(&li, (size_t)4 * h->cols * sizeof(int));
        if (e != cudaSuccess) { cudaFree(ws); cudaFree(li); CU(e); }
        h->ws = ws; h->lists = li;
    }
    ls.LA[0] = h->lists; ls.LA[1] = h->lists + S; ls.LT[0] = h->lists + 2 * S; ls.LT[1] = h->lists + 3 * S;
    const int cnt0[4] = {B, 0, 0, 0};
    CU(cudaMemcpyAsync(h->cnt, cnt0, sizeof cnt0, cudaMemcpyHostToDevice, st));
    kmpc_init_kernel<<<nblocks(B), KMPC_TPB, 0, st>>>(c, io, h->ws, S, ls.LA[0]);
    h->launches++;
    int ub = B;  // upper bound of the number of unfinished instances (non-increasing over the trips)
    int t = 0;
    const bool trace = getenv("KMPC_TRACE") != NULL;  // debugging aid: per-trip device time + active counts on stderr
    cudaEvent_t tev[2] = {NULL, NULL};
    if (trace) { cudaEventCreate(&tev[0]); cudaEventCreate(&tev[1]); }
    for (;; ++t) {
        const int p = t & 1, q = t % KMPC_LOOKAHEAD;
        if (t >= KMPC_LOOKAHEAD) {
            CU(cudaEventSynchronize(h->evq[q]));
            const int *hc = h->h_cnt + 4 * q;  // counts after trip t - LOOKAHEAD (parity p: its "next" lists are 1-p)
            const int active = hc[1 - p] + hc[2 + (1 - p)];
            if (active == 0) break;
            ub = active;
        }
        const int g = nblocks(ub);
        if (trace) cudaEventRecord(tev[0], st);
        if (O > 0) {
            kmpc_sweep_kernel<true><<<g, KMPC_TPB, 0, st>>>(c, io, h->ws, S, ls, p);
            kmpc_rollout_kernel<true><<<g, KMPC_TPB, 0, st>>>(c, h->ws, S, ls, p);
            kmpc_trial_kernel<true><<<g, KMPC_TPB, 0, st>>>(c, io, h->ws, S, ls, p);
        } else {
            kmpc_sweep_kernel<false><<<g, KMPC_TPB, 0, st>>>(c, io, h->ws, S, ls, p);
            kmpc_rollout_kernel<false><<<g, KMPC_TPB, 0, st>>>(c, h->ws, S, ls, p);
            kmpc_trial_kernel<false><<<g, KMPC_TPB, 0, st>>>(c, io, h->ws, S, ls, p);
        }
        h->launches += 3;
        if (trace) {
            int hc[4]; float ms = 0;
            cudaEventRecord(tev[1], st); cudaEventSynchronize(tev[1]); cudaEventElapsedTime(&ms, tev[0], tev[1]);
            cudaMemcpy(hc, h->cnt, sizeof hc, cudaMemcpyDeviceToHost);
            fprintf(stderr, "kmpc trip %4d  %8.3f ms  sweep %6d  trial %6d  -> next sweep %6d trial %6d\n", t, ms, hc[p], hc[2 + p], hc[1 - p], hc[2 + 1 - p]);
        }
        CU(cudaMemcpyAsync(h->h_cnt + 4 * q, h->cnt, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaEventRecord(h->evq[q], st));
        // the lists of parity p are consumed: empty them for trip t+1's appends
        CU(cudaMemsetAsync(h->cnt + p, 0, sizeof(int), st));
        CU(cudaMemsetAsync(h->cnt + 2 + p, 0, sizeof(int), st));
    }
    CU(cudaGetLastError());
    if (trace) { cudaEventDestroy(tev[0]); cudaEventDestroy(tev[1]); }
    h->last_host_trips = t;
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

static SolveArgs solve_args(const double *x_cur, const double *goal, const double *X0, const double *U0, const double *obs, int O, int stagewise,
                            double obs_radius, const double *obs_radii, double inflation, double *X_out, double *U_out, double *obj_out,
                            int32_t *status_out, int32_t *iters_out, const int32_t *active) {
    SolveArgs a;
    a.x_cur = x_cur; a.goal = goal; a.X0 = X0; a.U0 = U0; a.obs = obs; a.orad = obs_radii; a.O = O; a.stagewise = stagewise;
    a.obs_radius = obs_radius; a.inflation = inflation; a.X_out = X_out; a.U_out = U_out; a.obj = obj_out; a.status = status_out;
    a.iters = iters_out; a.active = active; a.cost = NULL;
    return a;
}

extern "C" int kmpc_solve(kmpc_handle *h, int B, const double *x_cur, const double *goal, const double *X0, const double *U0,
                          const double *obs_centers, int O, double obs_radius, const double *obs_radii, double inflation, double *X_out,
                          double *U_out, double *obj_out, int32_t *status_out, int32_t *iters_out, void *cuda_stream) {
    return solve_impl(h, B, solve_args(x_cur, goal, X0, U0, obs_centers, O, 0, obs_radius, obs_radii, inflation, X_out, U_out, obj_out, status_out,
                                       iters_out, NULL), cuda_stream);
}

extern "C" int kmpc_solve_tracks(kmpc_handle *h, int B, const double *x_cur, const double *goal, const double *X0, const double *U0,
                                 const double *obs_tracks, int O, double obs_radius, const double *obs_radii, double inflation, double *X_out,
                                 double *U_out, double *obj_out, int32_t *status_out, int32_t *iters_out, void *cuda_stream) {
    return solve_impl(h, B, solve_args(x_cur, goal, X0, U0, obs_tracks, O, 1, obs_radius, obs_radii, inflation, X_out, U_out, obj_out, status_out,
                                       iters_out, NULL), cuda_stream);
}

// pinned + device staging of kmpc_solve_host; all six buffers are committed to the handle together or not at all
static int ensure_staging(kmpc_handle *h) {
    if (h->h_iout) return 0;
    const kmpc_config *cf = &h->cfg;
    const size_t Bm = cf->B_max, N = cf->N, O = cf->O_max;
    const size_t in_doubles = Bm * (6 + 5 * N + 3 + 3 * O), out_doubles = Bm * (5 * N + 3 + 1);
    double *d_in = NULL, *d_out = NULL, *h_in = NULL, *h_out = NULL;
    int32_t *d_iout = NULL, *h_iout = NULL;
    cudaError_t e = cudaMalloc(&d_in, in_doubles * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_out, out_doubles * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_iout, Bm * 2 * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaHostAlloc(&h_in, in_doubles * sizeof(double), cudaHostAllocPortable | cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostAlloc(&h_out, out_doubles * sizeof(double), cudaHostAllocPortable | cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostAlloc(&h_iout, Bm * 2 * sizeof(int32_t), cudaHostAllocPortable | cudaHostAllocMapped);
    if (e != cudaSuccess) {
        cudaFree(d_in); cudaFree(d_out); cudaFree(d_iout);
        if (h_in) cudaFreeHost(h_in);
        if (h_out) cudaFreeHost(h_out);
        if (h_iout) cudaFreeHost(h_iout);
        CU(e);
    }
    h->in_doubles = in_doubles; h->out_doubles = out_doubles;
    h->d_in = d_in; h->d_out = d_out; h->d_iout = d_iout; h->h_in = h_in; h->h_out = h_out; h->h_iout = h_iout;
    return 0;
}

// Host-pointer solve.  X_out == NULL: results stay in the handle's pinned buffers (kmpc_host_result).  When ext_* are given
// (kmpc_solve_host_into) the solver writes into caller-owned pinned memory instead and nothing is copied afterwards.
static int solve_host_impl(kmpc_handle *h, int B, const double *x_cur, const double *goal, const double *X0, const double *U0,
                           const double *obs_centers, int O, double obs_radius, const double *obs_radii, double inflation, double *X_out,
                           double *U_out, double *obj_out, int32_t *status_out, int32_t *iters_out, int pinned_out, int do_sync) {
    if (!h) return fail(NULL, KMPC_E_BADARG, "kmpc_solve_host: NULL handle%s", "");
    if (B < 0 || B > h->cfg.B_max) return fail(h, KMPC_E_BADARG, "kmpc_solve_host: B outside [0, B_max]%s", "");
    if (O < 0 || O > h->cfg.O_max || (O > 0 && !obs_centers)) return fail(h, KMPC_E_BADARG, "kmpc_solve_host: bad obstacle arguments%s", "");
    if ((X0 == NULL) != (U0 == NULL)) return fail(h, KMPC_E_BADARG, "kmpc_solve_host: X0 and U0 must both be given or both be NULL%s", "");
    if (B == 0) return 0;
    if (!x_cur || !goal || ((X_out == NULL) != (U_out == NULL))) return fail(h, KMPC_E_BADARG, "kmpc_solve_host: NULL required pointer%s", "");
    if (pinned_out && !X_out) return fail(h, KMPC_E_BADARG, "kmpc_solve_host_into: X_out and U_out are required%s", "");
    KMPC_ON_DEVICE(h);
    int rc = ensure_staging(h);
    if (rc) return rc;
    const size_t N = h->cfg.N, b = B;
    const size_t nX = b * 3 * (N + 1), nU = b * 2 * N;
    // pack inputs into the pinned buffer: [x_cur | goal | X0 | U0 | obs | radii]
    size_t o = 0;
    memcpy(h->h_in + o, x_cur, b * 3 * sizeof(double)); o += b * 3;
    memcpy(h->h_in + o, goal, b * 3 * sizeof(double)); o += b * 3;
    size_t oX = 0, oU = 0, oO = 0, oR = 0;
    if (X0) { oX = o; memcpy(h->h_in + o, X0, nX * sizeof(double)); o += nX; oU = o; memcpy(h->h_in + o, U0, nU * sizeof(double)); o += nU; }
    if (O) { oO = o; memcpy(h->h_in + o, obs_centers, b * 2 * O * sizeof(double)); o += b * 2 * O; }
    if (O && obs_radii) { oR = o; memcpy(h->h_in + o, obs_radii, b * O * sizeof(double)); o += b * O; }
    // Inputs: one H2D copy of the packed buffer -- or, for a handful of instances (the single-agent call of agent.py:139-152), none at
    // all: the kernel reads the few hundred bytes straight from the mapped pinned buffer, which saves a stream operation (~5 us) per call.
    const bool zc_in = b * (6 + (X0 ? 5 * N + 3 : 0) + 3 * (size_t)O) <= 4096;
    double *din = zc_in ? h->h_in : h->d_in;
    if (!zc_in) CU(cudaMemcpyAsync(h->d_in, h->h_in, o * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    // Results: the warp solver writes each finished instance straight into pinned host memory (device-visible under unified
    // addressing), so the 81 MB of D2H traffic at B = 65,536 rides over PCIe while later instances are still being solved; the
    // thread solver (piecemeal 8-byte writes) goes through device staging and one copy.
    const bool direct = warp_path(h, O, 0) && getenv("KMPC_STAGED_D2H") == NULL;
    if (pinned_out && !direct) return fail(h, KMPC_E_BADARG, "kmpc_solve_host_into: needs the warp solver (N <= 63, obstacle rows within shared memory)%s", "");
    double *dX, *dU, *dobj;
    int32_t *dst, *dit;
    if (pinned_out) { dX = X_out; dU = U_out; dobj = obj_out; dst = status_out; dit = iters_out; }
    else { dX = direct ? h->h_out : h->d_out; dU = dX + nX; dobj = dX + nX + nU; dst = direct ? h->h_iout : h->d_iout; dit = dst + b; }
    rc = solve_impl(h, B, solve_args(din, din + b * 3, X0 ? din + oX : NULL, X0 ? din + oU : NULL, O ? din + oO : NULL, O, 0,
                                     obs_radius, (O && obs_radii) ? din + oR : NULL, inflation, dX, dU, dobj, dst, dit, NULL), h->stream);
    if (rc) return rc;
    if (!direct) {
        CU(cudaMemcpyAsync(h->h_out, h->d_out, (nX + nU + b) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaMemcpyAsync(h->h_iout, h->d_iout, b * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    }
    if (!do_sync) return 0;
    CU(cudaStreamSynchronize(h->stream));
    if (pinned_out) return 0;
    h->host_B = B;
    if (!X_out) return 0;  // zero-copy: the caller reads the pinned buffers through kmpc_host_result
    memcpy(X_out, h->h_out, nX * sizeof(double));
    memcpy(U_out, h->h_out + nX, nU * sizeof(double));
    if (obj_out) memcpy(obj_out, h->h_out + nX + nU, b * sizeof(double));
    if (status_out) memcpy(status_out, h->h_iout, b * sizeof(int32_t));
    if (iters_out) memcpy(iters_out, h->h_iout + b, b * sizeof(int32_t));
    return 0;
}

extern "C" int kmpc_solve_host(kmpc_handle *h, int B, const double *x_cur, const double *goal, const double *X0, const double *U0,
                               const double *obs_centers, int O, double obs_radius, const double *obs_radii, double inflation, double *X_out,
                               double *U_out, double *obj_out, int32_t *status_out, int32_t *iters_out) {
    return solve_host_impl(h, B, x_cur, goal, X0, U0, obs_centers, O, obs_radius, obs_radii, inflation, X_out, U_out, obj_out, status_out,
                           iters_out, 0, 1);
}

extern "C" int kmpc_solve_host_into(kmpc_handle *h, int B, const double *x_cur, const double *goal, const double *X0, const double *U0,
                                    const double *obs_centers, int O, double obs_radius, const double *obs_radii, double inflation,
                                    double *X_pinned, double *U_pinned, double *obj_pinned, int32_t *status_pinned, int32_t *iters_pinned) {
    return solve_host_impl(h, B, x_cur, goal, X0, U0, obs_centers, O, obs_radius, obs_radii, inflation, X_pinned, U_pinned, obj_pinned,
                           status_pinned, iters_pinned, 1, 0);
}

extern "C" int kmpc_host_sync(kmpc_handle *h) {
    if (!h) return fail(NULL, KMPC_E_BADARG, "kmpc_host_sync: NULL handle%s", "");
    KMPC_ON_DEVICE(h);
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int kmpc_host_result(kmpc_handle *h, const double **X, const double **U, const double **obj, const int32_t **status,
                                const int32_t **iters) {
    if (!h) return fail(NULL, KMPC_E_BADARG, "kmpc_host_result: NULL handle%s", "");
    if (!h->h_out || h->host_B <= 0) return fail(h, KMPC_E_BADARG, "kmpc_host_result: no kmpc_solve_host result on this handle%s", "");
    const size_t N = h->cfg.N, b = h->host_B, nX = b * 3 * (N + 1), nU = b * 2 * N;
    if (X) *X = h->h_out;
    if (U) *U = h->h_out + nX;
    if (obj) *obj = h->h_out + nX + nU;
    if (status) *status = h->h_iout;
    if (iters) *iters = h->h_iout + b;
    return 0;
}

// ---- buffers shared between the devices / processes of one box (the gather of a sharded batch) ----
extern "C" int kmpc_pinned_alloc(size_t bytes, void **out) {
    if (!out || bytes == 0) return KMPC_E_BADARG;
    *out = NULL;
    return cudaHostAlloc(out, bytes, cudaHostAllocPortable | cudaHostAllocMapped) == cudaSuccess ? 0 : KMPC_E_NOMEM;
}
extern "C" int kmpc_pinned_free(void *p) { return (!p || cudaFreeHost(p) == cudaSuccess) ? 0 : KMPC_E_CUDA; }

extern "C" int kmpc_shared_buffer_create(kmpc_handle *h, size_t bytes, void **dptr, unsigned char *ipc_handle_out) {
    if (!h || !dptr || bytes == 0) return fail(h, KMPC_E_BADARG, "kmpc_shared_buffer_create: bad arguments%s", "");
    KMPC_ON_DEVICE(h);
    void *p = NULL;
    CU(cudaMalloc(&p, bytes));
    if (ipc_handle_out) {
        cudaIpcMemHandle_t mh;
        cudaError_t e = cudaIpcGetMemHandle(&mh, p);
        if (e != cudaSuccess) { cudaFree(p); CU(e); }
        static_assert(sizeof(cudaIpcMemHandle_t) == KMPC_IPC_HANDLE_BYTES, "ipc handle size");
        memcpy(ipc_handle_out, &mh, sizeof mh);
    }
    *dptr = p;
    return 0;
}

extern "C" int kmpc_shared_buffer_open(kmpc_handle *h, const unsigned char *ipc_handle, void **dptr) {
    if (!h || !dptr || !ipc_handle) return fail(h, KMPC_E_BADARG, "kmpc_shared_buffer_open: bad arguments%s", "");
    KMPC_ON_DEVICE(h);
    cudaIpcMemHandle_t mh;
    memcpy(&mh, ipc_handle, sizeof mh);
    CU(cudaIpcOpenMemHandle(dptr, mh, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

extern "C" int kmpc_shared_buffer_close(kmpc_handle *h, void *dptr, int owner) {
    if (!h) return fail(NULL, KMPC_E_BADARG, "kmpc_shared_buffer_close: NULL handle%s", "");
    if (!dptr) return 0;
    KMPC_ON_DEVICE(h);
    if (owner) { CU(cudaFree(dptr)); } else { CU(cudaIpcCloseMemHandle(dptr)); }
    return 0;
}

// peer access from this handle's device to `peer_device` (same process): afterwards kmpc_solve on this handle may be given
// result pointers that live on the peer -- every finished instance is then written over NVLink straight into the gather buffer
extern "C" int kmpc_enable_peer(kmpc_handle *h, int peer_device) {
    if (!h) return fail(NULL, KMPC_E_BADARG, "kmpc_enable_peer: NULL handle%s", "");
    if (peer_device == h->device) return 0;
    KMPC_ON_DEVICE(h);
    int can = 0;
    CU(cudaDeviceCanAccessPeer(&can, h->device, peer_device));
    if (!can) return fail(h, KMPC_E_CUDA, "kmpc_enable_peer: no peer access between the devices%s", "");
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
    CU(e);
    return 0;
}

extern "C" int kmpc_agent_handoff(kmpc_handle *h, int B, const double *X, const double *U, double *x_cur, double *applied_out,
                                  void *cuda_stream) {
    if (!h) return fail(NULL, KMPC_E_BADARG, "kmpc_agent_handoff: NULL handle%s", "");
    if (B < 0 || !X || !U || !x_cur) return fail(h, KMPC_E_BADARG, "kmpc_agent_handoff: bad arguments%s", "");
    if (B == 0) return 0;
    KMPC_ON_DEVICE(h);
    kmpc_handoff_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)cuda_stream>>>(B, h->cfg.N, h->cfg.layout, X, U, x_cur, applied_out, NULL, NULL, 0.0, 0.0);
    CU(cudaGetLastError());
    h->launches++;
    return 0;
}

#ifdef KMPC_SCHED_TRACE
// tuning builds only: device buffer (4 x u64 per instance) the warp kernel records its schedule in; NULL switches it off
extern "C" int kmpc_debug_sched_trace(void *buf) {
    unsigned long long *p = (unsigned long long *)buf;
    return cudaMemcpyToSymbol(g_sched, &p, sizeof p) == cudaSuccess ? 0 : -1;
}
#endif

#ifdef KMPC_PHASE_TIMING
// tuning builds only: read and reset the phase timers of kmpc_warp.cuh
extern "C" int kmpc_debug_phase_cycles(double *out) {
    unsigned long long h[KMPC_NPHASE], z[KMPC_NPHASE] = {0};
    if (cudaMemcpyFromSymbol(h, g_phase_cycles, sizeof h) != cudaSuccess) return -1;
    cudaMemcpyToSymbol(g_phase_cycles, z, sizeof z);
    for (int i = 0; i < KMPC_NPHASE; ++i) out[i] = (double)h[i];
    return KMPC_NPHASE;
}
#endif

extern "C" int kmpc_select_obstacles(kmpc_handle *h, int B, int M, const double *x_cur, const double *cand_centers,
                                     const double *cand_radius, double sensor_radius, int literal, int O, double pad_x, double pad_y,
                                     double *obs_out, int32_t *count_out, int32_t *index_out, double *radius_out, void *cuda_stream) {
    if (!h) return fail(NULL, KMPC_E_BADARG, "kmpc_select_obstacles: NULL handle%s", "");
    if (B < 0 || M < 0 || O < 1 || O > KMPC_SEL_MAX_O) return fail(h, KMPC_E_BADARG, "kmpc_select_obstacles: need B, M >= 0 and 1 <= O <= 32%s", "");
    if (B == 0) return 0;
    if (!x_cur || !obs_out || (M > 0 && (!cand_centers || !cand_radius))) return fail(h, KMPC_E_BADARG, "kmpc_select_obstacles: NULL required pointer%s", "");
    KMPC_ON_DEVICE(h);
    kmpc_select_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)cuda_stream>>>(B, M, h->cfg.layout, x_cur, cand_centers, 2, cand_radius, sensor_radius,
                                                                               literal, O, O, 0, pad_x, pad_y, obs_out, count_out, index_out, radius_out);
    CU(cudaGetLastError());
    h->launches++;
    return 0;
}

extern "C" int kmpc_predict_tracks(kmpc_handle *h, int B, int O, int M, const int32_t *index, const double *state, const double *lin_vel,
                                  const double *ang_vel, double dt, int literal, double pad_x, double pad_y, double *tracks_out,
                                  void *cuda_stream) {
    if (!h) return fail(NULL, KMPC_E_BADARG, "kmpc_predict_tracks: NULL handle%s", "");
    if (B < 0 || O < 1 || M < 0 || (!index && O > M)) return fail(h, KMPC_E_BADARG, "kmpc_predict_tracks: need B, M >= 0, O >= 1 (and O <= M without an index)%s", "");
    if (B == 0) return 0;
    if (!tracks_out || (M > 0 && (!state || !lin_vel || !ang_vel))) return fail(h, KMPC_E_BADARG, "kmpc_predict_tracks: NULL required pointer%s", "");
    KMPC_ON_DEVICE(h);
    const size_t n = (size_t)B * O;
    kmpc_predict_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)cuda_stream>>>(B, O, O, 0, M, h->cfg.N, h->cfg.layout, index, state, lin_vel,
                                                                                          ang_vel, dt, literal, pad_x, pad_y, tracks_out);
    CU(cudaGetLastError());
    h->launches++;
    return 0;
}

// closed loops: where step s records what every agent's solve cost (trips), and -- from the second step on -- the previous step's record
// as the queue-order key of this one (kmpc_order_hint_kernel).  KMPC_ORDER_NATURAL switches the ordering off as everywhere else.
static int loop_cost_buffers(kmpc_handle *h, int s, int32_t **cost_now) {
    if (!h->cost_buf) {
        CU(cudaMalloc(&h->cost_buf, (size_t)2 * h->cols * sizeof(int32_t)));
        CU(cudaMemset(h->cost_buf, 0, (size_t)2 * h->cols * sizeof(int32_t)));
    }
    *cost_now = h->cost_buf + (size_t)(s & 1) * h->cols;
    h->order_hint = (s > 0 && getenv("KMPC_NO_ORDER_HINT") == NULL) ? h->cost_buf + (size_t)((s - 1) & 1) * h->cols : NULL;
    return 0;
}

extern "C" int kmpc_closed_loop(kmpc_handle *h, int B, int steps, double *x_cur, const double *goal, double *X, double *U,
                                double *applied_log, int32_t *iters_log, int32_t *status_log, int32_t *active, double goal_radius,
                                double agent_radius, void *cuda_stream) {
    if (!h) return fail(NULL, KMPC_E_BADARG, "kmpc_closed_loop: NULL handle%s", "");
    if (B < 0 || B > h->cfg.B_max || steps < 0) return fail(h, KMPC_E_BADARG, "kmpc_closed_loop: bad B or steps%s", "");
    if (B == 0 || steps == 0) return 0;
    if (!x_cur || !goal || !X || !U) return fail(h, KMPC_E_BADARG, "kmpc_closed_loop: NULL required pointer%s", "");
    KMPC_ON_DEVICE(h);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    for (int s = 0; s < steps; ++s) {
        // in place: every instance reads its own warm-start rows before it writes its result rows
        SolveArgs sa = solve_args(x_cur, goal, X, U, NULL, 0, 0, 0.0, NULL, 0.0, X, U, NULL, status_log ? status_log + (size_t)s * B : NULL,
                                  iters_log ? iters_log + (size_t)s * B : NULL, active);
        int rc = loop_cost_buffers(h, s, &sa.cost);
        if (rc) return rc;
        rc = solve_impl(h, B, sa, cuda_stream);
        h->order_hint = NULL;
        if (rc) return rc;
        kmpc_handoff_kernel<<<(B + 255) / 256, 256, 0, st>>>(B, h->cfg.N, h->cfg.layout, X, U, x_cur,
                                                            applied_log ? applied_log + (size_t)s * B * 2 : NULL, goal, active, goal_radius, agent_radius);
        CU(cudaGetLastError());
        h->launches++;
    }
    return 0;
}

extern "C" int kmpc_environment_loop(kmpc_handle *h, int B, int steps, double *x_cur, const double *goal, double *X, double *U, int M,
                                     const double *cand_centers, const double *cand_radius, int O, int Md, const double *dyn_state,
                                     const double *dyn_radius, const double *dyn_lin_vel, const double *dyn_ang_vel, int Od, int use_tracks,
                                     double track_dt, int literal_heading, double sensor_radius, int literal, double inflation, double pad_x,
                                     double pad_y, double *applied_log, int32_t *iters_log, int32_t *status_log, int32_t *count_log,
                                     int32_t *dyn_count_log, int32_t *active, double goal_radius, double agent_radius, void *cuda_stream) {
    if (!h) return fail(NULL, KMPC_E_BADARG, "kmpc_environment_loop: NULL handle%s", "");
    if (B < 0 || B > h->cfg.B_max || steps < 0 || M < 0 || Md < 0) return fail(h, KMPC_E_BADARG, "kmpc_environment_loop: bad B, steps, M or Md%s", "");
    if (O < 0 || Od < 0 || O + Od < 1 || O + Od > h->cfg.O_max || O > KMPC_SEL_MAX_O || Od > KMPC_SEL_MAX_O)
        return fail(h, KMPC_E_BADARG, "kmpc_environment_loop: need 1 <= O + Od <= O_max, each <= 32%s", "");
    if (B == 0 || steps == 0) return 0;
    if (!x_cur || !goal || !X || !U || (M > 0 && O > 0 && (!cand_centers || !cand_radius)) ||
        (Md > 0 && Od > 0 && (!dyn_state || !dyn_radius)) || (use_tracks && Md > 0 && Od > 0 && (!dyn_lin_vel || !dyn_ang_vel)))
        return fail(h, KMPC_E_BADARG, "kmpc_environment_loop: NULL required pointer%s", "");
    KMPC_ON_DEVICE(h);
    const int Ot = O + Od, N = h->cfg.N;
    const bool tracks = use_tracks && Od > 0;
    if (!h->env_rad) {   // kept circles of every agent: centres (or N-column tracks), radii, the dynamic candidates' indices
        double *eo = NULL, *er = NULL; int32_t *ei = NULL;
        cudaError_t e = cudaMalloc(&eo, (size_t)h->cols * h->cfg.O_max * 2 * sizeof(double) * (size_t)(N + 1));
        if (e == cudaSuccess) e = cudaMalloc(&er, (size_t)h->cols * h->cfg.O_max * sizeof(double));
        if (e == cudaSuccess) e = cudaMalloc(&ei, (size_t)h->cols * h->cfg.O_max * sizeof(int32_t));
        if (e != cudaSuccess) { cudaFree(eo); cudaFree(er); cudaFree(ei); CU(e); }
        h->env_obs = eo; h->env_rad = er; h->env_idx = ei;
    }
    double *cen = h->env_obs;                                          // [B][Ot][2] current centres
    double *trk = h->env_obs + (size_t)h->cols * h->cfg.O_max * 2;      // [B][Ot][N][2] tracks (use_tracks)
    cudaStream_t st = (cudaStream_t)cuda_stream;
    for (int s = 0; s < steps; ++s) {
        // ROSEnvironment.step (environment.py:39-80): sensor filters (static, then dynamic) -> EgoAgent.step (solve + hand-off) -> at-goal test
        if (O > 0) {
            kmpc_select_kernel<<<(B + 127) / 128, 128, 0, st>>>(B, M, h->cfg.layout, x_cur, cand_centers, 2, cand_radius, sensor_radius, literal, O, Ot, 0,
                                                               pad_x, pad_y, cen, count_log ? count_log + (size_t)s * B : NULL, NULL, h->env_rad);
            CU(cudaGetLastError());
            h->launches++;
        }
        if (Od > 0) {   // dynamic obstacles are filtered by where they are now (dynamic_obstacle.py: the centre of their Circle)
            kmpc_select_kernel<<<(B + 127) / 128, 128, 0, st>>>(B, Md, h->cfg.layout, x_cur, dyn_state, 3, dyn_radius, sensor_radius, literal, Od, Ot, O,
                                                               pad_x, pad_y, cen, dyn_count_log ? dyn_count_log + (size_t)s * B : NULL, h->env_idx, h->env_rad);
            CU(cudaGetLastError());
            h->launches++;
        }
        const double *obs = cen;
        if (tracks) {
            // static slots: N equal columns; dynamic slots: the constant-velocity prediction of the kept obstacle (dynamic_obstacle.py:20-37)
            if (O > 0) {
                kmpc_repeat_kernel<<<(unsigned)(((size_t)B * O + 127) / 128), 128, 0, st>>>(B, O, Ot, N, h->cfg.layout, cen, trk);
                h->launches++;
            }
            kmpc_predict_kernel<<<(unsigned)(((size_t)B * Od + 127) / 128), 128, 0, st>>>(B, Od, Ot, O, Md, N, h->cfg.layout, h->env_idx, dyn_state,
                                                                                         dyn_lin_vel, dyn_ang_vel, track_dt, literal_heading, pad_x, pad_y, trk);
            CU(cudaGetLastError());
            h->launches++;
            obs = trk;
        }
        SolveArgs sa = solve_args(x_cur, goal, X, U, obs, Ot, tracks ? 1 : 0, 0.0, h->env_rad, inflation, X, U, NULL,
                                  status_log ? status_log + (size_t)s * B : NULL, iters_log ? iters_log + (size_t)s * B : NULL, active);
        int rc = loop_cost_buffers(h, s, &sa.cost);
        if (rc) return rc;
        rc = solve_impl(h, B, sa, cuda_stream);
        h->order_hint = NULL;
        if (rc) return rc;
        kmpc_handoff_kernel<<<(B + 255) / 256, 256, 0, st>>>(B, h->cfg.N, h->cfg.layout, X, U, x_cur,
                                                            applied_log ? applied_log + (size_t)s * B * 2 : NULL, goal, active, goal_radius, agent_radius);
        CU(cudaGetLastError());
        h->launches++;
    }
    return 0;
}

#include "kmpc_map.inl"   // occupancy map -> packed circles (host code)

extern "C" int kmpc_set_queue_order(kmpc_handle *h, int mode) {
    if (!h || (mode != KMPC_ORDER_NATURAL && mode != KMPC_ORDER_PRIOR)) return KMPC_E_BADARG;
    h->order_mode = mode;
    return 0;
}

extern "C" int kmpc_set_timing(kmpc_handle *h, int enable) {
    if (!h) return KMPC_E_BADARG;
    h->timing = enable ? 1 : 0;
    return 0;
}

extern "C" int kmpc_get_stats(kmpc_handle *h, kmpc_stats *out) {
    if (!h || !out) return KMPC_E_BADARG;
    out->last_kernel_ms = h->last_ms; out->launches = h->launches; out->slots = h->cols; out->blocks = h->last_host_trips;
    out->threads_per_block = KMPC_TPB; out->sm_count = h->sm_count; out->trips = h->last_trips; out->warp_path = h->last_path;
    return 0;
}

extern "C" int kmpc_measure_fp64_peak(kmpc_handle *h, double *tflops_out) {
    if (!h || !tflops_out) return KMPC_E_BADARG;
    KMPC_ON_DEVICE(h);
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
