/*
 * kmpc.h -- C ABI of the B200-native batched unicycle-MPC solver (libkmpc.so).
 *
 * This is the drop-in boundary for the ONE hot path of rtarun1/kiss-mpc: the per-step NLP solve
 *     mpc/optimizer.py:319-400   MotionPlanner.solve(...)          (the call made by mpc/agent.py:139-152)
 *     mpc/optimizer.py:354       ca.nlpsol("solver", "ipopt", ...) (rebuilt on every call)
 *     mpc/optimizer.py:375-391   solver(x0=..., lbx, ubx, lbg, ubg, p=[current_state; goal_state])
 * Every entry point below names the reference interface it replaces.  Plain C types only: no torch, no C++.
 * All compute entry points launch hand-written sm_100a CUDA kernels; there is NO CPU fallback in this library.
 *
 * Conventions
 *   - float64 everywhere (the reference computes in float64 through CasADi/IPOPT).
 *   - B independent problem instances ("agents") per call; one warp solves one instance, the serial Riccati recursions of a
 *     block's instances run side by side on one warp; a thread-per-instance solver is the fall-back for N > 63.
 *   - I/O layouts, selected per handle by kmpc_config.layout:
 *       KMPC_LAYOUT_INSTANCE_MAJOR (0): x_cur[B][3], goal[B][3], X[B][3][N+1], U[B][2][N], obs[B][O][2]
 *            == what B stacked reference calls hold: states_matrix (3,N+1), controls_matrix (2,N) (optimizer.py:392-400)
 *       KMPC_LAYOUT_BATCH_MINOR    (1): x_cur[3][B], goal[3][B], X[3][N+1][B], U[2][N][B], obs[O][2][B]
 *            structure-of-arrays with the instance index fastest (fully coalesced loads/stores on the device).
 *   - per-instance solver outcome uses IPOPT's ApplicationReturnStatus numbering (the reference discards it,
 *     optimizer.py:375-400 reads only solution["x"]):
 *         0 Solve_Succeeded, 2 Infeasible_Problem_Detected (the restoration phase ended at a stationary point of the
 *         constraint violation), 4 Diverging_Iterates, -1 Maximum_Iterations_Exceeded, -2 Restoration_Failed (restoration
 *         entered at an almost feasible point or could not make progress), -3 Error_In_Step_Computation,
 *         -13 Invalid_Number_Detected, -199 Internal_Error (the 512-entry filter is full; never seen by the oracle either).
 *   - functions return 0 on success, a negative KMPC_E_* code otherwise; nothing throws across the ABI.
 */
#ifndef KMPC_H
#define KMPC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KMPC_VERSION 200

#define KMPC_LAYOUT_INSTANCE_MAJOR 0
#define KMPC_LAYOUT_BATCH_MINOR 1

#define KMPC_COST_README 0       /* README.md:15-27: W_v- min(0,v)^2 + W_v+ max(0,v)^2 */
#define KMPC_COST_CODE_LITERAL 1 /* optimizer.py:91-96: 300 * fmin(v, 0), linear */

#define KMPC_E_BADARG (-1)
#define KMPC_E_CUDA (-2)
#define KMPC_E_NOMEM (-3)
#define KMPC_E_NODEVICE (-4)

#define KMPC_NO_BOUND 1e19 /* |bound| >= 1e19 means "no bound" (IPOPT nlp_lower/upper_bound_inf) */

/* Problem + solver options.  Replaces MotionPlanner.__init__(time_step, horizon) (optimizer.py:40-77), the hard-coded
 * weights (optimizer.py:57-60), the IPOPT option dict (optimizer.py:344-352) and the bound tuples that
 * get_optimization_variable_bounds builds per call (optimizer.py:111-156). */
typedef struct kmpc_config {
    int32_t N;          /* horizon (optimizer.py:40)                                                       */
    int32_t O_max;      /* obstacle slots per instance (0 = none); optimizer.py:198-258                     */
    int32_t cost_mode;  /* KMPC_COST_*                                                                      */
    int32_t goal_k_lo;  /* goal cost over k = goal_k_lo..goal_k_hi: README 1..N, code 1..N-1 (optimizer.py:80) */
    int32_t goal_k_hi;
    int32_t max_iter;   /* optimizer.py:346 (2000)                                                          */
    int32_t B_max;      /* largest batch kmpc_solve will be called with (sizes the device workspace)        */
    int32_t layout;     /* KMPC_LAYOUT_*                                                                    */
    int32_t device;     /* CUDA device ordinal                                                              */
    int32_t reserved;
    double T;           /* time_step (optimizer.py:40)                                                      */
    double W[3];        /* optimizer.py:57   diag(100,100,50)                                               */
    double Wv_neg;      /* optimizer.py:59   300                                                            */
    double Wv_pos;      /* README.md:24      0                                                              */
    double Ww;          /* optimizer.py:60   10                                                             */
    double lo[4];       /* lower bounds of x, y, v, omega (theta is free: optimizer.py:114-115)             */
    double hi[4];
    double tol;         /* IPOPT tol (1e-8; optimizer.py:348 sets acceptable_tol to the same value)         */
} kmpc_config;

typedef struct kmpc_handle kmpc_handle;

/* Library identity / sizing (no reference equivalent). */
int kmpc_version(void);
size_t kmpc_workspace_bytes(const kmpc_config *cfg);

/* Replaces MotionPlanner(time_step, horizon) (optimizer.py:40; called once per agent at agent.py:62):
 * binds a device and sizes the solver for B_max instances (the thread-per-instance fall-back solver allocates its HBM workspace,
 * kmpc_workspace_bytes, on first use; the warp solver needs a few MB of scratch).  One handle = one device = one host thread at a time. */
int kmpc_create(const kmpc_config *cfg, kmpc_handle **out);
void kmpc_destroy(kmpc_handle *h);
const char *kmpc_last_error(const kmpc_handle *h);

/* Replaces MotionPlanner.solve (optimizer.py:319-400) for B instances at once.  DEVICE pointers owned by the caller;
 * asynchronous on `cuda_stream` (a cudaStream_t, NULL = default stream).
 *   x_cur, goal           current_state / goal_state          (optimizer.py:390  p = [current_state; goal_state])
 *   X0, U0                states_matrix / controls_matrix     (optimizer.py:376-385 primal warm start); both NULL = the
 *                         cold start of agent.py:59-60 (X = tile(x_cur), U = 0) synthesised on chip
 *   obs_centers, O        static/dynamic obstacle circle centres (optimizer.py:217-221), 0 <= O <= O_max; NULL iff O == 0
 *   obs_radius, obs_radii radius subtracted from the centre distance of every obstacle slot.  The reference keeps ONE radius per obstacle
 *                         class: static_obstacles[0].radius for the static columns, dynamic_obstacles[0].radius for the dynamic
 *                         ones (optimizer.py:231-250).  obs_radii ([B][O]; batch-minor [O][B]) carries a radius per instance and
 *                         slot, so the caller writes the first static radius into the static slots and the first dynamic radius
 *                         into the dynamic slots; obs_radii == NULL: every slot uses the scalar obs_radius.
 *   inflation             inflation_radius = lower bound of the distance rows (optimizer.py:254-258; agent.py:149)
 *   X_out, U_out          the returned (3,N+1) / (2,N) matrices (optimizer.py:392-400)
 *   obj_out, status_out, iters_out   new outputs (the reference never reads IPOPT's stats); each may be NULL. */
int kmpc_solve(kmpc_handle *h, int B, const double *x_cur, const double *goal, const double *X0, const double *U0,
               const double *obs_centers, int O, double obs_radius, const double *obs_radii, double inflation, double *X_out,
               double *U_out, double *obj_out, int32_t *status_out, int32_t *iters_out, void *cuda_stream);

/* The same solve with circle centres that move with the stage: obs_tracks holds, per instance and obstacle, a track of N
 * centres ([B][O][N][2]; SoA layout [O][N][2][B]); column t is the centre paired with X_{t+1}.  This is what the reference's
 * per-obstacle constraint path builds from DynamicObstacle.states_matrix (dynamic_obstacle.py:47-56 called from
 * optimizer.py:207-215; the vectorised path optimizer.py:217-250 keeps only the current centre).  A static obstacle is a track
 * of N equal columns.  Everything else as kmpc_solve. */
int kmpc_solve_tracks(kmpc_handle *h, int B, const double *x_cur, const double *goal, const double *X0, const double *U0,
                      const double *obs_tracks, int O, double obs_radius, const double *obs_radii, double inflation, double *X_out,
                      double *U_out, double *obj_out, int32_t *status_out, int32_t *iters_out, void *cuda_stream);

/* Same call with HOST pointers (what a ctypes/NumPy caller such as the reference's agent.py holds): stages through the
 * handle's pinned buffers, copies host->device, solves, copies device->host and synchronises before returning. */
int kmpc_solve_host(kmpc_handle *h, int B, const double *x_cur, const double *goal, const double *X0, const double *U0,
                    const double *obs_centers, int O, double obs_radius, const double *obs_radii, double inflation, double *X_out,
                    double *U_out, double *obj_out, int32_t *status_out, int32_t *iters_out);

/* The host path of ONE SHARD of a batch that is split over several devices (one handle per device): inputs are host
 * pointers as in kmpc_solve_host, the result pointers are slices of ONE caller-owned pinned buffer (kmpc_pinned_alloc) shared by
 * all the handles -- every device writes its finished instances straight into its slice, so the gather of the sharded batch costs
 * no extra copy.  Returns once the work is queued on the handle's own stream; kmpc_host_sync(h) waits for it.  obj / status /
 * iters may be NULL.  Needs the warp solver (N <= 63). */
int kmpc_solve_host_into(kmpc_handle *h, int B, const double *x_cur, const double *goal, const double *X0, const double *U0,
                         const double *obs_centers, int O, double obs_radius, const double *obs_radii, double inflation,
                         double *X_pinned, double *U_pinned, double *obj_pinned, int32_t *status_pinned, int32_t *iters_pinned);
int kmpc_host_sync(kmpc_handle *h);
int kmpc_pinned_alloc(size_t bytes, void **out); /* portable, device-mapped pinned host memory (visible to every device of the box) */
int kmpc_pinned_free(void *p);

/* Gather buffer of a batch sharded over the GPUs of one box, one PROCESS per GPU (the final result gather of SURVEY 8e; the
 * reference has no counterpart -- one NLP per agent, no coupling, optimizer.py:375-391).  The root rank creates the buffer on its
 * device and hands the 64-byte handle to the other ranks (any transport: torch.distributed broadcast); they open it and pass
 * pointers into it as the X_out / U_out / ... of kmpc_solve: every finished instance is then written over NVLink straight into
 * the root's memory -- compute and gather are one kernel, no collective.  kmpc_enable_peer does the same for several handles of
 * ONE process (peer access from h's device to peer_device). */
#define KMPC_IPC_HANDLE_BYTES 64
int kmpc_shared_buffer_create(kmpc_handle *h, size_t bytes, void **dptr, unsigned char *ipc_handle_out /* [64] or NULL */);
int kmpc_shared_buffer_open(kmpc_handle *h, const unsigned char *ipc_handle /* [64] */, void **dptr);
int kmpc_shared_buffer_close(kmpc_handle *h, void *dptr, int owner);
int kmpc_enable_peer(kmpc_handle *h, int peer_device);

/* Zero-copy variant of the host path: when kmpc_solve_host is called with X_out == U_out == NULL the results are left in
 * the handle's pinned staging buffers; this call returns their addresses (layout as kmpc_config.layout, sized for the
 * B of that call).  The pointers stay valid until the next solve on, or the destruction of, the handle.  Replaces the
 * `.full()` copies of optimizer.py:392-400 for callers that consume the result before the next solve. */
int kmpc_host_result(kmpc_handle *h, const double **X, const double **U, const double **obj, const int32_t **status,
                     const int32_t **iters);

/* Batched EgoAgent.step hand-off (agent.py:139-155 + agent.py:70-72): after a solve, on the device,
 *   applied[b] = U[:,0]  (agent.py:154-155),  x_cur[b] <- X[:,1] (the "perfect model" state hand-off, agent.py:70-72);
 * X/U stay in place as the next solve's UNSHIFTED warm start (agent.py:139-145).  Device pointers, handle layout. */
int kmpc_agent_handoff(kmpc_handle *h, int B, const double *X, const double *U, double *x_cur, double *applied_out,
                       void *cuda_stream);

/* Replaces the sensor filter of ROSEnvironment.step (environment.py:48-65) for B agents at once: of M candidate circles
 * (cand_centers[M][2], cand_radius[M], shared by all agents) every agent keeps those whose distance to its state is
 * <= sensor_radius (agent.py:101 default 5), nearest first, at most O of them.  Distance = Obstacle.calculate_distance
 * (obstacle.py:22-23 -> geometry.py:44): literal != 0 gives the formula as written, ||(p - c) - r||_2 (the radius is subtracted
 * from both components); literal == 0 the intended ||p - c||_2 - r.  Candidates at exactly equal distance
 * collapse to the LAST one (the reference keys a dict by distance, environment.py:48-51).  obs_out has the handle's obstacle
 * layout ([B][O][2] / [O][2][B]); unused slots get (pad_x, pad_y) -- pick a point far outside the workspace, its rows stay
 * inactive; count_out[B] (may be NULL) is the number of real obstacles per agent; index_out[B][O] (may be NULL) the
 * candidate index kept in every slot, -1 for padding; radius_out[B][O] (may be NULL; layout as obs_radii) gets, in every slot,
 * the radius of the NEAREST kept circle -- the list handed to the planner is sorted nearest-first and the planner reads the
 * radius of its first element for the whole class (optimizer.py:231-245).  Device pointers, asynchronous. */
int kmpc_select_obstacles(kmpc_handle *h, int B, int M, const double *x_cur, const double *cand_centers, const double *cand_radius,
                          double sensor_radius, int literal, int O, double pad_x, double pad_y, double *obs_out, int32_t *count_out,
                          int32_t *index_out, double *radius_out, void *cuda_stream);

/* Replaces DynamicObstacle._get_predicted_states_matrix (dynamic_obstacle.py:20-37) for the obstacles every agent kept:
 * M moving obstacles (state[M][3] = x, y, heading; lin_vel[M]; ang_vel[M]), index[B][O] = which obstacle sits in each
 * agent's slot (kmpc_select_obstacles' index_out; -1 = empty slot -> the padding point in every column; NULL = slot o is
 * obstacle o for every agent).  Column 0 of a track is the current position, column t = column t-1 advanced by
 * [v cos(a) dt, v sin(a) dt, omega dt] (dt = 0.1 in the reference, :21); literal != 0 keeps a = deg2rad(heading) on a heading
 * that is already in radians (:24-25), literal == 0 uses the heading itself.  tracks_out: N columns per slot in the
 * obs_tracks layout of kmpc_solve_tracks (N = the handle's horizon).  Device pointers, asynchronous. */
int kmpc_predict_tracks(kmpc_handle *h, int B, int O, int M, const int32_t *index, const double *state, const double *lin_vel,
                        const double *ang_vel, double dt, int literal, double pad_x, double pad_y, double *tracks_out,
                        void *cuda_stream);

/* Device-resident receding-horizon loop: `steps` repetitions of EgoAgent.step (agent.py:130-155) for B agents without
 * leaving the device:   solve(x_cur, goal, warm start = previous X, U UNSHIFTED, agent.py:139-145) -> applied = U[:,0]
 * (agent.py:154-155) -> x_cur <- X[:,1] (agent.py:70-72).  X, U are in/out (start values = the first warm start, e.g. the
 * cold start X = tile(x_cur), U = 0 of agent.py:59-60; on return the last solution), x_cur is in/out.  Optional per-step
 * records (NULL to skip): applied_log[steps][B][2] (handle layout per step), iters_log[steps][B], status_log[steps][B].
 * active (int32 [B], in/out, or NULL): agents with active[b] == 0 are not solved (status_log = 1000, iters_log = 0) and
 * keep their state -- the reference stops stepping an agent once its final goal is reached (environment.py:31-33);
 * with goal_radius > 0 the mask is refreshed after every step with Agent.at_goal (agent.py:78-80):
 * ||(goal_xy - p_xy) - agent_radius||_2 - goal_radius <= 0, the literal formula of geometry.py:44 (agent_radius = 0 gives
 * the Euclidean distance).  Asynchronous on `cuda_stream`; device pointers.  No obstacle rows in this entry point
 * (kmpc_environment_loop has them). */
int kmpc_closed_loop(kmpc_handle *h, int B, int steps, double *x_cur, const double *goal, double *X, double *U,
                     double *applied_log, int32_t *iters_log, int32_t *status_log, int32_t *active, double goal_radius,
                     double agent_radius, void *cuda_stream);

/* kmpc_closed_loop with the environment's sensor filters in the loop -- `steps` repetitions of ROSEnvironment.step
 * (environment.py:39-80) for B agents on the device.  Every step each agent keeps the (at most O) nearest of the M static candidate
 * circles within its sensor radius (environment.py:48-56) and the (at most Od) nearest of the Md dynamic obstacles
 * (environment.py:57-65; dyn_state[Md][3] = x, y, heading, dyn_radius[Md]; they are filtered by their current centre), solves with
 * the static slots followed by the dynamic slots as obstacle rows (optimizer.py:198-258; radius per class = the nearest kept
 * circle's, optimizer.py:231-250; inflation = agent radius + 0.1, agent.py:149; unused slots at (pad_x, pad_y)), hands off
 * (agent.py:139-155) and refreshes the at-goal mask (agent.py:78-80).  use_tracks == 0: every circle enters with its current centre,
 * as the reference's vectorised constraint path reads it (optimizer.py:217-221); use_tracks != 0: a dynamic slot is paired stage by
 * stage with the constant-velocity prediction of its obstacle (kmpc_predict_tracks: dyn_lin_vel / dyn_ang_vel [Md], track_dt,
 * literal_heading; dynamic_obstacle.py:20-37, :47-56) and a static slot is a track of N equal columns.  Obstacles themselves do
 * not move between steps (the reference never advances them).  count_log / dyn_count_log [steps][B] (may be NULL): how many real
 * static / dynamic circles each agent saw; the other arguments are kmpc_closed_loop's. */
int kmpc_environment_loop(kmpc_handle *h, int B, int steps, double *x_cur, const double *goal, double *X, double *U, int M,
                          const double *cand_centers, const double *cand_radius, int O, int Md, const double *dyn_state,
                          const double *dyn_radius, const double *dyn_lin_vel, const double *dyn_ang_vel, int Od, int use_tracks,
                          double track_dt, int literal_heading, double sensor_radius, int literal, double inflation, double pad_x,
                          double pad_y, double *applied_log, int32_t *iters_log, int32_t *status_log, int32_t *count_log,
                          int32_t *dyn_count_log, int32_t *active, double goal_radius, double agent_radius, void *cuda_stream);

/* Occupancy map -> packed circles: the static-obstacle candidates of a map.  Replaces the stand-alone script
 * obstacle_handling/static_obstacle.py:12-56 (threshold at 127 :23, distance transform of the occupied region :35, then greedily the
 * largest inscribed circle, its disc blanked, until the largest remaining distance is below MIN_RADIUS :38-57) and returns what the
 * script only paints: centres (x, y in pixels, raster order of discovery) and integer radii.  The arithmetic is OpenCV's and is
 * restated to the bit (OpenCV 4.x's float32 5x5 chamfer transform with weights 1 / 1.4 / 2.1969, first maximum in raster order, filled midpoint circle), so the circle
 * list equals the script's on the same image.  HOST pointers; image[h][w] 8-bit grey; at most max_circles are written, *count_out is
 * the number found.  One-off preprocessing on the host (no device is touched). */
/* the distance map alone (static_obstacle.py:23-35): depth of every occupied pixel inside the occupied region, float32 [h][w] */
int kmpc_map_distance(const unsigned char *image, int w, int h, int threshold, float *dist_out);
int kmpc_map_to_circles(const unsigned char *image, int w, int h, int threshold, double min_radius, int max_circles,
                        int32_t *centers_out /*[max_circles][2]*/, int32_t *radii_out /*[max_circles]*/, int32_t *count_out);

/* Scheduling of the batch inside kmpc_solve (no reference equivalent; results never depend on it).  The solver kernel is
 * persistent: warps pull instances from a queue, and interior-point iteration counts differ by more than 8x between instances,
 * so an instance that is fetched late and runs long sets the end of the launch.  KMPC_ORDER_PRIOR (default) hands the
 * instances out in descending order of a geometric prior of their iteration count (bearing of the goal from the start
 * heading, heading change, goal distance; csrc/kmpc_order_prior.h); KMPC_ORDER_NATURAL in index order. */
#define KMPC_ORDER_NATURAL 0
#define KMPC_ORDER_PRIOR 1
int kmpc_set_queue_order(kmpc_handle *h, int mode);

/* Measurement helpers (no reference equivalent). */
typedef struct kmpc_stats {
    double last_kernel_ms;    /* device time of the solver kernel of the last kmpc_solve on this handle (CUDA events on its stream) */
    int64_t launches;         /* kernels launched by this handle since creation */
    int32_t slots;            /* workspace columns (one per instance) */
    int32_t blocks;           /* host-driven solver trips (3 launches each) of the last solve */
    int32_t threads_per_block;
    int32_t sm_count;
    int64_t trips;            /* total solver-loop trips of the last solve (sum over threads), if timing enabled */
    int32_t warp_path;        /* 1: the last solve ran the warp kernel; 0: the thread-per-instance fall-back */
    int32_t reserved;
} kmpc_stats;
int kmpc_set_timing(kmpc_handle *h, int enable); /* enable: kmpc_solve records events + syncs to fill last_kernel_ms */
int kmpc_get_stats(kmpc_handle *h, kmpc_stats *out);
/* FP64 FMA-pipe micro-benchmark on the handle's device: returns measured DFMA TFLOP/s (the roofline denominator of
 * the solver kernel; MEASURED_PEAKS.json has no FP64 entry). */
int kmpc_measure_fp64_peak(kmpc_handle *h, double *tflops_out);

#ifdef __cplusplus
}
#endif
#endif /* KMPC_H */
