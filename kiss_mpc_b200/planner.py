"""Host side of the B200 batched MPC solver: the reference's planner interface over the C ABI (include/kmpc.h).

  * ``MotionPlanner``         drop-in for mpc/optimizer.py:39-400 (same constructor, same ``solve`` keywords and
                              return shapes), so mpc/agent.py:62/:139-152 can use it unchanged.  B = 1, NumPy in/out.
  * ``BatchedMotionPlanner``  the same solve for B instances at once on torch CUDA tensors (or NumPy / CPU tensors,
                              staged through pinned memory inside the library).
  * ``PlannerConfig``         the NLP the reference builds: README form by default, the
                              code-literal form of optimizer.py:79-156 via ``PlannerConfig.code_literal``.

PyTorch is used for device memory and streams only.  All arithmetic runs in libkmpc.so (hand-written sm_100a CUDA);
there is no CPU fallback: without the library or without a GPU these classes raise ``KmpcError``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, replace
from typing import NamedTuple, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import KmpcConfig, KmpcError, KmpcStats

INF = float("inf")

# IPOPT ApplicationReturnStatus values the solver can return (kmpc.h)
STATUS_NAMES = {0: "Solve_Succeeded", -1: "Maximum_Iterations_Exceeded", -2: "Restoration_Failed",
                -3: "Error_In_Step_Computation", 4: "Diverging_Iterates", -13: "Invalid_Number_Detected"}


@dataclass(frozen=True)
class PlannerConfig:
    """Problem + options.  Defaults: README.md:15-66 form, EgoAgent bounds (agent.py:104-106), optimizer.py:57-60 weights,
    optimizer.py:344-352 options."""
    N: int = 30                                        # horizon (optimizer.py:40)
    T: float = 0.1                                     # time_step (optimizer.py:40)
    W: Tuple[float, float, float] = (100.0, 100.0, 50.0)   # optimizer.py:57
    Wv_neg: float = 300.0                              # optimizer.py:59
    Wv_pos: float = 0.0                                # README.md:24
    Ww: float = 10.0                                   # optimizer.py:60
    cost_mode: str = "readme"                          # "readme" | "code_literal" (optimizer.py:91-96)
    goal_range: str = "readme"                         # "readme": k = 1..N (README.md:17) | "code": k = 1..N-1 (optimizer.py:80)
    x_bounds: Tuple[float, float] = (-20.0, 20.0)      # optimizer.py:114-115 / agent.py:106
    y_bounds: Tuple[float, float] = (-20.0, 20.0)      # README.md:61-66 (code: unbounded)
    v_bounds: Tuple[float, float] = (-0.2, 0.5)        # agent.py:104
    w_bounds: Tuple[float, float] = (-0.5, 0.5)        # agent.py:105
    O_max: int = 0                                     # obstacle slots per instance
    tol: float = 1e-8                                  # IPOPT tol; optimizer.py:348 acceptable_tol
    max_iter: int = 2000                               # optimizer.py:346

    @staticmethod
    def code_literal(**kw) -> "PlannerConfig":
        """The NLP exactly as optimizer.py writes it: goal cost k=1..N-1, linear 300*fmin(v,0), only x bounded."""
        return PlannerConfig(cost_mode="code_literal", goal_range="code", y_bounds=(-INF, INF), **kw)

    def to_c(self, B_max: int, layout: int, device: int) -> KmpcConfig:
        if self.cost_mode not in ("readme", "code_literal") or self.goal_range not in ("readme", "code"):
            raise ValueError("cost_mode must be 'readme'|'code_literal', goal_range 'readme'|'code'")
        c = KmpcConfig()
        c.N, c.O_max = int(self.N), int(self.O_max)
        c.cost_mode = _lib.COST_README if self.cost_mode == "readme" else _lib.COST_CODE_LITERAL
        c.goal_k_lo, c.goal_k_hi = 1, (self.N if self.goal_range == "readme" else self.N - 1)
        c.max_iter, c.B_max, c.layout, c.device = int(self.max_iter), int(B_max), int(layout), int(device)
        c.T = float(self.T)
        c.W = (C.c_double * 3)(*[float(v) for v in self.W])
        c.Wv_neg, c.Wv_pos, c.Ww = float(self.Wv_neg), float(self.Wv_pos), float(self.Ww)
        lo = [self.x_bounds[0], self.y_bounds[0], self.v_bounds[0], self.w_bounds[0]]
        hi = [self.x_bounds[1], self.y_bounds[1], self.v_bounds[1], self.w_bounds[1]]
        c.lo = (C.c_double * 4)(*[max(float(v), -1e20) for v in lo])
        c.hi = (C.c_double * 4)(*[min(float(v), 1e20) for v in hi])
        c.tol = float(self.tol)
        return c


class SolveResult(NamedTuple):
    states: object      # [B,3,N+1]  (optimizer.py:392-395)
    controls: object    # [B,2,N]    (optimizer.py:396-399)
    objective: object   # [B] unscaled objective value
    status: object      # [B] int32, IPOPT ApplicationReturnStatus numbering
    iters: object       # [B] int32 interior-point iterations


def _torch():
    import torch
    return torch


class BatchedMotionPlanner:
    """B independent MotionPlanner.solve calls in one kernel launch.

    Tensor layout "instance_major" (default): states [B,3,N+1], controls [B,2,N] -- B stacked reference matrices.
    Layout "batch_minor": states [3,N+1,B], controls [2,N,B], x/goal [3,B], obstacles [O,2,B] (coalesced device I/O).
    """

    def __init__(self, config: PlannerConfig = PlannerConfig(), max_batch: int = 65536, device: int = 0,
                 layout: str = "instance_major"):
        self.config = config
        self.max_batch = int(max_batch)
        self.device = int(device)
        self.layout = {"instance_major": _lib.LAYOUT_INSTANCE_MAJOR, "batch_minor": _lib.LAYOUT_BATCH_MINOR}[layout]
        self._L = _lib.load()
        self._h = C.c_void_p()
        cc = config.to_c(self.max_batch, self.layout, self.device)
        rc = self._L.kmpc_create(C.byref(cc), C.byref(self._h))
        if rc != 0:
            msg = self._L.kmpc_last_error(None)
            self._h = C.c_void_p()
            raise KmpcError(f"kmpc_create failed (rc={rc}): {msg.decode() if msg else ''}")

    # -- lifetime -------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.kmpc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_queue_order(self, prior: bool = True):
        """Order in which the persistent solver kernel takes the instances of a batch: likely-long instances first (a geometric
        prior of the iteration count, default) or index order.  Scheduling only; results are identical."""
        _lib.check(self._L.kmpc_set_queue_order(self._h, 1 if prior else 0), self._h, "kmpc_set_queue_order")

    # -- shapes ---------------------------------------------------------------------------------
    def _shapes(self, B: int, O: int):
        N = self.config.N
        if self.layout == _lib.LAYOUT_INSTANCE_MAJOR:
            return (B, 3), (B, 3, N + 1), (B, 2, N), (B, O, 2), (B, 2)
        return (3, B), (3, N + 1, B), (2, N, B), (O, 2, B), (2, B)

    def _batch_of(self, x) -> int:
        return int(x.shape[0] if self.layout == _lib.LAYOUT_INSTANCE_MAJOR else x.shape[-1])

    # -- solve ----------------------------------------------------------------------------------
    def solve(self, current_state, goal_state, states_matrix=None, controls_matrix=None, obstacles=None,
              obstacle_radius: float = 0.3, inflation_radius: float = 0.0, copy: bool = True) -> SolveResult:
        """current_state/goal_state [B,3]; states_matrix [B,3,N+1] and controls_matrix [B,2,N] = primal warm start
        (both None: the cold start of agent.py:59-60); obstacles [B,O,2] circle centres, or [B,O,N,2] centre tracks (column t
        paired with X_{t+1}, dynamic_obstacle.py:47-56; kmpc_solve_tracks).  CUDA tensors stay on the
        device (asynchronous on the current torch stream); NumPy arrays / CPU tensors go through kmpc_solve_host.
        Host path only: ``copy=False`` returns NumPy views of the planner's pinned result buffers (no 80 MB memcpy at
        B = 65,536); they are overwritten by the next solve on this planner."""
        torch = _torch()
        if isinstance(current_state, torch.Tensor) and current_state.is_cuda:
            return self._solve_device(current_state, goal_state, states_matrix, controls_matrix, obstacles,
                                      obstacle_radius, inflation_radius)
        if obstacles is not None and np.ndim(obstacles) == 4:   # tracks: staged through torch, solved by kmpc_solve_tracks
            dev = torch.device("cuda", self.device)
            up = lambda a: None if a is None else torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
            r = self._solve_device(up(current_state), up(goal_state), up(states_matrix), up(controls_matrix), up(obstacles),
                                   obstacle_radius, inflation_radius)
            return SolveResult(*[t.cpu().numpy() for t in r])
        return self._solve_host(current_state, goal_state, states_matrix, controls_matrix, obstacles, obstacle_radius,
                                inflation_radius, copy)

    def _check_O(self, obstacles):
        if obstacles is None:
            return 0
        O = int(obstacles.shape[1] if self.layout == _lib.LAYOUT_INSTANCE_MAJOR else obstacles.shape[0])
        if O == 0:
            return 0
        if O > self.config.O_max:
            raise ValueError(f"{O} obstacles > O_max={self.config.O_max} of this planner")
        return O

    def _solve_device(self, x, goal, X0, U0, obs, obs_radius, inflation) -> SolveResult:
        torch = _torch()
        B = self._batch_of(x)
        O = self._check_O(obs)
        sx, sX, sU, sO, _ = self._shapes(B, O)
        dev = torch.device("cuda", self.device)

        def prep(t, shape, name):
            if t is None:
                return None
            if t.device != dev or t.dtype != torch.float64:
                raise ValueError(f"{name}: expected float64 tensor on {dev}")
            if tuple(t.shape) != shape:
                raise ValueError(f"{name}: expected shape {shape}, got {tuple(t.shape)}")
            return t.contiguous()

        x, goal = prep(x, sx, "current_state"), prep(goal, sx, "goal_state")
        X0, U0 = prep(X0, sX, "states_matrix"), prep(U0, sU, "controls_matrix")
        tracks = O > 0 and obs.dim() == 4
        if tracks:   # [B,O,N,2] / [O,N,2,B]
            sO = (B, O, self.config.N, 2) if self.layout == _lib.LAYOUT_INSTANCE_MAJOR else (O, self.config.N, 2, B)
        obs = prep(obs, sO, "obstacles") if O else None
        if (X0 is None) != (U0 is None):
            raise ValueError("states_matrix and controls_matrix must both be given or both be None")
        with torch.cuda.device(dev):
            Xo = torch.empty(sX, dtype=torch.float64, device=dev)
            Uo = torch.empty(sU, dtype=torch.float64, device=dev)
            obj = torch.empty(B, dtype=torch.float64, device=dev)
            st = torch.empty(B, dtype=torch.int32, device=dev)
            it = torch.empty(B, dtype=torch.int32, device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
            fn = self._L.kmpc_solve_tracks if tracks else self._L.kmpc_solve
            rc = fn(self._h, B, p(x), p(goal), p(X0), p(U0), p(obs), O, float(obs_radius), float(inflation),
                    p(Xo), p(Uo), p(obj), p(st), p(it), C.c_void_p(stream))
        _lib.check(rc, self._h, "kmpc_solve_tracks" if tracks else "kmpc_solve")
        return SolveResult(Xo, Uo, obj, st, it)

    def _solve_host(self, x, goal, X0, U0, obs, obs_radius, inflation, copy=True) -> SolveResult:
        def np64(a):
            if a is None:
                return None
            if hasattr(a, "detach"):
                a = a.detach().cpu().numpy()
            return np.ascontiguousarray(a, dtype=np.float64)

        x, goal, X0, U0, obs = np64(x), np64(goal), np64(X0), np64(U0), np64(obs)
        B = self._batch_of(x)
        O = self._check_O(obs)
        sx, sX, sU, sO, _ = self._shapes(B, O)
        for a, s, n in ((x, sx, "current_state"), (goal, sx, "goal_state"), (X0, sX, "states_matrix"),
                        (U0, sU, "controls_matrix"), (obs if O else None, sO, "obstacles")):
            if a is not None and a.shape != s:
                raise ValueError(f"{n}: expected shape {s}, got {a.shape}")
        if (X0 is None) != (U0 is None):
            raise ValueError("states_matrix and controls_matrix must both be given or both be None")
        p = lambda a: None if a is None else C.c_void_p(a.ctypes.data)
        if not copy:
            rc = self._L.kmpc_solve_host(self._h, B, p(x), p(goal), p(X0), p(U0), p(obs) if O else None, O, float(obs_radius),
                                         float(inflation), None, None, None, None, None)
            _lib.check(rc, self._h, "kmpc_solve_host")
            ptr = [C.c_void_p() for _ in range(5)]
            _lib.check(self._L.kmpc_host_result(self._h, *[C.byref(q) for q in ptr]), self._h, "kmpc_host_result")

            def view(q, shape, ct, dt):
                n = int(np.prod(shape))
                return np.frombuffer((ct * n).from_address(q.value), dtype=dt).reshape(shape)

            return SolveResult(view(ptr[0], sX, C.c_double, np.float64), view(ptr[1], sU, C.c_double, np.float64),
                               view(ptr[2], (B,), C.c_double, np.float64), view(ptr[3], (B,), C.c_int32, np.int32),
                               view(ptr[4], (B,), C.c_int32, np.int32))
        Xo = np.empty(sX); Uo = np.empty(sU); obj = np.empty(B); st = np.empty(B, np.int32); it = np.empty(B, np.int32)
        rc = self._L.kmpc_solve_host(self._h, B, p(x), p(goal), p(X0), p(U0), p(obs) if O else None, O, float(obs_radius),
                                     float(inflation), p(Xo), p(Uo), p(obj), p(st), p(it))
        _lib.check(rc, self._h, "kmpc_solve_host")
        return SolveResult(Xo, Uo, obj, st, it)

    # -- closed loop (agent.py:139-155) -------------------------------------------------------------
    def agent_handoff(self, states, controls, current_state, applied=None):
        """In place on the device: applied <- U[:,0] (agent.py:154-155), current_state <- X[:,1] (agent.py:70-72)."""
        torch = _torch()
        B = self._batch_of(current_state)
        stream = torch.cuda.current_stream(torch.device("cuda", self.device)).cuda_stream
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        rc = self._L.kmpc_agent_handoff(self._h, B, p(states), p(controls), p(current_state), p(applied), C.c_void_p(stream))
        _lib.check(rc, self._h, "kmpc_agent_handoff")

    def select_obstacles(self, current_state, centers, radii, sensor_radius: float = 5.0, slots: Optional[int] = None,
                         literal_distance: bool = True, pad_center=(1.0e6, 1.0e6), return_index: bool = False):
        """Batched sensor filter of ROSEnvironment.step (environment.py:48-65): per agent the candidate circles
        (centers [M,2], radii [M], CUDA tensors) within `sensor_radius` (agent.py:101), nearest first, at most `slots`
        (default O_max).  Returns (obstacles [B,slots,2] ready for ``solve(obstacles=...)``, count [B]); unused slots hold
        `pad_center`, whose rows stay inactive.  literal_distance: geometry.py:44 as written (True) or ||p-c|| - r.
        return_index: also return index [B,slots] int32, the candidate kept in every slot (-1 = padding)."""
        torch = _torch()
        dev = torch.device("cuda", self.device)
        B = self._batch_of(current_state)
        O = int(slots if slots is not None else self.config.O_max)
        M = int(centers.shape[0])
        _, _, _, sO, _ = self._shapes(B, O)
        out = torch.empty(sO, dtype=torch.float64, device=dev)
        cnt = torch.empty(B, dtype=torch.int32, device=dev)
        idx = torch.empty((B, O), dtype=torch.int32, device=dev) if return_index else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        rc = self._L.kmpc_select_obstacles(self._h, B, M, p(current_state.contiguous()), p(centers.contiguous()), p(radii.contiguous()),
                                           float(sensor_radius), 1 if literal_distance else 0, O, float(pad_center[0]), float(pad_center[1]),
                                           p(out), p(cnt), p(idx), C.c_void_p(stream))
        _lib.check(rc, self._h, "kmpc_select_obstacles")
        return (out, cnt, idx) if return_index else (out, cnt)

    def predict_tracks(self, batch: int, obstacle_state, linear_velocity, angular_velocity, index=None, slots: Optional[int] = None,
                       dt: float = 0.1, literal_heading: bool = True, pad_center=(1.0e6, 1.0e6)):
        """Batched DynamicObstacle._get_predicted_states_matrix (dynamic_obstacle.py:20-37): N-column constant-velocity tracks
        of the moving obstacles every agent kept.  obstacle_state [M,3] (x, y, heading), linear_velocity [M], angular_velocity
        [M] (CUDA float64); index [B,slots] int32 from ``select_obstacles(..., return_index=True)`` (None: slot o = obstacle o
        for every agent).  literal_heading keeps the reference's deg2rad of a radian heading (:24-25).  Returns tracks
        [B,slots,N,2] ready for ``solve(obstacles=...)``."""
        torch = _torch()
        dev = torch.device("cuda", self.device)
        M = int(obstacle_state.shape[0])
        O = int(slots if slots is not None else (index.shape[1] if index is not None else M))
        N = self.config.N
        shape = (batch, O, N, 2) if self.layout == _lib.LAYOUT_INSTANCE_MAJOR else (O, N, 2, batch)
        out = torch.empty(shape, dtype=torch.float64, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        p = lambda t: None if t is None else C.c_void_p(t.contiguous().data_ptr())
        rc = self._L.kmpc_predict_tracks(self._h, int(batch), O, M, p(index), p(obstacle_state), p(linear_velocity), p(angular_velocity),
                                         float(dt), 1 if literal_heading else 0, float(pad_center[0]), float(pad_center[1]), p(out),
                                         C.c_void_p(stream))
        _lib.check(rc, self._h, "kmpc_predict_tracks")
        return out

    def closed_loop(self, current_state, goal_state, steps: int, states_matrix=None, controls_matrix=None,
                    log_applied: bool = True, log_iters: bool = True, goal_radius: float = 0.0, agent_radius: float = 0.0,
                    active=None, obstacle_centers=None, obstacle_radii=None, sensor_radius: float = 5.0, slots: Optional[int] = None,
                    obstacle_radius: float = 0.3, inflation_radius: float = 0.5, literal_distance: bool = True,
                    pad_center=(1.0e6, 1.0e6)):
        """`steps` receding-horizon steps of EgoAgent.step (agent.py:130-155) for all B agents on the device
        (kmpc_closed_loop): warm start = previous solution unshifted, x <- X[:,1], applied = U[:,0].
        current_state [B,3] is advanced in place.  Returns (states, controls, applied_log [steps,B,2] | None,
        iters_log [steps,B] | None, status_log [steps,B]).
        goal_radius > 0: an agent that satisfies Agent.at_goal (agent.py:78-80, literal distance of geometry.py:44 with
        agent_radius; 0 = Euclidean) after a step is not solved again (status 1000 in the log), as the reference's
        environment stops stepping it (environment.py:31-33); `active` (int32 [B]) carries that mask in and out.
        obstacle_centers [M,2] / obstacle_radii [M] (CUDA float64): the whole ROSEnvironment.step (environment.py:39-80,
        kmpc_environment_loop) -- every step each agent keeps the `slots` nearest circles within `sensor_radius` and solves with
        them as obstacle rows; the per-step obstacle counts are left in ``self.last_obstacle_counts`` [steps,B]."""
        torch = _torch()
        dev = torch.device("cuda", self.device)
        B = self._batch_of(current_state)
        sx, sX, sU, _, sA = self._shapes(B, 0)
        N = self.config.N
        if states_matrix is None:            # agent.py:59-60
            if self.layout == _lib.LAYOUT_INSTANCE_MAJOR:
                states_matrix = current_state[:, :, None].repeat(1, 1, N + 1).contiguous()
            else:
                states_matrix = current_state[:, None, :].repeat(1, N + 1, 1).contiguous()
            controls_matrix = torch.zeros(sU, dtype=torch.float64, device=dev)
        X, U = states_matrix.contiguous(), controls_matrix.contiguous()
        applied = torch.empty((steps,) + sA, dtype=torch.float64, device=dev) if log_applied else None
        iters = torch.empty((steps, B), dtype=torch.int32, device=dev) if log_iters else None
        status = torch.empty((steps, B), dtype=torch.int32, device=dev)
        if active is None and goal_radius > 0:
            active = torch.ones(B, dtype=torch.int32, device=dev)
        self.last_active = active
        stream = torch.cuda.current_stream(dev).cuda_stream
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        if obstacle_centers is not None:
            O = int(slots if slots is not None else self.config.O_max)
            counts = torch.empty((steps, B), dtype=torch.int32, device=dev)
            cen, rad = obstacle_centers.contiguous(), obstacle_radii.contiguous()
            rc = self._L.kmpc_environment_loop(self._h, B, int(steps), p(current_state), p(goal_state.contiguous()), p(X), p(U),
                                               int(cen.shape[0]), p(cen), p(rad), float(sensor_radius), 1 if literal_distance else 0, O,
                                               float(obstacle_radius), float(inflation_radius), float(pad_center[0]), float(pad_center[1]),
                                               p(applied), p(iters), p(status), p(counts), p(active), float(goal_radius),
                                               float(agent_radius), C.c_void_p(stream))
            _lib.check(rc, self._h, "kmpc_environment_loop")
            self.last_obstacle_counts = counts
            return X, U, applied, iters, status
        rc = self._L.kmpc_closed_loop(self._h, B, int(steps), p(current_state), p(goal_state.contiguous()), p(X), p(U), p(applied),
                                      p(iters), p(status), p(active), float(goal_radius), float(agent_radius), C.c_void_p(stream))
        _lib.check(rc, self._h, "kmpc_closed_loop")
        return X, U, applied, iters, status

    # -- measurement --------------------------------------------------------------------------------
    def set_timing(self, enable: bool):
        self._L.kmpc_set_timing(self._h, 1 if enable else 0)

    def stats(self) -> dict:
        s = KmpcStats()
        _lib.check(self._L.kmpc_get_stats(self._h, C.byref(s)), self._h, "kmpc_get_stats")
        return {k: getattr(s, k) for k, _ in KmpcStats._fields_}

    def measure_fp64_peak(self) -> float:
        v = C.c_double()
        _lib.check(self._L.kmpc_measure_fp64_peak(self._h, C.byref(v)), self._h, "kmpc_measure_fp64_peak")
        return v.value


class MotionPlanner:
    """Drop-in for the reference's ``MotionPlanner`` (mpc/optimizer.py:39): same constructor (optimizer.py:40), same
    ``solve`` keyword arguments (optimizer.py:319-333, called at agent.py:139-152) and return value
    (``(states (3,N+1), controls (2,N))`` float64 ndarrays, optimizer.py:400).  The solve runs on the GPU.

    ``problem_form``: "readme" (default; README.md:15-66: goal cost k=1..N, squared velocity penalty, x and y bounded by
    ``state_bounds``) or "code_literal" (optimizer.py as written: k=1..N-1, 300*fmin(v,0), only x bounded).
    After each call ``last_status`` / ``last_iterations`` / ``last_objective`` hold what IPOPT's stats would (the reference
    never reads them, optimizer.py:375-400).
    ``use_obstacle_tracks``: False (default) reads only the current centre of every obstacle, as the reference's vectorised
    constraint path does (optimizer.py:217-221); True pairs X_{t+1} with column t of a dynamic obstacle's ``states_matrix``
    (what DynamicObstacle.calculate_symbolic_matrix_distance builds, dynamic_obstacle.py:47-56) when it has >= N columns.
    """

    def __init__(self, time_step: float, horizon: int, problem_form: str = "readme", device: int = 0,
                 use_obstacle_tracks: bool = False):
        self.use_obstacle_tracks = bool(use_obstacle_tracks)
        self.time_step = float(time_step)
        self.horizon = int(horizon)
        self.num_states, self.num_controls = 3, 2          # optimizer.py:44-55
        self.problem_form = problem_form
        self.device = device
        self._planner: Optional[BatchedMotionPlanner] = None
        self._key = None
        self.last_status = None
        self.last_iterations = None
        self.last_objective = None

    @staticmethod
    def _centers(obstacles) -> list:
        # optimizer.py:217-221: only Circle geometry (.center, .radius) is read; duck-typed, no casadi import
        out = []
        for ob in obstacles:
            g = getattr(ob, "geometry", ob)
            out.append((tuple(np.asarray(g.center, dtype=float).reshape(-1)[:2]), float(g.radius)))
        return out

    def solve(self, current_state, current_linear_velocity=None, current_angular_velocity=None, goal_state=None,
              states_matrix=None, controls_matrix=None, state_bounds=(-20.0, 20.0), linear_velocity_bounds=(-0.2, 0.5),
              angular_velocity_bounds=(-0.5, 0.5), static_obstacles=(), dynamic_obstacles=(), inflation_radius=None):
        # current_linear_velocity / current_angular_velocity are accepted and unused, as in the reference
        N = self.horizon
        obs = self._centers(list(static_obstacles) + list(dynamic_obstacles))
        O = len(obs)
        sb = (float(state_bounds[0]), float(state_bounds[1]))
        key = (sb, tuple(map(float, linear_velocity_bounds)), tuple(map(float, angular_velocity_bounds)), O)
        if self._planner is None or key[:3] != self._key[:3] or O > self._key[3]:
            base = PlannerConfig() if self.problem_form == "readme" else PlannerConfig.code_literal()
            cfg = replace(base, N=N, T=self.time_step, x_bounds=sb, v_bounds=key[1], w_bounds=key[2], O_max=O,
                          y_bounds=sb if self.problem_form == "readme" else (-INF, INF))
            if self._planner is not None:
                self._planner.close()
            self._planner = BatchedMotionPlanner(cfg, max_batch=1, device=self.device)
            self._key = key
        x = np.asarray(current_state, dtype=np.float64).reshape(1, 3)
        g = np.asarray(goal_state, dtype=np.float64).reshape(1, 3)
        X0 = None if states_matrix is None else np.asarray(states_matrix, dtype=np.float64).reshape(1, 3, N + 1)
        U0 = None if controls_matrix is None else np.asarray(controls_matrix, dtype=np.float64).reshape(1, 2, N)
        if (X0 is None) != (U0 is None):
            raise ValueError("states_matrix and controls_matrix must both be given")
        centers = np.array([c for c, _ in obs], dtype=np.float64).reshape(1, O, 2) if O else None
        if O and self.use_obstacle_tracks:
            centers = np.ascontiguousarray(np.repeat(centers[:, :, None, :], N, axis=2))          # static: N equal columns
            for j, ob in enumerate(dynamic_obstacles, start=len(static_obstacles)):
                sm = getattr(ob, "states_matrix", None)
                if sm is not None and np.shape(sm)[1] >= N:
                    centers[0, j] = np.asarray(sm, dtype=np.float64)[:2, :N].T
        radius = obs[0][1] if O else 0.0                      # optimizer.py:231-245: the first obstacle's radius for all
        infl = float(inflation_radius) if inflation_radius is not None else 0.0   # optimizer.py:362
        r = self._planner.solve(x, g, X0, U0, centers, radius, infl)
        self.last_status = int(r.status[0]); self.last_iterations = int(r.iters[0]); self.last_objective = float(r.objective[0])
        return np.array(r.states[0]), np.array(r.controls[0])


# ---- multi-GPU: batch slices, one process per GPU, no collective inside the iteration ----------------
def shard_range(B: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of a B-instance batch owned by `rank` (ceil split; trailing ranks may be empty)."""
    per = -(-B // world)
    lo = min(B, rank * per)
    return lo, min(B, lo + per)


def gather_results(local: SolveResult, B: int, group=None) -> Optional[SolveResult]:
    """The one exchange of the sharded path: gather every rank's slice on rank 0 (torch.distributed; NCCL on GPUs,
    gloo in the CPU tests).  Slices are padded to the ceil-split size so a plain all_gather works for ragged B."""
    torch = _torch()
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = -(-B // world)
    out = []
    for t in local:
        t = torch.as_tensor(t)
        pad = torch.zeros((per,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad, group=group)
        if rank == 0:
            parts = [bufs[r][: shard_range(B, r, world)[1] - shard_range(B, r, world)[0]] for r in range(world)]
            out.append(torch.cat(parts, 0))
    return SolveResult(*out) if rank == 0 else None
