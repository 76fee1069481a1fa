"""Host side of the B200 batched MPC solver: the reference's planner interface over the C ABI (include/kmpc.h).

  * ``MotionPlanner``         drop-in for mpc/optimizer.py:39-400 (same constructor, same ``solve`` keywords and
                              return shapes), so mpc/agent.py:62/:139-152 can use it unchanged.  B = 1, NumPy in/out.
  * ``BatchedMotionPlanner``  the same solve for B instances at once on torch CUDA tensors (or NumPy / CPU tensors,
                              staged through pinned memory inside the library).
  * ``ShardedMotionPlanner``  one batch split over several GPUs of the box by contiguous slices (one handle + stream per device,
                              one call, one result buffer; no collective -- instances are independent, optimizer.py:375-391).
  * ``PlannerConfig``         the NLP the reference builds: README form by default, the
                              code-literal form of optimizer.py:79-156 via ``PlannerConfig.code_literal``.

PyTorch is used for device memory and streams only.  All arithmetic runs in libkmpc.so (hand-written sm_100a CUDA);
there is no CPU fallback: without the library or without a GPU these classes raise ``KmpcError``.
"""
from __future__ import annotations

import ctypes as C
import warnings
from dataclasses import dataclass, replace
from typing import NamedTuple, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import KmpcConfig, KmpcError, KmpcStats

INF = float("inf")

# IPOPT ApplicationReturnStatus values the solver can return (kmpc.h)
STATUS_NAMES = {0: "Solve_Succeeded", 2: "Infeasible_Problem_Detected", -1: "Maximum_Iterations_Exceeded", -2: "Restoration_Failed",
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

    # -- argument checks: everything handed to the C ABI as a raw address is validated here --------
    def _dev(self, t, shape, name, dtype=None, optional=False):
        """`t` as a contiguous tensor of `dtype` (float64) on this planner's device with exactly `shape`; ValueError otherwise
        (the kernels read and write raw addresses: a float32, CPU or strided tensor would be an out-of-bounds access)."""
        torch = _torch()
        if t is None:
            if optional:
                return None
            raise ValueError(f"{name}: required")
        dtype = dtype or torch.float64
        dev = torch.device("cuda", self.device)
        if not isinstance(t, torch.Tensor) or t.device != dev or t.dtype != dtype:
            raise ValueError(f"{name}: expected a {dtype} tensor on {dev}")
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
        return t.contiguous()

    def _inplace(self, t, shape, name, dtype=None, optional=False):
        """Like _dev for buffers the library writes: must already be contiguous (a copy would swallow the result)."""
        r = self._dev(t, shape, name, dtype, optional)
        if r is not None and r.data_ptr() != t.data_ptr():
            raise ValueError(f"{name}: must be contiguous (it is updated in place)")
        return r

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
              obstacle_radius=0.3, inflation_radius: float = 0.0, copy: bool = True) -> SolveResult:
        """current_state/goal_state [B,3]; states_matrix [B,3,N+1] and controls_matrix [B,2,N] = primal warm start
        (both None: the cold start of agent.py:59-60); obstacles [B,O,2] circle centres, or [B,O,N,2] centre tracks (column t
        paired with X_{t+1}, dynamic_obstacle.py:47-56; kmpc_solve_tracks); obstacle_radius: one float for every circle, or an
        array [B,O] (batch-minor [O,B]) with a radius per instance and slot -- the reference keeps one radius per obstacle
        class, the first static / first dynamic obstacle's (optimizer.py:231-250).  CUDA tensors stay on the
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
            rad = obstacle_radius if np.ndim(obstacle_radius) == 0 else up(obstacle_radius)
            r = self._solve_device(up(current_state), up(goal_state), up(states_matrix), up(controls_matrix), up(obstacles),
                                   rad, inflation_radius)
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
        x, goal = self._dev(x, sx, "current_state"), self._dev(goal, sx, "goal_state")
        X0, U0 = self._dev(X0, sX, "states_matrix", optional=True), self._dev(U0, sU, "controls_matrix", optional=True)
        tracks = O > 0 and obs.dim() == 4
        if tracks:   # [B,O,N,2] / [O,N,2,B]
            sO = (B, O, self.config.N, 2) if self.layout == _lib.LAYOUT_INSTANCE_MAJOR else (O, self.config.N, 2, B)
        obs = self._dev(obs, sO, "obstacles") if O else None
        rad_s, rad_t = (float(obs_radius), None) if np.ndim(obs_radius) == 0 else (0.0, obs_radius)
        if rad_t is not None:
            rad_t = self._dev(rad_t, (B, O) if self.layout == _lib.LAYOUT_INSTANCE_MAJOR else (O, B), "obstacle_radius") if O else None
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
            rc = fn(self._h, B, p(x), p(goal), p(X0), p(U0), p(obs), O, rad_s, p(rad_t), float(inflation),
                    p(Xo), p(Uo), p(obj), p(st), p(it), C.c_void_p(stream))
        _lib.check(rc, self._h, "kmpc_solve_tracks" if tracks else "kmpc_solve")
        return SolveResult(Xo, Uo, obj, st, it)

    def _host_args(self, x, goal, X0, U0, obs, obs_radius):
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
        rad_s, rad_a = (float(obs_radius), None) if np.ndim(obs_radius) == 0 else (0.0, np64(obs_radius) if O else None)
        for a, s, n in ((x, sx, "current_state"), (goal, sx, "goal_state"), (X0, sX, "states_matrix"),
                        (U0, sU, "controls_matrix"), (obs if O else None, sO, "obstacles"),
                        (rad_a, (B, O) if self.layout == _lib.LAYOUT_INSTANCE_MAJOR else (O, B), "obstacle_radius")):
            if a is not None and a.shape != s:
                raise ValueError(f"{n}: expected shape {s}, got {a.shape}")
        if (X0 is None) != (U0 is None):
            raise ValueError("states_matrix and controls_matrix must both be given or both be None")
        return x, goal, X0, U0, (obs if O else None), O, rad_s, rad_a, B

    def _solve_host(self, x, goal, X0, U0, obs, obs_radius, inflation, copy=True) -> SolveResult:
        x, goal, X0, U0, obs, O, rad_s, rad_a, B = self._host_args(x, goal, X0, U0, obs, obs_radius)
        _, sX, sU, _, _ = self._shapes(B, O)
        p = lambda a: None if a is None else C.c_void_p(a.ctypes.data)
        if not copy:
            rc = self._L.kmpc_solve_host(self._h, B, p(x), p(goal), p(X0), p(U0), p(obs), O, rad_s, p(rad_a),
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
        rc = self._L.kmpc_solve_host(self._h, B, p(x), p(goal), p(X0), p(U0), p(obs), O, rad_s, p(rad_a),
                                     float(inflation), p(Xo), p(Uo), p(obj), p(st), p(it))
        _lib.check(rc, self._h, "kmpc_solve_host")
        return SolveResult(Xo, Uo, obj, st, it)

    # -- closed loop (agent.py:139-155) -------------------------------------------------------------
    def agent_handoff(self, states, controls, current_state, applied=None):
        """In place on the device: applied <- U[:,0] (agent.py:154-155), current_state <- X[:,1] (agent.py:70-72)."""
        torch = _torch()
        B = self._batch_of(current_state)
        sx, sX, sU, _, sA = self._shapes(B, 0)
        states, controls = self._dev(states, sX, "states"), self._dev(controls, sU, "controls")
        current_state = self._inplace(current_state, sx, "current_state")
        applied = self._inplace(applied, sA, "applied", optional=True)
        stream = torch.cuda.current_stream(torch.device("cuda", self.device)).cuda_stream
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        rc = self._L.kmpc_agent_handoff(self._h, B, p(states), p(controls), p(current_state), p(applied), C.c_void_p(stream))
        _lib.check(rc, self._h, "kmpc_agent_handoff")

    def select_obstacles(self, current_state, centers, radii, sensor_radius: float = 5.0, slots: Optional[int] = None,
                         literal_distance: bool = True, pad_center=(1.0e6, 1.0e6), return_index: bool = False,
                         return_radius: bool = False):
        """Batched sensor filter of ROSEnvironment.step (environment.py:48-65): per agent the candidate circles
        (centers [M,2], radii [M], CUDA tensors) within `sensor_radius` (agent.py:101), nearest first, at most `slots`
        (default O_max).  Returns (obstacles [B,slots,2] ready for ``solve(obstacles=...)``, count [B]); unused slots hold
        `pad_center`, whose rows stay inactive.  literal_distance: geometry.py:44 as written (True) or ||p-c|| - r.
        return_index: also return index [B,slots] int32, the candidate kept in every slot (-1 = padding).
        return_radius: also return radius [B,slots] -- in every slot the radius of the agent's NEAREST kept circle, which is what
        the planner subtracts for the whole class (optimizer.py:231-245) -- ready for ``solve(obstacle_radius=...)``."""
        torch = _torch()
        dev = torch.device("cuda", self.device)
        B = self._batch_of(current_state)
        O = int(slots if slots is not None else self.config.O_max)
        M = int(centers.shape[0])
        sx, _, _, sO, _ = self._shapes(B, O)
        current_state = self._dev(current_state, sx, "current_state")
        centers, radii = self._dev(centers, (M, 2), "centers"), self._dev(radii, (M,), "radii")
        out = torch.empty(sO, dtype=torch.float64, device=dev)
        cnt = torch.empty(B, dtype=torch.int32, device=dev)
        idx = torch.empty((B, O), dtype=torch.int32, device=dev) if return_index else None
        rad = torch.empty((B, O) if self.layout == _lib.LAYOUT_INSTANCE_MAJOR else (O, B), dtype=torch.float64, device=dev) if return_radius else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        rc = self._L.kmpc_select_obstacles(self._h, B, M, p(current_state), p(centers), p(radii),
                                           float(sensor_radius), 1 if literal_distance else 0, O, float(pad_center[0]), float(pad_center[1]),
                                           p(out), p(cnt), p(idx), p(rad), C.c_void_p(stream))
        _lib.check(rc, self._h, "kmpc_select_obstacles")
        r = (out, cnt)
        if return_index:
            r += (idx,)
        if return_radius:
            r += (rad,)
        return r

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
        obstacle_state = self._dev(obstacle_state, (M, 3), "obstacle_state")
        linear_velocity, angular_velocity = self._dev(linear_velocity, (M,), "linear_velocity"), self._dev(angular_velocity, (M,), "angular_velocity")
        index = self._dev(index, (int(batch), O), "index", dtype=torch.int32, optional=True)
        shape = (batch, O, N, 2) if self.layout == _lib.LAYOUT_INSTANCE_MAJOR else (O, N, 2, batch)
        out = torch.empty(shape, dtype=torch.float64, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        rc = self._L.kmpc_predict_tracks(self._h, int(batch), O, M, p(index), p(obstacle_state), p(linear_velocity), p(angular_velocity),
                                         float(dt), 1 if literal_heading else 0, float(pad_center[0]), float(pad_center[1]), p(out),
                                         C.c_void_p(stream))
        _lib.check(rc, self._h, "kmpc_predict_tracks")
        return out

    def closed_loop(self, current_state, goal_state, steps: int, states_matrix=None, controls_matrix=None,
                    log_applied: bool = True, log_iters: bool = True, goal_radius: float = 0.0, agent_radius: float = 0.0,
                    active=None, obstacle_centers=None, obstacle_radii=None, sensor_radius: float = 5.0, slots: Optional[int] = None,
                    inflation_radius: float = 0.5, literal_distance: bool = True, pad_center=(1.0e6, 1.0e6),
                    dynamic_states=None, dynamic_radii=None, dynamic_linear_velocity=None, dynamic_angular_velocity=None,
                    dynamic_slots: int = 0, use_tracks: bool = False, track_dt: float = 0.1, literal_heading: bool = True):
        """`steps` receding-horizon steps of EgoAgent.step (agent.py:130-155) for all B agents on the device
        (kmpc_closed_loop): warm start = previous solution unshifted, x <- X[:,1], applied = U[:,0].
        current_state [B,3] is advanced in place.  Returns (states, controls, applied_log [steps,B,2] | None,
        iters_log [steps,B] | None, status_log [steps,B]).
        goal_radius > 0: an agent that satisfies Agent.at_goal (agent.py:78-80, literal distance of geometry.py:44 with
        agent_radius; 0 = Euclidean) after a step is not solved again (status 1000 in the log), as the reference's
        environment stops stepping it (environment.py:31-33); `active` (int32 [B]) carries that mask in and out.
        obstacle_centers [M,2] / obstacle_radii [M] and/or dynamic_states [Md,3] / dynamic_radii [Md] (CUDA float64): the whole
        ROSEnvironment.step (environment.py:39-80, kmpc_environment_loop) -- every step each agent keeps the `slots` nearest static
        circles and the `dynamic_slots` nearest dynamic obstacles within `sensor_radius` and solves with them as obstacle rows,
        radius per class = the nearest kept one's (optimizer.py:231-250); use_tracks pairs every dynamic slot stage by stage with
        its constant-velocity prediction (dynamic_obstacle.py:20-37).  Per-step obstacle counts are left in
        ``self.last_obstacle_counts`` / ``self.last_dynamic_counts`` [steps,B]."""
        torch = _torch()
        dev = torch.device("cuda", self.device)
        B = self._batch_of(current_state)
        sx, sX, sU, _, sA = self._shapes(B, 0)
        N = self.config.N
        current_state = self._inplace(current_state, sx, "current_state")
        goal_state = self._dev(goal_state, sx, "goal_state")
        if (states_matrix is None) != (controls_matrix is None):
            raise ValueError("states_matrix and controls_matrix must both be given or both be None")
        if states_matrix is None:            # agent.py:59-60
            if self.layout == _lib.LAYOUT_INSTANCE_MAJOR:
                states_matrix = current_state[:, :, None].repeat(1, 1, N + 1).contiguous()
            else:
                states_matrix = current_state[:, None, :].repeat(1, N + 1, 1).contiguous()
            controls_matrix = torch.zeros(sU, dtype=torch.float64, device=dev)
        X, U = self._dev(states_matrix, sX, "states_matrix"), self._dev(controls_matrix, sU, "controls_matrix")
        applied = torch.empty((steps,) + sA, dtype=torch.float64, device=dev) if log_applied else None
        iters = torch.empty((steps, B), dtype=torch.int32, device=dev) if log_iters else None
        status = torch.empty((steps, B), dtype=torch.int32, device=dev)
        if active is None and goal_radius > 0:
            active = torch.ones(B, dtype=torch.int32, device=dev)
        active = self._inplace(active, (B,), "active", dtype=torch.int32, optional=True)
        self.last_active = active
        stream = torch.cuda.current_stream(dev).cuda_stream
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        if obstacle_centers is not None or dynamic_states is not None:
            M = Md = O = Od = 0
            cen = rad = dst = drad = dlv = dav = None
            if obstacle_centers is not None:
                M = int(obstacle_centers.shape[0])
                O = int(slots if slots is not None else self.config.O_max - int(dynamic_slots))
                cen, rad = self._dev(obstacle_centers, (M, 2), "obstacle_centers"), self._dev(obstacle_radii, (M,), "obstacle_radii")
            if dynamic_states is not None:
                Md, Od = int(dynamic_states.shape[0]), int(dynamic_slots)
                if Od < 1:
                    raise ValueError("dynamic_states needs dynamic_slots >= 1")
                dst, drad = self._dev(dynamic_states, (Md, 3), "dynamic_states"), self._dev(dynamic_radii, (Md,), "dynamic_radii")
                dlv = self._dev(dynamic_linear_velocity, (Md,), "dynamic_linear_velocity", optional=not use_tracks)
                dav = self._dev(dynamic_angular_velocity, (Md,), "dynamic_angular_velocity", optional=not use_tracks)
            counts = torch.empty((steps, B), dtype=torch.int32, device=dev) if O else None
            dcounts = torch.empty((steps, B), dtype=torch.int32, device=dev) if Od else None
            rc = self._L.kmpc_environment_loop(self._h, B, int(steps), p(current_state), p(goal_state), p(X), p(U),
                                               M, p(cen), p(rad), O, Md, p(dst), p(drad), p(dlv), p(dav), Od, 1 if use_tracks else 0,
                                               float(track_dt), 1 if literal_heading else 0, float(sensor_radius), 1 if literal_distance else 0,
                                               float(inflation_radius), float(pad_center[0]), float(pad_center[1]),
                                               p(applied), p(iters), p(status), p(counts), p(dcounts), p(active), float(goal_radius),
                                               float(agent_radius), C.c_void_p(stream))
            _lib.check(rc, self._h, "kmpc_environment_loop")
            self.last_obstacle_counts, self.last_dynamic_counts = counts, dcounts
            return X, U, applied, iters, status
        rc = self._L.kmpc_closed_loop(self._h, B, int(steps), p(current_state), p(goal_state), p(X), p(U), p(applied),
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
    never reads them, optimizer.py:375-400, and applies whatever iterate comes back).  ``on_failure`` says what this class
    does when the status is not Solve_Succeeded: "warn" (default: a RuntimeWarning, the iterate is still returned as the
    reference would), "raise" (KmpcError) or "ignore".
    ``use_obstacle_tracks``: False (default) reads only the current centre of every obstacle, as the reference's vectorised
    constraint path does (optimizer.py:217-221); True pairs X_{t+1} with column t of a dynamic obstacle's ``states_matrix``
    (what DynamicObstacle.calculate_symbolic_matrix_distance builds, dynamic_obstacle.py:47-56) when it has >= N columns.
    """

    def __init__(self, time_step: float, horizon: int, problem_form: str = "readme", device: int = 0,
                 use_obstacle_tracks: bool = False, on_failure: str = "warn"):
        if on_failure not in ("warn", "raise", "ignore"):
            raise ValueError("on_failure must be 'warn', 'raise' or 'ignore'")
        self.use_obstacle_tracks = bool(use_obstacle_tracks)
        self.on_failure = on_failure
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
        stat, dyn = self._centers(list(static_obstacles)), self._centers(list(dynamic_obstacles))
        obs = stat + dyn
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
            for j, ob in enumerate(dynamic_obstacles, start=len(stat)):
                sm = getattr(ob, "states_matrix", None)
                if sm is not None and np.shape(sm)[1] >= N:
                    centers[0, j] = np.asarray(sm, dtype=np.float64)[:2, :N].T
        # optimizer.py:231-250: ONE radius per obstacle class -- the first static obstacle's for every static column, the first
        # dynamic obstacle's for every dynamic column
        radius = np.array([[stat[0][1]] * len(stat) + ([dyn[0][1]] * len(dyn) if dyn else [])], dtype=np.float64) if O else 0.0
        infl = float(inflation_radius) if inflation_radius is not None else 0.0   # optimizer.py:362
        r = self._planner.solve(x, g, X0, U0, centers, radius, infl)
        self.last_status = int(r.status[0]); self.last_iterations = int(r.iters[0]); self.last_objective = float(r.objective[0])
        if self.last_status != 0 and self.on_failure != "ignore":
            msg = (f"MotionPlanner.solve: solver status {self.last_status} ({STATUS_NAMES.get(self.last_status, '?')}) after "
                   f"{self.last_iterations} iterations; the returned trajectory is the last iterate, not a solution")
            if self.on_failure == "raise":
                raise KmpcError(msg)
            warnings.warn(msg, RuntimeWarning, stacklevel=2)
        return np.array(r.states[0]), np.array(r.controls[0])


# ---- multi-GPU: batch slices, no collective inside the iteration -------------------------------------
def shard_range(B: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of a B-instance batch owned by `rank` (ceil split; trailing ranks may be empty)."""
    per = -(-B // world)
    lo = min(B, rank * per)
    return lo, min(B, lo + per)


class _PinnedBlock:
    """One portable, device-mapped pinned host allocation carved into NumPy views (the result buffer every device of a
    ShardedMotionPlanner writes its slice into)."""

    def __init__(self, L, fields):
        self._L = L
        self.offsets, off = {}, 0
        for name, shape, dt in fields:
            off = (off + 255) // 256 * 256
            self.offsets[name] = (off, shape, np.dtype(dt))
            off += int(np.prod(shape)) * np.dtype(dt).itemsize
        self.nbytes = max(off, 256)
        self.ptr = C.c_void_p()
        if L.kmpc_pinned_alloc(self.nbytes, C.byref(self.ptr)) != 0:
            raise KmpcError(f"kmpc_pinned_alloc({self.nbytes}) failed")
        self._raw = (C.c_uint8 * self.nbytes).from_address(self.ptr.value)

    def view(self, name):
        off, shape, dt = self.offsets[name]
        n = int(np.prod(shape))
        return np.frombuffer(self._raw, dtype=dt, count=n, offset=off).reshape(shape)

    def addr(self, name, index=0, stride=0):
        off, _, dt = self.offsets[name]
        return C.c_void_p(self.ptr.value + off + index * stride * dt.itemsize)

    def free(self):
        if self.ptr is not None and self.ptr.value:
            self._raw = None
            self._L.kmpc_pinned_free(self.ptr)
            self.ptr = C.c_void_p()


class ShardedMotionPlanner:
    """One batch over several GPUs of the box in ONE call from ONE process: device g solves the contiguous slice
    ``shard_range(B, g, G)`` with its own handle and stream; there is no collective -- the reference solves one NLP per agent
    with no coupling (optimizer.py:375-391), so the only exchange is the result gather, and that is done by the solver kernels
    themselves: every finished instance is written straight into its slice of one shared buffer (pinned host memory for
    ``solve`` on NumPy inputs, the first device's memory over NVLink for ``solve_device``).  ``devices`` may name a device more
    than once (several handles on one GPU: what the single-GPU tests use).  Instance-major layout.  Per instance the results are
    bit-identical to BatchedMotionPlanner's (same kernels, same arithmetic; only the queue each instance sits in differs)."""

    def __init__(self, config: PlannerConfig = PlannerConfig(), max_batch: int = 65536, devices: Sequence[int] = (0,)):
        self.config, self.max_batch, self.devices = config, int(max_batch), [int(d) for d in devices]
        if not self.devices:
            raise ValueError("devices must name at least one GPU")
        G = len(self.devices)
        self._per = -(-self.max_batch // G)
        self.planners = [BatchedMotionPlanner(config, max_batch=self._per, device=d) for d in self.devices]
        self._L = self.planners[0]._L
        N = config.N
        self._host = _PinnedBlock(self._L, [("X", (self.max_batch, 3, N + 1), np.float64), ("U", (self.max_batch, 2, N), np.float64),
                                            ("obj", (self.max_batch,), np.float64), ("status", (self.max_batch,), np.int32),
                                            ("iters", (self.max_batch,), np.int32)])
        self._dev_out = None
        self._peer_ok = False

    def close(self):
        for p in getattr(self, "planners", []):
            p.close()
        if getattr(self, "_host", None) is not None:
            self._host.free()
            self._host = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def shards(self, B: int):
        return [shard_range(B, g, len(self.devices)) for g in range(len(self.devices))]

    def solve(self, current_state, goal_state, states_matrix=None, controls_matrix=None, obstacles=None, obstacle_radius=0.3,
              inflation_radius: float = 0.0) -> SolveResult:
        """NumPy in -> NumPy views of the planner's pinned result buffer (valid until the next solve): the host API of
        BatchedMotionPlanner.solve(copy=False) across all the devices.  Every device stages its slice of the inputs, solves, and
        writes its results into its slice of the one buffer; the call returns when all devices are done."""
        first = self.planners[0]
        x, goal, X0, U0, obs, O, rad_s, rad_a, B = first._host_args(current_state, goal_state, states_matrix, controls_matrix,
                                                                    obstacles, obstacle_radius)
        if B > self.max_batch:
            raise ValueError(f"batch {B} > max_batch {self.max_batch}")
        N = self.config.N
        p = lambda a, lo: None if a is None else C.c_void_p(a[lo:].ctypes.data)
        H = self._host
        work = []
        for pl, (lo, hi) in zip(self.planners, self.shards(B)):
            if hi <= lo:
                continue
            rc = self._L.kmpc_solve_host_into(pl._h, hi - lo, p(x, lo), p(goal, lo), p(X0, lo), p(U0, lo), p(obs, lo), O, rad_s, p(rad_a, lo),
                                              float(inflation_radius), H.addr("X", lo, 3 * (N + 1)), H.addr("U", lo, 2 * N), H.addr("obj", lo, 1),
                                              H.addr("status", lo, 1), H.addr("iters", lo, 1))
            _lib.check(rc, pl._h, "kmpc_solve_host_into")
            work.append(pl)
        for pl in work:
            _lib.check(self._L.kmpc_host_sync(pl._h), pl._h, "kmpc_host_sync")
        return SolveResult(H.view("X")[:B], H.view("U")[:B], H.view("obj")[:B], H.view("status")[:B], H.view("iters")[:B])

    def solve_device(self, current_state: Sequence, goal_state: Sequence, states_matrix: Optional[Sequence] = None,
                     controls_matrix: Optional[Sequence] = None) -> SolveResult:
        """Device-resident form: ``current_state[g]`` / ``goal_state[g]`` are float64 CUDA tensors [B_g,3] already on
        ``devices[g]`` (slice g of the batch, sizes as ``shards(B)``); the results of all slices are gathered in tensors on
        ``devices[0]``: the solver kernels of the other devices write them there over NVLink peer access (no copy kernel, no
        collective).  Asynchronous: each device works on its current torch stream; synchronise the devices before reading."""
        torch = _torch()
        G = len(self.devices)
        sizes = [int(t.shape[0]) for t in current_state]
        B, N = sum(sizes), self.config.N
        if [hi - lo for lo, hi in self.shards(B)] != sizes:
            raise ValueError(f"shard sizes {sizes} do not match shards({B}) = {self.shards(B)}")
        if not self._peer_ok:
            for pl in self.planners[1:]:
                _lib.check(self._L.kmpc_enable_peer(pl._h, self.devices[0]), pl._h, "kmpc_enable_peer")
            self._peer_ok = True
        d0 = torch.device("cuda", self.devices[0])
        if self._dev_out is None or self._dev_out[0].shape[0] < B:
            cap = max(B, self.max_batch)
            self._dev_out = (torch.empty((cap, 3, N + 1), dtype=torch.float64, device=d0), torch.empty((cap, 2, N), dtype=torch.float64, device=d0),
                             torch.empty(cap, dtype=torch.float64, device=d0), torch.empty(cap, dtype=torch.int32, device=d0),
                             torch.empty(cap, dtype=torch.int32, device=d0))
        Xo, Uo, obj, st, it = self._dev_out
        ev0 = None
        if G > 1:   # the gather buffers must not be overwritten before earlier work on devices[0]'s stream is done with them
            ev0 = torch.cuda.Event(); ev0.record(torch.cuda.current_stream(d0))
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        for g, (pl, (lo, hi)) in enumerate(zip(self.planners, self.shards(B))):
            if hi <= lo:
                continue
            dev = torch.device("cuda", pl.device)
            sx, sX, sU, _, _ = pl._shapes(hi - lo, 0)
            x, gl = pl._dev(current_state[g], sx, f"current_state[{g}]"), pl._dev(goal_state[g], sx, f"goal_state[{g}]")
            X0 = pl._dev(states_matrix[g], sX, f"states_matrix[{g}]") if states_matrix is not None else None
            U0 = pl._dev(controls_matrix[g], sU, f"controls_matrix[{g}]") if controls_matrix is not None else None
            stream = torch.cuda.current_stream(dev)
            if ev0 is not None and pl.device != self.devices[0]:
                stream.wait_event(ev0)
            rc = self._L.kmpc_solve(pl._h, hi - lo, p(x), p(gl), p(X0), p(U0), None, 0, 0.0, None, 0.0, p(Xo[lo:hi]), p(Uo[lo:hi]),
                                    p(obj[lo:hi]), p(st[lo:hi]), p(it[lo:hi]), C.c_void_p(stream.cuda_stream))
            _lib.check(rc, pl._h, "kmpc_solve")
        return SolveResult(Xo[:B], Uo[:B], obj[:B], st[:B], it[:B])

    def synchronize(self):
        torch = _torch()
        for d in set(self.devices):
            torch.cuda.synchronize(torch.device("cuda", d))


# ---- one process per GPU (torchrun): the same gather, through a buffer on the root rank's device --------------------
RESULT_FIELDS = ("states", "controls", "objective", "status", "iters")


def result_layout(B: int, N: int):
    """Byte offsets of the five result arrays of a B-instance batch inside one gather buffer (256-byte aligned)."""
    out, off = {}, 0
    for name, per, size in (("states", 3 * (N + 1), 8), ("controls", 2 * N, 8), ("objective", 1, 8), ("status", 1, 4), ("iters", 1, 4)):
        off = (off + 255) // 256 * 256
        out[name] = (off, per, size)
        off += B * per * size
    return out, (off + 255) // 256 * 256


class RankGather:
    """Result gather of a batch sharded over the ranks of a torch.distributed job on one box (one process per GPU).  The root
    rank owns ONE buffer on its device; every other rank maps it (CUDA IPC, NVLink peer access) and hands the solver pointers
    into it, so each rank's kernel writes its finished instances straight into the root's memory: compute and gather are one
    kernel, nothing is exchanged afterwards and NCCL is not involved ("no NCCL in the iteration").
    ``transport="nccl"`` keeps the results local and gathers them with one torch.distributed.gather of a packed byte tensor."""

    def __init__(self, planner: BatchedMotionPlanner, B: int, group=None, root: int = 0, transport: str = "ipc"):
        import torch.distributed as dist
        torch = _torch()
        self.planner, self.B, self.group, self.root, self.transport = planner, int(B), group, root, transport
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.N = planner.config.N
        self.lo, self.hi = shard_range(self.B, self.rank, self.world)
        self.layout, self.nbytes = result_layout(self.B, self.N)
        self._L, self._ptr, self._owner = planner._L, C.c_void_p(), self.rank == root
        dev = torch.device("cuda", planner.device)
        if transport == "ipc":
            hb = torch.zeros(_lib.IPC_HANDLE_BYTES, dtype=torch.uint8, device=dev)
            if self._owner:
                raw = (C.c_uint8 * _lib.IPC_HANDLE_BYTES)()
                _lib.check(self._L.kmpc_shared_buffer_create(planner._h, self.nbytes, C.byref(self._ptr), raw), planner._h,
                           "kmpc_shared_buffer_create")
                hb.copy_(torch.tensor(list(raw), dtype=torch.uint8))
            dist.broadcast(hb, src=root, group=group)
            if not self._owner:
                raw = (C.c_uint8 * _lib.IPC_HANDLE_BYTES)(*hb.cpu().tolist())
                _lib.check(self._L.kmpc_shared_buffer_open(planner._h, raw, C.byref(self._ptr)), planner._h, "kmpc_shared_buffer_open")
        elif transport == "nccl":
            self._local = torch.empty(self.nbytes, dtype=torch.uint8, device=dev)   # same layout; only this rank's slices are used
            self._ptr = C.c_void_p(self._local.data_ptr())
            self._owner = False
            per = -(-self.B // self.world)
            self._pack_bytes = result_layout(per, self.N)[1]
            self._send = torch.empty(self._pack_bytes, dtype=torch.uint8, device=dev)
            self._recv = [torch.empty(self._pack_bytes, dtype=torch.uint8, device=dev) for _ in range(self.world)] if self.rank == root else None
        else:
            raise ValueError("transport must be 'ipc' or 'nccl'")

    def _addr(self, name, index):
        off, per, size = self.layout[name]
        return C.c_void_p(self._ptr.value + off + index * per * size)

    def solve(self, current_state, goal_state, states_matrix=None, controls_matrix=None):
        """This rank's slice (CUDA tensors [hi-lo, ...] on its device) -> results in the gather buffer.  Asynchronous on the
        current torch stream.  With the NCCL transport the gather collective is issued here as well."""
        torch = _torch()
        pl, n = self.planner, self.hi - self.lo
        dev = torch.device("cuda", pl.device)
        stream = torch.cuda.current_stream(dev).cuda_stream
        if n > 0:
            sx, sX, sU, _, _ = pl._shapes(n, 0)
            x, g = pl._dev(current_state, sx, "current_state"), pl._dev(goal_state, sx, "goal_state")
            X0, U0 = pl._dev(states_matrix, sX, "states_matrix", optional=True), pl._dev(controls_matrix, sU, "controls_matrix", optional=True)
            p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
            rc = self._L.kmpc_solve(pl._h, n, p(x), p(g), p(X0), p(U0), None, 0, 0.0, None, 0.0, self._addr("states", self.lo),
                                    self._addr("controls", self.lo), self._addr("objective", self.lo), self._addr("status", self.lo),
                                    self._addr("iters", self.lo), C.c_void_p(stream))
            _lib.check(rc, pl._h, "kmpc_solve")
        if self.transport == "nccl":
            import torch.distributed as dist
            per = -(-self.B // self.world)
            lay, _ = result_layout(per, self.N)
            for name in RESULT_FIELDS:   # pack this rank's slices back to back
                off, pr, size = self.layout[name]
                o2 = lay[name][0]
                nb = n * pr * size
                if nb:
                    self._send[o2:o2 + nb].copy_(self._local[off + self.lo * pr * size: off + self.lo * pr * size + nb])
            dist.gather(self._send, self._recv, dst=self.root, group=self.group)
            if self.rank == self.root:
                for r in range(self.world):
                    lo, hi = shard_range(self.B, r, self.world)
                    for name in RESULT_FIELDS:
                        off, pr, size = self.layout[name]
                        o2 = lay[name][0]
                        nb = (hi - lo) * pr * size
                        if nb and r != self.root:
                            self._local[off + lo * pr * size: off + lo * pr * size + nb].copy_(self._recv[r][o2:o2 + nb])

    def result(self) -> Optional[SolveResult]:
        """Root rank: the gathered batch as torch views of the buffer (after a device synchronise + barrier); other ranks: None."""
        if self.rank != self.root:
            return None
        torch = _torch()
        dev = torch.device("cuda", self.planner.device)
        if self.transport == "nccl":
            raw = self._local
        else:
            raw = _cuda_view(self._ptr.value, self.nbytes, dev)
        B, N = self.B, self.N
        out = []
        for name, shape, dt in (("states", (B, 3, N + 1), torch.float64), ("controls", (B, 2, N), torch.float64),
                                ("objective", (B,), torch.float64), ("status", (B,), torch.int32), ("iters", (B,), torch.int32)):
            off, per, size = self.layout[name]
            out.append(raw[off: off + B * per * size].view(dt).reshape(shape))
        return SolveResult(*out)

    def close(self):
        if self.transport == "ipc" and self._ptr.value:
            self._L.kmpc_shared_buffer_close(self.planner._h, self._ptr, 1 if self._owner else 0)
            self._ptr = C.c_void_p()


def _cuda_view(ptr: int, nbytes: int, device):
    """uint8 torch tensor over raw device memory owned by the library (__cuda_array_interface__; no copy)."""
    torch = _torch()

    class _Raw:
        __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}

    with torch.cuda.device(device):
        return torch.as_tensor(_Raw(), device=device)


def gather_results(local: SolveResult, B: int, group=None, root: int = 0) -> Optional[SolveResult]:
    """Gather-to-root of per-rank result slices through torch.distributed (gloo on CPU tensors, NCCL on CUDA tensors): ONE
    ``gather`` of a packed byte tensor per call, padded to the ceil-split size so ragged B works.  The GPU bench does not need it
    (RankGather writes into the root's memory from inside the solver kernel); it serves hosts without peer access and the CPU tests."""
    torch = _torch()
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = -(-B // world)
    parts = [torch.as_tensor(t) for t in local]
    n = parts[0].shape[0]
    dev = parts[0].device
    rows = [t.reshape(n, -1).contiguous().view(torch.uint8).reshape(n, -1) for t in parts]   # bytes per instance, field by field
    widths = [r.shape[1] for r in rows]
    send = torch.zeros((per, sum(widths)), dtype=torch.uint8, device=dev)
    if n:
        send[:n] = torch.cat(rows, 1)
    recv = [torch.empty_like(send) for _ in range(world)] if rank == root else None
    dist.gather(send, recv, dst=root, group=group)
    if rank != root:
        return None
    full = torch.cat([recv[r][: shard_range(B, r, world)[1] - shard_range(B, r, world)[0]] for r in range(world)], 0)
    out, o = [], 0
    for t, w in zip(parts, widths):
        out.append(full[:, o:o + w].contiguous().view(t.dtype).reshape((B,) + tuple(t.shape[1:])))
        o += w
    return SolveResult(*out)
