"""Real-CasADi golden runner -- TEST INFRASTRUCTURE ONLY (same rule as oracle.py).

Runs the reference's own ``MotionPlanner.solve`` (mpc/optimizer.py:319-400: ca.nlpsol("ipopt") rebuilt per call, :354, solved at
:375-391) on seeded instances and stores what IPOPT returned as golden vectors (.npz): the pin the CPU oracle is still missing
("parity unpinned").  It needs ``casadi`` (requirements.txt:1 pins 3.7.1) and the reference tree; neither the build
container nor the GPU box has casadi, so everything here is guarded: ``available()`` is False and tests/test_casadi_golden.py
skips.  On any machine with ``pip install casadi==3.7.1`` and a checkout of rtarun1/kiss-mpc:

    KMPC_REFERENCE=/path/to/kiss-mpc python -m oracle.casadi_runner            # writes tests/golden/casadi_*.npz
    python -m pytest tests/test_casadi_golden.py                              # oracle vs IPOPT on those vectors

HEAD of the reference cannot run.  The reference classes are imported from its tree UNMODIFIED and
subclassed with the minimal repairs, nothing else:
  C-1  optimizer.py:337-340  solve() calls get_symbolic_constraints without the obstacle arguments its signature (:260-266) needs
  C-2  optimizer.py:268-271  get_symbolic_constraints passes velocities to get_symbolic_state_constrains, which takes none (:163-165)
  C-3  optimizer.py:359-364  solve() passes velocity bounds to get_constraints_bounds(inflation_radius, num_obstacles) (:284-288);
                             the obstacles are never forwarded into g although their bounds are appended (:363)
  C-4  optimizer.py:223-227  the subtraction of the state terms sits on its own lines (dangling unary minus): the distance rows are
                             constants; repaired to the intended |centre - p_k| - radius (README.md:78-81)
``form="readme"`` additionally swaps in the README cost / bounds (README.md:15-66: goal cost over k = 1..N, squared asymmetric
velocity penalty, y bounded like x) -- labelled as such in the golden file; ``form="code"`` is the reference NLP as written.
"""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np

REF_DEFAULT = os.environ.get("KMPC_REFERENCE", "/root/reference")

# IPOPT return_status strings (what CasADi's stats() reports) -> ApplicationReturnStatus numbers used throughout this repo
STATUS_CODES = {"Solve_Succeeded": 0, "Solved_To_Acceptable_Level": 1, "Infeasible_Problem_Detected": 2,
                "Search_Direction_Becomes_Too_Small": 3, "Diverging_Iterates": 4, "Maximum_Iterations_Exceeded": -1,
                "Restoration_Failed": -2, "Error_In_Step_Computation": -3, "Invalid_Number_Detected": -13}


def available(ref: str = REF_DEFAULT) -> bool:
    try:
        import casadi  # noqa: F401
    except Exception:
        return False
    return os.path.exists(os.path.join(ref, "mpc", "optimizer.py"))


def load_repaired(ref: str = REF_DEFAULT, form: str = "code"):
    """Returns (RepairedMotionPlanner class, reference geometry module, record) -- record["solver"] is the last nlpsol object."""
    import casadi as ca
    if ref not in sys.path:
        sys.path.insert(0, ref)
    opt = importlib.import_module("mpc.optimizer")
    geo = importlib.import_module("obstacle_handling.geometry")
    record = {"solver": None}
    real_nlpsol = ca.nlpsol

    def nlpsol(*a, **k):                     # the reference discards IPOPT's stats (optimizer.py:375-400): keep the solver object
        record["solver"] = real_nlpsol(*a, **k)
        return record["solver"]

    opt.ca.nlpsol = nlpsol

    class Repaired(opt.MotionPlanner):
        _stat, _dyn = (), ()

        def solve(self, *args, static_obstacles=[], dynamic_obstacles=[], **kw):
            self._stat, self._dyn = list(static_obstacles), list(dynamic_obstacles)          # C-3: remembered for g
            return super().solve(*args, static_obstacles=static_obstacles, dynamic_obstacles=dynamic_obstacles, **kw)

        def get_symbolic_constraints(self, current_linear_velocity=None, current_angular_velocity=None, static_obstacles=None,
                                     dynamic_obstacles=None):                                    # C-1 (+ C-3: obstacles forwarded)
            return super().get_symbolic_constraints(current_linear_velocity, current_angular_velocity,
                                                    list(self._stat) if static_obstacles is None else static_obstacles,
                                                    list(self._dyn) if dynamic_obstacles is None else dynamic_obstacles)

        def get_symbolic_state_constrains(self, current_linear_velocity=None, current_angular_velocity=None):   # C-2
            return super().get_symbolic_state_constrains()

        def get_constraints_bounds(self, inflation_radius=0, num_obstacles=0, **_ignored):                     # C-3
            return super().get_constraints_bounds(inflation_radius=inflation_radius, num_obstacles=num_obstacles)

        def get_symbolic_obstacle_constraints(self, static_obstacles, dynamic_obstacles):                      # C-4
            obstacles = list(static_obstacles) + list(dynamic_obstacles)
            cols = []
            for j, ob in enumerate(obstacles):     # (N x O): column o = |c_o - p_k| - radius of the obstacle's class, k = 1..N
                cls = static_obstacles if j < len(static_obstacles) else dynamic_obstacles
                c = ob.geometry.center
                dx = c[0] - self.symbolic_states_matrix[0, 1:]
                dy = c[1] - self.symbolic_states_matrix[1, 1:]
                cols.append((ca.sqrt(dx ** 2 + dy ** 2) - cls[0].geometry.radius).T)
            return ca.horzcat(*cols)

    if form == "readme":
        class Readme(Repaired):
            def get_symbolic_goal_cost(self):                       # README.md:17: t = 1..N (the code stops at N-1, optimizer.py:80)
                err = self.symbolic_states_matrix[:, 1:] - ca.repmat(self.symbolic_terminal_states_vector[3:], 1, self.horizon)
                return ca.sum2(ca.sum1(ca.diag(ca.DM([100.0, 100.0, 50.0])) @ (err * err)))

            def get_symbolic_negative_linear_velocity_cost(self):   # README.md:23-24: W_v- min(0, v)^2  (W_v+ = 0)
                v = self.symbolic_controls_matrix[0, :]
                return 300.0 * ca.sum2(ca.fmin(v, 0) ** 2)

            def get_optimization_variable_bounds(self, state_bounds, linear_velocity_bounds, angular_velocity_bounds):
                lo, hi = super().get_optimization_variable_bounds(state_bounds=state_bounds, linear_velocity_bounds=linear_velocity_bounds,
                                                                  angular_velocity_bounds=angular_velocity_bounds)
                lo, hi = np.array(lo.full()).reshape(-1), np.array(hi.full()).reshape(-1)
                for k in range(self.horizon + 1):                   # README.md:61-66: y bounded like x
                    lo[3 * k + 1], hi[3 * k + 1] = state_bounds[0], state_bounds[1]
                return ca.DM(lo), ca.DM(hi)

        return Readme, geo, record
    return Repaired, geo, record


def run_batch(x_cur, goal, N=30, T=0.1, form="code", obs=None, obs_static=0, radii=(0.3, 0.3), inflation=0.5, ref: str = REF_DEFAULT,
              v_bounds=(-0.2, 0.5), w_bounds=(-0.5, 0.5), state_bounds=(-20.0, 20.0)):
    """Cold-start solves (agent.py:59-60) of B instances through the repaired reference planner.  obs [B,O,2]: the first
    `obs_static` circles of every instance are static obstacles (radius radii[0]), the rest dynamic (radii[1])."""
    cls, geo, record = load_repaired(ref, form)
    B = len(x_cur)
    X = np.zeros((B, 3, N + 1)); U = np.zeros((B, 2, N)); st = np.zeros(B, np.int32); it = np.zeros(B, np.int32)

    class Ob:
        def __init__(self, c, r):
            self.geometry = geo.Circle(center=c, radius=r)

    for i in range(B):
        mp = cls(time_step=T, horizon=N)
        stat = [] if obs is None else [Ob(c, radii[0]) for c in obs[i][:obs_static]]
        dyn = [] if obs is None else [Ob(c, radii[1]) for c in obs[i][obs_static:]]
        X0 = np.tile(x_cur[i], (N + 1, 1)).T; U0 = np.zeros((2, N))
        X[i], U[i] = mp.solve(current_state=x_cur[i], current_linear_velocity=0.0, current_angular_velocity=0.0, goal_state=goal[i],
                              states_matrix=X0, controls_matrix=U0, state_bounds=state_bounds, linear_velocity_bounds=v_bounds,
                              angular_velocity_bounds=w_bounds, static_obstacles=stat, dynamic_obstacles=dyn,
                              inflation_radius=inflation if obs is not None else None)
        s = record["solver"].stats()
        st[i] = STATUS_CODES.get(s["return_status"], -100); it[i] = s["iter_count"]
    return X, U, st, it


def main():
    import casadi
    from kiss_mpc_b200.synthetic import cfg1_instance, make_batch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, "tests", "golden")
    x1, g1 = cfg1_instance()
    for form in ("code", "readme"):
        b = make_batch(63, seed=1002)
        x = np.concatenate([x1, b["x_cur"]]); g = np.concatenate([g1, b["goal"]])
        X, U, st, it = run_batch(x, g, form=form)
        np.savez(os.path.join(out, f"casadi_{form}_box.npz"), x_cur=x, goal=g, X=X, U=U, status=st, iters=it, form=form,
                 casadi_version=casadi.__version__, repairs="App. C-1..C-4 only" + ("; README cost/bounds" if form == "readme" else ""))
        bo = make_batch(32, seed=1004, O=6)
        X, U, st, it = run_batch(bo["x_cur"], bo["goal"], form=form, obs=bo["obs"], obs_static=4, radii=(0.1, 0.3))
        np.savez(os.path.join(out, f"casadi_{form}_obs.npz"), x_cur=bo["x_cur"], goal=bo["goal"], obs=bo["obs"], obs_static=4, radii=np.array([0.1, 0.3]),
                 inflation=0.5, X=X, U=U, status=st, iters=it, form=form, casadi_version=casadi.__version__)
        print(form, "status counts", dict(zip(*np.unique(st, return_counts=True))), "mean iterations", it.mean())


if __name__ == "__main__":
    if not available():
        sys.exit("casadi and/or the reference tree are not available here (see the module docstring)")
    main()
