"""Test helper: an object with MotionPlanner.solve's keyword interface (optimizer.py:319-333) that answers from the CPU
oracle.  Used to drive kiss_mpc_b200.Model on CPU and as the checker of the GPU-driven Model."""
from dataclasses import replace

import numpy as np


class OraclePlanner:
    def __init__(self, oracle_mod, time_step, horizon, **cfg_kw):
        self.ok, self.N, self.T, self.cfg_kw = oracle_mod, int(horizon), float(time_step), cfg_kw
        self.calls = []

    def solve(self, current_state, current_linear_velocity=None, current_angular_velocity=None, goal_state=None,
              states_matrix=None, controls_matrix=None, state_bounds=(-20.0, 20.0), linear_velocity_bounds=(-0.2, 0.5),
              angular_velocity_bounds=(-0.5, 0.5), static_obstacles=(), dynamic_obstacles=(), inflation_radius=None):
        obs = list(static_obstacles) + list(dynamic_obstacles)
        cfg = self.ok.OracleConfig(N=self.N, T=self.T, linsolve="dense", x_bounds=state_bounds, y_bounds=state_bounds,
                                   v_bounds=linear_velocity_bounds, w_bounds=angular_velocity_bounds, O=len(obs),
                                   obs_radius=obs[0].geometry.radius if obs else 0.3, inflation=inflation_radius or 0.0)
        cfg = replace(cfg, **self.cfg_kw)
        cen = np.array([o.geometry.center for o in obs], dtype=float).reshape(1, len(obs), 2) if obs else None
        r = self.ok.solve(cfg, np.asarray(current_state, float).reshape(1, 3), np.asarray(goal_state, float).reshape(1, 3),
                          X0=np.asarray(states_matrix, float)[None], U0=np.asarray(controls_matrix, float)[None], obs=cen)
        self.calls.append(dict(current_state=np.array(current_state, float), goal_state=np.array(goal_state, float),
                               n_static=len(static_obstacles), n_dynamic=len(dynamic_obstacles), status=int(r.status[0]),
                               first_static=(static_obstacles[0] if static_obstacles else None)))
        return r.X[0].copy(), r.U[0].copy()
