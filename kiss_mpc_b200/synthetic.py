"""Seeded synthetic MPC workloads (SURVEY.md 8d "Common synthetic inputs").

Start p_x,p_y ~ U(-5,5), theta ~ U(-pi,pi); goal = start + r(cos b, sin b), r ~ U(0.5,6), b ~ U(-pi,pi),
theta_g ~ U(-pi,pi).  Bounds/weights are the reference defaults (agent.py:104-106, optimizer.py:57-60).
Obstacles (cfg 4): O circles of radius 0.3, inflation 0.5, centres uniform in the start-goal bounding box inflated by
2 m, rejection-sampled so that |c - start| >= r_o + I + 0.3 (feasible start).
"""
from __future__ import annotations

import numpy as np


def make_batch(B: int, seed: int = 1000, O: int = 0, obs_radius: float = 0.3, inflation: float = 0.5):
    """Returns dict(x_cur[B,3], goal[B,3], obs[B,O,2] or None), float64, C-contiguous."""
    rng = np.random.default_rng(seed)
    x_cur = np.empty((B, 3)); goal = np.empty((B, 3))
    x_cur[:, 0:2] = rng.uniform(-5.0, 5.0, size=(B, 2))
    x_cur[:, 2] = rng.uniform(-np.pi, np.pi, size=B)
    r = rng.uniform(0.5, 6.0, size=B); b = rng.uniform(-np.pi, np.pi, size=B)
    goal[:, 0] = x_cur[:, 0] + r * np.cos(b); goal[:, 1] = x_cur[:, 1] + r * np.sin(b)
    goal[:, 2] = rng.uniform(-np.pi, np.pi, size=B)
    obs = None
    if O > 0:
        lo = np.minimum(x_cur[:, :2], goal[:, :2]) - 2.0
        hi = np.maximum(x_cur[:, :2], goal[:, :2]) + 2.0
        obs = np.empty((B, O, 2))
        need = np.ones((B, O), bool)
        dmin = obs_radius + inflation + 0.3
        while need.any():
            cand = lo[:, None, :] + rng.uniform(size=(B, O, 2)) * (hi - lo)[:, None, :]
            obs[need] = cand[need]
            d = np.linalg.norm(obs - x_cur[:, None, :2], axis=2)
            need = d < dmin
    return {"x_cur": x_cur, "goal": goal, "obs": obs}


def make_tracks(obs, N: int, seed: int = 0, max_speed: float = 0.3, dt: float = 0.1):
    """Moving-obstacle tracks [B,O,N,2] from the centres of ``make_batch``: every circle drifts with a constant velocity
    (speed ~ U(0, max_speed) m/s, direction ~ U(-pi, pi)); column t = centre + t dt v is the centre paired with X_{t+1}
    (the straight-line case of dynamic_obstacle.py:20-37)."""
    rng = np.random.default_rng(seed)
    B, O = obs.shape[:2]
    sp = rng.uniform(0.0, max_speed, size=(B, O)); hd = rng.uniform(-np.pi, np.pi, size=(B, O))
    vel = np.stack([sp * np.cos(hd), sp * np.sin(hd)], axis=2)
    return np.ascontiguousarray(obs[:, :, None, :] + (np.arange(N) * dt)[None, None, :, None] * vel[:, :, None, :])


def cfg1_instance():
    """BASELINE config 0: single agent, start (0,0,pi/2) (ros2interface.py:30-31), goal (2,3,0) (SURVEY 8d cfg 1)."""
    return np.array([[0.0, 0.0, np.pi / 2]]), np.array([[2.0, 3.0, 0.0]])
