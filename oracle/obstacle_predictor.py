"""CPU restatement of the dynamic-obstacle predictor (obstacle_handling/dynamic_obstacle.py:20-37) -- TEST INFRASTRUCTURE ONLY.

    _predict_state(state):   dt = 0.1
        [x + v cos(deg2rad(theta)) dt,  y + v sin(deg2rad(theta)) dt,  theta + omega dt]          (:20-28)
    _get_predicted_states_matrix(horizon): column 0 = current state, column t = _predict_state(column t-1)   (:30-37)

The reference applies np.deg2rad to a heading that is already in radians (default orientation np.deg2rad(90), :8);
literal=True keeps that, literal=False uses the heading as radians.  Column t of the track is the centre paired with
X_{t+1} by calculate_symbolic_matrix_distance (:47-56).  Pure-Python loop, as the reference runs it."""
import numpy as np


def predict_track(state, linear_velocity, angular_velocity, horizon, dt=0.1, literal=True):
    """(3, horizon) predicted states of ONE obstacle; state = (x, y, heading)."""
    out = np.zeros((3, horizon))
    out[:, 0] = np.asarray(state, float)
    for t in range(1, horizon):
        x, y, th = out[:, t - 1]
        a = np.deg2rad(th) if literal else th
        out[:, t] = [x + linear_velocity * np.cos(a) * dt, y + linear_velocity * np.sin(a) * dt, th + angular_velocity * dt]
    return out


def predict_tracks(states, linear_velocity, angular_velocity, horizon, dt=0.1, literal=True):
    """[M, horizon, 2] centre tracks of M obstacles (states [M,3], velocities [M])."""
    states = np.asarray(states, float).reshape(-1, 3)
    return np.stack([predict_track(states[m], float(linear_velocity[m]), float(angular_velocity[m]), horizon, dt, literal)[:2].T
                     for m in range(len(states))]) if len(states) else np.zeros((0, horizon, 2))
