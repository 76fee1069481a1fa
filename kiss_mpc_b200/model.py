"""``Model`` -- the single-agent stepping object the reference's ROS node drives (``from mpc.model import Model``,
ros2interface.py:19) and whose source HEAD does not ship.  Rebuilt from its call sites and from the two classes it stood
for: ``EgoAgent`` (mpc/agent.py:8-155: buffers, hand-off, ``at_goal``, ``reset``) and ``ROSEnvironment``
(mpc/environment.py:8-92: sensor filter, waypoint advance).

Call sites it has to satisfy (ros2interface.py):
  :28-38    Model(id, initial_position, initial_orientation, horizon=7, use_warm_start=True, planning_time_step=0.8,
                  linear_velocity_bounds=(-0.3, 0.3), angular_velocity_bounds=(-0.3, 0.3), waypoints=[])
  :55       model.step()                       one planner solve + hand-off + waypoint bookkeeping
  :59-60    model.linear_velocity / model.angular_velocity      the control to publish (U[:, 0], agent.py:154-155)
  :65       model.states_matrix                                 the predicted states to visualise
  :93-107   model.initial_state = odom pose;  model.reset(matrices_only=True)
  :172-174  model.waypoints = ndarray; model.waypoint_index = 0; model.update_goal(model.current_waypoint())

This is host glue above the hot path (B = 1): every solve goes through ``MotionPlanner`` (the C ABI, GPU only).  A different
planner object with the same ``solve`` keywords can be injected (the tests use one to check the bookkeeping on CPU).
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np


def literal_circle_distance(center, radius, point) -> float:
    """Circle.calculate_distance as written (geometry.py:38-44): the radius is subtracted from BOTH components."""
    return float(np.linalg.norm(np.array(np.asarray(point, dtype=float)[:2] - np.asarray(center, dtype=float)) - radius))


class Model:
    def __init__(self, id, initial_position, initial_orientation, horizon: int = 50, use_warm_start: bool = False,
                 planning_time_step: float = 0.041, linear_velocity_bounds=(-0.2, 0.5), angular_velocity_bounds=(-0.5, 0.5),
                 waypoints=None, radius: float = 0.4, state_bounds=(-20.0, 20.0), sensor_radius: float = 5.0,
                 static_obstacles: Sequence = (), dynamic_obstacles: Sequence = (), state_override: bool = False,
                 planner=None, problem_form: str = "readme", device: int = 0):
        # defaults: EgoAgent (agent.py:91-106); radius has no default in the reference -- 0.4 gives the inflation 0.5 of agent.py:149
        assert horizon > 0                                                   # agent.py:29
        self.id = id
        self.radius = float(radius)
        self.center = np.array(initial_position, dtype=np.float64)           # Circle(center=initial_position) agent.py:36
        self.sensor_radius = sensor_radius
        self.initial_state = np.array([*initial_position, initial_orientation], dtype=np.float64)    # agent.py:38
        self.goal_state = self.initial_state                                 # agent.py:39-43 (no goal yet)
        self.horizon = int(horizon)
        self.time_step = float(planning_time_step)
        self.linear_velocity_bounds = tuple(linear_velocity_bounds)
        self.angular_velocity_bounds = tuple(angular_velocity_bounds)
        self.state_bounds = tuple(state_bounds)
        self.initial_linear_velocity = self.initial_angular_velocity = 0.0
        self.linear_velocity = self.angular_velocity = 0.0
        self.states_matrix = np.tile(self.initial_state, (self.horizon + 1, 1)).T          # agent.py:59
        self.controls_matrix = np.zeros((2, self.horizon))                                 # agent.py:60
        self.use_warm_start = use_warm_start
        self.goal_radius = 0.5                                                             # agent.py:65
        self.static_obstacles = list(static_obstacles)
        self.dynamic_obstacles = list(dynamic_obstacles)
        self.state_override = bool(state_override)
        self.waypoints = waypoints if waypoints is not None else []
        self.waypoint_index = 0
        if planner is None:
            from .planner import MotionPlanner
            planner = MotionPlanner(time_step=self.time_step, horizon=self.horizon, problem_form=problem_form, device=device)
        self.planner = planner                                                             # agent.py:62
        if len(self.waypoints):
            self.update_goal(self.current_waypoint())                                      # environment.py:18

    # ---- agent.py:67-92 ----
    def update_goal(self, goal):
        self.goal_state = np.asarray(goal, dtype=np.float64) if goal is not None else self.initial_state

    @property
    def state(self):
        return self.states_matrix[:, 1]                                                    # agent.py:70-72

    @property
    def at_goal(self) -> bool:
        return literal_circle_distance(self.center, self.radius, self.goal_state) - self.goal_radius <= 0   # agent.py:78-80

    def reset(self, matrices_only: bool = False, to_initial_state: bool = True):
        self.states_matrix = np.tile(self.initial_state if to_initial_state else self.state, (self.horizon + 1, 1)).T
        self.controls_matrix = np.zeros((2, self.horizon))
        if not matrices_only:
            self.linear_velocity = self.initial_linear_velocity
            self.angular_velocity = self.initial_angular_velocity

    # ---- environment.py:22-34 ----
    def current_waypoint(self):
        return self.waypoints[self.waypoint_index] if self.waypoint_index < len(self.waypoints) else None

    @property
    def final_goal_reached(self) -> bool:
        return self.waypoint_index == len(self.waypoints) - 1 and self.at_goal

    def _sensor_filter(self, obstacles):
        # environment.py:48-65: dict keyed by distance (equal keys keep the later obstacle), ascending, within the sensor radius
        st = self.state
        by_distance = {literal_circle_distance(ob.geometry.center, ob.geometry.radius, st): ob for ob in obstacles}
        return [by_distance[d] for d in sorted(by_distance.keys()) if d <= self.sensor_radius]

    # ---- environment.py:39-80 + agent.py:130-155 ----
    def step(self):
        self.goal_radius = 0.5                                                             # environment.py:40-45 (both branches)
        static = self._sensor_filter(self.static_obstacles)
        dynamic = self._sensor_filter(self.dynamic_obstacles)
        current = self.initial_state if self.state_override else self.state
        self.states_matrix, self.controls_matrix = self.planner.solve(
            current_state=current, current_linear_velocity=self.linear_velocity, current_angular_velocity=self.angular_velocity,
            goal_state=self.goal_state, states_matrix=self.states_matrix, controls_matrix=self.controls_matrix,
            state_bounds=self.state_bounds, linear_velocity_bounds=self.linear_velocity_bounds,
            angular_velocity_bounds=self.angular_velocity_bounds, inflation_radius=self.radius + 0.1,
            static_obstacles=static, dynamic_obstacles=dynamic)
        self.center = np.array(self.initial_state[:2] if self.state_override else self.state[:2], dtype=np.float64)   # agent.py:153
        self.linear_velocity = float(self.controls_matrix[0, 0])                          # agent.py:154
        self.angular_velocity = float(self.controls_matrix[1, 0])                         # agent.py:155
        if len(self.waypoints) and self.at_goal and not self.final_goal_reached:          # environment.py:77-80
            self.waypoint_index += 1
            self.update_goal(self.current_waypoint())
