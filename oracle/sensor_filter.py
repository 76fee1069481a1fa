"""CPU restatement of the sensor filter in ROSEnvironment.step (mpc/environment.py:48-65) -- TEST INFRASTRUCTURE ONLY.

    static_obstacles_dict = {obstacle.calculate_distance(self.agent.state): obstacle for obstacle in self.static_obstacles}
    filtered = [static_obstacles_dict[d] for d in sorted(static_obstacles_dict.keys()) if d <= self.agent.sensor_radius]

with Obstacle.calculate_distance -> Circle.calculate_distance (obstacle_handling/geometry.py:38-44):
    np.linalg.norm(np.array(distance_to[:2] - center) - self.radius)        # the radius is subtracted from BOTH components
Pure-Python loops, as the reference runs them (one agent at a time)."""
import numpy as np


def circle_distance(state, center, radius, literal=True):
    if literal:                                                         # geometry.py:44 as written
        return float(np.linalg.norm(np.array(state[:2] - center) - radius))
    return float(np.linalg.norm(state[:2] - center) - radius)           # the intended distance


def sensor_filter(state, centers, radii, sensor_radius, literal=True):
    """Indices of the candidates the reference would pass to the planner, in its order (environment.py:48-56)."""
    d = {}
    for m in range(len(centers)):                                       # dict keyed by distance: equal keys keep the later one
        d[circle_distance(state, centers[m], radii[m], literal)] = m
    return [d[k] for k in sorted(d.keys()) if k <= sensor_radius]
