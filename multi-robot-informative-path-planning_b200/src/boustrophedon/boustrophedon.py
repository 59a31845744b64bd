"""Lawn-mower baselines: drop-in for the reference's src/boustrophedon/boustrophedon.py.

Path generation is host code (deterministic, trivial); the planners matter here as CALLERS of the
accelerated `PointSourceField.predict_spatial_field` (boustrophedon.py:65,143 -- config C1).
"""
import numpy as np

from mripp_b200.src.boustrophedon.bspline import bspline


def _zigzag(x_coords, y_down, y_up):
    """Control polygon: up one column, down the next."""
    cv = []
    for k, x in enumerate(x_coords):
        lo, hi = (y_down, y_up) if k % 2 == 0 else (y_up, y_down)
        cv.append([x, lo])
        cv.append([x, hi])
    return np.array(cv)


def _thin(p, spacing, budget):
    """Waypoints at least `spacing` apart along the sampled curve until `budget` is spent
    (boustrophedon.py:47-60).  Returns (waypoints, index of the last visited sample)."""
    wp = [p[0]]
    covered = 0.0
    i = 0
    for i in range(1, len(p)):
        gap = np.linalg.norm(p[i] - wp[-1])
        step = np.linalg.norm(p[i] - p[i - 1])
        if covered + step > budget:
            break
        if gap >= spacing:
            wp.append(p[i])
        covered += step
    return wp, i


class Boustrophedon:
    """boustrophedon.py:9-65."""

    def __init__(self, scenario, d_waypoint_distance=2.5, budget=375, line_spacing=5):
        self.scenario = scenario
        self.workspace_size = scenario.workspace_size
        self.d_waypoint_distance = d_waypoint_distance
        self.budget = budget
        self.full_path = None
        self.obs_wp = None
        self.name = "Boustrophedon"
        self.line_spacing = line_spacing

    def run(self):
        ws = self.workspace_size
        xs = np.arange(ws[0] + 0.5, ws[1], self.line_spacing)
        if xs[-1] + 0.5 != ws[1]:
            xs = np.append(xs, ws[0] - 0.5)  # sic: boustrophedon.py:35
        p = bspline(_zigzag(xs, ws[2] + 0.5, ws[3] - 0.5), n=1000, degree=2)
        wp, last = _thin(p, self.d_waypoint_distance, self.budget)
        self.obs_wp = np.array(wp)
        self.full_path = p[: last + 1].T
        self.measurements = self.scenario.simulate_measurements(self.obs_wp)
        return self.scenario.predict_spatial_field(self.obs_wp, self.measurements)


class MR_Boustrophedon:
    """boustrophedon.py:68-143."""

    def __init__(self, scenario, num_agents=1, d_waypoint_distance=2.5, budget=375, line_spacing=4):
        self.scenario = scenario
        self.workspace_size = scenario.workspace_size
        self.d_waypoint_distance = d_waypoint_distance
        self.total_budget = budget
        self.num_agents = num_agents
        self.line_spacing = line_spacing
        self.agents_full_path = [[] for _ in range(num_agents)]
        self.agents_measurements = [[] for _ in range(num_agents)]
        self.agents_obs_wp = [[] for _ in range(num_agents)]
        self.name = "MR_Boustrophedon"
        ws = self.workspace_size
        if num_agents > 1:
            self.agent_positions = [np.array([i * ws[1] / num_agents, ws[0] + 0.5]) for i in range(num_agents)]
        else:
            self.agent_positions = [np.array([ws[0] + 0.5, ws[2] + 0.5])]

    def run(self):
        ws = self.workspace_size
        for i in range(self.num_agents):
            x_start = self.agent_positions[i][0]
            x_end = ws[1] - 0.5 if i == self.num_agents - 1 else self.agent_positions[i + 1][0] - self.line_spacing
            xs = np.arange(x_start, x_end, self.line_spacing)
            if xs[-1] + 0.5 != x_end:
                xs = np.append(xs, x_end)
            p = bspline(_zigzag(xs, 0.5 + ws[2], ws[3] - 0.5), n=1000, degree=2)
            wp, last = _thin(p, self.d_waypoint_distance, self.total_budget)
            self.agents_obs_wp[i] = np.array(wp)
            self.agents_full_path[i] = p[: last + 1]
        return self.finalize_all_agents()

    def finalize_all_agents(self):
        self.obs_wp = np.vstack(self.agents_obs_wp)
        self.full_path = np.vstack(self.agents_full_path).T
        self.measurements = self.scenario.simulate_measurements(self.obs_wp)
        return self.scenario.predict_spatial_field(self.obs_wp, self.measurements)


# the name the shipped config lists (config/example_setup.json:46)
MultiAgentBoustrophedon = MR_Boustrophedon
