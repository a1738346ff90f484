"""Test doubles shaped like the objects the reference's examples hand to ``pi_mpc.MPPI``.

``/root/reference`` does not exist on the GPU box, so the drop-in path (``compat/pi_mpc`` ->
``models.resolve`` -> ``_Reference*`` bindings) is driven there by stand-ins that carry the SAME class names and
the SAME attribute layout as the reference's objects - device tensors where the reference has device tensors,
numpy arrays where it has numpy - filled from the fixtures recorded from the live reference
(tests/golden/env_*.npz). They contain no reference arithmetic: the ``dynamics`` / ``cost_function`` methods
exist to be passed as bound methods and raise if anything calls them (the engine resolves them to device models
and must never execute Python on the rollout path).

Attribute sources (reference file:line):
  ObstacleMap / LaneMap   src/envs/obstacle_map_2d.py:60-100,164-166; src/envs/lane_map_2d.py:55-61,84-88
  RacingEnv               src/envs/racing_env.py:24-116
  racing_controller       example/racing.py:16-108 (solver built BEFORE the weights / maps are assigned)
  Navigation2DEnv         src/envs/navigation_2d.py:26-90
  GoalInDangerZoneEnv     src/envs/goal_in_danger_zone.py:40-110
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch

from oracle import fixtures as fx


class _GridMap:
    def __init__(self, grid: np.ndarray, cell_size: float, device, dtype=torch.float32):
        self._device, self._dtype = torch.device(device), dtype
        self._map = np.asarray(grid, dtype=np.float64)  # the reference keeps a float64 numpy grid
        self._cell_size = cell_size
        self._cell_map_origin = np.array([self._map.shape[0] / 2, self._map.shape[1] / 2]).astype(int)
        self._torch_cell_map_origin = torch.from_numpy(self._cell_map_origin).to(self._device, self._dtype)
        xr, yr = cell_size * self._map.shape[0], cell_size * self._map.shape[1]
        self.x_lim, self.y_lim = [-xr / 2, xr / 2], [-yr / 2, yr / 2]
        self._map_torch = None

    def convert_to_torch(self) -> torch.Tensor:
        self._map_torch = torch.from_numpy(self._map).to(self._device, self._dtype)
        return self._map_torch


class ObstacleMap(_GridMap):
    pass


class LaneMap(_GridMap):
    def __init__(self, grid, cell_size, device, dtype=torch.float32):
        super().__init__(grid, cell_size, device, dtype)
        self.convert_to_torch()  # lane_map_2d.py:84-88 builds the torch grid in its constructor


def _never(*_a, **_k):
    raise AssertionError("the engine must not call the Python dynamics / cost on the rollout path")


class RacingEnv:
    def __init__(self, device=torch.device("cuda"), dtype=torch.float32, seed: int = 42):
        self._device, self._dtype = torch.device(device), dtype
        e = fx.load_env_racing()
        self.u_min = torch.tensor(e.u_min, device=self._device, dtype=dtype)
        self.u_max = torch.tensor(e.u_max, device=self._device, dtype=dtype)
        self.L = torch.tensor(1, device=self._device, dtype=dtype)
        self.V_MAX = torch.tensor(8.0, device=self._device, dtype=dtype)
        self.racing_center_path = e.center_path.to(self._device, dtype)
        self.map_size, self.cell_size = (80, 80), 0.1
        self._lane_map = LaneMap(e.lane, self.cell_size, self._device, dtype)
        self._obstacle_map = ObstacleMap(e.obstacle, self.cell_size, self._device, dtype)
        self._obstacle_map.convert_to_torch()
        self._robot_state = e.start_state.to(self._device, dtype)
        self._fixture = e

    def dynamics(self, state, action):
        _never()


class racing_controller:
    """Constructor order of example/racing.py:16-58: the solver first, then the weights, then ``None`` maps."""

    def __init__(self, env, MPPI, horizon=25, num_samples=4000, **solver_kw):
        self.current_path_index = 0
        self.solver = MPPI(horizon=horizon, num_samples=num_samples, dim_state=4, dim_control=2,
                           dynamics=env.dynamics, cost_func=self.cost_function, u_min=env.u_min, u_max=env.u_max,
                           sigmas=torch.tensor([0.5, 0.1]), lambda_=1.0, **solver_kw)
        self.env = env
        self.Qc, self.Ql, self.Qv, self.Qo, self.Qin, self.Qdin = 2.0, 3.0, 2.0, 10000.0, 0.01, 0.5
        self.reference_path = None
        self.obstacle_map = None
        self.lane_map = None

    def update(self, state, racing_center_path):
        import mppi_playground_b200 as eng

        # calc_ref_trajectory's job (example/racing.py:161-218), by the host twin that is checked bit for bit
        # against the reference in tests/test_reference_live.py; note solver._horizon (racing.py:77)
        self.reference_path, self.current_path_index = eng.racing_reference_path(
            torch.as_tensor(state).to(racing_center_path.device), racing_center_path, self.current_path_index,
            self.solver._horizon, v_max=8.0)
        return self.solver.forward(state=state)

    def get_top_samples(self, num_samples=300):
        return self.solver.get_top_samples(num_samples=num_samples)

    def set_cost_map(self, obstacle_map, lane_map):
        self.obstacle_map, self.lane_map = obstacle_map, lane_map

    def cost_function(self, state, action, info):
        _never()


class Navigation2DEnv:
    def __init__(self, device=torch.device("cuda"), dtype=torch.float32, seed: int = 42):
        self._device, self._dtype = torch.device(device), dtype
        e = fx.load_env_navigation2d()
        self._obstacle_map = ObstacleMap(e.obstacle, e.cell, self._device, dtype)
        self._obstacle_map.convert_to_torch()
        self._start_pos = e.start_state[:2].to(self._device, dtype)
        self._goal_pos = torch.tensor(e.goal, device=self._device, dtype=dtype)
        self._robot_state = e.start_state.to(self._device, dtype)
        self.u_min = torch.tensor(e.u_min, device=self._device, dtype=dtype)
        self.u_max = torch.tensor(e.u_max, device=self._device, dtype=dtype)

    def dynamics(self, state, action):
        _never()

    def cost_function(self, state, action, info):
        _never()


class GoalInDangerZoneEnv:
    def __init__(self, goal=(-2.5, 6.0), center=(0.0, 0.0), radius=10.0):
        self._v_min, self._v_max, self._omega_min, self._omega_max, self._dt = -1.0, 1.0, -1.0, 1.0, 0.1
        self._goal = np.array(goal, dtype=np.float64)
        self._danger_zone = SimpleNamespace(center=np.array(center, dtype=np.float64), radius=radius)

    def parallel_step(self, state, action):
        _never()

    def parallel_cost(self, state, action, info):
        _never()
