"""CPU tests of the SURVEY 8f rows 2 and 4 checkers: the oracle's control-step epilogue against vectors recorded
from the live reference (tests/golden/epilogue_*.npz, oracle/gen_golden.py:epilogue_cases) and the map painters +
the product's host shape conversions (mppi_playground_b200/maps.py) against the reference's own grids
(env_racing.npz / env_navigation2d.npz, painted by the reference's loops from the shapes in env_shapes.npz)."""
import json
import os

import numpy as np
import pytest
import torch

from mppi_playground_b200 import maps
from oracle import fixtures as fx
from oracle import mppi_oracle as mo

EPILOGUE_CASES = ["epilogue_racing", "epilogue_navigation2d"]


def _same_build(z):
    return json.loads(str(z["versions"]))["torch"] == torch.__version__


@pytest.mark.parametrize("name", EPILOGUE_CASES)
def test_oracle_epilogue_reproduces_reference(name):
    z = np.load(os.path.join(fx.GOLDEN_DIR, f"{name}.npz"))
    cfg = json.loads(str(z["cfg"]))
    model = fx.oracle_model(cfg["model"])
    obstacle = model.obstacle if cfg["model"] == "racing" else model.grid
    goal, thr = z["goal"], float(z["goal_threshold"])
    exact = _same_build(z)
    # closed loop: env.step on the solve's first action, collision flags of its predicted trajectory
    for s in range(len(z["state"])):
        nxt, reached = mo.env_step(model, torch.from_numpy(z["state"][s]), torch.from_numpy(z["action_seq"][s][0]),
                                   goal, thr)
        if exact:
            np.testing.assert_array_equal(nxt.numpy(), z["next_state"][s])
        else:
            np.testing.assert_allclose(nxt.numpy(), z["next_state"][s], rtol=1e-5, atol=1e-5)
        assert reached == bool(z["is_goal"][s])
        coll = mo.collision_check(obstacle, torch.from_numpy(z["state_seq"][s]))
        np.testing.assert_array_equal(coll.numpy(), z["collisions"][s])
    # probes: goal flags on both sides of the threshold, actions beyond the env bounds, positions beyond the map
    assert 0 < int(z["probe_goal"].sum()) < len(z["probe_goal"])
    for st, act, want_next, want_goal in zip(z["probe_state"], z["probe_action"], z["probe_next"], z["probe_goal"]):
        nxt, reached = mo.env_step(model, torch.from_numpy(st), torch.from_numpy(act), goal, thr)
        np.testing.assert_allclose(nxt.numpy(), want_next, rtol=0 if exact else 1e-5, atol=0 if exact else 1e-5)
        assert reached == bool(want_goal)
    coll = mo.collision_check(obstacle, torch.from_numpy(z["coll_probe_in"]))
    np.testing.assert_array_equal(coll.numpy(), z["coll_probe_out"])
    assert 0 < coll.sum() < coll.numel()


def _shapes():
    return np.load(os.path.join(fx.GOLDEN_DIR, "env_shapes.npz"))


def racing_obstacle_raster():
    z = _shapes()
    r = maps.ObstacleRaster(map_size=tuple(int(v) for v in z["racing_map_size"]), cell_size=float(z["racing_cell"]))
    for c, rad in zip(z["racing_circle_centers"], z["racing_circle_radii"]):
        r.add_circle_obstacle(c, float(rad))
    return r


def racing_lane_raster():
    z = _shapes()
    env = fx.load_env_racing()
    return maps.LaneRaster(env.center_path.numpy().astype(np.float64), float(z["racing_lane_width"]),
                           map_size=tuple(int(v) for v in z["racing_map_size"]), cell_size=float(z["racing_cell"]))


def navigation_obstacle_raster():
    z = _shapes()
    r = maps.ObstacleRaster(map_size=tuple(int(v) for v in z["nav_map_size"]), cell_size=float(z["nav_cell"]))
    for c, rad in zip(z["nav_circle_centers"], z["nav_circle_radii"]):
        r.add_circle_obstacle(c, float(rad))
    for c, (w, h) in zip(z["nav_rect_centers"], z["nav_rect_wh"]):
        r.add_rectangle_obstacle(c, float(w), float(h))
    return r


def test_painted_obstacle_maps_equal_the_reference_grids():
    env = fx.load_env_racing()
    r = racing_obstacle_raster()
    assert (r.width, r.height) == env.obstacle.shape and list(r.origin) == env.origin[0]
    grid = mo.paint_obstacle_map(r.width, r.height, r.discs, r.rects)
    np.testing.assert_array_equal(grid, env.obstacle)
    assert int(grid.sum()) == 18602
    nav = fx.load_env_navigation2d()
    r2 = navigation_obstacle_raster()
    assert (r2.width, r2.height) == nav.obstacle.shape and list(r2.origin) == nav.origin
    grid2 = mo.paint_obstacle_map(r2.width, r2.height, r2.discs, r2.rects)
    np.testing.assert_array_equal(grid2, nav.obstacle)
    assert int(grid2.sum()) == 5019 and len(r2.rects) == 7 and len(r2.discs) == 7


def test_painted_lane_map_equals_the_reference_grid():
    env = fx.load_env_racing()
    r = racing_lane_raster()
    assert (r.width, r.height) == env.lane.shape and list(r.origin) == env.origin[1]
    cells = [(x, y) for x, y, _ in r.discs]
    grid = mo.paint_lane_map(r.width, r.height, cells, r.r2)
    np.testing.assert_array_equal(grid, env.lane)
    assert int(grid.sum()) == 445529
    # the integer threshold is exactly what the fp64 EDT comparison accepts
    m = (float(_shapes()["racing_lane_width"]) / 2) / r.cell_size
    assert np.sqrt(np.float64(r.r2)) <= m < np.sqrt(np.float64(r.r2 + 1))


def test_circle_painting_clips_onto_the_border_like_the_reference():
    """obstacle_map_2d.py:121-122 clips INDICES: a disc that leaves the map smears onto the border row."""
    r = maps.ObstacleRaster(map_size=(4, 4), cell_size=0.5)  # 8 x 8 cells
    r.add_circle_obstacle(np.array([-2.4, 0.1]), 1.2)  # centre cell (-1, 4): partly outside
    r.add_rectangle_obstacle(np.array([1.9, 1.9]), 1.0, 3.0)
    grid = mo.paint_obstacle_map(r.width, r.height, r.discs, r.rects)
    # brute force, literally the reference's double loop
    want = np.zeros((8, 8))
    (cx, cy, r2), = r.discs
    rad = int(np.sqrt(r2))
    for i in range(-rad, rad + 1):
        for j in range(-rad, rad + 1):
            if i * i + j * j <= r2:
                want[np.clip(cx + i, 0, 7), np.clip(cy + j, 0, 7)] = 1
    (x0, x1, y0, y1), = r.rects
    want[x0:x1, y0:y1] = 1
    np.testing.assert_array_equal(grid, want)
    assert want[0].sum() > 0
