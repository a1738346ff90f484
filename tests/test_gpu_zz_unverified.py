"""GPU tests written after round 2's GPU budget was spent: non-strict xfail until their first GPU run (a pass shows
as XPASS, a failure cannot hide). The file name sorts after every verified GPU test file on purpose."""
import numpy as np
import pytest

from engine_util import build_engine
from oracle import mppi_oracle as mo

pytestmark = pytest.mark.gpu


# Written after this round's GPU budget was spent: the kernel branches below (index clipping at the map border,
# lane discs cut by the border) were checked through a Python twin of the per-cell formula against the oracle
# painter (300 random maps) and the painter against the LIVE reference (tests/test_reference_live.py), but this
# test itself has not run on a GPU yet - hence non-strict xfail: a pass shows as XPASS, a failure cannot hide.
@pytest.mark.xfail(strict=False, reason="first GPU run of this test is the round-end suite")
def test_device_rasteriser_border_cases_match_the_oracle_painter():
    from mppi_playground_b200 import maps

    rng = np.random.default_rng(11)
    _, solver = build_engine(dict(model="navigation2d", horizon=10, num_samples=256, sigmas=[0.5, 0.5], lambda_=1.0))
    for trial in range(4):
        size = [(12, 8), (4, 4), (20, 6), (8, 14)][trial]
        cell = [0.07, 0.5, 0.11, 0.05][trial]
        r = maps.ObstacleRaster(map_size=size, cell_size=cell)
        for _ in range(10):
            r.add_circle_obstacle(rng.uniform(-0.6, 0.6, size=2) * np.array(size), float(rng.uniform(0.2, 1.5)))
        for _ in range(6):
            r.add_rectangle_obstacle(rng.uniform(-0.6, 0.6, size=2) * np.array(size), float(rng.uniform(0.3, 3.0)),
                                     float(rng.uniform(0.3, 3.0)))
        grid = solver.rasterise_map(0, r, want_grid=True)
        want = mo.paint_obstacle_map(r.width, r.height, r.discs, r.rects)
        np.testing.assert_array_equal(grid.cpu().numpy(), want)
        assert want[0].sum() + want[-1].sum() + want[:, 0].sum() + want[:, -1].sum() > 0  # the border is involved
        # lane mode: a centre line that runs off the map, discs cut by the border
        t = np.linspace(0.0, 1.0, 200)
        lane = np.stack([(t - 0.3) * size[0] * 1.2, np.sin(6 * t) * size[1] * 0.55, np.zeros_like(t)], axis=1)
        lr = maps.LaneRaster(lane, lane_width=float(rng.uniform(0.4, 2.0)), map_size=size, cell_size=cell)
        lgrid = solver.rasterise_map(0, lr, want_grid=True)
        lwant = mo.paint_lane_map(lr.width, lr.height, [(x, y) for x, y, _ in lr.discs], lr.r2)
        np.testing.assert_array_equal(lgrid.cpu().numpy(), lwant)
        assert 0 < lwant.sum() < lwant.size


@pytest.mark.xfail(strict=False, reason="first GPU run of this test is the round-end suite")
def test_navigation2d_with_a_map_larger_than_shared_memory_matches_the_oracle():
    """The global-memory instantiation for the one-map model (the racing twin of this test is verified:
    tests/test_gpu_epilogue.py::test_maps_larger_than_shared_memory_use_the_global_path)."""
    import torch

    import mppi_playground_b200 as eng
    from engine_util import ParityStats, assert_parity

    rng = np.random.default_rng(4)
    W = 3000
    cell, origin, lim = 0.01, (1500, 1500), (-15.0, 15.0, -15.0, 15.0)
    grid = np.zeros((W, W), dtype=np.float32)
    for _ in range(300):
        cx, cy, r = rng.integers(0, W), rng.integers(0, W), rng.integers(10, 80)
        grid[max(cx - r, 0): cx + r, max(cy - r, 0): cy + r] = 1.0
    model = eng.Navigation2DModel(grid, cell, origin, u_min=(0.0, -1.0), u_max=(2.0, 1.0), goal=(9.0, 9.0), lim=lim)
    solver = eng.MPPI(horizon=40, num_samples=4096, dim_state=3, dim_control=2, dynamics=model.dynamics,
                      cost_func=model.cost_func, u_min=model.u_min, u_max=model.u_max, sigmas=torch.tensor([0.5, 0.5]),
                      lambda_=1.0)
    omodel = mo.Navigation2DModel(mo.GridMap(torch.from_numpy(grid), cell, origin), u_min=(0.0, -1.0), u_max=(2.0, 1.0),
                                  goal=(9.0, 9.0), lim=lim)
    oracle = mo.OracleMPPI(horizon=40, num_samples=4096, dim_state=3, dim_control=2, dynamics=omodel.dynamics,
                           cost_func=omodel.cost, u_min=[0.0, -1.0], u_max=[2.0, 1.0], sigmas=[0.5, 0.5], lambda_=1.0,
                           burn_constructor_draw=False)
    state = torch.tensor([-9.0, -9.0, 0.7])
    for s in range(2):
        noise = solver.sampler_noise().cpu()
        action, states = solver.forward(state)
        assert solver.launch_info()["smem_bytes"] < 60000  # nothing staged
        tr = oracle.forward(state, noise=noise)
        st = ParityStats(solver._costs.cpu().numpy(), tr.costs.numpy(), action.cpu().numpy(), tr.action_seq.numpy(),
                         states.cpu().numpy(), tr.state_seq.numpy(), 1.0, 1.0)
        assert_parity(st)
        oracle.prev_action_seq = action.cpu().clone()
        state = states[0, 1].cpu().clone()


@pytest.mark.xfail(strict=False, reason="first GPU run of this test is the round-end suite")
def test_device_resident_control_loop_tracks_the_oracle_loop():
    """The three pieces either side of the solve chained on the device like example/racing.py:229-237 chains them on
    the host: reference path (mppi_refpath_update) -> solve -> epilogue (env.step, collision flags, top samples),
    the next state never leaving the GPU; the oracle runs the same loop on the engine's noise. Each piece has its
    own verified test; this one checks that they compose."""
    import torch

    import mppi_playground_b200 as eng
    from engine_util import build_oracle
    from oracle import fixtures as fx

    cfg = dict(model="racing", horizon=40, num_samples=4096, sigmas=[0.5, 0.1], lambda_=1.0, use_sg_filter=True)
    model, solver = build_engine(cfg)
    omodel, oracle = build_oracle(cfg, burn_constructor_draw=False)
    env = fx.load_env_racing()
    gen = eng.RacingReferencePath(env.center_path, 40, v_max=env.v_max)
    goal = (float(env.center_path[-1][0]), float(env.center_path[-1][1]))
    state_dev, state_host, cind = env.start_state.clone().cuda(), env.start_state.clone(), 0
    for step in range(5):
        model.reference_path_tensor = gen.update(state_dev)
        ref, cind = eng.racing_reference_path(state_host, env.center_path, cind, 40, v_max=env.v_max)
        omodel.reference_path = ref
        np.testing.assert_allclose(model.reference_path_tensor.cpu().numpy(), ref.numpy(), rtol=0, atol=1e-3)
        noise = solver.sampler_noise().cpu()
        action, states = solver.forward(state_dev)
        nxt, reached, coll, (traj, w) = solver.step_epilogue(action, states, state=state_dev, goal=goal,
                                                             goal_threshold=1.0, top_n=100)
        tr = oracle.forward(state_host, noise=noise)
        onxt, oreached = mo.env_step(omodel, state_host, tr.action_seq[0], goal, 1.0)
        np.testing.assert_allclose(action.cpu().numpy(), tr.action_seq.numpy(), rtol=0, atol=2e-3)
        np.testing.assert_allclose(nxt.cpu().numpy(), onxt.numpy(), rtol=0, atol=2e-3)
        assert bool(reached) == oreached and tuple(coll.shape) == (1, 41) and tuple(traj.shape) == (100, 41, 4)
        assert bool((w[:-1] >= w[1:]).all())
        # keep the two loops on the same trajectory (they differ by the parity tolerance per step)
        oracle.prev_action_seq = action.cpu().clone()
        oracle.history = solver._actions_history_for_sg.cpu().clone()
        state_dev, state_host = nxt, nxt.cpu().clone()
