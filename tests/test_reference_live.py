"""Checks against the LIVE reference (only where /root/reference exists, i.e. the build
container; skipped on the GPU box): the engine's callable resolution works on the real objects
the examples pass, and the fixtures / host twins agree with the reference's own code."""
import numpy as np
import pytest
import torch

from oracle import ref_harness as rh

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference checkout not present")


def test_racing_objects_resolve_and_match_fixtures():
    from mppi_playground_b200 import _capi, models
    from oracle import fixtures as fx

    env, ctl, _ = rh.make_racing()
    b = models.resolve(env.dynamics, ctl.cost_function, 4, 2)
    assert b.model_id == _capi.MODEL_RACING
    fixture = fx.load_env_racing()
    want = [*np.ravel(list(zip(fixture.u_min, fixture.u_max))), fixture.wheelbase, fixture.v_max, *fixture.lim, 0.1,
            *fixture.Q]
    assert b.params() == pytest.approx(want)
    (og, oc, ox, oy), (lg, lc, lx, ly) = b.maps()
    np.testing.assert_array_equal(og.numpy(), fixture.obstacle)
    np.testing.assert_array_equal(lg.numpy(), fixture.lane)
    assert (oc, ox, oy, lc, lx, ly) == (0.1, 400.0, 400.0, 0.1, 400.0, 400.0)
    # the per-solve reference path is read from the live controller
    ctl.reference_path, _ = ctl.calc_ref_trajectory(env._robot_state, env.racing_center_path, 0, 25, DL=0.1,
                                                    lookahead_distance=3, reference_path_interval=0.85)
    assert b.reference_path() is ctl.reference_path


def test_navigation2d_objects_resolve():
    from mppi_playground_b200 import _capi, models

    env, _ = rh.make_navigation2d()
    b = models.resolve(env.dynamics, env.cost_function, 3, 2)
    assert b.model_id == _capi.MODEL_NAVIGATION2D
    assert b.params() == pytest.approx([0.0, 2.0, -1.0, 1.0, 9.0, 9.0, -10.0, 10.0, -10.0, 10.0, 0.1, 10000.0])
    assert b.maps()[0][1:] == (0.1, 100.0, 100.0)


@pytest.mark.parametrize("example,names,want", [("pendulum", ["dynamics", "cost_function"], "pendulum"),
                                                ("cartpole", ["dynamics", "stage_cost"], "cartpole"),
                                                ("mountaincar", ["dynamics", "cost_func"], "mountaincar"),
                                                ("mujoco_cartpole", ["dynamics", "cost_func"],
                                                 "cartpole_continuous")])
def test_example_closures_resolve(example, names, want):
    from mppi_playground_b200 import models

    dyn, cost = rh.extract_closures(example, names)
    ds = {"pendulum": 2, "cartpole": 4, "mountaincar": 2, "mujoco_cartpole": 4}[example]
    assert models.resolve(dyn, cost, ds, 1).name == want


def test_reference_path_twin_matches_calc_ref_trajectory():
    import mppi_playground_b200 as eng

    env, ctl, _ = rh.make_racing()
    state, cind_ref, cind = env._robot_state.clone(), 0, 0
    for _ in range(3):
        want, cind_ref = ctl.calc_ref_trajectory(state, env.racing_center_path, cind_ref, 80, DL=0.1,
                                                 lookahead_distance=3, reference_path_interval=0.85)
        got, cind = eng.racing_reference_path(state, env.racing_center_path, cind, 80, v_max=float(env.V_MAX))
        assert cind == cind_ref
        np.testing.assert_array_equal(got.numpy(), want.numpy())
        state = torch.tensor([want[5, 0], want[5, 1], want[5, 2], 4.0])


def test_oracle_matches_live_reference_on_a_fresh_seed():
    """Beyond the recorded fixtures: a new seed / config, reference and oracle side by side."""
    from oracle import mppi_oracle as mo

    ns = rh.load_reference()
    dyn, cost = rh.extract_closures("cartpole", ["dynamics", "stage_cost"])
    kw = dict(horizon=12, num_samples=300, dim_state=4, dim_control=1, u_min=torch.tensor([-3.0]),
              u_max=torch.tensor([3.0]), sigmas=torch.tensor([1.0]), lambda_="ESSPS", exploration=0.1,
              use_sg_filter=True, seed=7)
    ref = ns.MPPI(dynamics=dyn, cost_func=cost, **kw)
    m = mo.CartpoleModel()
    okw = {k: (v.tolist() if torch.is_tensor(v) else v) for k, v in kw.items()}
    ora = mo.OracleMPPI(dynamics=m.dynamics, cost_func=m.cost, **okw)
    state = torch.tensor([0.0, 0.2, 0.03, -0.1])
    for _ in range(3):
        a, s = ref.forward(state.clone())
        tr = ora.forward(state.clone(), noise=ref._action_noises)
        np.testing.assert_array_equal(tr.action_seq.numpy(), a.numpy())
        np.testing.assert_array_equal(tr.state_seq.numpy(), s.numpy())
        assert tr.lam == ref._lambda
        state = s[0, 1].clone()


class _HostOnlyMPPI:
    """Stands in for the engine's MPPI where there is no GPU: runs the real host half of its constructor
    (mppi_playground_b200.mppi.host_setup: callable resolution + MppiConfig) on whatever the example passes."""

    def __init__(self, **kw):
        from mppi_playground_b200.mppi import host_setup

        kw.pop("device", None), kw.pop("dtype", None)
        self._horizon = kw["horizon"]
        self.setup = host_setup(**kw)


def test_real_racing_example_constructs_through_the_dropin_in_its_own_order():
    """example/racing.py:16-58 builds MPPI(cost_func=self.cost_function, ...) BEFORE it assigns Qc..Qdin,
    reference_path and the maps: resolution must not depend on those attributes, the
    constructor must tolerate their absence, and the first solve must see the real values."""
    from mppi_playground_b200 import _capi
    from oracle import fixtures as fx

    ns = rh.load_reference()
    mod = rh.load_example_with_solver("racing", _HostOnlyMPPI)
    with rh._cwd(rh.REFERENCE_ROOT):
        env = ns.RacingEnv()
    ctl = mod.racing_controller(env, debug=False)  # raises NotImplementedError / AttributeError before the fix
    hs = ctl.solver.setup
    assert hs.binding.model_id == _capi.MODEL_RACING and hs.cfg.num_model_params == _capi.RACING_NUM_PARAMS
    assert hs.params[11:] == [0.0] * 6  # the weights did not exist yet: placeholders
    assert hs.binding.reference_path() is None
    with pytest.raises(ValueError, match="must be set"):
        hs.binding.maps()
    # what forward() reads on the first solve (racing.py:227 set_cost_map, :73-81 reference path)
    ctl.set_cost_map(env._obstacle_map, env._lane_map)
    fixture = fx.load_env_racing()
    assert hs.binding.params(strict=True)[11:] == pytest.approx(fixture.Q)
    assert len(hs.binding.maps()) == 2
    ctl.reference_path, _ = ctl.calc_ref_trajectory(env._robot_state, env.racing_center_path, 0, ctl.solver._horizon,
                                                    DL=0.1, lookahead_distance=3, reference_path_interval=0.85)
    assert tuple(hs.binding.reference_path().shape) == (26, 4)
    del ctl.Qc  # a weight that is still missing at solve time is the reference's own AttributeError
    with pytest.raises(AttributeError):
        hs.binding.params(strict=True)


def test_real_navigation2d_example_arguments_resolve():
    from mppi_playground_b200 import _capi
    from mppi_playground_b200.mppi import host_setup

    env, _ = rh.make_navigation2d()
    hs = host_setup(horizon=30, num_samples=3000, dim_state=3, dim_control=2, dynamics=env.dynamics,
                    cost_func=env.cost_function, u_min=env.u_min, u_max=env.u_max, sigmas=torch.tensor([0.5, 0.5]),
                    lambda_="ESSPS")  # example/navigation2d.py:16-27
    assert hs.binding.model_id == _capi.MODEL_NAVIGATION2D and hs.cfg.lambda_mode == _capi.LAMBDA_ESSPS
    assert hs.cfg.essps_target_ess == 300.0


def test_reference_shaped_doubles_carry_what_the_live_objects_carry():
    """tests/reference_shapes.py stands in for the reference's objects on the GPU box (tests/test_gpu_dropin.py).
    Here, where the reference is importable, the doubles are held against the real thing: same class names, same
    attribute values for everything the bindings read, and the bindings produce identical parameter blocks /
    grids / geometry from either."""
    import reference_shapes as rs
    from mppi_playground_b200 import models

    env, ctl, _ = rh.make_racing()
    d_env = rs.RacingEnv(device="cpu")
    assert type(d_env).__name__ == type(env).__name__ and rs.racing_controller.__name__ == type(ctl).__name__
    for name in ("u_min", "u_max", "L", "V_MAX", "racing_center_path", "_robot_state"):
        np.testing.assert_array_equal(getattr(d_env, name).numpy(), getattr(env, name).numpy(), err_msg=name)
    for m_name in ("_obstacle_map", "_lane_map"):
        live, dbl = getattr(env, m_name), getattr(d_env, m_name)
        assert type(dbl).__name__ == type(live).__name__
        assert dbl._cell_size == live._cell_size and list(dbl._cell_map_origin) == list(live._cell_map_origin)
        np.testing.assert_array_equal(dbl._map_torch.numpy(), live._map_torch.numpy())
        if m_name == "_obstacle_map":
            assert dbl.x_lim == live.x_lim and dbl.y_lim == live.y_lim

    class _NoSolver:
        def __init__(self, **kw):
            self._horizon = kw["horizon"]

    d_ctl = rs.racing_controller(d_env, _NoSolver)
    d_ctl.set_cost_map(d_env._obstacle_map, d_env._lane_map)
    live_b = models.resolve(env.dynamics, ctl.cost_function, 4, 2)
    dbl_b = models.resolve(d_env.dynamics, d_ctl.cost_function, 4, 2)
    assert type(live_b) is type(dbl_b) and live_b.params() == dbl_b.params()
    for (g0, *geo0), (g1, *geo1) in zip(live_b.maps(), dbl_b.maps()):
        assert geo0 == geo1 and torch.equal(g0, g1)
    # the double's update() produces the reference's own look-ahead path
    want, ind = ctl.calc_ref_trajectory(env._robot_state, env.racing_center_path, 0, 25, DL=0.1,
                                        lookahead_distance=3, reference_path_interval=0.85)
    d_ctl.solver.forward = lambda state: None
    d_ctl.update(d_env._robot_state, d_env.racing_center_path)
    np.testing.assert_array_equal(d_ctl.reference_path.numpy(), want.numpy())
    assert d_ctl.current_path_index == ind

    nav, _ = rh.make_navigation2d()
    d_nav = rs.Navigation2DEnv(device="cpu")
    assert type(d_nav).__name__ == type(nav).__name__
    for name in ("u_min", "u_max", "_goal_pos", "_robot_state"):
        np.testing.assert_array_equal(getattr(d_nav, name).numpy(), getattr(nav, name).numpy(), err_msg=name)
    lb = models.resolve(nav.dynamics, nav.cost_function, 3, 2)
    db = models.resolve(d_nav.dynamics, d_nav.cost_function, 3, 2)
    assert lb.params() == db.params() and lb.maps()[0][1:] == db.maps()[0][1:]
    assert torch.equal(lb.maps()[0][0], db.maps()[0][0])


def test_live_reference_mpo_lambda_moves_under_one_ulp_of_cost_noise():
    """The same experiment as tests/test_oracle_golden.py::test_mpo_lambda_moves_under_one_ulp_of_cost_noise, on
    the LIVE reference's MPPI: one ulp on every stage cost moves the reference's own MPO lambda by > 1e-3."""
    from types import SimpleNamespace

    from oracle import fixtures as fx
    from test_oracle_golden import mpo_trajectory_under_ulp_noise

    case = fx.load_case("navigation2d_mpo_expl")
    env, ns = rh.make_navigation2d()
    kw = {k: v for k, v in fx.solver_kwargs(case.cfg).items()}

    class _Ref:  # adapts the reference's forward() to the (lam_next, action_seq) record the helper reads
        def __init__(self, cost):
            self.m = ns.MPPI(dim_state=3, dim_control=2, dynamics=env.dynamics, cost_func=cost, u_min=env.u_min,
                             u_max=env.u_max, sigmas=torch.tensor(case.cfg["sigmas"]), **kw)

        def forward(self, state, noise):
            import unittest.mock as um

            # inject the recorded noise: the reference draws through self._noise_distribution.rsample (mppi.py:261)
            with um.patch.object(self.m._noise_distribution, "rsample", lambda sample_shape: noise.clone()):
                a, _ = self.m.forward(state.clone())
            return SimpleNamespace(lam_next=float(self.m._lambda), action_seq=a.detach())

    def make(wrap):
        return None, _Ref(wrap(env.cost_function))

    base_l, base_a = mpo_trajectory_under_ulp_noise(make, case, "base")
    np.testing.assert_array_equal(base_l, case.lam_next)  # the unperturbed run IS the recorded golden run
    np.testing.assert_array_equal(base_a, case.action_seq)
    l, a = mpo_trajectory_under_ulp_noise(make, case, "up")
    assert float(np.max(np.abs(l - base_l) / np.abs(base_l))) > 1e-3
    assert float(np.max(np.abs(a - base_a))) > 3e-4


def test_map_rasters_from_live_map_objects_equal_the_live_grids():
    """§8f row 4 on the live objects: the host shape conversions of mppi_playground_b200/maps.py (the product
    half) fed from the reference's own map objects, painted by the oracle's restatement, equal the grids the
    reference's loops / distance transform produced - including a fresh map the fixtures never saw (another seed,
    shapes that leave the map)."""
    import sys

    from mppi_playground_b200 import maps
    from oracle import mppi_oracle as mo

    env, _, _ = rh.make_racing()
    om, lm = env._obstacle_map, env._lane_map
    r = maps.ObstacleRaster.from_obstacle_map(om, env.map_size)
    assert (r.width, r.height) == om._map.shape and list(r.origin) == list(om._cell_map_origin)
    assert r.x_lim == om.x_lim and r.y_lim == om.y_lim
    np.testing.assert_array_equal(mo.paint_obstacle_map(r.width, r.height, r.discs, r.rects), om._map)
    lane = maps.LaneRaster(env.racing_center_path.numpy(), env.line_width * 0.8, env.map_size, env.cell_size)
    assert list(lane.origin) == list(lm._cell_map_origin) and lane.x_lim == lm.x_lim
    np.testing.assert_array_equal(mo.paint_lane_map(lane.width, lane.height, [(x, y) for x, y, _ in lane.discs], lane.r2),
                                  lm._map)
    # a fresh reference map: circles and rectangles partly outside, odd cell size
    ObstacleMap = sys.modules["envs.obstacle_map_2d"].ObstacleMap
    rng = np.random.default_rng(11)
    live = ObstacleMap(map_size=(12, 8), cell_size=0.07, device=torch.device("cpu"))
    mine = maps.ObstacleRaster(map_size=(12, 8), cell_size=0.07)
    for _ in range(12):
        c, rad = rng.uniform(-7, 7, size=2), float(rng.uniform(0.2, 1.5))
        live.add_circle_obstacle(c, rad)
        mine.add_circle_obstacle(c, rad)
    for _ in range(8):
        c, w, h = rng.uniform(-7, 7, size=2), float(rng.uniform(0.3, 3.0)), float(rng.uniform(0.3, 3.0))
        live.add_rectangle_obstacle(c, w, h)
        mine.add_rectangle_obstacle(c, w, h)
    assert (mine.width, mine.height) == live._map.shape
    np.testing.assert_array_equal(mo.paint_obstacle_map(mine.width, mine.height, mine.discs, mine.rects), live._map)
    assert live._map[0].sum() + live._map[-1].sum() > 0  # something did smear onto a border row


def test_oracle_epilogue_matches_live_env_on_fresh_states():
    """§8f row 2 beyond the recorded vectors: env.step / collision_check of the live envs against the oracle's
    restatement on random states, actions beyond the bounds and positions beyond the map."""
    from oracle import fixtures as fx
    from oracle import mppi_oracle as mo

    rng = np.random.default_rng(5)
    for make, model, lim, thr in ((lambda: rh.make_racing()[0], fx.oracle_racing_model(), 40.0, 1.0),
                                  (lambda: rh.make_navigation2d()[0], fx.oracle_navigation2d_model(), 10.0, 0.5)):
        env = make()
        ds = model.dim_state
        goal = env._goal_pos.numpy()
        for _ in range(40):
            st = np.zeros(ds, dtype=np.float32)
            st[:2] = rng.uniform(-1.1 * lim, 1.1 * lim, size=2) if rng.random() < 0.5 else goal + rng.uniform(-1.5, 1.5, 2) * thr
            st[2:] = rng.uniform(-4.0, 4.0, size=ds - 2)
            act = (rng.uniform(-1, 1, size=2) * 4.0).astype(np.float32)
            env._robot_state = torch.tensor(st)
            want_next, want_goal = env.step(torch.tensor(act))
            nxt, reached = mo.env_step(model, torch.tensor(st), torch.tensor(act), goal, thr)
            np.testing.assert_array_equal(nxt.numpy(), want_next.numpy())
            assert reached == bool(want_goal)
        traj = torch.tensor(rng.uniform(-1.2 * lim, 1.2 * lim, size=(2, 300, ds)).astype(np.float32))
        obstacle = model.obstacle if ds == 4 else model.grid
        np.testing.assert_array_equal(mo.collision_check(obstacle, traj).numpy(),
                                      env.collision_check(state=traj).numpy())
