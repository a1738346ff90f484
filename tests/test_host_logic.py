"""Host-side logic that needs no GPU: callable -> device-model resolution, the
vectorised racing reference path, sample-shard bookkeeping."""
import numpy as np
import pytest
import torch

import mppi_playground_b200 as eng
from mppi_playground_b200 import _capi, models
from oracle import fixtures as fx
from oracle import mppi_oracle as mo


@pytest.mark.parametrize("oracle_cls,want", [(mo.PendulumModel, "pendulum"), (mo.CartpoleModel, "cartpole"),
                                             (mo.MountainCarModel, "mountaincar"),
                                             (mo.CartpoleContinuousModel, "cartpole_continuous")])
def test_example_closures_are_fingerprinted(oracle_cls, want):
    m = oracle_cls()  # same arithmetic as the example closures (pinned by the golden tests)
    b = models.resolve(m.dynamics, m.cost, m.dim_state, m.dim_control)
    assert b.name == want


def test_scripted_closure_is_fingerprinted():
    @torch.jit.script
    def dynamics(state: torch.Tensor, action: torch.Tensor) -> torch.Tensor:
        th = state[:, 0].view(-1, 1)
        thdot = state[:, 1].view(-1, 1)
        u = torch.clamp(action[:, 0].view(-1, 1), -2, 2)
        newthdot = thdot + (-15.0 * torch.sin(th + torch.pi) + 3.0 * u) * 0.05
        newth = th + newthdot * 0.05
        return torch.cat((newth, torch.clamp(newthdot, -8, 8)), dim=1)

    def cost(state, action, info):
        return (((state[:, 0] + torch.pi) % (2 * torch.pi)) - torch.pi) ** 2 + 0.1 * state[:, 1] ** 2

    assert models.resolve(dynamics, cost, 2, 1).name == "pendulum"


def test_unknown_callables_raise_instead_of_falling_back():
    with pytest.raises(NotImplementedError, match="no CPU"):
        models.resolve(lambda s, a: s * 0.5, lambda s, a, i: s[:, 0], 2, 1)
    m = mo.PendulumModel()
    with pytest.raises(NotImplementedError):  # right dynamics, foreign cost
        models.resolve(m.dynamics, lambda s, a, i: s[:, 0] ** 2, 2, 1)


def test_reference_objects_are_recognised_by_class():
    env_fx = fx.load_env_racing()

    class FakeMap:
        def __init__(self, grid, cell, origin):
            self._map_torch, self._cell_size, self._cell_map_origin = torch.from_numpy(grid), cell, np.array(origin)
            self.x_lim, self.y_lim = [-40.0, 40.0], [-40.0, 40.0]

    class RacingEnv:  # same class name and attributes as src/envs/racing_env.py
        def __init__(self):
            self.u_min, self.u_max = torch.tensor([-2.0, -0.25]), torch.tensor([2.0, 0.25])
            self.L, self.V_MAX = torch.tensor(1.0), torch.tensor(8.0)
            self._obstacle_map = FakeMap(env_fx.obstacle, 0.1, (400, 400))

        def dynamics(self, state, action):
            raise AssertionError("never called on the host")

    class racing_controller:  # example/racing.py:16-58
        def __init__(self):
            self.Qc, self.Ql, self.Qv, self.Qo, self.Qin, self.Qdin = 2.0, 3.0, 2.0, 10000.0, 0.01, 0.5
            self.reference_path, self.obstacle_map, self.lane_map = None, None, None

        def cost_function(self, state, action, info):
            raise AssertionError("never called on the host")

    env, ctl = RacingEnv(), racing_controller()
    b = models.resolve(env.dynamics, ctl.cost_function, 4, 2)
    assert b.model_id == _capi.MODEL_RACING
    assert b.params() == pytest.approx([-2.0, 2.0, -0.25, 0.25, 1.0, 8.0, -40.0, 40.0, -40.0, 40.0, 0.1, 2.0, 3.0,
                                        2.0, 10000.0, 0.01, 0.5])
    with pytest.raises(ValueError, match="must be set"):
        b.maps()  # maps arrive later through set_cost_map (example/racing.py:227)
    ctl.obstacle_map = env._obstacle_map
    ctl.lane_map = FakeMap(env_fx.lane, 0.1, (400, 400))
    assert len(b.maps()) == 2 and b.maps()[1][1:] == (0.1, 400.0, 400.0)
    ctl.Qc = 5.0
    assert b.params()[11] == 5.0  # weights are re-read every solve
    with pytest.raises(ValueError):
        models.resolve(env.dynamics, ctl.cost_function, 3, 2)


def test_descriptor_params_follow_the_documented_layout():
    env = fx.load_env_racing()
    m = eng.RacingModel(env.obstacle, env.lane, cell_size=env.cell, origin=env.origin, u_min=env.u_min,
                        u_max=env.u_max, wheelbase=env.wheelbase, v_max=env.v_max, lim=env.lim)
    assert len(m.params()) == _capi.RACING_NUM_PARAMS
    nav = fx.load_env_navigation2d()
    n = eng.Navigation2DModel(nav.obstacle, nav.cell, nav.origin, u_min=nav.u_min, u_max=nav.u_max, goal=nav.goal,
                              lim=nav.lim)
    assert len(n.params()) == _capi.NAV2D_NUM_PARAMS and n.params()[4:6] == [9.0, 9.0]
    assert models.resolve(m.dynamics, m.cost_func, 4, 2) is m
    with pytest.raises(ValueError):
        models.resolve(m.dynamics, n.cost_func, 4, 2)


def test_racing_reference_path_matches_the_oracle_restatement():
    env = fx.load_env_racing()
    state, cind = env.start_state.clone(), 0
    for step in range(5):
        want, wi = mo.racing_reference_path(state, env.center_path, cind, 80, v_max=env.v_max)
        got, gi = eng.racing_reference_path(state, env.center_path, cind, 80, v_max=env.v_max)
        assert gi == wi
        np.testing.assert_array_equal(got.numpy(), want.numpy())
        cind = gi
        state = torch.tensor([want[3, 0], want[3, 1], want[3, 2], 5.0])
    # end of the course: every target speed drops to zero (example/racing.py:213-216)
    got, _ = eng.racing_reference_path(state, env.center_path, len(env.center_path) - 40, 80)
    want, _ = mo.racing_reference_path(state, env.center_path, len(env.center_path) - 40, 80)
    np.testing.assert_array_equal(got.numpy(), want.numpy())
    assert float(got[:, 3].abs().max()) == 0.0


def test_golden_refpath_reproduced():
    """The recorded reference paths come from the reference's calc_ref_trajectory."""
    case, env = fx.load_case("racing_sg"), fx.load_env_racing()
    cind = 0
    for s in range(case.n_solves):
        got, cind = eng.racing_reference_path(torch.from_numpy(case.state[s]), env.center_path, cind, 80,
                                              v_max=env.v_max)
        np.testing.assert_array_equal(got.numpy(), case.refpath[s])


@pytest.mark.parametrize("K,world", [(65536, 8), (1000, 3), (7, 8), (1048576, 8)])
def test_shard_bounds_partition_the_samples(K, world):
    spans = [eng.shard_bounds(K, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == K
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        eng.shard_bounds(K, world, world)


def test_goal_in_danger_zone_env_is_recognised_and_goal_is_reread():
    class Zone:
        center, radius = [0.0, 0.0], 10.0

    class GoalInDangerZoneEnv:  # attribute names of src/envs/goal_in_danger_zone.py:79-99
        _v_min, _v_max, _omega_min, _omega_max, _dt = -1.0, 1.0, -1.0, 1.0, 0.1
        _danger_zone = Zone()
        _goal = np.array([1.0, 2.0])

        def parallel_step(self, obs, action):
            raise AssertionError("never called on the host")

        def parallel_cost(self, obs, action, info):
            raise AssertionError("never called on the host")

    env = GoalInDangerZoneEnv()
    b = models.resolve(env.parallel_step, env.parallel_cost, 7, 2)
    assert b.model_id == _capi.MODEL_GOAL_IN_DANGER_ZONE and len(b.params()) == _capi.GOAL_ZONE_NUM_PARAMS
    env._goal = np.array([-3.0, 4.0])  # env.reset() draws a new goal
    assert b.params()[5:7] == [-3.0, 4.0]
