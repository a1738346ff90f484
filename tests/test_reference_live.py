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
