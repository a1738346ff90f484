"""The drop-in path on the GPU: ``from pi_mpc.mppi import MPPI`` (compat shim) handed objects shaped exactly like
the reference's (tests/reference_shapes.py: classes named RacingEnv / racing_controller / Navigation2DEnv /
GoalInDangerZoneEnv, CUDA tensors where the reference keeps CUDA tensors), driven in the order
example/racing.py:221-237, example/navigation2d.py:12-44, example/goal_in_danger_zone.py:30-63 and
example/pendulum.py:58-76 drive them, and checked against the CPU oracle on the engine's own noise."""
import os
import sys

import numpy as np
import pytest
import torch

import reference_shapes as rs
from engine_util import ParityStats, assert_parity, build_oracle, tol_for

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def MPPI():
    """``pi_mpc.mppi.MPPI`` resolved the way a user of the reference would get the engine: the compat directory
    first on sys.path."""
    compat = os.path.join(ROOT, "mppi_playground_b200", "compat")
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "pi_mpc" or k.startswith("pi_mpc.")}
    sys.path.insert(0, compat)
    try:
        from pi_mpc.mppi import MPPI as cls  # noqa: WPS433
        import pi_mpc

        assert pi_mpc.MPPI is cls and pi_mpc.__file__.startswith(compat)
        yield cls
    finally:
        sys.path.remove(compat)
        for k in [k for k in sys.modules if k == "pi_mpc" or k.startswith("pi_mpc.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def _parity(solver, oracle, action, states, noise, state, lam_mode):
    tr = oracle.forward(torch.as_tensor(np.asarray(state.cpu() if torch.is_tensor(state) else state),
                                        dtype=torch.float32), noise=noise)
    used, _ = solver._lambdas()
    st = ParityStats(solver._costs.cpu().numpy(), tr.costs.numpy(), action.cpu().numpy(), tr.action_seq.numpy(),
                     states.cpu().numpy(), tr.state_seq.numpy(), used, tr.lam)
    assert_parity(st, tol=tol_for(lam_mode))
    oracle.prev_action_seq = action.cpu().clone()
    return tr


def test_racing_example_objects_through_the_pi_mpc_shim(MPPI):
    """example/racing.py:221-237 - controller built before its weights / maps exist, ``set_cost_map`` afterwards,
    ``update`` (reads ``solver._horizon``), ``get_top_samples(300)``; CUDA state, CUDA grids, CUDA path."""
    env = rs.RacingEnv()  # device="cuda" like the reference's default
    assert env.u_min.is_cuda and env._obstacle_map._map_torch.is_cuda and env.racing_center_path.is_cuda
    controller = rs.racing_controller(env, MPPI)
    assert controller.obstacle_map is None and controller.solver._horizon == 25
    with pytest.raises(ValueError, match="must be set"):  # racing.py:83-90: maps not set yet
        controller.update(env._robot_state, env.racing_center_path)
    controller.current_path_index = 0
    controller.set_cost_map(env._obstacle_map, env._lane_map)
    solver = controller.solver
    cfg = dict(model="racing", horizon=25, num_samples=4000, sigmas=[0.5, 0.1], lambda_=1.0)
    omodel, oracle = build_oracle(cfg, burn_constructor_draw=False)
    state = env._robot_state
    for step in range(3):
        noise = solver.sampler_noise().cpu()
        action_seq, state_seq = controller.update(state, env.racing_center_path)
        assert action_seq.is_cuda and tuple(action_seq.shape) == (25, 2) and tuple(state_seq.shape) == (1, 26, 4)
        omodel.reference_path = controller.reference_path.cpu()
        _parity(solver, oracle, action_seq, state_seq, noise, state, 1.0)
        top_samples, top_weights = controller.get_top_samples(num_samples=300)
        assert tuple(top_samples.shape) == (300, 26, 4) and tuple(top_weights.shape) == (300,)
        w = top_weights.cpu().numpy()
        assert np.all(np.diff(w) <= 0) and w[0] > 0
        otraj, ow = oracle.get_top_samples(300)
        np.testing.assert_allclose(w, ow.numpy(), rtol=5e-3, atol=1e-7)
        if abs(float(ow[0] - ow[1])) > 1e-3 * float(ow[0]):
            np.testing.assert_allclose(top_samples[0].cpu().numpy(), otraj[0].numpy(), rtol=0, atol=5e-4)
        state = state_seq[0, 1]  # stays a CUDA tensor, like env.step's return (racing_env.py:142-163)
    # cost weights are live attributes of the controller (racing.py:41-46): a change reaches the next solve
    controller.Qv, omodel.Qv = 9.0, 9.0
    noise = solver.sampler_noise().cpu()
    a, s = controller.update(state, env.racing_center_path)
    omodel.reference_path = controller.reference_path.cpu()
    _parity(solver, oracle, a, s, noise, state, 1.0)
    # replacing the maps after construction (set_cost_map with new objects) re-uploads the grids
    e = env._fixture
    empty = rs.ObstacleMap(np.zeros_like(e.obstacle), 0.1, "cuda")
    empty.convert_to_torch()
    controller.set_cost_map(empty, env._lane_map)
    omodel.obstacle.grid = torch.zeros_like(omodel.obstacle.grid)
    noise = solver.sampler_noise().cpu()
    a, s = controller.update(state, env.racing_center_path)
    omodel.reference_path = controller.reference_path.cpu()
    _parity(solver, oracle, a, s, noise, state, 1.0)


def test_navigation2d_example_objects_through_the_pi_mpc_shim(MPPI):
    """example/navigation2d.py:12-44: bound methods of Navigation2DEnv, lambda_="ESSPS", CUDA everything."""
    env = rs.Navigation2DEnv()
    solver = MPPI(horizon=30, num_samples=3000, dim_state=3, dim_control=2, dynamics=env.dynamics,
                  cost_func=env.cost_function, u_min=env.u_min, u_max=env.u_max, sigmas=torch.tensor([0.5, 0.5]),
                  lambda_="ESSPS")
    cfg = dict(model="navigation2d", horizon=30, num_samples=3000, sigmas=[0.5, 0.5], lambda_="ESSPS")
    _, oracle = build_oracle(cfg, burn_constructor_draw=False)
    state = env._robot_state
    for _ in range(3):
        noise = solver.sampler_noise().cpu()
        action_seq, state_seq = solver.forward(state=state)
        _parity(solver, oracle, action_seq, state_seq, noise, state, "ESSPS")
        top_samples, top_weights = solver.get_top_samples(num_samples=300)
        assert tuple(top_samples.shape) == (300, 31, 3) and bool((top_weights[:-1] >= top_weights[1:]).all())
        state = state_seq[0, 1]
    assert solver._lambda > 0.0  # README: controller._lambda is read after solves


def test_goal_in_danger_zone_example_objects_through_the_pi_mpc_shim(MPPI):
    """example/goal_in_danger_zone.py:30-63: parallel_step / parallel_cost, CPU float32 observation tensor,
    get_top_samples(100), a new goal between solves (env.reset draws one)."""
    env = rs.GoalInDangerZoneEnv(goal=(-2.5, 6.0))
    solver = MPPI(horizon=30, num_samples=3000, dim_state=7, dim_control=2, dynamics=env.parallel_step,
                  cost_func=env.parallel_cost, u_min=torch.tensor([-1.0, -1.0]), u_max=torch.tensor([1.0, 1.0]),
                  sigmas=torch.tensor([0.5, 0.5]), lambda_=1.0)
    cfg = dict(model="goal_in_danger_zone", horizon=30, num_samples=3000, u_min=[-1.0, -1.0], u_max=[1.0, 1.0],
               sigmas=[0.5, 0.5], lambda_=1.0, goal=[-2.5, 6.0], center=[0.0, 0.0], radius=10.0)
    omodel, oracle = build_oracle(cfg, burn_constructor_draw=False)
    obs = np.array([-14.0, 3.0, 0.4, -2.5 + 14.0, 6.0 - 3.0, 14.0, -3.0])
    for step in range(3):
        if step == 2:  # the goal moves: the binding re-reads it every solve
            env._goal = np.array([4.0, -1.0])
            omodel.goal = [4.0, -1.0]
            obs[3], obs[4] = 4.0 - obs[0], -1.0 - obs[1]
        state = torch.tensor(obs, dtype=torch.float32)  # goal_in_danger_zone.py:48
        noise = solver.sampler_noise().cpu()
        action_seq, predicted = solver.forward(state=state)
        _parity(solver, oracle, action_seq, predicted, noise, state, 1.0)
        top_samples, top_weights = solver.get_top_samples(num_samples=100)
        assert tuple(top_samples.shape) == (100, 31, 7)
        obs = predicted[0, 1].cpu().numpy().astype(np.float64)


def test_pendulum_example_closures_through_the_pi_mpc_shim(MPPI):
    """example/pendulum.py:17-76: module-level style closures (fingerprinted), float64 numpy state, __call__."""
    from mppi_playground_b200.models import ModelBindingWarning

    def angle_normalize(x):
        return ((x + torch.pi) % (2 * torch.pi)) - torch.pi

    def dynamics(state, action):
        th, thdot = state[:, 0].view(-1, 1), state[:, 1].view(-1, 1)
        u = torch.clamp(action[:, 0].view(-1, 1), -2, 2)
        newthdot = thdot + (-15.0 * torch.sin(th + torch.pi) + 3.0 * u) * 0.05
        return torch.cat((th + newthdot * 0.05, torch.clamp(newthdot, -8, 8)), dim=1)

    def cost_function(state, action, info):
        return angle_normalize(state[:, 0]) ** 2 + 0.1 * state[:, 1] ** 2

    with pytest.warns(ModelBindingWarning, match="pendulum"):
        solver = MPPI(horizon=15, num_samples=1000, dim_state=2, dim_control=1, dynamics=dynamics,
                      cost_func=cost_function, u_min=torch.tensor([-2.0]), u_max=torch.tensor([2.0]),
                      sigmas=torch.tensor([1.0]), lambda_="ESSPS")
    cfg = dict(model="pendulum", horizon=15, num_samples=1000, u_min=[-2.0], u_max=[2.0], sigmas=[1.0],
               lambda_="ESSPS")
    _, oracle = build_oracle(cfg, burn_constructor_draw=False)
    state = np.array([3.14, 0.0])  # float64 numpy, like env.unwrapped.state.copy() (pendulum.py:73)
    for _ in range(3):
        noise = solver.sampler_noise().cpu()
        action_seq, state_seq = solver(state)  # README.md:191 calls the module
        _parity(solver, oracle, action_seq, state_seq, noise, state, "ESSPS")
        state = state_seq[0, 1].cpu().numpy().astype(np.float64)
