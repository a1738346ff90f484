"""Restated scalar searches / MPO closed form (what the device port follows) vs the
third-party originals the reference calls (scipy.optimize, torch autograd + Adam)."""
import math

import numpy as np
import pytest
import torch
from scipy.optimize import brentq, minimize_scalar

from oracle import fixtures as fx
from oracle import mppi_oracle as mo


def _costs(seed, n=4096, scale=30.0):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n, generator=g) * scale + 5.0


@pytest.mark.parametrize("seed,delta", [(0, 0.01), (1, 0.2), (2, 0.5), (3, 0.9)])
def test_bounded_brent_matches_scipy(seed, delta):
    costs = _costs(seed)
    solver = fx.build_oracle(fx.load_case("navigation2d_lbps"))[1]
    solver.lbps_delta = delta
    f = lambda lam: solver.lbps_objective(lam, costs)  # noqa: E731
    res = minimize_scalar(f, bounds=(0.01, 10.0), method="bounded")
    x, n = mo.bounded_brent(f, 0.01, 10.0)
    assert n == res.nfev
    assert x == pytest.approx(res.x, rel=0, abs=1e-12)


@pytest.mark.parametrize("seed,target", [(0, 400.0), (1, 1000.0), (2, 50.0)])
def test_brentq_restated_matches_scipy(seed, target):
    costs = _costs(seed)
    f = lambda lam: mo.OracleMPPI.ess(costs, lam) - target  # noqa: E731
    root, r = brentq(f, 0.01, 10.0, full_output=True)
    x, n = mo.brentq_restated(f, 0.01, 10.0)
    assert n == r.function_calls
    assert x == pytest.approx(root, rel=0, abs=1e-12)


@pytest.mark.parametrize("name", ["cartpole_mpo", "navigation2d_mpo_expl"])
def test_mpo_device_form_tracks_autograd(name):
    """Closed-form gradient + scalar Adam follow torch autograd / torch.optim.Adam
    (and therefore the reference's lambda trajectory) to ~1e-4 relative."""
    case = fx.load_case(name)
    rho, m, v = 0.0, 0.0, 0.0
    for s in range(case.n_solves):
        costs = torch.from_numpy(case.costs[s])
        g = mo.mpo_gradient_device_form(costs, rho)
        p = torch.zeros(1, requires_grad=True)
        with torch.no_grad():
            p.fill_(rho)
        tau = torch.nn.functional.softplus(p)
        (tau * (0.1 + torch.logsumexp(-costs / tau, dim=0))).backward()
        assert g == pytest.approx(p.grad.item(), rel=2e-3)
        rho, m, v = mo.adam_scalar_step(rho, m, v, s + 1, g)
        assert math.exp(rho) == pytest.approx(case.lam_next[s], rel=2e-4)


def test_savgol_coeffs_known_values():
    c = mo.savgol_coeffs(5, 3).numpy()
    np.testing.assert_allclose(c, [-3 / 35, 12 / 35, 17 / 35, 12 / 35, -3 / 35], atol=1e-6)
    with pytest.raises(ValueError):
        mo.savgol_coeffs(4, 3)


def test_lbps_can_land_inside_the_bracket():
    """Guards against only ever testing the lambda_max corner (the nav2d golden case sits there)."""
    costs = _costs(5, scale=3.0)
    solver = fx.build_oracle(fx.load_case("navigation2d_lbps"))[1]
    solver.lbps_delta = 0.9
    x, _ = mo.bounded_brent(lambda lam: solver.lbps_objective(lam, costs), 0.01, 10.0)
    assert 0.02 < x < 9.9
