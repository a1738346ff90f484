"""TEST INFRASTRUCTURE ONLY - CPU oracle for the MPPI solve path.

A plain torch-on-CPU restatement of one ``pi_mpc.MPPI.forward`` solve of
kohonda/mppi_playground and of the five built-in env models. It exists to
*check* the CUDA engine (tests/, ``__graft_entry__.smoke()``) and to be timed
as the CPU baseline (``bench.py`` ``cpu_baseline`` / ``--impl reference``).
The product path never imports it: ``mppi_playground_b200`` fails loudly when
its CUDA library is missing instead of falling back to this file.

Parity pinning: the reference ships no golden vectors or known-answer tests
for this path (its tests/ never import pi_mpc), so this oracle is pinned
against outputs of the reference itself, recorded by ``oracle/gen_golden.py``
in the build container into ``tests/golden/*.npz`` and asserted (bit-exact on
the same torch build) by ``tests/test_oracle_golden.py``.

Every function cites the reference lines it restates (paths relative to the
reference root). The op order is kept identical on purpose: the reference is
fp32 ATen op by op with no fused multiply-add, and parity of the discontinuous
occupancy costs depends on that.

Third-party arithmetic on the path that is not reference source:
  * torch (2.11.0 here; reference lock 2.9.1): normal_(), softmax, conv1d,
    linalg.pinv, remainder - called directly, not restated.
  * scipy.optimize 1.18.1 (reference lock 1.15.3/1.17.0): minimize_scalar
    (method="bounded") and brentq drive the LBPS / ESSPS search in the
    reference (src/pi_mpc/mppi.py:344-348, 366-370). The oracle calls scipy
    itself; ``bounded_brent`` / ``brentq_restated`` below restate the two
    published algorithms (Brent 1973, ch. 5 "localmin" and ch. 4 "zero") in
    the form the device port follows, and are tested against scipy.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, Optional, Tuple

import numpy as np
import torch

# ----------------------------------------------------------------------------
# shared helpers
# ----------------------------------------------------------------------------


def wrap_angle(x: torch.Tensor) -> torch.Tensor:
    """((x + pi) % 2pi) - pi with torch's floored remainder.

    src/envs/racing_env.py:20-22 (identical copies in navigation_2d.py:18-20,
    example/pendulum.py:10-12, cartpole.py:10-12, mountaincar.py:10-12).
    """
    return ((x + torch.pi) % (2 * torch.pi)) - torch.pi


class GridMap:
    """Occupancy-grid lookup, src/envs/obstacle_map_2d.py:168-200 and the
    identical body in src/envs/lane_map_2d.py:90-122.

    ``grid`` is the [W, H] fp32 0/1 map (x is the slow axis), ``cell`` the
    cell size in metres, ``origin`` the (ox, oy) cell holding world (0, 0).
    """

    def __init__(self, grid: torch.Tensor, cell: float, origin) -> None:
        self.grid = grid.to(torch.float32).contiguous()
        self.cell = float(cell)
        self.origin = torch.tensor([float(origin[0]), float(origin[1])], dtype=torch.float32)

    def lookup(self, pos: torch.Tensor) -> torch.Tensor:
        """pos [..., 2] -> occupancy [...]; out-of-bounds reads as 1.0."""
        w, h = self.grid.shape
        idx = torch.round(pos / self.cell + self.origin).long()  # :179-180 (true division, half-even)
        ix, iy = idx[..., 0], idx[..., 1]
        oob = (ix < 0) | (ix >= w) | (iy < 0) | (iy >= h)  # :183-190
        occ = self.grid[ix.clamp(0, w - 1), iy.clamp(0, h - 1)]  # :191-195
        occ = occ.clone()
        occ[oob] = 1.0  # :198
        return occ

    def packed_bits(self) -> np.ndarray:
        """Row-padded bit-packing the CUDA engine uses (bit iy of row ix)."""
        return pack_grid_bits(self.grid.numpy())


def pack_grid_bits(grid: np.ndarray) -> np.ndarray:
    w, h = grid.shape
    words = (h + 31) // 32
    out = np.zeros((w, words), dtype=np.uint32)
    g = grid != 0
    for j in range(h):
        out[:, j // 32] |= g[:, j].astype(np.uint32) << np.uint32(j % 32)
    return out


# ----------------------------------------------------------------------------
# env models: dynamics(state[K,ds], action[K,du]) -> [K,ds]
#             cost(state[K,ds], action[K,du], info) -> [K]
# ----------------------------------------------------------------------------


class PendulumModel:
    """example/pendulum.py:17-47 (gymnasium Pendulum-v1 equations)."""

    name = "pendulum"
    dim_state, dim_control = 2, 1

    def dynamics(self, state, action):
        th = state[:, 0].view(-1, 1)
        thdot = state[:, 1].view(-1, 1)
        u = torch.clamp(action[:, 0].view(-1, 1), -2, 2)  # :26-27
        # :28-35  -3*g/(2*l) = -15.0, 3/(m*l^2) = 3.0, dt = 0.05
        newthdot = thdot + (-15.0 * torch.sin(th + torch.pi) + 3.0 * u) * 0.05
        newth = th + newthdot * 0.05  # :36 uses the unclamped rate
        newthdot = torch.clamp(newthdot, -8, 8)  # :37
        return torch.cat((newth, newthdot), dim=1)

    def cost(self, state, action, info):
        return wrap_angle(state[:, 0]) ** 2 + 0.1 * state[:, 1] ** 2  # :42-47


class CartpoleModel:
    """example/cartpole.py:17-81 (gymnasium CartPole-v1 equations, bang-bang force)."""

    name = "cartpole"
    dim_state, dim_control = 4, 1

    def dynamics(self, state, action):
        x = state[:, 0].view(-1, 1)
        x_dt = state[:, 1].view(-1, 1)
        theta = state[:, 2].view(-1, 1)
        theta_dt = state[:, 3].view(-1, 1)
        total_mass = 0.1 + 1.0  # :32
        polemass_length = 0.1 * 0.5  # :34
        a = action[:, 0].view(-1, 1)
        force = torch.zeros_like(a)  # :41-44
        force[a >= 0] = 10.0
        force[a < 0] = -10.0
        costheta = torch.cos(theta)
        sintheta = torch.sin(theta)
        temp = (force + polemass_length * theta_dt**2 * sintheta) / total_mass  # :49
        thetaacc = (9.8 * sintheta - costheta * temp) / (
            0.5 * (4.0 / 3.0 - 0.1 * costheta**2 / total_mass)
        )  # :50-52
        xacc = temp - polemass_length * thetaacc * costheta / total_mass  # :53
        newx = x + 0.02 * x_dt  # :55-58
        newx_dt = x_dt + 0.02 * xacc
        newtheta = theta + 0.02 * theta_dt
        newtheta_dt = theta_dt + 0.02 * thetaacc
        newx = torch.clamp(newx, -2.4, 2.4)  # :60-65
        lim = 12 * 2 * torch.pi / 360
        newtheta = torch.clamp(newtheta, -lim, lim)
        return torch.cat((newx, newx_dt, newtheta, newtheta_dt), dim=1)

    def cost(self, state, action, info):
        x, theta, theta_dt = state[:, 0], state[:, 2], state[:, 3]
        return wrap_angle(theta) ** 2 + 0.1 * theta_dt**2 + 0.1 * x**2  # :71-81


class MountainCarModel:
    """example/mountaincar.py:17-55.

    Quirk kept on purpose: ``velocity +=`` / ``position +=`` (:34, :36) write
    through the views into the caller's state row, so the solver's stored
    S[:, t] becomes (unclamped p', unclamped v') *before* the cost loop reads
    it. The oracle performs the same in-place writes.
    """

    name = "mountaincar"
    dim_state, dim_control = 2, 1

    def dynamics(self, state, action):
        position = state[:, 0].view(-1, 1)
        velocity = state[:, 1].view(-1, 1)
        force = torch.clamp(action[:, 0].view(-1, 1), -1.0, 1.0)
        velocity += force * 0.0015 - 0.0025 * torch.cos(3 * position)  # in place
        velocity = torch.clamp(velocity, -0.07, 0.07)
        position += velocity  # in place
        position = torch.clamp(position, -1.2, 0.6)
        return torch.cat((position, velocity), dim=1)

    def cost(self, state, action, info):
        return (0.45 - state[:, 0]) ** 2  # :45-55


class Navigation2DModel:
    """src/envs/navigation_2d.py:218-279 (unicycle + goal distance + occupancy)."""

    name = "navigation2d"
    dim_state, dim_control = 3, 2

    def __init__(self, grid: GridMap, u_min=(0.0, -1.0), u_max=(2.0, 1.0), goal=(9.0, 9.0),
                 lim=(-10.0, 10.0, -10.0, 10.0), dt: float = 0.1, obstacle_weight: float = 10000.0):
        self.grid = grid
        self.u_min = torch.tensor(u_min, dtype=torch.float32)
        self.u_max = torch.tensor(u_max, dtype=torch.float32)
        self.goal = torch.tensor(goal, dtype=torch.float32)
        self.lim = lim
        self.dt = dt
        self.obstacle_weight = obstacle_weight

    def dynamics(self, state, action):
        x = state[:, 0].view(-1, 1)
        y = state[:, 1].view(-1, 1)
        theta = state[:, 2].view(-1, 1)
        v = torch.clamp(action[:, 0].view(-1, 1), self.u_min[0], self.u_max[0])  # :235
        omega = torch.clamp(action[:, 1].view(-1, 1), self.u_min[1], self.u_max[1])  # :236
        theta = wrap_angle(theta)  # :237
        new_x = x + v * torch.cos(theta) * self.dt  # :239-241
        new_y = y + v * torch.sin(theta) * self.dt
        new_theta = wrap_angle(theta + omega * self.dt)
        xl = torch.tensor(self.lim[:2], dtype=torch.float32)  # :244-251
        yl = torch.tensor(self.lim[2:], dtype=torch.float32)
        return torch.cat([torch.clamp(new_x, xl[0], xl[1]), torch.clamp(new_y, yl[0], yl[1]), new_theta], dim=1)

    def cost(self, state, action, info):
        goal_cost = torch.norm(state[:, :2] - self.goal, dim=1)  # :269
        occ = self.grid.lookup(state[:, :2].unsqueeze(1)).squeeze(1)  # :271-275
        return goal_cost + self.obstacle_weight * occ  # :277


class RacingModel:
    """Kinematic bicycle src/envs/racing_env.py:327-372 + the controller cost
    example/racing.py:110-159. ``reference_path`` [T+1, 4] = (x, y, yaw,
    v_target) is replaced before every solve (example/racing.py:73-81)."""

    name = "racing"
    dim_state, dim_control = 4, 2

    def __init__(self, obstacle: GridMap, lane: GridMap, u_min=(-2.0, -0.25), u_max=(2.0, 0.25),
                 wheelbase: float = 1.0, v_max: float = 8.0, lim=(-40.0, 40.0, -40.0, 40.0), dt: float = 0.1,
                 Qc=2.0, Ql=3.0, Qv=2.0, Qo=10000.0, Qin=0.01, Qdin=0.5):
        self.obstacle, self.lane = obstacle, lane
        self.u_min = torch.tensor(u_min, dtype=torch.float32)
        self.u_max = torch.tensor(u_max, dtype=torch.float32)
        self.L = torch.tensor(wheelbase, dtype=torch.float32)
        self.V_MAX = torch.tensor(v_max, dtype=torch.float32)
        self.lim, self.dt = lim, dt
        self.Qc, self.Ql, self.Qv, self.Qo, self.Qin, self.Qdin = Qc, Ql, Qv, Qo, Qin, Qdin
        self.reference_path: Optional[torch.Tensor] = None

    def dynamics(self, state, action):
        x = state[:, 0].view(-1, 1)
        y = state[:, 1].view(-1, 1)
        theta = state[:, 2].view(-1, 1)
        v = state[:, 3].view(-1, 1)
        accel = torch.clamp(action[:, 0].view(-1, 1), self.u_min[0], self.u_max[0])  # :345
        steer = torch.clamp(action[:, 1].view(-1, 1), self.u_min[1], self.u_max[1])  # :346
        theta = wrap_angle(theta)  # :347
        dx = v * torch.cos(theta)  # :349-352
        dy = v * torch.sin(theta)
        dtheta = v * torch.tan(steer) / self.L
        new_x = x + dx * self.dt  # :354-357
        new_y = y + dy * self.dt
        new_theta = wrap_angle(theta + dtheta * self.dt)
        new_v = v + accel * self.dt
        xl = torch.tensor(self.lim[:2], dtype=torch.float32)  # :360-368
        yl = torch.tensor(self.lim[2:], dtype=torch.float32)
        return torch.cat(
            [torch.clamp(new_x, xl[0], xl[1]), torch.clamp(new_y, yl[0], yl[1]), new_theta,
             torch.clamp(new_v, -self.V_MAX, self.V_MAX)], dim=1)

    def cost(self, state, action, info):
        ref = self.reference_path
        t = info["t"]
        prev_action = info["prev_action"]
        sx = state[:, 0] - ref[t, 0]
        sy = state[:, 1] - ref[t, 1]
        ec = torch.sin(ref[t, 2]) * sx - torch.cos(ref[t, 2]) * sy  # racing.py:127-131
        el = -torch.cos(ref[t, 2]) * sx - torch.sin(ref[t, 2]) * sy  # :132-136
        path_cost = self.Qc * ec.pow(2) + self.Ql * el.pow(2)  # :138
        velocity_cost = self.Qv * (state[:, 3] - ref[t, 3]).pow(2)  # :141-143
        pos = state[:, :2].unsqueeze(1)
        occ = self.obstacle.lookup(pos).squeeze(1)  # :146-150
        occ += self.lane.lookup(pos).squeeze(1)
        occ = self.Qo * occ  # :151
        input_cost = self.Qin * action.pow(2).sum(dim=1)  # :154
        input_cost += self.Qdin * (action - prev_action).pow(2).sum(dim=1)  # :155
        return path_cost + velocity_cost + occ + input_cost  # :157


class CartpoleContinuousModel:
    """example/mujoco_cartpole.py:20-80: the cartpole equations with pole mass 1.0, the continuous
    action as the force (no bang-bang) and |x| <= 1."""

    name = "cartpole_continuous"
    dim_state, dim_control = 4, 1

    def dynamics(self, state, action):
        x = state[:, 0].view(-1, 1)
        x_dt = state[:, 1].view(-1, 1)
        theta = state[:, 2].view(-1, 1)
        theta_dt = state[:, 3].view(-1, 1)
        force = action[:, 0].view(-1, 1)  # :33
        total_mass = 1.0 + 1.0  # :38
        polemass_length = 1.0 * 0.5  # :40
        costheta = torch.cos(theta)
        sintheta = torch.sin(theta)
        temp = (force + polemass_length * theta_dt**2 * sintheta) / total_mass  # :46
        thetaacc = (9.8 * sintheta - costheta * temp) / (0.5 * (4.0 / 3.0 - 1.0 * costheta**2 / total_mass))  # :47-49
        xacc = temp - polemass_length * thetaacc * costheta / total_mass  # :50
        newx = x + 0.02 * x_dt  # :52-55
        newx_dt = x_dt + 0.02 * xacc
        newtheta = theta + 0.02 * theta_dt
        newtheta_dt = theta_dt + 0.02 * thetaacc
        newx = torch.clamp(newx, -1.0, 1.0)  # :57-62
        lim = 12 * 2 * torch.pi / 360
        newtheta = torch.clamp(newtheta, -lim, lim)
        return torch.cat((newx, newx_dt, newtheta, newtheta_dt), dim=1)

    def cost(self, state, action, info):
        x, theta, theta_dt = state[:, 0], state[:, 2], state[:, 3]
        return wrap_angle(theta) ** 2 + 0.1 * theta_dt**2 + 0.1 * x**2  # :68-80


class GoalInDangerZoneModel:
    """src/envs/goal_in_danger_zone.py:113-156 (parallel_step / parallel_cost): unicycle whose 7-dim
    observation carries the vectors to the goal and to the danger-zone centre."""

    name = "goal_in_danger_zone"
    dim_state, dim_control = 7, 2

    def __init__(self, goal=(3.0, -4.0), center=(0.0, 0.0), radius=10.0, v_lim=(-1.0, 1.0), w_lim=(-1.0, 1.0),
                 dt=0.1, collision_cost=1000.0):
        self.goal, self.center, self.radius = list(goal), list(center), radius
        self.v_lim, self.w_lim, self.dt, self.collision_cost = v_lim, w_lim, dt, collision_cost
        self.u_min = torch.tensor([v_lim[0], w_lim[0]])
        self.u_max = torch.tensor([v_lim[1], w_lim[1]])

    def dynamics(self, obs, action):
        x = obs[:, 0].view(-1, 1)
        y = obs[:, 1].view(-1, 1)
        theta = obs[:, 2].view(-1, 1)
        v = torch.clamp(action[:, 0].view(-1, 1), self.v_lim[0], self.v_lim[1])  # :118-119
        omega = torch.clamp(action[:, 1].view(-1, 1), self.w_lim[0], self.w_lim[1])
        theta = wrap_angle(theta + omega * self.dt)  # :124
        new_x = x + v * torch.cos(theta) * self.dt  # :126-127
        new_y = y + v * torch.sin(theta) * self.dt
        pos = torch.cat((new_x, new_y), dim=-1)
        vec_to_goal = torch.tensor(self.goal, dtype=obs.dtype) - pos  # :129-134
        vec_to_center = torch.tensor(self.center, dtype=obs.dtype) - pos
        return torch.cat((new_x, new_y, theta, vec_to_goal, vec_to_center), dim=-1)

    def cost(self, obs, action, info):
        cost = torch.norm(obs[:, 3:5], dim=-1)  # :145
        is_collided = torch.norm(obs[:, 5:7], dim=-1) < self.radius  # :153
        cost += is_collided.float() * self.collision_cost  # :154
        return cost


def racing_reference_path(state: torch.Tensor, path: torch.Tensor, cind: int, horizon: int, v_max: float = 8.0,
                          DL: float = 0.1, lookahead_distance: float = 3.0,
                          reference_path_interval: float = 0.85) -> Tuple[torch.Tensor, int]:
    """example/racing.py:161-218 restated with a vectorised nearest-point search
    (the reference loops over N points in Python; argmin of the same hypot)."""
    n = len(path)
    xref = torch.zeros((horizon + 1, 4), dtype=path.dtype)
    d = np.hypot(path[:, 0].numpy() - state[0].item(), path[:, 1].numpy() - state[1].item())
    ind = max(cind, int(np.argmin(d)))  # first minimum, like min(range, key=...)
    travel = lookahead_distance
    for i in range(horizon + 1):
        travel += reference_path_interval
        dind = int(round(travel / DL))
        if ind + dind < n:
            xref[i, :3] = path[ind + dind]
            xref[i, 3] = v_max
        else:
            xref[i, :3] = path[-1]
            xref[:, 3] = 0.0
    return xref, ind


# ----------------------------------------------------------------------------
# scalar searches (restated from scipy; tested against scipy in tests/)
# ----------------------------------------------------------------------------


def bounded_brent(func: Callable[[float], float], x1: float, x2: float, xatol: float = 1e-5,
                  maxiter: int = 500) -> Tuple[float, int]:
    """Brent's bounded scalar minimiser ("localmin"), as scipy.optimize
    minimize_scalar(method="bounded") runs it (defaults xatol=1e-5, maxiter=500).
    Returns (argmin, number of function evaluations)."""
    sqrt_eps = math.sqrt(2.2e-16)
    golden_mean = 0.5 * (3.0 - math.sqrt(5.0))
    a, b = x1, x2
    fulc = a + golden_mean * (b - a)
    nfc, xf = fulc, fulc
    rat = e = 0.0
    x = xf
    fx = func(x)
    num = 1
    ffulc = fnfc = fx
    xm = 0.5 * (a + b)
    tol1 = sqrt_eps * abs(xf) + xatol / 3.0
    tol2 = 2.0 * tol1
    while abs(xf - xm) > (tol2 - 0.5 * (b - a)):
        golden = True
        if abs(e) > tol1:  # try a parabolic step
            golden = False
            r = (xf - nfc) * (fx - ffulc)
            q = (xf - fulc) * (fx - fnfc)
            p = (xf - fulc) * q - (xf - nfc) * r
            q = 2.0 * (q - r)
            if q > 0.0:
                p = -p
            q = abs(q)
            r = e
            e = rat
            if abs(p) < abs(0.5 * q * r) and p > q * (a - xf) and p < q * (b - xf):
                rat = (p + 0.0) / q
                x = xf + rat
                if (x - a) < tol2 or (b - x) < tol2:
                    si = np.sign(xm - xf) + ((xm - xf) == 0)
                    rat = tol1 * si
            else:
                golden = True
        if golden:
            e = (a - xf) if xf >= xm else (b - xf)
            rat = golden_mean * e
        si = np.sign(rat) + (rat == 0)
        x = xf + si * max(abs(rat), tol1)
        fu = func(x)
        num += 1
        if fu <= fx:
            if x >= xf:
                a = xf
            else:
                b = xf
            fulc, ffulc = nfc, fnfc
            nfc, fnfc = xf, fx
            xf, fx = x, fu
        else:
            if x < xf:
                a = x
            else:
                b = x
            if fu <= fnfc or nfc == xf:
                fulc, ffulc = nfc, fnfc
                nfc, fnfc = x, fu
            elif fu <= ffulc or fulc == xf or fulc == nfc:
                fulc, ffulc = x, fu
        xm = 0.5 * (a + b)
        tol1 = sqrt_eps * abs(xf) + xatol / 3.0
        tol2 = 2.0 * tol1
        if num >= maxiter:
            break
    return float(xf), num


def brentq_restated(func: Callable[[float], float], xa: float, xb: float, xtol: float = 2e-12,
                    rtol: float = 8.881784197001252e-16, maxiter: int = 100) -> Tuple[float, int]:
    """Brent's root bracketing method ("zero") with scipy.optimize.brentq's
    defaults. Returns (root, function evaluations)."""
    xpre, xcur = xa, xb
    xblk = fblk = spre = scur = 0.0
    fpre, fcur = func(xpre), func(xcur)
    calls = 2
    if fpre == 0:
        return xpre, calls
    if fcur == 0:
        return xcur, calls
    if (fpre < 0) == (fcur < 0):
        raise ValueError("f(a) and f(b) must have different signs")
    for _ in range(maxiter):
        if fpre != 0 and fcur != 0 and ((fpre < 0) != (fcur < 0)):
            xblk, fblk = xpre, fpre
            spre = scur = xcur - xpre
        if abs(fblk) < abs(fcur):
            xpre, xcur, xblk = xcur, xblk, xcur
            fpre, fcur, fblk = fcur, fblk, fcur
        delta = (xtol + rtol * abs(xcur)) / 2
        sbis = (xblk - xcur) / 2
        if fcur == 0 or abs(sbis) < delta:
            return xcur, calls
        if abs(spre) > delta and abs(fcur) < abs(fpre):
            if xpre == xblk:
                stry = -fcur * (xcur - xpre) / (fcur - fpre)  # secant
            else:  # inverse quadratic
                dpre = (fpre - fcur) / (xpre - xcur)
                dblk = (fblk - fcur) / (xblk - xcur)
                stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre))
            if 2 * abs(stry) < min(abs(spre), 3 * abs(sbis) - delta):
                spre, scur = scur, stry
            else:
                spre = scur = sbis
        else:
            spre = scur = sbis
        xpre, fpre = xcur, fcur
        if abs(scur) > delta:
            xcur += scur
        else:
            xcur += delta if sbis > 0 else -delta
        fcur = func(xcur)
        calls += 1
    return xcur, calls


def savgol_coeffs(window: int, order: int) -> torch.Tensor:
    """src/pi_mpc/mppi.py:568-596: first row of pinv(vander(-h..h, order+1))."""
    if window % 2 == 0 or window <= order:
        raise ValueError("window_size must be odd and greater than poly_order.")
    h = (window - 1) // 2
    idx = torch.arange(-h, h + 1, dtype=torch.float32)
    return torch.linalg.pinv(torch.vander(idx, N=order + 1, increasing=True))[0]


def savgol_apply(y: torch.Tensor, coeffs: torch.Tensor) -> torch.Tensor:
    """src/pi_mpc/mppi.py:598-620: flip-pad both ends, 'valid' cross-correlation."""
    p = len(coeffs) // 2
    ypad = torch.cat([y[:p].flip(0), y, y[-p:].flip(0)])
    return torch.conv1d(ypad.view(1, 1, -1), coeffs.view(1, 1, -1), padding="valid").view(-1)


def mpo_gradient_device_form(costs: torch.Tensor, rho: float, epsilon: float = 0.1) -> float:
    """d loss / d rho of the MPO temperature loss, written the way the CUDA
    finalize step evaluates it so that it reproduces torch's fp32 autograd
    (src/pi_mpc/mppi.py:391-397) including its rounding of logsumexp:

        x_k = fl32(-c_k / tau),  e_k = exp(x_k - xmax),  S = sum e_k,  Sc = sum e_k c_k
        LSE32 = fl32(fl32(log S) + xmax)           (what torch.logsumexp returns)
        grad_tau = fl32(eps + LSE32) + exp(xmax - LSE32) * Sc / tau
        grad_rho = grad_tau * sigmoid(rho)

    exp(xmax - LSE32) is 1/S up to the rounding of LSE32; that factor is where
    the reference's ~1e-2 relative gradient noise comes from, so it is kept."""
    f32 = np.float32
    tau = torch.nn.functional.softplus(torch.tensor([rho], dtype=torch.float32))
    x = (-costs) / tau
    xmax = x.max()
    e = torch.exp(x - xmax)
    S = e.double().sum().item()
    Sc = (e.double() * costs.double()).sum().item()
    lse32 = f32(f32(math.log(S)) + f32(xmax.item()))
    term1 = float(f32(f32(epsilon) + lse32))
    term2 = math.exp(float(xmax.item()) - float(lse32)) * Sc / float(tau.item())
    return (term1 + term2) / (1.0 + math.exp(-rho))


def adam_scalar_step(rho: float, m: float, v: float, step: int, g: float, lr: float = 0.2, b1: float = 0.9,
                     b2: float = 0.999, eps: float = 1e-8) -> Tuple[float, float, float]:
    """torch.optim.Adam single-tensor update for one scalar (defaults as the
    reference uses, src/pi_mpc/mppi.py:200). Returns (rho, m, v); rho is kept fp32."""
    m = m + (g - m) * (1 - b1)
    v = b2 * v + (1 - b2) * g * g
    bc1 = 1 - b1**step
    bc2 = 1 - b2**step
    rho = float(np.float32(rho - (lr / bc1) * m / (math.sqrt(v) / math.sqrt(bc2) + eps)))
    return rho, m, v


# ----------------------------------------------------------------------------
# the solver
# ----------------------------------------------------------------------------


@dataclass
class SolveTrace:
    """Everything a parity test wants to look at after one solve."""

    noise: torch.Tensor  # [K,T,du]  sigma * eps, as used
    perturbed: torch.Tensor  # [K,T,du]  clamped samples
    costs: torch.Tensor  # [K]
    lam: float  # lambda the weights were computed with
    lam_next: float  # lambda carried to the next solve (MPO) / == lam otherwise
    weights: torch.Tensor  # [K]
    raw_action_seq: torch.Tensor  # [T,du] before the SG filter
    action_seq: torch.Tensor  # [T,du]
    state_seq: torch.Tensor  # [1,T+1,ds]
    func_evals: int = 0
    extra: Dict = field(default_factory=dict)


class OracleMPPI:
    """One-to-one restatement of ``MPPI`` (src/pi_mpc/mppi.py:16-620) on CPU.

    ``dynamics`` / ``cost_func`` follow the reference's callable contract, so
    the oracle also runs with the reference's own callables and vice versa.
    ``forward(state, noise=...)`` accepts an injected ``[K,T,du]`` noise tensor
    (sigma already applied, i.e. what ``rsample`` returns) for parity tests.
    """

    def __init__(self, horizon, num_samples, dim_state, dim_control, dynamics, cost_func, u_min, u_max, sigmas,
                 lambda_, lbps_delta=0.01, essps_target_ess=None, lambda_min=0.01, lambda_max=10.0,
                 exploration=0.0, use_sg_filter=False, sg_window_size=5, sg_poly_order=3, seed=42,
                 burn_constructor_draw=True, emulate_dead_work=False):
        torch.manual_seed(seed)  # :93
        self.T, self.K, self.ds, self.du = horizon, num_samples, dim_state, dim_control
        self.dynamics, self.cost_func = dynamics, cost_func
        self.u_min = torch.as_tensor(u_min, dtype=torch.float32).clone()
        self.u_max = torch.as_tensor(u_max, dtype=torch.float32).clone()
        self.sigmas = torch.as_tensor(sigmas, dtype=torch.float32).clone()
        assert self.u_min.shape == (dim_control,) and self.u_max.shape == (dim_control,)
        assert self.sigmas.shape == (dim_control,)
        self.exploration = exploration
        self.use_sg = use_sg_filter
        self.emulate_dead_work = emulate_dead_work
        if burn_constructor_draw:  # :146-148 advances the global RNG once
            self._draw_noise()
        self.prev_action_seq = torch.zeros(horizon, dim_control)  # :157
        self.coeffs = savgol_coeffs(sg_window_size, sg_poly_order)  # :160-162
        self.history = torch.zeros(horizon - 1, dim_control)  # :163-165
        self.weights = torch.zeros(num_samples)
        self.state_seq_batch = torch.zeros(num_samples, horizon + 1, dim_state)
        self.lbps_delta = lbps_delta
        self.essps_target = essps_target_ess if essps_target_ess is not None else num_samples / 10  # :185-187
        self.lambda_min, self.lambda_max = lambda_min, lambda_max
        if lambda_ == "MPO":  # :191-200
            self.mode, self.lam = "MPO", 1.0
            self.mpo_epsilon = 0.1
            self.rho = 0.0  # log_temperature = log(1.0)
            self._mpo_rho = torch.nn.Parameter(torch.log(torch.tensor([1.0])))
            self._mpo_opt = torch.optim.Adam([self._mpo_rho], lr=0.2)
        elif lambda_ in ("LBPS", "ESSPS"):
            self.mode, self.lam = lambda_, lambda_
        elif isinstance(lambda_, float):
            self.mode, self.lam = None, lambda_
        else:
            raise ValueError("lambda_ must be 'MPO', 'LBPS', 'ESSPS', or a float value.")  # :207-210

    # -- pieces ---------------------------------------------------------------
    def _draw_noise(self) -> torch.Tensor:
        """MultivariateNormal(0, diag sigma^2).rsample([K]) (:139-148, :261-263):
        torch draws eps = empty([K,T,du]).normal_() and applies scale_tril = diag(sigma)."""
        eps = torch.empty(self.K, self.T, self.du).normal_()
        return eps * self.sigmas

    def reset(self):  # :212-221
        self.prev_action_seq = torch.zeros(self.T, self.du)
        self.history = torch.zeros(self.T - 1, self.du)

    @staticmethod
    def ess(costs: torch.Tensor, lam: float) -> float:  # :526-532
        w = torch.softmax(-costs / lam, dim=0)
        return 1.0 / torch.sum(w**2).item()

    def lbps_objective(self, lam: float, costs: torch.Tensor) -> float:  # :534-557
        w = torch.softmax(-costs / lam, dim=0)
        ess = 1.0 / torch.sum(w**2).item()
        expected_return = -torch.sum(w * costs).item()
        cost_range = (costs.max() - costs.min()).item()
        penalty = cost_range * math.sqrt((1 - self.lbps_delta) / self.lbps_delta) / math.sqrt(ess)
        return -(expected_return - penalty)

    def rollout(self, state: torch.Tensor, action_seqs: torch.Tensor) -> torch.Tensor:  # :508-524
        out = torch.zeros(action_seqs.shape[0], self.T + 1, self.ds)
        out[:, 0, :] = state
        for t in range(self.T):
            out[:, t + 1, :] = self.dynamics(out[:, t, :], action_seqs[:, t, :])
        return out

    def mpo_step(self, costs: torch.Tensor) -> float:
        """One Adam(lr=0.2) step on rho for loss = tau*(eps + logsumexp(-c/tau)),
        tau = softplus(rho); then lambda = exp(rho) (:387-398).

        The reference's fp32 autograd gradient is dominated by a cancellation
        (eps + LSE ~ -1e3 against E_w[c]/tau ~ +1e3), so the oracle calls torch's
        autograd and torch.optim.Adam (third-party, not reference source) to stay
        bit-exact; ``mpo_gradient_device_form`` restates the same arithmetic in
        the closed form the CUDA finalize step evaluates."""
        self._mpo_opt.zero_grad()
        tau = torch.nn.functional.softplus(self._mpo_rho)
        loss = tau * (self.mpo_epsilon + torch.mean(torch.logsumexp(-costs / tau, dim=0)))
        loss.backward()
        self._mpo_opt.step()
        self.rho = self._mpo_rho.item()
        return torch.exp(self._mpo_rho).item()

    # -- one solve --------------------------------------------------------------
    def forward(self, state, noise: Optional[torch.Tensor] = None, info: Optional[Dict] = None) -> SolveTrace:
        info = {} if info is None else info
        state = torch.as_tensor(np.asarray(state) if not torch.is_tensor(state) else state).to(torch.float32)
        assert state.shape == (self.ds,)  # :247
        K, T = self.K, self.T
        mean = self.prev_action_seq.clone()  # :255 (no time shift of the warm start)
        noise = self._draw_noise() if noise is None else noise.to(torch.float32)  # :261-263
        thr = int(K * (1 - self.exploration))  # :266
        perturbed = torch.clamp(torch.cat([mean + noise[:thr], noise[thr:]]), self.u_min, self.u_max)  # :267-275

        S = self.state_seq_batch
        S[:, 0, :] = state.repeat(K, 1)  # :280
        for t in range(T):  # :282-286
            S[:, t + 1, :] = self.dynamics(S[:, t, :], perturbed[:, t, :])

        stage = torch.zeros(K, T)  # :291-316
        if self.emulate_dead_work:
            dead = torch.zeros(K, T)
            inv_cov = torch.zeros(T, self.du, self.du)
            inv_cov[1:] = torch.diag(1.0 / self.sigmas**2)
        for t in range(T):
            p = t - 1 if t > 0 else 0
            info["prev_state"] = S[:, p, :]
            info["prev_action"] = perturbed[:, p, :]
            info["initial_state"] = S[:, 0, :]
            info["t"] = t
            stage[:, t] = self.cost_func(S[:, t, :], perturbed[:, t, :], info)
            if self.emulate_dead_work:  # :312-316, computed then discarded by the reference
                dead[:, t] = mean[t] @ inv_cov[t] @ perturbed[:, t].T
        info["prev_state"] = S[:, -2, :]  # :318-319; info["t"], info["prev_action"] stay stale
        terminal = self.cost_func(S[:, -1, :], torch.zeros(K, self.du), info)  # :320-328
        costs = torch.sum(stage, dim=1) + terminal  # :333-336

        evals = 0
        if self.mode == "LBPS":  # :341-349
            from scipy.optimize import minimize_scalar

            res = minimize_scalar(lambda l: self.lbps_objective(l, costs), bounds=(self.lambda_min, self.lambda_max),
                                  method="bounded")
            self.lam, evals = res.x, res.nfev
        elif self.mode == "ESSPS":  # :351-370
            from scipy.optimize import brentq

            e_lo, e_hi = self.ess(costs, self.lambda_min), self.ess(costs, self.lambda_max)
            if self.essps_target <= e_lo:
                self.lam = self.lambda_min
            elif self.essps_target >= e_hi:
                self.lam = self.lambda_max
            else:
                self.lam, r = brentq(lambda l: self.ess(costs, l) - self.essps_target, self.lambda_min,
                                     self.lambda_max, full_output=True)
                evals = r.function_calls
        lam_used = float(self.lam)

        self.weights = torch.softmax(-costs / self.lam, dim=0)  # :376
        raw = torch.sum(self.weights.view(K, 1, 1) * perturbed, dim=0)  # :381-384
        if self.mode == "MPO":
            self.lam = self.mpo_step(costs)  # :387-398

        opt = raw
        if self.use_sg:  # :423-443
            y = torch.cat([self.history, raw], dim=0)
            filt = torch.zeros_like(y)
            for d in range(self.du):
                filt[:, d] = savgol_apply(y[:, d], self.coeffs)
            opt = filt[-T:]
        state_seq = self.rollout(state, opt.repeat(1, 1, 1))  # :448-449
        self.prev_action_seq = opt  # :452
        self.history = torch.cat([self.history[1:], opt[0].view(1, -1)])  # :455-458
        return SolveTrace(noise=noise, perturbed=perturbed, costs=costs, lam=lam_used, lam_next=float(self.lam),
                          weights=self.weights, raw_action_seq=raw, action_seq=opt, state_seq=state_seq,
                          func_evals=evals)

    def get_top_samples(self, n: int):  # :462-487
        assert n <= self.K
        idx = torch.topk(self.weights, n).indices
        w = self.weights[idx]
        order = torch.argsort(w, descending=True)
        return self.state_seq_batch[idx][order], w[order]


# ----------------------------------------------------------------------------
# control-step epilogue (SURVEY 8f "next" row 2): what the reference's loops run between two solves
# ----------------------------------------------------------------------------


def env_step(model, state: torch.Tensor, action: torch.Tensor, goal, goal_threshold: float):
    """``RacingEnv.step`` src/envs/racing_env.py:142-163 (== ``Navigation2DEnv.step`` navigation_2d.py:97-117):
    clamp the executed action to the env bounds, one batch-1 dynamics step, goal test. ``model`` is one of the
    oracle models above (its ``dynamics`` is the env's); returns ``(next_state [ds], is_goal_reached bool)``."""
    u = action
    if hasattr(model, "u_min"):
        u = torch.clamp(action, model.u_min, model.u_max)  # :151 / :105
    nxt = model.dynamics(state.view(1, -1), u.view(1, -1)).squeeze(0)  # :153-155
    goal_t = torch.as_tensor(goal, dtype=torch.float32)
    reached = bool(torch.norm(nxt[:2] - goal_t) < goal_threshold)  # :158-161
    return nxt, reached


def collision_check(obstacle: GridMap, state_seq: torch.Tensor) -> torch.Tensor:
    """``RacingEnv.collision_check`` src/envs/racing_env.py:374-384 (== navigation_2d.py:281-291): obstacle-map
    value of every predicted position. ``state_seq`` [B, L, ds] -> [B, L] (the reference's ``squeeze(1)`` is a
    no-op for L > 1)."""
    return obstacle.lookup(state_seq[:, :, :2])


# ----------------------------------------------------------------------------
# map construction (SURVEY 8f "next" row 4), numpy restatements of the reference's painters
# ----------------------------------------------------------------------------


def paint_obstacle_map(width: int, height: int, discs, rects) -> np.ndarray:
    """``ObstacleMap.add_circle_obstacle`` / ``add_rectangle_obstacle`` src/envs/obstacle_map_2d.py:103-160 on a
    zero [width, height] grid. ``discs`` = (centre cell x, y, radius_occ**2) per circle, ``rects`` = the clipped
    (x_init, x_end, y_init, y_end) per rectangle - the host conversions of mppi_playground_b200/maps.py (the
    reference's fp64 expressions, checked against the live reference in tests/test_reference_live.py)."""
    grid = np.zeros((width, height))
    for cx, cy, r2 in discs:
        r = int(math.isqrt(int(r2)))
        i = np.arange(-r, r + 1)
        ii, jj = np.meshgrid(i, i, indexing="ij")
        hit = ii**2 + jj**2 <= r2  # :120
        xs = np.clip(cx + ii[hit], 0, width - 1)  # :121-122 (indices are clipped, not dropped)
        ys = np.clip(cy + jj[hit], 0, height - 1)
        grid[xs, ys] = 1  # :123
    for x0, x1, y0, y1 in rects:
        grid[x0:x1, y0:y1] = 1  # :156
    return grid


def paint_lane_map(width: int, height: int, cells, r2: int) -> np.ndarray:
    """``LaneMap.populate_map`` src/envs/lane_map_2d.py:68-88: ones, the in-map centre-line ``cells`` zeroed, then
    ``distance_transform_edt(map) <= max_distance`` -> 0 else 1, with the threshold expressed as the largest
    accepted integer squared cell distance ``r2`` (maps.LaneRaster). Brute-force nearest-cell distance in exact
    integer arithmetic (the EDT's value is sqrt of exactly this integer)."""
    cells = np.asarray(cells, dtype=np.int64).reshape(-1, 2)
    best = np.full((width, height), np.iinfo(np.int64).max, dtype=np.int64)
    xs = np.arange(width, dtype=np.int64)[:, None]
    ys = np.arange(height, dtype=np.int64)[None, :]
    r = int(math.isqrt(max(int(r2), 0)))
    for cx, cy in cells:  # only the (2r+1)^2 window around a centre cell can pass the threshold
        x0, x1 = max(cx - r, 0), min(cx + r + 1, width)
        y0, y1 = max(cy - r, 0), min(cy + r + 1, height)
        d2 = (xs[x0:x1] - cx) ** 2 + (ys[:, y0:y1] - cy) ** 2
        np.minimum(best[x0:x1, y0:y1], d2, out=best[x0:x1, y0:y1])
    return np.where(best <= r2, 0, 1).astype(np.float64)
