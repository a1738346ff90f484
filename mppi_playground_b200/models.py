"""Host-side model descriptors and callable -> device-model resolution.

The reference's operator API for the hot path is two Python callables,
``dynamics(state[K,ds], action[K,du])`` and ``cost_func(state, action, info)``
(src/pi_mpc/mppi.py:30-31). The engine cannot run Python inside a CUDA kernel,
so ``resolve`` maps the callables an example passes onto one of the built-in
``__device__`` models (csrc/mppi_models.cuh):

* bound methods of the reference's env / controller objects are recognised by
  the class of ``__self__`` (``RacingEnv`` + ``racing_controller``,
  ``Navigation2DEnv``); their parameters, occupancy grids and the per-solve
  reference path are read from those live objects, like the reference does;
* the closures of example/{pendulum,cartpole,mountaincar}.py are all called
  ``dynamics`` - they are fingerprinted by evaluating them once on a fixed
  16-row probe batch on the CPU and matching the result against the host
  formulas below (a few dozen flops at construction, not a solve path);
* the ``*Model`` descriptor classes of this module resolve to themselves, so
  the engine is usable without the reference installed;
* anything else raises ``NotImplementedError``: there is no CPU fallback.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence, Tuple

import torch

from . import _capi

MapSpec = Tuple[torch.Tensor, float, float, float]  # grid[W,H] fp32, cell, origin_x, origin_y


class Binding:
    """What the engine needs to know about one model instance."""

    model_id: int = -1
    name: str = ""
    dim_state: int = 0
    dim_control: int = 0

    def params(self, strict: bool = True) -> List[float]:
        """Model parameter block. ``strict=False`` (solver construction): attributes the caller only sets
        after building the solver (example/racing.py:24-46) are filled with placeholders; every solve re-reads
        them with ``strict=True``."""
        return []

    def maps(self) -> List[Optional[MapSpec]]:
        return []

    def reference_path(self) -> Optional[torch.Tensor]:
        return None


class _Descriptor(Binding):
    """Base of the stand-alone model descriptors. ``dynamics`` / ``cost_func``
    exist so that ``MPPI(dynamics=m.dynamics, cost_func=m.cost_func, ...)``
    reads like the reference's examples; the arithmetic itself runs in the
    CUDA kernel, never here."""

    def dynamics(self, state, action):  # pragma: no cover - marker only
        raise NotImplementedError(
            f"{type(self).__name__}.dynamics is evaluated inside the CUDA rollout kernel; there is no host path")

    def cost_func(self, state, action, info):  # pragma: no cover - marker only
        raise NotImplementedError(
            f"{type(self).__name__}.cost_func is evaluated inside the CUDA rollout kernel; there is no host path")


class PendulumModel(_Descriptor):
    """example/pendulum.py:17-47."""

    model_id, name, dim_state, dim_control = _capi.MODEL_PENDULUM, "pendulum", 2, 1


class CartpoleModel(_Descriptor):
    """example/cartpole.py:17-81."""

    model_id, name, dim_state, dim_control = _capi.MODEL_CARTPOLE, "cartpole", 4, 1


class MountainCarModel(_Descriptor):
    """example/mountaincar.py:17-55."""

    model_id, name, dim_state, dim_control = _capi.MODEL_MOUNTAINCAR, "mountaincar", 2, 1


class CartpoleContinuousModel(_Descriptor):
    """example/mujoco_cartpole.py:20-80 (pole mass 1.0, continuous force, |x| <= 1)."""

    model_id, name, dim_state, dim_control = _capi.MODEL_CARTPOLE_CONTINUOUS, "cartpole_continuous", 4, 1


class GoalInDangerZoneModel(_Descriptor):
    """src/envs/goal_in_danger_zone.py:113-156: unicycle, 7-dim observation (x, y, theta, vector to the goal,
    vector to the danger-zone centre); cost = distance to goal + 1000 inside the zone. ``goal`` may be
    reassigned between solves (the env draws a new one at every reset)."""

    model_id, name, dim_state, dim_control = _capi.MODEL_GOAL_IN_DANGER_ZONE, "goal_in_danger_zone", 7, 2

    def __init__(self, goal, center=(0.0, 0.0), radius: float = 10.0, v_lim=(-1.0, 1.0), w_lim=(-1.0, 1.0),
                 dt: float = 0.1, collision_cost: float = 1000.0):
        self.goal, self.center, self.radius = list(goal), list(center), float(radius)
        self.v_lim, self.w_lim, self.dt, self.collision_cost = tuple(v_lim), tuple(w_lim), dt, collision_cost
        self.u_min = torch.tensor([v_lim[0], w_lim[0]], dtype=torch.float32)
        self.u_max = torch.tensor([v_lim[1], w_lim[1]], dtype=torch.float32)

    def params(self, strict: bool = True):
        return [self.v_lim[0], self.v_lim[1], self.w_lim[0], self.w_lim[1], self.dt, float(self.goal[0]),
                float(self.goal[1]), float(self.center[0]), float(self.center[1]), self.radius, self.collision_cost]


class Navigation2DModel(_Descriptor):
    """src/envs/navigation_2d.py:218-279: unicycle, goal distance + occupancy cost."""

    model_id, name, dim_state, dim_control = _capi.MODEL_NAVIGATION2D, "navigation2d", 3, 2

    def __init__(self, obstacle_grid, cell_size: float, origin: Sequence[float], u_min=(0.0, -1.0),
                 u_max=(2.0, 1.0), goal=(9.0, 9.0), lim=(-10.0, 10.0, -10.0, 10.0), dt: float = 0.1,
                 obstacle_weight: float = 10000.0):
        self.obstacle_grid = torch.as_tensor(obstacle_grid, dtype=torch.float32).contiguous()
        self.cell_size, self.origin = float(cell_size), (float(origin[0]), float(origin[1]))
        self.u_min, self.u_max = torch.tensor(u_min, dtype=torch.float32), torch.tensor(u_max, dtype=torch.float32)
        self.goal, self.lim, self.dt, self.obstacle_weight = tuple(goal), tuple(lim), dt, obstacle_weight

    def params(self, strict: bool = True):
        lo, hi = self.u_min.tolist(), self.u_max.tolist()
        return [lo[0], hi[0], lo[1], hi[1], self.goal[0], self.goal[1], *self.lim, self.dt, self.obstacle_weight]

    def maps(self):
        return [(self.obstacle_grid, self.cell_size, *self.origin)]


class RacingModel(_Descriptor):
    """Kinematic bicycle (src/envs/racing_env.py:327-372) with the racing
    controller's cost (example/racing.py:110-159). Set ``reference_path``
    ([T+1,4]: x, y, yaw, v_target) before every solve (racing.py:73-81)."""

    model_id, name, dim_state, dim_control = _capi.MODEL_RACING, "racing", 4, 2

    def __init__(self, obstacle_grid, lane_grid, cell_size=(0.1, 0.1), origin=((400, 400), (400, 400)),
                 u_min=(-2.0, -0.25), u_max=(2.0, 0.25), wheelbase: float = 1.0, v_max: float = 8.0,
                 lim=(-40.0, 40.0, -40.0, 40.0), dt: float = 0.1, Qc=2.0, Ql=3.0, Qv=2.0, Qo=10000.0, Qin=0.01,
                 Qdin=0.5):
        self.obstacle_grid = torch.as_tensor(obstacle_grid, dtype=torch.float32).contiguous()
        self.lane_grid = torch.as_tensor(lane_grid, dtype=torch.float32).contiguous()
        self.cell_size, self.origin = tuple(cell_size), tuple(tuple(o) for o in origin)
        self.u_min, self.u_max = torch.tensor(u_min, dtype=torch.float32), torch.tensor(u_max, dtype=torch.float32)
        self.wheelbase, self.v_max, self.lim, self.dt = wheelbase, v_max, tuple(lim), dt
        self.Qc, self.Ql, self.Qv, self.Qo, self.Qin, self.Qdin = Qc, Ql, Qv, Qo, Qin, Qdin
        self.reference_path_tensor: Optional[torch.Tensor] = None

    def params(self, strict: bool = True):
        lo, hi = self.u_min.tolist(), self.u_max.tolist()
        return [lo[0], hi[0], lo[1], hi[1], self.wheelbase, self.v_max, *self.lim, self.dt, self.Qc, self.Ql,
                self.Qv, self.Qo, self.Qin, self.Qdin]

    def maps(self):
        return [(self.obstacle_grid, float(self.cell_size[0]), float(self.origin[0][0]), float(self.origin[0][1])),
                (self.lane_grid, float(self.cell_size[1]), float(self.origin[1][0]), float(self.origin[1][1]))]

    def reference_path(self):
        return self.reference_path_tensor


def racing_reference_path(state: torch.Tensor, path: torch.Tensor, cind: int, horizon: int, v_max: float = 8.0,
                          DL: float = 0.1, lookahead_distance: float = 3.0,
                          reference_path_interval: float = 0.85) -> Tuple[torch.Tensor, int]:
    """Look-ahead reference for the racing cost, the job of
    ``racing_controller.calc_ref_trajectory`` (example/racing.py:161-218):
    nearest centre-line point (never behind ``cind``), then one row every
    ``reference_path_interval`` metres starting ``lookahead_distance`` ahead.
    Vectorised on whatever device ``path`` lives on; returns ([T+1,4], index)."""
    d2 = (path[:, 0] - state[0].to(path.device)) ** 2 + (path[:, 1] - state[1].to(path.device)) ** 2
    ind = max(int(cind), int(torch.argmin(d2).item()))
    n = path.shape[0]
    # the reference accumulates `travel += interval` in fp64 and rounds half-to-even per row;
    # the running sum is kept (81 scalar adds) so that x.5 cases land on the same index
    dind = torch.tensor(reference_index_offsets(horizon, DL, lookahead_distance, reference_path_interval),
                        dtype=torch.long)
    idx = ind + dind
    beyond = idx >= n
    xref = torch.zeros(horizon + 1, 4, dtype=path.dtype, device=path.device)
    xref[:, :3] = path[idx.clamp(max=n - 1).to(path.device)]
    xref[:, 3] = 0.0 if bool(beyond.any()) else v_max
    return xref, ind


def reference_index_offsets(horizon: int, DL: float = 0.1, lookahead_distance: float = 3.0,
                            reference_path_interval: float = 0.85) -> List[int]:
    """int(round(travel / DL)) per reference row, with the reference's running fp64 sum (racing.py:205-208)."""
    travel, out = float(lookahead_distance), []
    for _ in range(horizon + 1):
        travel += reference_path_interval
        out.append(int(round(travel / DL)))
    return out


class RacingReferencePath:
    """Device-resident ``calc_ref_trajectory`` (example/racing.py:161-218): keeps the centre line and the
    carried path index on the GPU; ``update(state)`` returns the [T+1,4] reference path for a device state
    with one small kernel and no host synchronisation, ready to be handed to the solver."""

    def __init__(self, center_path: torch.Tensor, horizon: int, v_max: float = 8.0, DL: float = 0.1,
                 lookahead_distance: float = 3.0, reference_path_interval: float = 0.85, device=None):
        import ctypes as C

        self._lib = _capi.load()
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.horizon = horizon
        path = torch.as_tensor(center_path).detach().to("cpu", torch.float32).contiguous()
        assert path.ndim == 2 and path.shape[1] == 3
        offs = reference_index_offsets(horizon, DL, lookahead_distance, reference_path_interval)
        arr = (C.c_int32 * len(offs))(*offs)
        h = C.c_void_p()
        _capi.check(self._lib.mppi_refpath_create(self.device.index, path.data_ptr(), path.shape[0], arr, len(offs),
                                                  float(v_max), C.byref(h)))
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._lib.mppi_refpath_destroy(h)
            except Exception:
                pass

    def update(self, state: torch.Tensor) -> torch.Tensor:
        st = state.detach().to(self.device, torch.float32).contiguous()
        out = torch.empty(self.horizon + 1, 4, device=self.device, dtype=torch.float32)
        _capi.check(self._lib.mppi_refpath_update(self._h, st.data_ptr(), out.data_ptr(),
                                                  torch.cuda.current_stream(self.device).cuda_stream))
        return out

    @property
    def path_index(self) -> int:
        import ctypes as C

        v = C.c_int32()
        _capi.check(self._lib.mppi_refpath_index(self._h, -1, C.byref(v),
                                                 torch.cuda.current_stream(self.device).cuda_stream))
        return v.value

    @path_index.setter
    def path_index(self, value: int) -> None:
        _capi.check(self._lib.mppi_refpath_index(self._h, int(value), None,
                                                 torch.cuda.current_stream(self.device).cuda_stream))


# ---- bindings onto the reference's live objects -----------------------------------------


def _lim4(obstacle_map) -> List[float]:
    return [float(obstacle_map.x_lim[0]), float(obstacle_map.x_lim[1]), float(obstacle_map.y_lim[0]),
            float(obstacle_map.y_lim[1])]


def _map_spec(m) -> MapSpec:
    """(grid, cell, ox, oy) of a reference ObstacleMap / LaneMap
    (src/envs/obstacle_map_2d.py:76-89,164-166; lane_map_2d.py:55-61,84-88)."""
    if getattr(m, "_map_torch", None) is None:
        raise ValueError("cost map has no torch grid yet (ObstacleMap.convert_to_torch() not called)")
    return (m._map_torch, float(m._cell_size), float(m._cell_map_origin[0]), float(m._cell_map_origin[1]))


class _ReferenceRacing(Binding):
    model_id, name, dim_state, dim_control = _capi.MODEL_RACING, "racing", 4, 2

    def __init__(self, env, controller):
        self.env, self.ctl = env, controller
        # env constants live in (possibly CUDA) tensors: read them ONCE - a float() of a device tensor is a
        # device synchronisation, and these do not change after RacingEnv.__init__ (racing_env.py:37-42)
        self._env_params = [float(env.u_min[0]), float(env.u_max[0]), float(env.u_min[1]), float(env.u_max[1]),
                            float(env.L), float(env.V_MAX), *_lim4(env._obstacle_map),
                            0.1]  # delta_t default, racing_env.py:328

    _WEIGHTS = ("Qc", "Ql", "Qv", "Qo", "Qin", "Qdin")

    def params(self, strict: bool = True):
        # the cost weights are plain Python attributes of the controller, re-read every solve (racing.py:41-46).
        # example/racing.py builds the solver BEFORE it assigns them (racing.py:24-46): at construction
        # (strict=False) missing weights are placeholders; at solve time a missing weight is the same
        # AttributeError the reference's cost_function would raise.
        c = self.ctl
        if strict:
            return [*self._env_params, *(float(getattr(c, n)) for n in self._WEIGHTS)]
        return [*self._env_params, *(float(getattr(c, n, 0.0)) for n in self._WEIGHTS)]

    def maps(self):
        c = self.ctl  # set_cost_map runs after the solver is built (example/racing.py:106-108, 227)
        om, lm = getattr(c, "obstacle_map", None), getattr(c, "lane_map", None)
        if om is None or lm is None:  # example/racing.py:83-90 raises the same way
            raise ValueError("reference path, obstacle map, and lane map must be set before calling solve method.")
        return [_map_spec(om), _map_spec(lm)]

    def map_identity(self):
        return (id(getattr(self.ctl, "obstacle_map", None)), id(getattr(self.ctl, "lane_map", None)))

    def reference_path(self):
        return getattr(self.ctl, "reference_path", None)


class _ReferenceNavigation2D(Binding):
    model_id, name, dim_state, dim_control = _capi.MODEL_NAVIGATION2D, "navigation2d", 3, 2

    def __init__(self, env):
        self.env = env
        e = env  # constants of Navigation2DEnv.__init__ (navigation_2d.py:53-71), read once (device tensors)
        self._params = [float(e.u_min[0]), float(e.u_max[0]), float(e.u_min[1]), float(e.u_max[1]),
                        float(e._goal_pos[0]), float(e._goal_pos[1]), *_lim4(e._obstacle_map), 0.1,
                        10000.0]  # navigation_2d.py:219,277

    def params(self, strict: bool = True):
        return list(self._params)

    def maps(self):
        return [_map_spec(self.env._obstacle_map)]

    def map_identity(self):
        return (id(self.env._obstacle_map),)


class _ReferenceGoalInDangerZone(Binding):
    model_id, name, dim_state, dim_control = _capi.MODEL_GOAL_IN_DANGER_ZONE, "goal_in_danger_zone", 7, 2

    def __init__(self, env):
        self.env = env

    def params(self, strict: bool = True):
        e = self.env  # plain Python / numpy attributes; the goal changes at every env.reset()
        return [float(e._v_min), float(e._v_max), float(e._omega_min), float(e._omega_max), float(e._dt),
                float(e._goal[0]), float(e._goal[1]), float(e._danger_zone.center[0]),
                float(e._danger_zone.center[1]), float(e._danger_zone.radius), 1000.0]  # goal_in_danger_zone.py:154


# ---- behavioural fingerprints of the example closures -------------------------------------


def _wrap(x):
    return ((x + torch.pi) % (2 * torch.pi)) - torch.pi


class ModelBindingWarning(UserWarning):
    """Raised (as a warning) when Python callables were replaced by a built-in device model."""


def _probe_inputs(ds: int, du: int):
    """Probe batch for the behavioural fingerprint: 5 magnitudes x 16 rows, wide enough to drive every
    clamp / saturation of the built-in closures (pendulum: |u| > 2, |thdot| > 8, |theta| > pi; cartpole:
    |x| > 2.4, |theta| > 12 deg; mountaincar: both walls, |v| > 0.07, |u| > 1) as well as their interiors."""
    g = torch.Generator().manual_seed(1234)
    base = torch.tensor({2: [4.0, 10.0], 4: [3.0, 2.5, 0.3, 3.0]}.get(ds, [2.0] * ds)[:ds])
    states, actions = [], []
    for scale, a_scale in ((1.0, 6.0), (0.3, 1.5), (0.05, 0.3), (0.008, 6.0), (1.0, 1.0)):
        states.append((torch.rand(16, ds, generator=g) - 0.5) * 2.0 * scale * base)
        actions.append((torch.rand(16, du, generator=g) - 0.5) * 2.0 * a_scale)
    return torch.cat(states), torch.cat(actions)


def _pendulum_host(s, a):
    u = a[:, 0].clamp(-2, 2)
    thd = s[:, 1] + (-15.0 * torch.sin(s[:, 0] + torch.pi) + 3.0 * u) * 0.05
    return torch.stack([s[:, 0] + thd * 0.05, thd.clamp(-8, 8)], 1), _wrap(s[:, 0]) ** 2 + 0.1 * s[:, 1] ** 2


def _cartpole_host(s, a):
    x, xd, th, thd = s.unbind(1)
    force = torch.where(a[:, 0] >= 0, 10.0, -10.0)
    ct, st = torch.cos(th), torch.sin(th)
    temp = (force + 0.05 * thd**2 * st) / 1.1
    thacc = (9.8 * st - ct * temp) / (0.5 * (4.0 / 3.0 - 0.1 * ct**2 / 1.1))
    xacc = temp - 0.05 * thacc * ct / 1.1
    lim = 12 * 2 * math.pi / 360
    nxt = torch.stack([(x + 0.02 * xd).clamp(-2.4, 2.4), xd + 0.02 * xacc, (th + 0.02 * thd).clamp(-lim, lim),
                       thd + 0.02 * thacc], 1)
    return nxt, _wrap(th) ** 2 + 0.1 * thd**2 + 0.1 * x**2


def _cartpole_continuous_host(s, a):
    x, xd, th, thd = s.unbind(1)
    ct, st = torch.cos(th), torch.sin(th)
    temp = (a[:, 0] + 0.5 * thd**2 * st) / 2.0
    thacc = (9.8 * st - ct * temp) / (0.5 * (4.0 / 3.0 - ct**2 / 2.0))
    xacc = temp - 0.5 * thacc * ct / 2.0
    lim = 12 * 2 * math.pi / 360
    nxt = torch.stack([(x + 0.02 * xd).clamp(-1.0, 1.0), xd + 0.02 * xacc, (th + 0.02 * thd).clamp(-lim, lim),
                       thd + 0.02 * thacc], 1)
    return nxt, _wrap(th) ** 2 + 0.1 * thd**2 + 0.1 * x**2


def _mountaincar_host(s, a):
    p, v = s[:, 0], s[:, 1]
    v2 = (v + a[:, 0].clamp(-1, 1) * 0.0015 - 0.0025 * torch.cos(3 * p)).clamp(-0.07, 0.07)
    return torch.stack([(p + v2).clamp(-1.2, 0.6), v2], 1), (0.45 - p) ** 2


_CLOSURE_TWINS = [
    (PendulumModel, _pendulum_host),
    (CartpoleModel, _cartpole_host),
    (CartpoleContinuousModel, _cartpole_continuous_host),
    (MountainCarModel, _mountaincar_host),
]


def _fingerprint(dynamics: Callable, cost_func: Callable, ds: int, du: int) -> Optional[Binding]:
    """Match user closures against the host twins of the built-in models on the probe batch. The device
    costs depend on the state only, so the user's cost must return the same values for several ``t``, for
    the terminal-style call (zero action, stale ``t``, mppi.py:318-328) and for a different ``prev_action``:
    a closure with a time-dependent, terminal or action-dependent cost does not match and is rejected."""
    state, action = _probe_inputs(ds, du)
    zero = torch.zeros_like(action)

    def info(t, prev):
        return {"prev_action": prev.clone(), "t": t, "prev_state": state.clone(), "initial_state": state.clone()}

    try:
        with torch.no_grad():
            costs = [cost_func(state.clone(), action.clone(), info(0, action)),
                     cost_func(state.clone(), action.clone(), info(7, zero)),
                     cost_func(state.clone(), zero.clone(), info(48, -action))]
            nxt = dynamics(state.clone(), action.clone())  # clones: mountaincar's closure writes through its input
    except Exception:
        return None
    for cls, twin in _CLOSURE_TWINS:
        if (cls.dim_state, cls.dim_control) != (ds, du):
            continue
        want_next, want_cost = twin(state, action)
        if (tuple(nxt.shape) == tuple(want_next.shape) and torch.allclose(nxt, want_next, rtol=1e-4, atol=1e-5)
                and all(torch.allclose(c.reshape(-1), want_cost, rtol=1e-4, atol=1e-5) for c in costs)):
            import warnings

            warnings.warn(f"dynamics / cost_func match the built-in device model '{cls.name}' on an {len(state)}-row "
                          "behavioural probe; the solve runs that CUDA model - the Python callables are not "
                          "executed during rollouts", ModelBindingWarning, stacklevel=4)
            return cls()
    return None


def resolve(dynamics: Callable, cost_func: Callable, dim_state: int, dim_control: int) -> Binding:
    """Map the two callables of ``MPPI(...)`` onto a built-in device model."""
    d_self = getattr(dynamics, "__self__", None)
    c_self = getattr(cost_func, "__self__", None)
    binding: Optional[Binding] = None
    if isinstance(d_self, _Descriptor):
        if c_self is not d_self:
            raise ValueError("dynamics and cost_func must come from the same model descriptor")
        binding = d_self
    elif d_self is not None and type(d_self).__name__ == "RacingEnv":
        # The controller is recognised by its class / bound method, NOT by its attributes: example/racing.py
        # passes `self.cost_function` to MPPI(...) before it has assigned Qc..Qdin, reference_path or the maps
        # (racing.py:24-58), so none of them exist yet at this point.
        is_controller = (c_self is not None and c_self is not d_self and
                         (type(c_self).__name__ == "racing_controller" or
                          getattr(cost_func, "__name__", "") == "cost_function"))
        if not is_controller:
            raise NotImplementedError("RacingEnv.dynamics is only supported with racing_controller.cost_function")
        binding = _ReferenceRacing(d_self, c_self)
    elif d_self is not None and type(d_self).__name__ == "GoalInDangerZoneEnv":
        if c_self is not d_self or getattr(dynamics, "__name__", "") != "parallel_step":
            raise NotImplementedError("GoalInDangerZoneEnv is supported through parallel_step / parallel_cost")
        binding = _ReferenceGoalInDangerZone(d_self)
    elif d_self is not None and type(d_self).__name__ == "Navigation2DEnv":
        if c_self is not d_self:
            raise NotImplementedError("Navigation2DEnv.dynamics is only supported with Navigation2DEnv.cost_function")
        binding = _ReferenceNavigation2D(d_self)
    else:
        binding = _fingerprint(dynamics, cost_func, dim_state, dim_control)
    if binding is None:
        raise NotImplementedError(
            "dynamics/cost_func do not match a built-in device model (pendulum, cartpole, mountaincar, "
            "mujoco-cartpole, navigation2d, racing, goal-in-danger-zone). mppi_playground_b200 runs the rollout in a CUDA kernel and has no CPU "
            "fallback for arbitrary Python callables.")
    if (binding.dim_state, binding.dim_control) != (dim_state, dim_control):
        raise ValueError(f"model {binding.name} has dim_state={binding.dim_state}, dim_control={binding.dim_control};"
                         f" got {dim_state}/{dim_control}")
    return binding
