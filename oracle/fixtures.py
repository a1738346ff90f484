"""TEST INFRASTRUCTURE ONLY - loads tests/golden fixtures and builds oracle models.

Used by tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline leg.
The env fixtures (occupancy grids, racing centre line) were produced by
``oracle/gen_golden.py`` from the live reference; nothing here reads
/root/reference.
"""
from __future__ import annotations

import json
import os
from types import SimpleNamespace

import numpy as np
import torch

from . import mppi_oracle as mo

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _unpack(bits: np.ndarray, shape) -> np.ndarray:
    return np.unpackbits(bits, axis=1)[:, : int(shape[1])].astype(np.float32)


def load_env_racing() -> SimpleNamespace:
    z = np.load(os.path.join(GOLDEN_DIR, "env_racing.npz"))
    shape = z["shape"]
    return SimpleNamespace(
        obstacle=_unpack(z["obstacle_bits"], shape), lane=_unpack(z["lane_bits"], shape),
        cell=[float(c) for c in z["cell"]], origin=[[int(v) for v in o] for o in z["origin"]],
        lim=[float(v) for v in z["lim"]], center_path=torch.from_numpy(z["center_path"].copy()),
        start_state=torch.from_numpy(z["start_state"].copy()), u_min=z["u_min"].tolist(), u_max=z["u_max"].tolist(),
        wheelbase=float(z["wheelbase"]), v_max=float(z["v_max"]), Q=[float(q) for q in z["Q"]])


def load_env_navigation2d() -> SimpleNamespace:
    z = np.load(os.path.join(GOLDEN_DIR, "env_navigation2d.npz"))
    return SimpleNamespace(
        obstacle=_unpack(z["obstacle_bits"], z["shape"]), cell=float(z["cell"][0]),
        origin=[int(v) for v in z["origin"][0]], lim=[float(v) for v in z["lim"]],
        start_state=torch.from_numpy(z["start_state"].copy()), goal=z["goal"].tolist(), u_min=z["u_min"].tolist(),
        u_max=z["u_max"].tolist())


def oracle_racing_model(env=None) -> mo.RacingModel:
    env = env or load_env_racing()
    q = env.Q
    return mo.RacingModel(
        mo.GridMap(torch.from_numpy(env.obstacle), env.cell[0], env.origin[0]),
        mo.GridMap(torch.from_numpy(env.lane), env.cell[1], env.origin[1]), u_min=env.u_min, u_max=env.u_max,
        wheelbase=env.wheelbase, v_max=env.v_max, lim=tuple(env.lim), Qc=q[0], Ql=q[1], Qv=q[2], Qo=q[3], Qin=q[4],
        Qdin=q[5])


def oracle_navigation2d_model(env=None) -> mo.Navigation2DModel:
    env = env or load_env_navigation2d()
    return mo.Navigation2DModel(mo.GridMap(torch.from_numpy(env.obstacle), env.cell, env.origin), u_min=env.u_min,
                                u_max=env.u_max, goal=env.goal, lim=tuple(env.lim))


def oracle_model(name: str):
    if name == "pendulum":
        return mo.PendulumModel()
    if name == "cartpole":
        return mo.CartpoleModel()
    if name == "mountaincar":
        return mo.MountainCarModel()
    if name == "mujoco_cartpole":
        return mo.CartpoleContinuousModel()
    if name == "goal_in_danger_zone":
        return mo.GoalInDangerZoneModel()
    if name == "navigation2d":
        return oracle_navigation2d_model()
    if name == "racing":
        return oracle_racing_model()
    raise KeyError(name)


GOLDEN_CASES = ["mujoco_cartpole", "goal_in_danger_zone", "pendulum_c1", "pendulum_essps", "cartpole", "cartpole_mpo", "mountaincar", "navigation2d_lbps",
                "navigation2d_essps", "navigation2d_mpo_expl", "racing_sg", "racing_example"]


# one case per BASELINE.json size (configs[1..3]), recorded with the reference's native noise draws; the noise is
# kept as a digest and regenerated (oracle/gen_golden.py:noise_digest, full_size_cases)
FULL_SIZE_CASES = ["full_cartpole_c2", "full_navigation2d_c3", "full_racing_c4"]


def regenerate_noise(case: SimpleNamespace, solver, s: int):
    """Draw solve ``s``'s noise from the oracle's native sampler (must be called once per solve, in order, on a
    solver built with the constructor draw burnt) and hold it against the recorded digest. Returns the noise
    tensor, or None when this host's torch build produces a different normal_() stream than the recording."""
    noise = solver._draw_noise()
    flat = noise.numpy().reshape(-1)
    same = (np.array_equal(noise.numpy()[:4], case.noise_head[s]) and
            np.array_equal(flat[::4099][:4096], case.noise_sample[s]) and
            float(flat.astype(np.float64).sum()) == float(case.noise_sum[s]))
    return noise if same else None


def load_case(name: str) -> SimpleNamespace:
    z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    cfg = json.loads(str(z["cfg"]))
    d = {k: z[k] for k in z.files if k not in ("cfg", "versions")}
    return SimpleNamespace(name=name, cfg=cfg, n_solves=len(d["state"]), **d)


def solver_kwargs(cfg: dict) -> dict:
    """MPPI constructor kwargs (minus callables / bounds) recorded in a golden cfg."""
    keep = ("horizon", "num_samples", "lambda_", "lbps_delta", "essps_target_ess", "lambda_min", "lambda_max",
            "exploration", "use_sg_filter", "sg_window_size", "sg_poly_order", "seed")
    return {k: cfg[k] for k in keep if k in cfg}


def bounds_for(case_cfg: dict, model) -> tuple:
    if "u_min" in case_cfg:
        return case_cfg["u_min"], case_cfg["u_max"]
    return model.u_min.tolist(), model.u_max.tolist()


def build_oracle(case: SimpleNamespace):
    model = oracle_model(case.cfg["model"])
    if case.cfg["model"] == "goal_in_danger_zone":
        model = mo.GoalInDangerZoneModel(goal=case.cfg["goal"], center=case.cfg["center"], radius=case.cfg["radius"])
    u_min, u_max = bounds_for(case.cfg, model)
    solver = mo.OracleMPPI(dim_state=model.dim_state, dim_control=model.dim_control, dynamics=model.dynamics,
                           cost_func=model.cost, u_min=u_min, u_max=u_max, sigmas=case.cfg["sigmas"],
                           **solver_kwargs(case.cfg))
    return model, solver
