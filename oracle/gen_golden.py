"""TEST INFRASTRUCTURE ONLY - records golden vectors from the LIVE reference.

Run in the build container (the only place /root/reference exists):

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py

Imports the unmodified reference through ``oracle/ref_harness.py``, runs
closed-loop ``MPPI.forward`` solves on CPU for each case below and writes
``tests/golden/<case>.npz`` holding the inputs (state, injected noise =
``solver._action_noises``, reference path) and outputs (per-sample costs via a
recording wrapper around ``cost_func``, lambda, action_seq, state_seq, top
samples). It also writes the env fixtures (occupancy grids bit-packed, centre
line) that tests and bench.py need where the reference is absent.

Versions recorded in every file: torch / scipy / numpy of this container.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import scipy
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
VERSIONS = json.dumps({"torch": torch.__version__, "scipy": scipy.__version__, "numpy": np.__version__})


class CostRecorder:
    """Wraps a reference cost callable and keeps every per-call cost vector, so
    the per-sample total the reference forms at src/pi_mpc/mppi.py:333-336 can
    be rebuilt without touching the reference."""

    def __init__(self, fn):
        self.fn, self.calls = fn, []

    def __call__(self, state, action, info):
        c = self.fn(state, action, info)
        self.calls.append(c.detach().clone())
        return c

    def take_total(self, horizon: int) -> torch.Tensor:
        calls, self.calls = self.calls[: horizon + 1], self.calls[horizon + 1:]
        stage = torch.stack(calls[:horizon], dim=1)
        return torch.sum(stage, dim=1) + calls[horizon]


def noise_digest(noise: np.ndarray):
    """What a full-size case keeps of its noise tensor instead of the tensor itself (42 MB per solve at
    K=65536, T=80): the case is recorded with the reference's NATIVE draws (torch.manual_seed(seed), one
    constructor draw, one draw per solve - mppi.py:93,146,261), which the oracle's sampler reproduces on the
    same torch build; the digest (head rows, a strided sample, the fp64 sum) proves a regenerated stream is the
    recorded one before anything is compared against the recorded outputs."""
    flat = noise.reshape(-1)
    return noise[:4].copy(), flat[::4099][:4096].copy(), np.float64(flat.astype(np.float64).sum())


def run_case(name, solver, recorder, horizon, state0, advance, n_solves, pre_solve=None, extra_cfg=None, top_n=8,
             store_noise=True):
    rec = {k: [] for k in ("state", "noise", "costs", "lam", "lam_next", "action_seq", "state_seq", "top_traj",
                           "top_w", "refpath", "noise_head", "noise_sample", "noise_sum")}
    state = state0.clone()
    for s in range(n_solves):
        if pre_solve is not None:
            rec["refpath"].append(pre_solve(state).numpy().copy())
        lam_before = solver._lambda
        a, ss = solver.forward(state=state.clone())
        rec["state"].append(state.numpy().copy())
        if store_noise:
            rec["noise"].append(solver._action_noises.numpy().copy())
        else:
            h, sm, tot = noise_digest(solver._action_noises.numpy())
            rec["noise_head"].append(h), rec["noise_sample"].append(sm), rec["noise_sum"].append(tot)
        rec["costs"].append(recorder.take_total(horizon).numpy().copy())
        mode = getattr(solver, "_auto_lambda", None)
        # lambda the weights were formed with: MPO updates after the weights (mppi.py:376 vs :398)
        rec["lam"].append(float(lam_before) if mode in (None, "MPO") else float(solver._lambda))
        rec["lam_next"].append(float(solver._lambda))
        rec["action_seq"].append(a.detach().numpy().copy())
        rec["state_seq"].append(ss.detach().numpy().copy())
        tt, tw = solver.get_top_samples(top_n)
        rec["top_traj"].append(tt.detach().numpy().copy())
        rec["top_w"].append(tw.detach().numpy().copy())
        state = advance(state, a.detach())
    out = {k: np.stack(v) for k, v in rec.items() if len(v)}
    out["cfg"] = np.array(json.dumps(extra_cfg or {}))
    out["versions"] = np.array(VERSIONS)
    path = os.path.join(OUT, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1e3:.0f} kB, lam={rec['lam']}")


def closure_case(name, example, fn_names, cfg, state0, n_solves=3, **run_kw):
    ns = rh.load_reference()
    dyn, cost = rh.extract_closures(example, fn_names)
    recorder = CostRecorder(cost)
    kw = dict(cfg)
    solver = ns.MPPI(dynamics=dyn, cost_func=recorder, u_min=torch.tensor(kw.pop("u_min")),
                     u_max=torch.tensor(kw.pop("u_max")), sigmas=torch.tensor(kw.pop("sigmas")), **kw)

    def advance(state, a):
        return dyn(state.clone().view(1, -1), a[0].view(1, -1)).view(-1)

    run_case(name, solver, recorder, cfg["horizon"], torch.tensor(state0, dtype=torch.float32), advance, n_solves,
             extra_cfg=dict(cfg, model=example, state0=state0), **run_kw)


def nav2d_case(name, cfg, n_solves=3, **run_kw):
    env, ns = rh.make_navigation2d()
    recorder = CostRecorder(env.cost_function)
    kw = dict(cfg)
    solver = ns.MPPI(dim_state=3, dim_control=2, dynamics=env.dynamics, cost_func=recorder, u_min=env.u_min,
                     u_max=env.u_max, sigmas=torch.tensor(kw.pop("sigmas")), **kw)

    def advance(state, a):
        u = torch.clamp(a[0], env.u_min, env.u_max)
        return env.dynamics(state.view(1, -1), u.view(1, -1)).view(-1)

    run_case(name, solver, recorder, cfg["horizon"], env._robot_state.clone(), advance, n_solves,
             extra_cfg=dict(cfg, model="navigation2d"), **run_kw)


def racing_case(name, cfg, n_solves=3, **run_kw):
    env, ctl, ns = rh.make_racing()
    recorder = CostRecorder(ctl.cost_function)
    kw = dict(cfg)
    ctl.solver = ns.MPPI(dim_state=4, dim_control=2, dynamics=env.dynamics, cost_func=recorder, u_min=env.u_min,
                         u_max=env.u_max, sigmas=torch.tensor(kw.pop("sigmas")), **kw)

    def pre_solve(state):  # example/racing.py:73-81
        ctl.reference_path, ctl.current_path_index = ctl.calc_ref_trajectory(
            state, env.racing_center_path, ctl.current_path_index, ctl.solver._horizon, DL=0.1,
            lookahead_distance=3, reference_path_interval=0.85)
        return ctl.reference_path

    def advance(state, a):  # racing_env.py:142-163 (step = clamp + dynamics)
        u = torch.clamp(a[0], env.u_min, env.u_max)
        return env.dynamics(state.view(1, -1), u.view(1, -1)).view(-1)

    run_case(name, ctl.solver, recorder, cfg["horizon"], env._robot_state.clone(), advance, n_solves,
             pre_solve=pre_solve, extra_cfg=dict(cfg, model="racing"), **run_kw)


def goal_zone_case(name, cfg, goal, state0, n_solves=3):
    """GoalInDangerZoneEnv needs gymnasium to be constructed; its two batch methods are plain torch code,
    so they are compiled from the reference file where it lies (src/envs/goal_in_danger_zone.py:113-156)
    and bound to an object that carries exactly the attributes they read."""
    import ast
    import types

    ns = rh.load_reference()
    path = os.path.join(rh.REFERENCE_ROOT, "src", "envs", "goal_in_danger_zone.py")
    tree = ast.parse(open(path).read(), filename=path)
    env_ns = {"torch": torch, "np": np}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in ("parallel_step", "parallel_cost"):
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), env_ns)
    env = types.SimpleNamespace(_v_min=-1.0, _v_max=1.0, _omega_min=-1.0, _omega_max=1.0, _dt=0.1,
                                _goal=np.array(goal), _danger_zone=types.SimpleNamespace(center=[0.0, 0.0], radius=10.0))
    step = types.MethodType(env_ns["parallel_step"], env)
    cost = types.MethodType(env_ns["parallel_cost"], env)
    recorder = CostRecorder(cost)
    kw = dict(cfg)
    solver = ns.MPPI(dim_state=7, dim_control=2, dynamics=step, cost_func=recorder, u_min=torch.tensor([-1.0, -1.0]),
                     u_max=torch.tensor([1.0, 1.0]), sigmas=torch.tensor(kw.pop("sigmas")), **kw)

    def advance(state, a):
        return step(state.view(1, -1), a[0].view(1, -1)).view(-1)

    run_case(name, solver, recorder, cfg["horizon"], torch.tensor(state0, dtype=torch.float32), advance, n_solves,
             extra_cfg=dict(cfg, model="goal_in_danger_zone", goal=list(goal), center=[0.0, 0.0], radius=10.0,
                            state0=state0))


def env_fixtures():
    env, ctl, _ = rh.make_racing()
    om, lm = env._obstacle_map, env._lane_map
    np.savez_compressed(
        os.path.join(OUT, "env_racing.npz"),
        obstacle_bits=np.packbits(om._map_torch.numpy().astype(np.uint8), axis=1),
        lane_bits=np.packbits(lm._map_torch.numpy().astype(np.uint8), axis=1),
        shape=np.array(om._map_torch.shape), cell=np.array([om._cell_size, lm._cell_size]),
        origin=np.array([om._cell_map_origin, lm._cell_map_origin]),
        lim=np.array(om.x_lim + om.y_lim, dtype=np.float64),
        center_path=env.racing_center_path.numpy(), start_state=env._robot_state.numpy(),
        u_min=env.u_min.numpy(), u_max=env.u_max.numpy(), wheelbase=np.array(env.L.item()),
        v_max=np.array(env.V_MAX.item()),
        Q=np.array([ctl.Qc, ctl.Ql, ctl.Qv, ctl.Qo, ctl.Qin, ctl.Qdin]), versions=np.array(VERSIONS))
    env2, _ = rh.make_navigation2d()
    om2 = env2._obstacle_map
    np.savez_compressed(
        os.path.join(OUT, "env_navigation2d.npz"),
        obstacle_bits=np.packbits(om2._map_torch.numpy().astype(np.uint8), axis=1),
        shape=np.array(om2._map_torch.shape), cell=np.array([om2._cell_size]),
        origin=np.array([om2._cell_map_origin]), lim=np.array(om2.x_lim + om2.y_lim, dtype=np.float64),
        start_state=env2._robot_state.numpy(), goal=env2._goal_pos.numpy(), u_min=env2.u_min.numpy(),
        u_max=env2.u_max.numpy(), versions=np.array(VERSIONS))
    print("env fixtures written")


def epilogue_cases():
    """Control-step epilogue: closed loops driven exactly like example/racing.py:229-237 and
    example/navigation2d.py:36-44 - forward, env.step(action_seq[0]), env.collision_check(state_seq),
    get_top_samples - recorded per step, plus probes that reach what the short loops do not: goal flags on both
    sides of the threshold and occupancy flags over obstacles and beyond the map border."""
    rng = np.random.default_rng(7)

    def drive(env, solver, n_solves, pre_solve=None, top_n=64):
        rec = {k: [] for k in ("state", "noise", "refpath", "action_seq", "state_seq", "next_state", "is_goal",
                               "collisions", "top_traj", "top_w", "costs")}
        state = env._robot_state.clone()
        for _ in range(n_solves):
            if pre_solve is not None:
                rec["refpath"].append(pre_solve(state).numpy().copy())
            a, ss = solver.forward(state=state.clone())
            rec["state"].append(state.numpy().copy())
            rec["noise"].append(solver._action_noises.numpy().copy())
            rec["action_seq"].append(a.numpy().copy())
            rec["state_seq"].append(ss.numpy().copy())
            state, goal = env.step(a[0, :])
            rec["next_state"].append(state.numpy().copy())
            rec["is_goal"].append(bool(goal))
            rec["collisions"].append(env.collision_check(state=ss).numpy().copy())
            tt, tw = solver.get_top_samples(top_n)
            rec["top_traj"].append(tt.numpy().copy())
            rec["top_w"].append(tw.numpy().copy())
        return {k: np.stack(v) for k, v in rec.items() if len(v)}

    def probes(env, ds, du, lim):
        goal = env._goal_pos.numpy()
        thr = 1.0 if ds == 4 else 0.5  # racing_env.py:158 / navigation_2d.py:112
        st = np.zeros((24, ds), dtype=np.float32)
        st[:, :2] = goal + rng.uniform(-1.6 * thr, 1.6 * thr, size=(24, 2))
        st[:, 2] = rng.uniform(-3.0, 3.0, size=24)
        act = rng.uniform(-1.0, 1.0, size=(24, du)).astype(np.float32) * 3.0  # beyond the env bounds: step clamps
        nxt, reached = [], []
        for s_, a_ in zip(st, act):
            env._robot_state = torch.tensor(s_)
            n_, g_ = env.step(torch.tensor(a_))
            nxt.append(n_.numpy().copy()), reached.append(bool(g_))
        traj = np.zeros((1, 600, ds), dtype=np.float32)
        traj[0, :, :2] = rng.uniform(-1.15 * lim, 1.15 * lim, size=(600, 2))  # some positions beyond the border
        coll = env.collision_check(state=torch.tensor(traj)).numpy().copy()
        return dict(probe_state=st, probe_action=act, probe_next=np.stack(nxt), probe_goal=np.array(reached),
                    coll_probe_in=traj, coll_probe_out=coll, goal=goal, goal_threshold=np.float32(thr))

    env, ctl, ns = rh.make_racing()
    cfg = dict(horizon=25, num_samples=1024, sigmas=[0.5, 0.1], lambda_=1.0)
    kw = dict(cfg)
    ctl.solver = ns.MPPI(dim_state=4, dim_control=2, dynamics=env.dynamics, cost_func=ctl.cost_function,
                         u_min=env.u_min, u_max=env.u_max, sigmas=torch.tensor(kw.pop("sigmas")), **kw)

    def pre_solve(state):
        ctl.reference_path, ctl.current_path_index = ctl.calc_ref_trajectory(
            state, env.racing_center_path, ctl.current_path_index, ctl.solver._horizon, DL=0.1,
            lookahead_distance=3, reference_path_interval=0.85)
        return ctl.reference_path

    out = drive(env, ctl.solver, 4, pre_solve)
    out.update(probes(env, 4, 2, 40.0))
    out["cfg"] = np.array(json.dumps(dict(cfg, model="racing")))
    out["versions"] = np.array(VERSIONS)
    np.savez_compressed(os.path.join(OUT, "epilogue_racing.npz"), **out)
    print("epilogue_racing: goal flags", out["probe_goal"].sum(), "/ 24, occupied probes", int(out["coll_probe_out"].sum()))

    env2, ns = rh.make_navigation2d()
    cfg2 = dict(horizon=30, num_samples=768, sigmas=[0.5, 0.5], lambda_=0.5)
    kw = dict(cfg2)
    solver2 = ns.MPPI(dim_state=3, dim_control=2, dynamics=env2.dynamics, cost_func=env2.cost_function,
                      u_min=env2.u_min, u_max=env2.u_max, sigmas=torch.tensor(kw.pop("sigmas")), **kw)
    out = drive(env2, solver2, 4)
    out.update(probes(env2, 3, 2, 10.0))
    out["cfg"] = np.array(json.dumps(dict(cfg2, model="navigation2d")))
    out["versions"] = np.array(VERSIONS)
    np.savez_compressed(os.path.join(OUT, "epilogue_navigation2d.npz"), **out)
    print("epilogue_navigation2d: goal flags", out["probe_goal"].sum(), "/ 24, occupied probes",
          int(out["coll_probe_out"].sum()))


def shape_fixtures():
    """Map construction inputs: the shapes the reference envs paint their grids from
    (obstacle_map_2d.py:235-345 with default_rng(seed), racing_env.py:59-92, navigation_2d.py:34-51); the grids
    themselves are already in env_racing.npz / env_navigation2d.npz."""
    env, _, _ = rh.make_racing()
    om = env._obstacle_map
    env2, _ = rh.make_navigation2d()
    om2 = env2._obstacle_map
    np.savez_compressed(
        os.path.join(OUT, "env_shapes.npz"),
        racing_circle_centers=np.stack([c.center for c in om.circle_obs_list]),
        racing_circle_radii=np.array([c.radius for c in om.circle_obs_list]),
        racing_map_size=np.array(env.map_size), racing_cell=np.array(env.cell_size),
        racing_lane_width=np.array(env.line_width * 0.8),
        nav_circle_centers=np.stack([c.center for c in om2.circle_obs_list]),
        nav_circle_radii=np.array([c.radius for c in om2.circle_obs_list]),
        nav_rect_centers=np.stack([r.center for r in om2.rectangle_obs_list]),
        nav_rect_wh=np.array([[r.width, r.height] for r in om2.rectangle_obs_list]),
        nav_map_size=np.array([20, 20]), nav_cell=np.array(om2._cell_size), versions=np.array(VERSIONS))
    print("shape fixtures written")


def full_size_cases():
    """One recorded case per BASELINE.json size (configs[1], [2], [3]) from the live reference: native noise
    draws, the noise kept as a digest (see noise_digest), every one of the K costs kept."""
    closure_case("full_cartpole_c2", "cartpole", ["dynamics", "stage_cost"],
                 dict(horizon=50, num_samples=8192, dim_state=4, dim_control=1, u_min=[-3.0], u_max=[3.0],
                      sigmas=[1.0], lambda_=0.001), [0.0, 0.0, 0.05, 0.0], n_solves=2, store_noise=False)
    nav2d_case("full_navigation2d_c3", dict(horizon=60, num_samples=32768, sigmas=[0.5, 0.5], lambda_="LBPS"),
               n_solves=2, store_noise=False)
    racing_case("full_racing_c4", dict(horizon=80, num_samples=65536, sigmas=[0.5, 0.1], lambda_=1.0,
                                       use_sg_filter=True), n_solves=2, store_noise=False)


def main(only=None):
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    if only == "full":
        full_size_cases()
        return
    if only == "epilogue":
        epilogue_cases()
        shape_fixtures()
        return
    if only is None:
        env_fixtures()
    # BASELINE.json config 1, verbatim: the reference's own CPU-runnable case
    if only in (None, "extra"):
        closure_case("mujoco_cartpole", "mujoco_cartpole", ["dynamics", "cost_func"],
                     dict(horizon=50, num_samples=512, dim_state=4, dim_control=1, u_min=[-3.0], u_max=[3.0],
                          sigmas=[1.0], lambda_=1.0), [0.0, 0.0, 0.05, 0.0])  # example/mujoco_cartpole.py:95-106
        goal_zone_case("goal_in_danger_zone", dict(horizon=30, num_samples=768, sigmas=[0.5, 0.5], lambda_=1.0),
                       goal=(3.2, -4.1), state0=[12.0, 9.0, -2.4, 3.2 - 12.0, -4.1 - 9.0, -12.0, -9.0])
    if only == "extra":
        return
    closure_case("pendulum_c1", "pendulum", ["dynamics", "cost_function"],
                 dict(horizon=50, num_samples=1000, dim_state=2, dim_control=1, u_min=[-2.0], u_max=[2.0],
                      sigmas=[1.0], lambda_=1.0), [3.14, 0.0])
    closure_case("pendulum_essps", "pendulum", ["dynamics", "cost_function"],
                 dict(horizon=15, num_samples=1000, dim_state=2, dim_control=1, u_min=[-2.0], u_max=[2.0],
                      sigmas=[1.0], lambda_="ESSPS"), [2.5, 0.5])  # example/pendulum.py:58-69
    closure_case("cartpole", "cartpole", ["dynamics", "stage_cost"],
                 dict(horizon=50, num_samples=512, dim_state=4, dim_control=1, u_min=[-3.0], u_max=[3.0],
                      sigmas=[1.0], lambda_=0.001), [0.0, 0.0, 0.05, 0.0])
    closure_case("cartpole_mpo", "cartpole", ["dynamics", "stage_cost"],
                 dict(horizon=20, num_samples=512, dim_state=4, dim_control=1, u_min=[-3.0], u_max=[3.0],
                      sigmas=[1.0], lambda_="MPO", use_sg_filter=True), [0.0, 0.1, 0.05, -0.1], n_solves=5)
    closure_case("mountaincar", "mountaincar", ["dynamics", "cost_func"],
                 dict(horizon=100, num_samples=384, dim_state=2, dim_control=1, u_min=[-1.0], u_max=[1.0],
                      sigmas=[1.0], lambda_=0.1), [-0.5, 0.0])
    nav2d_case("navigation2d_lbps", dict(horizon=60, num_samples=512, sigmas=[0.5, 0.5], lambda_="LBPS"))
    nav2d_case("navigation2d_essps", dict(horizon=30, num_samples=768, sigmas=[0.5, 0.5], lambda_="ESSPS"))
    nav2d_case("navigation2d_mpo_expl",
               dict(horizon=30, num_samples=512, sigmas=[0.5, 0.5], lambda_="MPO", exploration=0.25), n_solves=4)
    racing_case("racing_sg", dict(horizon=80, num_samples=512, sigmas=[0.5, 0.1], lambda_=1.0, use_sg_filter=True),
                n_solves=4)
    racing_case("racing_example", dict(horizon=25, num_samples=1024, sigmas=[0.5, 0.1], lambda_=1.0), n_solves=3)


if __name__ == "__main__":
    # `extra`: only the later-added models; `full`: BASELINE sizes; `epilogue`: control-step epilogue + map shapes
    main(sys.argv[1] if len(sys.argv) > 1 else None)
