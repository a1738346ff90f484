#!/usr/bin/env python
"""Microseconds per control step of what runs BETWEEN two solves, racing K=65536 T=80:

  reference way  env.step + env.collision_check + solver.get_top_samples(300) as stock ATen ops on CUDA tensors
                 (the oracle's restatement of racing_env.py:142-163,374-384 and mppi.py:462-487 moved to the GPU:
                 topk over the K weights, gathers from a materialised [K,T+1,ds] state batch) - ~100 tiny launches;
  round-1 engine get_top_samples through a full cub radix sort of the K costs + a re-roll kernel (n = 1025 forces
                 that path now), env.step / collision_check left to the caller;
  round-2 engine mppi_step_epilogue: ONE launch (radix select + re-roll + env step + goal test + collision flags).

Also the map rasteriser: seconds of Python loops in the reference vs one kernel (800 x 800 racing maps).
CUDA events around back-to-back calls after warm-up. Writes gpurun_out/time_epilogue.json."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mppi_playground_b200 as eng  # noqa: E402
from engine_util import build_engine  # noqa: E402
from oracle import fixtures as fx  # noqa: E402
from oracle import mppi_oracle as mo  # noqa: E402
from test_oracle_epilogue_maps import racing_lane_raster, racing_obstacle_raster  # noqa: E402


def timed(fn, n=200, warm=20):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3, (time.perf_counter() - t0) / n * 1e6  # device us, wall us


cfg = dict(model="racing", horizon=80, num_samples=65536, sigmas=[0.5, 0.1], lambda_=1.0, use_sg_filter=True)
model, solver = build_engine(cfg)
env = fx.load_env_racing()
state = env.start_state.clone().cuda()
model.reference_path_tensor, _ = eng.racing_reference_path(state.cpu(), env.center_path, 0, 80, v_max=env.v_max)
action, states = solver.forward(state)
goal = (float(env.center_path[-1][0]), float(env.center_path[-1][1]))
out = {}

# bare C-ABI calls on preallocated buffers (the Python wrapper's allocations cost more than the kernels)
import ctypes as C  # noqa: E402

from mppi_playground_b200 import _capi  # noqa: E402

lib, h = solver._lib, solver._h
T, ds = 80, 4
nxt, flags = torch.empty(ds, device="cuda"), torch.empty(T + 2, device="cuda")
traj, w = torch.empty(300, T + 1, ds, device="cuda"), torch.empty(300, device="cuda")
cc, ci = torch.empty(1024, device="cuda"), torch.empty(1024, device="cuda", dtype=torch.int32)
seq = states.reshape(T + 1, ds).contiguous()


def args(top_n, step):
    a = _capi.MppiStepEpilogue()
    if step:
        a.d_state, a.d_action_seq, a.d_state_seq = state.data_ptr(), action.data_ptr(), seq.data_ptr()
        a.goal_x, a.goal_y, a.goal_threshold = goal[0], goal[1], 1.0
        a.d_next_state, a.d_flags = nxt.data_ptr(), flags.data_ptr()
    if top_n:
        a.top_n, a.d_top_traj, a.d_top_w = top_n, traj.data_ptr(), w.data_ptr()
    return a


for name, a in (("engine_step_epilogue_top300", args(300, True)), ("engine_step_and_flags_only", args(0, True)),
                ("engine_top_samples_300_only", args(300, False)), ("engine_top_samples_1_only", args(1, False))):
    dev_us, wall_us = timed(lambda: _capi.check(lib.mppi_step_epilogue(h, C.byref(a), None)))
    out[name] = {"device_us": dev_us, "wall_us": wall_us, "launches": lib.mppi_last_epilogue_launches(h)}
dev_us, wall_us = timed(lambda: _capi.check(lib.mppi_top_candidates(h, 300, cc.data_ptr(), ci.data_ptr(), None)))
out["engine_select_only_top_candidates_300"] = {"device_us": dev_us, "wall_us": wall_us}
dev_us, wall_us = timed(lambda: solver.step_epilogue(action, states, state=state, goal=goal, goal_threshold=1.0, top_n=300))
out["python_wrapper_step_epilogue_top300"] = {"device_us": dev_us, "wall_us": wall_us}
dev_us, wall_us = timed(lambda: solver.get_top_samples(1025))
out["engine_get_top_samples_1025_full_sort_round1_path"] = {"device_us": dev_us, "wall_us": wall_us}

# the reference's ATen sequence on CUDA tensors (what pi_mpc.MPPI + RacingEnv run with device='cuda')
omodel = fx.oracle_racing_model(env)
for g in (omodel.obstacle, omodel.lane):
    g.grid, g.origin = g.grid.cuda(), g.origin.cuda()
for k, v in list(vars(omodel).items()):  # bounds, limits, cost weights: every tensor attribute
    if torch.is_tensor(v):
        setattr(omodel, k, v.cuda())
    elif isinstance(v, (list, tuple)) and v and all(torch.is_tensor(x) for x in v):
        setattr(omodel, k, type(v)(x.cuda() for x in v))
weights = solver._weights
state_batch = torch.zeros(cfg["num_samples"], 81, 4, device="cuda")  # the reference keeps this from the solve
goal_t = torch.tensor(goal, device="cuda")


def dynamics_on_device(s, a):  # the op sequence of racing_env.py:341-370 (oracle RacingModel.dynamics) on CUDA
    x, y, theta, v = (s[:, i].view(-1, 1) for i in range(4))
    accel = torch.clamp(a[:, 0].view(-1, 1), omodel.u_min[0], omodel.u_max[0])
    steer = torch.clamp(a[:, 1].view(-1, 1), omodel.u_min[1], omodel.u_max[1])
    theta = mo.wrap_angle(theta)
    dx, dy = v * torch.cos(theta), v * torch.sin(theta)
    dtheta = v * torch.tan(steer) / omodel.L
    new_x, new_y = x + dx * omodel.dt, y + dy * omodel.dt
    new_theta = mo.wrap_angle(theta + dtheta * omodel.dt)
    new_v = v + accel * omodel.dt
    xl = torch.tensor(omodel.lim[:2], device="cuda")  # the reference builds these per call as well (:360-367)
    yl = torch.tensor(omodel.lim[2:], device="cuda")
    return torch.cat([torch.clamp(new_x, xl[0], xl[1]), torch.clamp(new_y, yl[0], yl[1]), new_theta,
                      torch.clamp(new_v, -omodel.V_MAX, omodel.V_MAX)], dim=1)


def reference_way():
    u = torch.clamp(action[0], omodel.u_min, omodel.u_max)
    nxt = dynamics_on_device(state.view(1, -1), u.view(1, -1)).squeeze(0)
    reached = torch.norm(nxt[:2] - goal_t) < 1.0
    coll = omodel.obstacle.lookup(states[:, :, :2])
    idx = torch.topk(weights, 300).indices
    top, tw = state_batch[idx], weights[idx]
    top = top[torch.argsort(tw, descending=True)]
    tw = tw[torch.argsort(tw, descending=True)]
    return nxt, reached, coll, top, tw


dev_us, wall_us = timed(reference_way, n=100, warm=10)
out["aten_on_cuda_reference_sequence"] = {"device_us": dev_us, "wall_us": wall_us}

# rasteriser: both racing maps on the device vs the oracle's (vectorised!) numpy painter on the host
obs, lane = racing_obstacle_raster(), racing_lane_raster()
t0 = time.perf_counter()
for _ in range(20):
    solver.rasterise_map(0, obs)
    solver.rasterise_map(1, lane)
torch.cuda.synchronize()
out["device_raster_both_racing_maps_ms"] = (time.perf_counter() - t0) / 20 * 1e3
t0 = time.perf_counter()
mo.paint_obstacle_map(obs.width, obs.height, obs.discs, obs.rects)
mo.paint_lane_map(lane.width, lane.height, [(x, y) for x, y, _ in lane.discs], lane.r2)
out["numpy_painter_both_racing_maps_ms"] = (time.perf_counter() - t0) * 1e3
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "time_epilogue.json"), "w"), indent=1)
