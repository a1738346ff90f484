#!/usr/bin/env python
"""Solve time of every BASELINE.json config on one GPU (context for profiles/; bench.py stays the
headline). CUDA events around back-to-back solves after warm-up, device-resident inputs."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mppi_playground_b200 as eng  # noqa: E402
from engine_util import build_engine  # noqa: E402
from oracle import fixtures as fx  # noqa: E402

CONFIGS = [
    ("C1 pendulum K=1000 T=50 lambda=1.0", dict(model="pendulum", horizon=50, num_samples=1000, u_min=[-2.0],
                                                 u_max=[2.0], sigmas=[1.0], lambda_=1.0), [3.14, 0.0]),
    ("C2 cartpole K=8192 T=50", dict(model="cartpole", horizon=50, num_samples=8192, u_min=[-3.0], u_max=[3.0],
                                     sigmas=[1.0], lambda_=0.001), [0.0, 0.0, 0.05, 0.0]),
    ("C3 navigation2d K=32768 T=60 LBPS", dict(model="navigation2d", horizon=60, num_samples=32768,
                                               sigmas=[0.5, 0.5], lambda_="LBPS"), None),
    ("C4 racing K=65536 T=80 SG", dict(model="racing", horizon=80, num_samples=65536, sigmas=[0.5, 0.1], lambda_=1.0,
                                       use_sg_filter=True), None),
    ("C5 cartpole K=1048576 T=50 (one GPU)", dict(model="cartpole", horizon=50, num_samples=1048576, u_min=[-3.0],
                                                  u_max=[3.0], sigmas=[1.0], lambda_=0.001), [0.0, 0.0, 0.05, 0.0]),
]
out = []
for name, cfg, s0 in CONFIGS:
    model, solver = build_engine(cfg)
    if cfg["model"] == "racing":
        env = fx.load_env_racing()
        state = env.start_state.clone()
        model.reference_path_tensor, _ = eng.racing_reference_path(state, env.center_path, 0, cfg["horizon"])
    elif cfg["model"] == "navigation2d":
        state = fx.load_env_navigation2d().start_state.clone()
    else:
        state = torch.tensor(s0)
    state = state.cuda()
    for _ in range(10):
        a, s = solver.forward(state)
        state = s[0, 1]
    n = 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        a, s = solver.forward(state)
        state = s[0, 1]
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    info = solver.launch_info()
    out.append({"config": name, "ms_per_solve": ms, "solves_per_s": 1e3 / ms, **info})
    print(f"{name:45s} {ms * 1e3:9.1f} us/solve  {1e3 / ms:10.0f} solves/s  {info}")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "time_configs.json"), "w"), indent=1)
