#!/usr/bin/env python
"""A handful of control-step epilogues after one racing solve (K=65536, T=80) - the command ncu wraps for the
epilogue kernels (profiles/): `ncu --set full -k regex:"epilogue|reroll|topn" python tools/epilogue_once.py`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mppi_playground_b200 as eng  # noqa: E402
from engine_util import build_engine  # noqa: E402
from oracle import fixtures as fx  # noqa: E402

cfg = dict(model="racing", horizon=80, num_samples=65536, sigmas=[0.5, 0.1], lambda_=1.0, use_sg_filter=True)
model, solver = build_engine(cfg)
env = fx.load_env_racing()
state = env.start_state.clone().cuda()
model.reference_path_tensor, _ = eng.racing_reference_path(state.cpu(), env.center_path, 0, 80, v_max=env.v_max)
action, states = solver.forward(state)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    out = solver.step_epilogue(action, states, state=state, goal=(0.0, 0.0), goal_threshold=1.0, top_n=300)
torch.cuda.synchronize()
print("ok", float(out[3][1][0]))
