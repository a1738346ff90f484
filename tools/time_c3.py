#!/usr/bin/env python
"""Per-launch timeline of one LBPS solve at BASELINE.json configs[2] (navigation2d K=32768 T=60): run under
   ncu --metrics gpu__time_duration.sum to see the three launches (costs | lambda search | reduce + finish)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from engine_util import build_engine  # noqa: E402
from oracle import fixtures as fx  # noqa: E402

lam = sys.argv[1] if len(sys.argv) > 1 else "LBPS"
cfg = dict(model="navigation2d", horizon=60, num_samples=32768, sigmas=[0.5, 0.5], lambda_=lam)
model, solver = build_engine(cfg)
state = fx.load_env_navigation2d().start_state.clone().cuda()
for _ in range(6):
    a, s = solver.forward(state)
    state = s[0, 1]
torch.cuda.synchronize()
print("lambda", solver._lambdas(), solver.launch_info())
