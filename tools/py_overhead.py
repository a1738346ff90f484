#!/usr/bin/env python
"""Host-side cost of one solve call: MPPI.forward (Python drop-in) vs mppi_solve (C ABI through ctypes),
racing K=65536 T=80. Prints CPU microseconds per call (enqueue only) and wall microseconds per solve."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import mppi_playground_b200 as eng  # noqa: E402
from engine_util import build_engine  # noqa: E402
from mppi_playground_b200 import _capi  # noqa: E402
from oracle import fixtures as fx  # noqa: E402

env = fx.load_env_racing()
for K in (1024, 65536):
    cfg = dict(bench.CFG, num_samples=K)
    model, solver = build_engine(cfg)
    state = env.start_state.cuda()
    model.reference_path_tensor = eng.racing_reference_path(env.start_state, env.center_path, 0, 80)[0].cuda()
    for _ in range(20):
        solver.forward(state)
    torch.cuda.synchronize()
    n = 2000
    t0 = time.perf_counter()
    for _ in range(n):
        a, s = solver.forward(state)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    lib, h = solver._lib, solver._h
    ref = model.reference_path_tensor
    sp = torch.cuda.current_stream().cuda_stream
    t3 = time.perf_counter()
    for _ in range(n):
        _capi.check(lib.mppi_solve(h, state.data_ptr(), ref.data_ptr(), None, a.data_ptr(), s.data_ptr(), sp))
    t4 = time.perf_counter()
    torch.cuda.synchronize()
    t5 = time.perf_counter()
    print(f"K={K}: forward() enqueue {1e6 * (t1 - t0) / n:.1f} us/call, wall {1e6 * (t2 - t0) / n:.1f} us/solve | "
          f"mppi_solve enqueue {1e6 * (t4 - t3) / n:.1f} us/call, wall {1e6 * (t5 - t3) / n:.1f} us/solve")
