#!/usr/bin/env python
"""Per-block phase timeline of the fused solve kernel at the bench workload (profiling aid).
   python tools/block_trace.py [--block-size N]   -> prints phase durations (us) across blocks."""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import mppi_playground_b200 as eng  # noqa: E402
from engine_util import build_engine  # noqa: E402
from mppi_playground_b200 import _capi  # noqa: E402
from oracle import fixtures as fx  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--block-size", type=int, default=0)
ap.add_argument("--solves", type=int, default=30)
ap.add_argument("--flush", action="store_true", help="write a 192 MiB buffer before the traced solve (cold L2, like "
                                                     "bench.py's timed steps)")
a = ap.parse_args()
env = fx.load_env_racing()
model, solver = build_engine(bench.CFG, block_size=a.block_size)
state, cind = env.start_state.clone(), 0
lib, h = solver._lib, solver._h
for s in range(a.solves):
    ref, cind = eng.racing_reference_path(state, env.center_path, cind, bench.HORIZON, v_max=env.v_max)
    model.reference_path_tensor = ref
    if s == a.solves - 1:
        _capi.check(lib.mppi_block_trace(h, 1, None, 0))
        if a.flush:
            torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device="cuda").zero_()
            torch.cuda.synchronize()
    _, seq = solver.forward(state)
    state = seq[0, 1].cpu()
info = solver.launch_info()
g = info["grid"]  # row 0: the finisher block, rows 1..: the workers
buf = (C.c_uint64 * (g * 24))()
_capi.check(lib.mppi_block_trace(h, 1, buf, g))
t = np.array(buf, dtype=np.uint64).reshape(g, 24).astype(np.int64)
t0 = t[:, 0].min()
rel = (t - t0) / 1e3
print("launch", info)
print("workers (blocks 1..):")
w = rel[1:]
for i, n in [(0, "start"), (1, "staged"), (2, "costs"), (3, "weights"), (4, "partial")]:
    col = w[:, i][t[1:, i] > 0]
    if len(col):
        print(f"  {n:9s} n={len(col):4d} min={col.min():8.2f} median={np.median(col):8.2f} max={col.max():8.2f} us")
d = w[:, 2] - w[:, 1]
print("  pass1 (staged->costs) per block: min %.2f median %.2f max %.2f us" % (d.min(), np.median(d), d.max()))
d = w[:, 4] - w[:, 3]
print("  pass2 (weights->partial) per block: min %.2f median %.2f max %.2f us" % (d.min(), np.median(d), d.max()))
print("finisher (block 0):")
for i, n in [(0, "start"), (1, "warm-up pass over"), (2, "all workers' tickets seen"), (3, "partials in shared memory"),
             (8, "  combine: header reductions done"), (9, "  combine: numerators done"),
             (5, "combined"), (7, "pre-rollout (SG, carry done)"), (10, "  rollout: controls / tan done"),
             (11, "  rollout: speed chain done (warp 0)"), (12, "  rollout: heading chain done (warp 1)"),
             (14, "  rollout: sin / cos done (warp 2)"), (13, "  rollout: position chains done, block joined"),
             (6, "finished")]:
    if t[0, i] > 0:
        print(f"  {n:40s} {rel[0, i]:8.2f} us")
print("  tail after the last worker's partial: %.2f us" % (rel[0, 6] - w[:, 4].max()))
if t[0, 15] > 0 and t[0, 12] > t[0, 10] > 0:
    cyc, ns = int(t[0, 15]), int(t[0, 12] - t[0, 10])
    print(f"  heading chain: {cyc} SM cycles in {ns} ns = {cyc / ns * 1e3:.0f} MHz, {cyc / bench.HORIZON:.1f} cycles per stage")

