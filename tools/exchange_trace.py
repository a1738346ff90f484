#!/usr/bin/env python
"""In-kernel timeline of the fused NVLink shard exchange, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/exchange_trace.py [--config c4|c5]

Every rank's finisher block stamps %globaltimer when it enters the exchange, when its partial has been stored into
every peer's mailbox, when all peers' flags have arrived and when the gathered partials are copied. Rank 0 prints
the per-rank numbers (us) of the last of 40 closed-loop solves and their summary as JSON."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mppi_playground_b200 import _capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c4", choices=sorted(bench.WORKLOADS))
ap.add_argument("--solves", type=int, default=40)
a = ap.parse_args()
real_stdout = os.dup(1)  # keep stdout for the JSON (NCCL prints its version banner there)
os.dup2(2, 1)
rank, local_rank, world = bench.dist_env()
device = torch.device("cuda", local_rank)
torch.cuda.set_device(device)
dist.init_process_group("nccl", device_id=device)
wl = bench.WORKLOADS[a.config]
model, solver = bench.make_engine(wl, device, process_group=dist.group.WORLD)
assert solver._fused, "the fused peer exchange is not connected"
T = wl["cfg"]["horizon"]
if wl["refpath"]:
    import mppi_playground_b200 as eng

    env = bench.load_racing_fixture()
    state = env["start_state"].clone()
    ref, _ = eng.racing_reference_path(state, env["center_path"], 0, T, v_max=env["v_max"])
    model.reference_path_tensor = ref.to(device)
else:
    state = torch.tensor(wl["state0"])
state = state.to(device)
lib, h = solver._lib, solver._h
gate = torch.zeros(1, device=device)
for s in range(a.solves):
    if s == a.solves - 1:
        _capi.check(lib.mppi_block_trace(h, 1, None, 0))
    solver.peer_barrier()  # ranks aligned on the device before every solve, like bench.py's timed steps
    solver.forward(state)
solver.check_exchange()
buf = (C.c_uint64 * 24)()
_capi.check(lib.mppi_block_trace(h, 1, buf, 1))
t = np.array(buf, dtype=np.int64)
mine = {"rank": rank, "workers_done_to_exchange_us": (t[16] - t[2]) / 1e3, "send_us": (t[17] - t[16]) / 1e3,
        "poll_and_gather_us": (t[18] - t[17]) / 1e3,
        "exchange_total_us": (t[19] - t[16]) / 1e3, "kernel_us": (t[6] - t[0]) / 1e3,
        "tail_after_exchange_us": (t[6] - t[19]) / 1e3}
allr = [None] * world
dist.all_gather_object(allr, mine)
if rank == 0:
    ex = [r["exchange_total_us"] for r in allr]
    out = {"config": a.config, "n_gpus": world, "ranks": allr,
           "exchange_total_us": {"min": min(ex), "median": float(np.median(ex)), "max": max(ex)},
           "note": "send = P tagged 8-byte words stored to every peer over NVLink (no fence); poll_and_gather = "
                   "waiting until every rank's words carry this solve's sequence number (includes the residual skew "
                   "between the ranks' kernels) while copying them out"}
    os.write(real_stdout, (json.dumps(out, indent=1) + "\n").encode())
dist.destroy_process_group()
