#!/usr/bin/env python
"""Multi-GPU check (run under torchrun, one rank per GPU):
     python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py
   The K samples are sharded over the ranks (NCCL all-gather of the shard partials); every rank must
   return the same sequences, and rank 0 compares them with an unsharded solve on its own GPU."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mppi_playground_b200 as eng  # noqa: E402
from engine_util import build_engine  # noqa: E402
from oracle import fixtures as fx  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
device = torch.device("cuda", local)
torch.cuda.set_device(device)
dist.init_process_group("nccl", device_id=device)
CASES = [
    dict(model="racing", horizon=80, num_samples=65536, sigmas=[0.5, 0.1], lambda_=1.0, use_sg_filter=True),
    dict(model="navigation2d", horizon=60, num_samples=32768, sigmas=[0.5, 0.5], lambda_="LBPS"),
    dict(model="cartpole", horizon=50, num_samples=1048576, u_min=[-3.0], u_max=[3.0], sigmas=[1.0], lambda_=0.001),
]
report = []
for cfg in CASES:
    model, sharded = build_engine(cfg, device=device, process_group=dist.group.WORLD)
    single_model, single = build_engine(cfg, device=device) if rank == 0 else (None, None)
    if cfg["model"] == "racing":
        env = fx.load_env_racing()
        state, cind = env.start_state.clone(), 0
    elif cfg["model"] == "navigation2d":
        state = fx.load_env_navigation2d().start_state.clone()
    else:
        state = torch.tensor([0.0, 0.0, 0.05, 0.0])
    worst, top_diff = 0.0, None
    for s in range(3):
        if cfg["model"] == "racing":
            ref, cind = eng.racing_reference_path(state, env.center_path, cind, cfg["horizon"], v_max=env.v_max)
            model.reference_path_tensor = ref
            if single_model is not None:
                single_model.reference_path_tensor = ref
        a, st = sharded.forward(state)
        gathered = [torch.empty_like(a) for _ in range(world)]
        dist.all_gather(gathered, a)
        assert all(torch.equal(g, gathered[0]) for g in gathered), "ranks disagree"
        if s == 0:
            # get_top_samples across the ranks (each rank's 300 best, one NCCL all-gather, merge, re-roll by global
            # sample id): after the first solve the shards' costs are bit-equal to the unsharded solve's, so the
            # winners' trajectories must be identical
            tt, tw = sharded.get_top_samples(300)
            both = [torch.empty_like(tt) for _ in range(world)]
            dist.all_gather(both, tt)
            assert all(torch.equal(g, both[0]) for g in both), "ranks disagree on the top samples"
        if rank == 0:
            a1, s1 = single.forward(state)
            if s == 0:
                t1, w1 = single.get_top_samples(300)
                top_diff = float((tt - t1).abs().max())
                assert float(((tw - w1).abs() / (w1.abs() + 1e-30)).max()) < 1e-3
            worst = max(worst, float((a - a1).abs().max()), float((st - s1).abs().max()))
            nxt = s1[0, 1].clone()
        else:
            nxt = torch.empty(st.shape[-1], device=device)
        dist.broadcast(nxt, 0)
        state = nxt.cpu()
    if rank == 0:
        report.append({"case": f"{cfg['model']}-{cfg['lambda_']}-K{cfg['num_samples']}", "world": world,
                       "max_abs_diff_vs_single_gpu": worst, "top_samples_300_max_abs_diff_vs_single_gpu": top_diff})
        assert worst < 2e-3 and top_diff == 0.0, report[-1]
if rank == 0:
    print(json.dumps(report))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"mgpu_check_n{world}.json"), "w") as f:
        json.dump(report, f)
dist.destroy_process_group()
