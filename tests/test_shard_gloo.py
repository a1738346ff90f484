"""N>1 host path on CPU: two gloo ranks each reduce their shard of the samples
(oracle arithmetic), exchange the shard partials exactly the way the engine's
sharded solve does (mppi_playground_b200.mppi.gather_shards / all_gather), and
must land on the single-process result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, case_name, out_dir):
    sys.path.insert(0, ROOT)
    from mppi_playground_b200.mppi import gather_shards, shard_bounds
    from oracle import fixtures as fx
    from oracle import shard_math as sm

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        case = fx.load_case(case_name)
        model, solver = fx.build_oracle(case)
        if hasattr(case, "refpath"):
            model.reference_path = torch.from_numpy(case.refpath[0])
        K = case.cfg["num_samples"]
        lo, hi = shard_bounds(K, world, rank)
        # every rank rolls only its own samples: a K_local-sample oracle on the shard's noise rows
        shard_cfg = dict(case.cfg, num_samples=hi - lo)
        shard_case = type(case)(**{**case.__dict__, "cfg": shard_cfg})
        smodel, ssolver = fx.build_oracle(shard_case)
        if hasattr(case, "refpath"):
            smodel.reference_path = torch.from_numpy(case.refpath[0])
        tr = ssolver.forward(torch.from_numpy(case.state[0]), noise=torch.from_numpy(case.noise[0][lo:hi]))
        # (a) variable-length gather used for the LBPS / ESSPS cost exchange
        all_costs = gather_shards(tr.costs, K, world)
        np.testing.assert_array_equal(all_costs.numpy(), case.costs[0])
        # (b) fixed-size partial exchange + combine
        lam = float(case.lam[0])
        part = torch.from_numpy(sm.shard_partial(tr.costs.numpy(), tr.perturbed.numpy(), lam))
        bufs = [torch.empty_like(part) for _ in range(world)]
        dist.all_gather(bufs, part)
        opt = sm.combine_partials(torch.stack(bufs).numpy()).reshape(case.cfg["horizon"], -1)
        np.save(os.path.join(out_dir, f"opt_{rank}.npy"), opt)
        # (c) get_top_samples across ranks: every rank's n best (cost, GLOBAL id) pairs - the contract of
        #     mppi_top_candidates: stable ascending order, padding (+inf, -1) when the shard is smaller than n -
        #     travel as one buffer (ids bit-reinterpreted), and the merge of the gathered lists by (cost, id) is the
        #     top-n of all K samples
        from mppi_playground_b200.mppi import gather_candidates

        n = 40
        order = torch.sort(tr.costs, stable=True).indices[:n]
        cost_l = torch.full((n,), float("inf"))
        ids_l = torch.full((n,), -1, dtype=torch.int32)
        cost_l[: len(order)], ids_l[: len(order)] = tr.costs[order], (order + lo).to(torch.int32)
        cc, ci = gather_candidates(cost_l, ids_l, world)
        assert cc.shape == (world * n,) and ci.dtype == torch.int32
        assert torch.equal(cc[rank * n:(rank + 1) * n], cost_l) and torch.equal(ci[rank * n:(rank + 1) * n], ids_l)
        key = np.lexsort((ci.numpy().astype(np.uint32), cc.numpy()))[:n]  # by cost, ties by (unsigned) id
        full = torch.sort(torch.from_numpy(case.costs[0]), stable=True)
        np.testing.assert_array_equal(ci.numpy()[key], full.indices[:n].numpy())
        np.testing.assert_array_equal(cc.numpy()[key], full.values[:n].numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case_name", ["racing_example", "pendulum_c1"])
def test_two_rank_sharded_reduction_equals_single_process(tmp_path, case_name):
    import socket

    from oracle import fixtures as fx

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, case_name, str(tmp_path)), nprocs=2, join=True)
    case = fx.load_case(case_name)
    a, b = np.load(tmp_path / "opt_0.npy"), np.load(tmp_path / "opt_1.npy")
    np.testing.assert_array_equal(a, b)  # every rank finishes with the same sequence
    # no SG filter in these cases -> the combined mean IS the reference's action_seq
    np.testing.assert_allclose(a, case.action_seq[0], rtol=2e-5, atol=2e-6)
