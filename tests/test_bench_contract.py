"""bench.py contract on the CPU: the reference arm (`--impl reference`) runs without a GPU and prints
one JSON line with the keys the driver reads; the GPU arm's static pieces are importable."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "3"],
                         cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "solves/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("MPPI solves/sec") and line["n_gpus"] == 1 and line["gpu_launches"] == 0
    assert line["value"] > 0 and abs(line["ms_per_step"] - 1e3 / line["value"]) < 1e-6 * line["ms_per_step"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and 1 <= cb["cores"] <= 32 and "K=65536" in cb["sample"] and cb["value"] == line["value"]
    assert line["warmup"] >= 1 and set(line["config"]) == {"workload"}  # same `config` dict as the GPU arm prints
    assert line["e2e"] == {"value": line["value"], "unit": "solves/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    assert "K=65536" in line["config"]["workload"] and "T=80" in line["config"]["workload"]


def test_gpu_arm_constants_and_fixture():
    sys.path.insert(0, ROOT)
    import bench

    assert bench.FLOPS_PER_SOLVE == 65536 * 80 * 97 + 65536 * 61
    assert bench.BYTES_PER_SOLVE == 4 * 65536 + 4 * 80 * 2 * 2 + 16 * 81 + 4 * 81 * 4 + 2 * 800 * 25 * 4
    assert (bench.H2D_BYTES, bench.D2H_BYTES) == (1312, 1936)
    env = bench.load_racing_fixture()
    assert env["obstacle"].shape == (800, 800) and int(env["obstacle"].sum()) == 18602
    assert int(env["lane"].sum()) == 445529 and tuple(env["center_path"].shape) == (3678, 3)
    assert bench.ncu_traffic() is None or bench.ncu_traffic() > 0
    # BASELINE.json configs[4]
    K, T, ds, du, flops, nbytes, h2d, d2h = bench.wl_numbers(bench.WORKLOADS["c5"])
    assert (K, T, ds, du) == (1048576, 50, 4, 1) and flops == K * T * 52 + K * 20 and (h2d, d2h) == (16, 200 + 816)


def test_same_gpu_reference_leg_keeps_every_tensor_on_its_device(monkeypatch):
    """bench.time_oracle_on_gpu (the oracle port with its tensors on the GPU, context for the CPU ratio): run on
    the `meta` device, where any tensor the oracle would leave on the CPU raises on first use."""
    import torch

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bench
    import mppi_playground_b200 as eng
    from oracle import fixtures as fx

    monkeypatch.setattr(torch.cuda, "synchronize", lambda device=None: None)
    wl = dict(bench.WORKLOADS["c4"])
    wl["cfg"] = dict(wl["cfg"], num_samples=256, u_min=[-2.0, -0.25], u_max=[2.0, 0.25])
    env = fx.load_env_racing()
    ref, _ = eng.racing_reference_path(env.start_state, env.center_path, 0, 80, v_max=env.v_max)
    out = bench.time_oracle_on_gpu(wl, torch.device("meta"), env.start_state.view(1, -1), ref.view(1, 81, 4),
                                   n_timed=1, n_warm=0)
    assert out["value"] > 0 and out["unit"] == "solves/s"
