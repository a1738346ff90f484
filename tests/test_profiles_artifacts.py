"""The committed evidence under profiles/ is complete and self-consistent (guards later rounds against
stale or missing artefacts). CPU only."""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def _load(name):
    with open(os.path.join(P, name)) as f:
        return json.load(f)


def test_bench_lines_carry_the_contract_keys():
    for n in (1, 2, 4, 8):
        d = _load(f"bench_r01_n{n}.json")
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
            assert k in d, (n, k)
        assert d["n_gpus"] == n and d["unit"] == "solves/s" and d["dtype"] == "f32" and d["vs_baseline"] is None
        assert abs(d["value"] - 1e3 / d["ms_per_step"]) < 1e-6 * d["value"]
        assert d["gpu_launches"] == d["steps"] * d["config"]["launch"]["launches_last_solve"]
        assert d["e2e"]["h2d_bytes_per_step"] == 1312 and d["e2e"]["d2h_bytes_per_step"] == 1936
        assert d["e2e"]["value"] != d["value"]  # measured separately, not a copy of the device-timed number
        r = d["roofline"]
        assert 0 < r["frac"] < 1 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    d1 = _load("bench_r01_n1.json")
    assert d1["cpu_baseline"]["kind"] == "port" and d1["cpu_baseline"]["value"] > 0
    assert d1["roofline"]["traffic"] is not None and d1["roofline"]["hbm"]["frac"] < 0.01
    ref = _load("bench_r01_reference.json")
    assert ref["impl"] == "reference" and ref["value"] > 0
    assert d1["e2e"]["value"] / ref["value"] > 1000  # the headline ratio this round: ~5.6e3


def test_ncu_summaries_match_the_traffic_record():
    md = open(os.path.join(P, "r01_solve_kernel_ncu.md")).read()
    assert "solve_kernel<Racing" in md and "UBLKCP" in md
    m = re.search(r"dram__bytes_read.sum` \| ([0-9.]+) \| Kbyte", md)
    assert m and abs(float(m.group(1)) * 1e3 - 292864) < 1.0  # (ncu_traffic.json now holds the round-2 capture)
    launches = open(os.path.join(P, "r01_launches_ncu.md")).read()
    assert "solve_kernel" in launches and "FillFunctor" in launches


def test_parity_and_multi_gpu_records():
    md = open(os.path.join(P, "parity_r01.md")).read()
    rows = [l for l in md.splitlines() if l.startswith(("| golden/", "| native/", "| edge/"))]
    assert len(rows) >= 28
    for l in rows:
        cells = [c.strip() for c in l.strip("|").split("|")]
        assert float(cells[3]) == 0.0, l  # no occupancy-cell flips on any recorded case
        assert float(cells[2]) < 2e-5, l  # per-sample cost bar
    for n in (2, 8):
        for case in _load(f"mgpu_check_r01_n{n}.json"):
            assert case["world"] == n and case["max_abs_diff_vs_single_gpu"] < 1e-5


# ---- round 2 ---------------------------------------------------------------------------------------------------
R02_BENCH = ["bench_r02_n1.json", "bench_r02_c5_n1.json", "bench_r02_c4_n2.json", "bench_r02_c5_n2.json",
             "bench_r02_c4_n8.json", "bench_r02_c5_n8.json"]


def test_round2_bench_lines_carry_the_contract_keys():
    for name in R02_BENCH:
        d = _load(name)
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
            assert k in d, (name, k)
        assert d["unit"] == "solves/s" and d["dtype"] == "f32" and d["vs_baseline"] is None and d["warmup"] >= 3
        assert abs(d["value"] - 1e3 / d["ms_per_step"]) < 1e-6 * d["value"]
        assert d["gpu_launches"] == d["steps"] * d["details"]["launch"]["launches_last_solve"]
        assert d["e2e"]["value"] != d["value"] and d["e2e"]["h2d_bytes_per_step"] > 0
        r = d["roofline"]
        assert 0 < r["frac"] < 1 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert "measured" in r["peak_source"] and r["fp32_microbench"]["ffma_tflops"] > 50
        assert d["clocks"]["samples"] > 0
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if d["n_gpus"] > 1:
            assert "peer barrier" in d["details"]["rank_alignment"]
    d1 = _load("bench_r02_n1.json")
    assert d1["n_gpus"] == 1 and d1["value"] > 17000 and d1["roofline"]["frac"] > 0.12  # round 1: 13.2k, 0.098
    assert d1["cpu_baseline"]["kind"] == "port" and d1["cpu_baseline"]["value"] > 0
    assert d1["details"]["control_step_epilogue_us"]["value"] < 60
    ref = _load("bench_r02_reference.json")
    assert ref["impl"] == "reference" and ref["value"] > 0 and ref["config"] == d1["config"]
    assert d1["e2e"]["value"] / ref["value"] > 3000
    # strong scaling as measured (the limiter is named in bench_r02_scaling.md)
    c5 = {n: _load(f)["value"] for n, f in ((1, "bench_r02_c5_n1.json"), (2, "bench_r02_c5_n2.json"),
                                            (8, "bench_r02_c5_n8.json"))}
    assert c5[2] / c5[1] > 1.6 and c5[8] / c5[1] > 4.0
    assert _load("bench_r02_c4_n8.json")["value"] > d1["value"]


def test_round2_ncu_summaries_and_timelines():
    md = open(os.path.join(P, "r02_solve_kernel_ncu.md")).read()
    assert "solve_kernel<Racing" in md and "UBLKCP" in md
    t = _load("ncu_traffic.json")
    m = re.search(r"dram__bytes_read.sum` \| ([0-9.]+) \| Kbyte", md)
    assert m and abs(float(m.group(1)) * 1e3 - t["dram_read"]) < 1.0
    assert "r02_solve_kernel_ncu.md" in t["source"]
    assert t["pipes"]["tensor_pipe_pct"] == 0.0 and t["pipes"]["warp_instructions_per_launch"] < 25e6  # round 1: 32.0 M
    launches = open(os.path.join(P, "r02_launches_ncu.md")).read()
    assert "solve_kernel<mppi::Racing, 0, 0, 2, 0>" in launches and "FillFunctor" in launches
    ep = open(os.path.join(P, "r02_epilogue_ncu.md")).read()
    for k in ("control_epilogue_kernel", "reroll_winners_kernel", "topn_select_kernel"):
        assert k in ep
    sass = open(os.path.join(P, "r02_pass1_sass_budget.md")).read()
    assert "FFMA2" in sass and "FMUL2" in sass
    trace = open(os.path.join(P, "r02_block_trace.txt")).read()
    assert "heading chain" in trace and "tail after the last worker's partial" in trace
    te = _load("time_epilogue_r02.json")
    assert te["engine_step_epilogue_top300"]["device_us"] < 0.2 * te["aten_on_cuda_reference_sequence"]["device_us"]
    assert te["engine_step_epilogue_top300"]["device_us"] < te["engine_get_top_samples_1025_full_sort_round1_path"]["device_us"]
    for n in (2, 8):
        for cfg in ("c4", "c5"):
            ex = _load(f"exchange_r02_{cfg}_n{n}.json")
            assert ex["n_gpus"] == n and len(ex["ranks"]) == n and ex["exchange_total_us"]["max"] < 15.0


def test_round2_parity_and_multi_gpu_records():
    md = open(os.path.join(P, "parity_r02.md")).read()
    rows = [l for l in md.splitlines() if l.startswith(("| golden", "| native/", "| edge/", "| fullsize/"))]
    assert len(rows) >= 35
    for l in rows:
        cells = [c.strip() for c in l.strip("|").split("|")]
        assert float(cells[3]) == 0.0, l  # no occupancy-cell flips on any recorded case
        assert float(cells[2]) < 2e-5, l  # per-sample cost bar
    assert any(l.startswith("| fullsize/cartpole-K1048576") for l in rows)
    for n in (2, 8):
        for case in _load(f"mgpu_check_r02_n{n}.json"):
            assert case["world"] == n and case["max_abs_diff_vs_single_gpu"] < 1e-5
            assert case["top_samples_300_max_abs_diff_vs_single_gpu"] == 0.0
    log = open(os.path.join(P, "r02_pytest_gpu.log")).read()
    assert re.search(r"9\d passed", log)
