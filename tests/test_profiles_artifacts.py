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
    t = _load("ncu_traffic.json")
    m = re.search(r"dram__bytes_read.sum` \| ([0-9.]+) \| Kbyte", md)
    assert m and abs(float(m.group(1)) * 1e3 - t["dram_read"]) < 1.0
    assert t["pipes"]["tensor_pipe_pct"] == 0.0 and 50 < t["pipes"]["issue_active_pct"] <= 100
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
