#!/usr/bin/env python
"""gpurun_out/parity_report.jsonl (written by tests/test_gpu_parity.py on the GPU box) -> profiles/parity_r01.md."""
import collections
import json
import sys

rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"

rows = [json.loads(l) for l in open("gpurun_out/parity_report.jsonl")]
agg = collections.OrderedDict()
for r in rows:
    a = agg.setdefault(r["test"], {"n": 0, "cost": 0.0, "flip": 0.0, "act": 0.0, "st": 0.0, "lam": 0.0})
    a["n"] += 1
    for k, src in (("cost", "cost_rel_max"), ("flip", "cost_flip_frac"), ("act", "action_err"), ("st", "state_err"),
                   ("lam", "lam_rel")):
        a[k] = max(a[k], float(r[src]))
out = [f"# Parity measured on B200 ({rnd})", "",
       "Source: `gpurun_out/parity_report.jsonl`, written by `tests/test_gpu_parity.py` on the GPU box (the last run of",
       "every test; regenerate with `python profiles/summarize_parity.py`). `golden/*`: engine fed the reference's recorded",
       "noise vs the reference's recorded outputs (`tests/golden/*.npz`, from the live reference). `native/*`: in-kernel",
       "Philox noise read back and given to the CPU oracle, closed loop. `edge/*`: odd shapes (T, K not multiples of the",
       "chunk / warp / block sizes). Worst value over the solves of each case; bars in `tests/engine_util.py`.", "",
       "| case | solves | cost rel (max) | cell-flip fraction | action_seq abs | state_seq abs | lambda rel |",
       "|---|---|---|---|---|---|---|"]
for name in sorted(agg):
    a = agg[name]
    out.append(f"| {name} | {a['n']} | {a['cost']:.1e} | {a['flip']:.1e} | {a['act']:.1e} | {a['st']:.1e} | {a['lam']:.1e} |")
out += ["", "Other GPU checks in the same suite (98 tests in round 2, `profiles/r02_pytest_gpu.log`): exhaustive self-tests",
        "(`tan_quarter == tanf` on |x|<=0.78, `sincos_bounded == sincosf` on |x|<=4, the `wrap_angle_*` family == `wrap_angle`, packed",
        "f32x2 forms == scalar forms, exact cell / wheelbase division: 0 mismatches over all fp32 inputs of each range), bounded and",
        "paired loops == general loop bit for bit, block-parallel tail rollout == serial `step()` bit for bit, host-buffer solve ==",
        "device-buffer solve bit for bit, sharded (2 and 3 shards, staged and fused peer exchange, peer barrier) == unsharded, fused",
        "exchange timeout -> exception, device reference path == host twin bit for bit, drop-in objects shaped like the reference's",
        "(`tests/test_gpu_dropin.py`), full-size golden cases recorded from the live reference (`golden-full/*`), all-K oracle",
        "comparison at every BASELINE size (`fullsize/*`), control-step epilogue vs the reference's recorded loop and probes, top-n",
        "select == stable sort (ids and costs bit for bit, K up to 1 048 576), sharded top samples == unsharded, device rasteriser ==",
        "the reference's grids bit for bit, 2000 x 2000 grids through the global-memory path vs the oracle, sampler statistics,",
        "Philox4x32-10 known answers. Multi-GPU: `profiles/mgpu_check_r02_n{2,8}.json`."]
open(f"profiles/parity_{rnd}.md", "w").write("\n".join(out) + "\n")
print(len(agg), "cases")
