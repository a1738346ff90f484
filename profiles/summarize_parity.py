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
out += ["", "Other GPU checks in the same suite (62 tests): exhaustive self-tests (`tan_quarter == tanf` on |x|<=0.78,",
        "`sincos_bounded == sincosf` on |x|<=4, `wrap_angle_bounded` / `wrap_angle_nonneg == wrap_angle`, exact cell / wheelbase",
        "division: 0 mismatches over all fp32 inputs of each range), bounded loop == general loop bit for bit, block-parallel",
        "tail rollout == serial `step()` bit for bit, host-buffer solve == device-buffer solve bit for bit, sharded (2 and 3",
        "shards, staged and fused peer exchange) == unsharded (costs bit-equal on the first solve, sequences to 2e-6), device",
        "reference path == host twin bit for bit, full-size properties at BASELINE configs 3/4/5 (oracle on a 4096-sample",
        "subset, fp64 recomputation of the weighted mean to 2e-6, `state_seq` == rollout of `action_seq`, determinism), sampler",
        "statistics (moments, KS, independence), Philox4x32-10 known answers. Multi-GPU: `profiles/mgpu_check_r01_n{2,8}.json`."]
open(f"profiles/parity_{rnd}.md", "w").write("\n".join(out) + "\n")
print(len(agg), "cases")
