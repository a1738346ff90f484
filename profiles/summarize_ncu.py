#!/usr/bin/env python
"""Turn gpurun_out/prof_<tag>.ncu-rep + launches_<tag>.csv into the committed summaries
   profiles/r01_solve_kernel_ncu.md, profiles/r01_launches_ncu.md and profiles/ncu_traffic.json (the
   DRAM bytes and pipe figures bench.py reports as roofline.traffic / ncu_pipes). Run in the build container:
   python profiles/summarize_ncu.py <tag> [<round prefix, e.g. r02>]"""
import collections
import csv
import io
import json
import subprocess
import sys

tag = sys.argv[1]
rnd = sys.argv[2] if len(sys.argv) > 2 else "r01"  # prefix of the committed summaries
rep, launches = f"gpurun_out/prof_{tag}.ncu-rep", f"gpurun_out/launches_{tag}.csv"

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
out = [f"# ncu --set full, {tag}: solve_kernel<Racing, inject=false, kFused>  (bench workload K=65536 T=80)", "",
       "Captured with `profiles/run_ncu.sh` (`--clock-control none`, launches inside bench.py's timed region).",
       "Numbers under a profiler are not bench values; they explain the bench value.", ""]
for li, r in enumerate(rows[2:]):
    out += [f"## captured launch {li}", "", "| metric | value | unit |", "|---|---|---|"]
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            out.append(f"| `{w}` | {r[i]} | {units[i]} |")
    stalls = []
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            if v > 0.03:
                stalls.append((v, h.replace("smsp__average_warps_issue_stalled_", "").replace(
                    "_per_issue_active.ratio", "")))
    out += ["", "warp stall reasons (warps per issue-active cycle): " +
            ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)), ""]
# dynamic opcode mix from the source page
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(srows) if r and r[0] == "Address"][0]
sh = srows[hi]
isrc, iex = sh.index("Source"), sh.index("Instructions Executed")
ops, tot = collections.Counter(), 0
import re
for r in srows[hi + 1:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) <= iex or not r[0].startswith("0x"):
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[isrc])
    n = int(r[iex])
    ops[m.group(2) if m else "?"] += n
    tot += n
out += ["## executed warp instructions by opcode (launch 0)", "", f"total {tot} "
        f"(= {tot * 32 / (65536 * 80):.0f} thread instructions per sample-timestep incl. pass 2 / reductions / tail)", "",
        "| opcode | warp instructions | share |", "|---|---|---|"]
for op, n in ops.most_common(24):
    out.append(f"| {op} | {n} | {100 * n / tot:.1f}% |")
sass = subprocess.run(["cuobjdump", "-sass", "mppi_playground_b200/libmppi_b200.so"], capture_output=True, text=True).stdout
out += ["", "## SASS evidence", "",
        f"`UBLKCP` (cp.async.bulk, TMA engine) occurrences in libmppi_b200.so: {sass.count('UBLKCP')}; "
        f"`SYNCS` (mbarrier): {sass.count('SYNCS')}; tensor-core mnemonics (`UTC*MMA`, `HMMA`): "
        f"{sass.count('UTCHMMA') + sass.count('HMMA')} (none by design: the path has no contraction)."]
open(f"profiles/{rnd}_solve_kernel_ncu.md", "w").write("\n".join(out) + "\n")


def _num(name, row=rows[2]):
    v = float(row[hdr.index(name)].replace(",", ""))
    u = units[hdr.index(name)]
    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)


rd, wr = _num("dram__bytes_read.sum"), _num("dram__bytes_write.sum")
json.dump({"dram_bytes_per_launch": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr),
           "source": f"profiles/{rnd}_solve_kernel_ncu.md (ncu --set full capture {tag}, dram__bytes_read.sum + "
                     "dram__bytes_write.sum, launch 0)",
           "pipes": {"issue_active_pct": _num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                     "fma_pipe_pct": _num("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                     "alu_pipe_pct": _num("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
                     "xu_pipe_pct": _num("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                     "tensor_pipe_pct": _num("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                     "warp_instructions_per_launch": _num("smsp__inst_executed.sum"),
                     "registers_per_thread": _num("launch__registers_per_thread")}},
          open("profiles/ncu_traffic.json", "w"), indent=1)

lrows = [r for r in csv.reader(open(launches)) if len(r) > 10]
lh = lrows[0]
ki, vi = lh.index("Kernel Name"), lh.index("Metric Value")
agg = collections.defaultdict(list)
for r in lrows[1:]:
    agg[r[ki]].append(float(r[vi].replace(",", "")))
tot_t = sum(sum(v) for v in agg.values())
lo = [f"# ncu launch list, {tag}: `python bench.py --steps 6 --warmup 3 --no-cpu-baseline`", "",
      "`ncu --metrics gpu__time_duration.sum --clock-control none` - every launch of the command, device time "
      "(cold-cache, serialised: compare SHARES).", "", "| kernel | launches | mean us | total us | share |",
      "|---|---|---|---|---|"]
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    lo.append(f"| `{k[:100]}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {sum(v) / 1e3:.1f} | {100 * sum(v) / tot_t:.1f}% |")
lo += ["", "The memset (`FillFunctor`) launches are bench.py's L2 flush between timed steps, the `pack_map` / "
       "`check_fastdiv` launches are one-time set-up (`mppi_set_map`). Within a solve the only kernel is `solve_kernel` (100%); "
       "`control_epilogue_kernel` / `topn_select_kernel` / `reroll_winners_kernel` are bench.py's control-step epilogue timing "
       "(details.control_step_epilogue_us), outside the solve metric."]
open(f"profiles/{rnd}_launches_ncu.md", "w").write("\n".join(lo) + "\n")
print("ok")
