#!/usr/bin/env python
"""Key figures of every launch in an ncu --set full report -> markdown (used for the control-step epilogue kernels):
   python profiles/summarize_ncu_kernels.py gpurun_out/prof_epilogue_r02h.ncu-rep "title" > profiles/r02_epilogue_ncu.md"""
import csv
import io
import subprocess
import sys

rep, title = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
print(f"# {title}\n")
print("`ncu --set full --clock-control none` (cold caches, serialised replays: these explain the CUDA-event numbers, "
      "they are not bench values).\n")
for li, r in enumerate(rows[2:]):
    print(f"## launch {li}: `{r[hdr.index('Kernel Name')][:110]}`\n")
    print("| metric | value | unit |\n|---|---|---|")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"| `{w}` | {r[i]} | {units[i]} |")
    st = []
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and r[i]:
            try:
                st.append((float(r[i].replace(",", "")), h.replace("smsp__average_warps_issue_stalled_", "").replace(
                    "_per_issue_active.ratio", "")))
            except ValueError:
                pass
    print("\nwarp stall reasons (warps per issue-active cycle): " +
          ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:6]) + "\n")
