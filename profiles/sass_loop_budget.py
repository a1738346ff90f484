#!/usr/bin/env python
"""Static issue-slot budget of the pass-1 loop of the headline kernel, from the SASS of the built library.

    python profiles/sass_loop_budget.py [path/to/libmppi_b200.so] > profiles/r02_pass1_sass_budget.md

Finds solve_kernel<Racing, inject=false, mode=fused>, takes its first long backward branch (the bounded pass-1
loop: one Philox4x32-10 call = 4 normals = 2 time steps per iteration) and counts instructions by opcode. With
the kernel issue/latency-bound (profiles/r01_solve_kernel_ncu.md: issue active 73 %, no pipe above 48 %), instructions per
step x steps x warps per scheduler / clock is the floor for pass 1 at this instruction mix.
"""
import collections
import os
import re
import subprocess
import sys

SYMBOL = "_ZN4mppi12solve_kernelINS_6RacingELb0ELi0ELi2ELb0EEEvNS_11SolveParamsE"  # <Racing, inject=false, kFused, SPT=2, staged maps>
SAMPLE_STEPS_PER_ITER = 4  # one Philox chunk per sample = 2 time steps, two samples per thread
GROUPS = [
    ("fp32 arithmetic (FADD/FMUL/FFMA/FMNMX/FSEL/FSETP)", ("FADD", "FMUL", "FFMA", "FMNMX", "FSEL", "FSETP", "HFMA2")),
    ("integer / logic (IMAD, LOP3, SHF, LEA, IADD3, VIADD, ISETP, MOV)",
     ("IMAD", "LOP3", "SHF", "LEA", "IADD3", "VIADD", "ISETP", "MOV", "VIMNMX", "SEL")),
    ("conversions (I2FP, F2I)", ("I2FP", "F2I", "I2F", "FRND")),
    ("MUFU (lg2, sqrt, sin, cos of Box-Muller)", ("MUFU",)),
    ("shared-memory loads (nominal, reference path, 2 map words)", ("LDS",)),
    ("constant-bank loads (LDCU/LDC of kernel params)", ("LDCU", "LDC")),
    ("control (BRA, BSSY, BSYNC)", ("BRA", "BSSY", "BSYNC")),
]


def loop_body(lib):
    """SASS instructions (predicates stripped) of the bounded pass-1 loop of the headline kernel, and its bounds."""
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", SYMBOL, lib], capture_output=True, text=True, check=True).stdout
    ins = [(int(m.group(1), 16), re.sub(r"^@!?U?P\d\s+", "", m.group(2)))
           for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", sass)]
    for addr, text in ins:
        m = re.search(r"BRA\s+0x([0-9a-f]+)", text)
        if m and addr - int(m.group(1), 16) > 0x1000:
            loop = (int(m.group(1), 16), addr)
            body = [t for a, t in ins if loop[0] <= a <= loop[1]]
            if sum(t.startswith(("FFMA2", "FMUL2", "FADD2")) for t in body) >= 40:  # the paired (packed fp32) loop
                return body, loop
    raise RuntimeError("pass-1 loop not found in " + lib)


DEFAULT_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mppi_playground_b200", "libmppi_b200.so")


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else DEFAULT_LIB
    body, loop = loop_body(lib)
    ops = collections.Counter(t.split()[0].split(".")[0] for t in body)
    steps_per_iter, T, warps_per_sched, mhz = 2, 80, 2, 1965.0  # an iteration = 2 time steps of 2 samples per lane
    per_step = len(body) / steps_per_iter
    per_sample_step = len(body) / SAMPLE_STEPS_PER_ITER
    packed = sum(ops[k] for k in ("FFMA2", "FMUL2", "FADD2"))
    print("# Pass-1 loop of `solve_kernel<Racing, SPT=2>`: static issue-slot budget (round 2, paired-sample loop)\n")
    print(f"`{os.path.basename(lib)}`, function `{SYMBOL}`, loop 0x{loop[0]:x}..0x{loop[1]:x}: **{len(body)} SASS "
          f"instructions per iteration** = {per_step:.1f} per warp and time step = **{per_sample_step:.1f} per "
          "sample-timestep** (every lane rolls two samples with packed fp32 - `FFMA2` / `FMUL2`, separate roundings "
          "per half, so parity-safe; one Philox call per sample feeds two steps). Round 1 (one sample per lane): 184.\n")
    print("| group | per iteration | per warp-step | share |\n|---|---|---|---|")
    seen = 0
    groups = [("packed fp32 (FFMA2 / FMUL2; an add is an fma by 1.0: ptxas would contract a plain packed mul+add)",
               ("FFMA2", "FMUL2", "FADD2"))] + GROUPS
    for name, keys in groups:
        n = sum(ops[k] for k in keys)
        seen += n
        print(f"| {name} | {n} | {n / steps_per_iter:.1f} | {100.0 * n / len(body):.0f} % |")
    print(f"| other | {len(body) - seen} | {(len(body) - seen) / steps_per_iter:.1f} | "
          f"{100.0 * (len(body) - seen) / len(body):.0f} % |\n")
    print("Opcode counts: " + ", ".join(f"{k} {v}" for k, v in ops.most_common()) + "\n")
    slots = per_step * T * warps_per_sched
    pipe = (len(body) + packed) / steps_per_iter * T * warps_per_sched  # a packed instruction holds the FMA pipe 2 cycles
    print(f"Issue floor: {per_step:.1f} instr per warp-step x T={T} x {warps_per_sched} warps per scheduler (K=65536: 1024 "
          f"paired warps on 128 worker SMs, 256-thread blocks) = {slots:,.0f} issue slots per scheduler = "
          f"**{slots / mhz:.1f} us at {mhz:.0f} MHz**; counting the second FMA-pipe cycle of the {packed} packed "
          f"instructions: {pipe:,.0f} cycles = {pipe / mhz:.1f} us. Measured pass 1 (in-kernel `%globaltimer`, "
          "`profiles/r02_block_trace.txt`, median block): 35.3 us, i.e. each warp sustains ~0.32 instructions per "
          "cycle and a scheduler with its two warps ~0.64 - the loop is bound by the dependent-issue latency of one "
          "warp's chain, not by issue slots (`profiles/r02_solve_kernel_ncu.md`: issue active 54 %, FMA pipe 48 %, "
          "ALU 41 %, XU 23 %; stalls: wait, math-pipe throttle, dispatch). More warps per scheduler are not available: "
          "K / 64 = 1024 warps is all the work there is, and the one-sample-per-lane form (184 per step, 4 warps per "
          "scheduler) is issue bound at 981 cycles per step against 852 here (csrc/mppi_engine.cu:paired_loop_pays).\n")
    print("What the count is made of: the arithmetic follows the reference's fp32 operation order without FMA "
          "contraction (`-fmad=false`; every FFMA is an explicit `fmaf` of the `sinf`/`cosf`/`tanf` polynomials or of "
          "the exact-division step), the sampler (two Philox4x32-10 calls + Box-Muller per iteration, ~135 integer "
          "instructions) and the per-lane pieces that cannot be packed (clamps, selects, float<->int conversions, "
          "the two grid-word loads per sample).")


if __name__ == "__main__":
    main()
