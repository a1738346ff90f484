#!/bin/bash
# Profiling recipe (run under gpurun on ONE B200; see /opt/skills/guides/B200_PROFILING.md).
#   bash profiles/run_ncu.sh <tag>        e.g. r01a
# Writes gpurun_out/launches_<tag>.csv (every launch with its device time) and
# gpurun_out/prof_<tag>.ncu-rep (--set full capture of the fused solve kernel).
set -u
TAG=${1:-r01}
CMD="python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 12 -c 2 \
    -o gpurun_out/prof_${TAG} -f $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/ | tail -8
