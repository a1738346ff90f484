#!/bin/bash
# Multi-GPU measurement batch (run under `gpurun --gpus N -- bash tools/final_multi_gpu.sh N TAG`): bench.py at the
# headline config (c4) and at BASELINE configs[4] (c5), the in-kernel exchange timeline of both, the N-GPU
# correctness check. Outputs under gpurun_out/.
set -u
N=${1:-2}
TAG=${2:-r02h}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
for c in c4 c5; do
  steps=1000; [ "$c" = c5 ] && steps=300
  timeout 300 $RUN --master-port 29511 bench.py --gpus $N --steps $steps --warmup 10 --config $c \
      > gpurun_out/${TAG}_bench_${c}_n${N}.json 2> gpurun_out/${TAG}_bench_${c}_n${N}.err
  cut -c1-230 gpurun_out/${TAG}_bench_${c}_n${N}.json
  timeout 150 $RUN --master-port 29512 tools/exchange_trace.py --config $c \
      > gpurun_out/${TAG}_exchange_${c}_n${N}.json 2> gpurun_out/${TAG}_exchange_${c}_n${N}.err
  python -c "
import json; d=json.load(open('gpurun_out/${TAG}_exchange_${c}_n${N}.json')); print(d['exchange_total_us']); [print(r) for r in d['ranks'][:3]]"
done
timeout 200 $RUN --master-port 29513 tools/mgpu_check.py > gpurun_out/${TAG}_mgpu_check_n${N}.log 2>&1
tail -4 gpurun_out/${TAG}_mgpu_check_n${N}.log
ls gpurun_out | grep -i mgpu_check | tail -3
