set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02h_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02h_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02h_smoke.log 2>&1; tail -1 gpurun_out/r02h_smoke.log
timeout 400 python bench.py > gpurun_out/r02h_bench_n1.json 2> gpurun_out/r02h_bench_n1.err; cut -c1-200 gpurun_out/r02h_bench_n1.json
timeout 200 python bench.py --config c5 --steps 300 --warmup 10 --no-cpu-baseline > gpurun_out/r02h_bench_c5_n1.json 2> gpurun_out/r02h_bench_c5_n1.err; cut -c1-200 gpurun_out/r02h_bench_c5_n1.json
timeout 300 python bench.py --impl reference --steps 4 --warmup 2 > gpurun_out/r02h_bench_reference.json 2> gpurun_out/r02h_bench_reference.err; cut -c1-300 gpurun_out/r02h_bench_reference.json
timeout 100 python tools/block_trace.py > gpurun_out/r02h_block_trace.txt 2>&1
timeout 100 python tools/block_trace.py --flush > gpurun_out/r02h_block_trace_flush.txt 2>&1
timeout 100 python tools/time_configs.py > gpurun_out/r02h_time_configs.txt 2>&1; cp gpurun_out/time_configs.json gpurun_out/r02h_time_configs.json
timeout 100 python tools/time_epilogue.py > gpurun_out/r02h_time_epilogue.json 2>&1
timeout 300 bash profiles/run_ncu.sh r02h > /dev/null 2>&1
timeout 250 ncu --set full --clock-control none --import-source on -k regex:"epilogue|reroll|topn" -s 4 -c 4 -o gpurun_out/prof_epilogue_r02h -f python tools/epilogue_once.py > gpurun_out/ncu_epilogue_r02h.log 2>&1
ls -la gpurun_out | grep r02h
