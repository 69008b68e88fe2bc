#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --seconds 0.5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_wn_quick.log 2>&1
tail -1 gpurun_out/bench_wn_quick.log | cut -c1-200
MMK_WN_TRACE_T=3000 MMK_WN_TRACE_FILE=gpurun_out/wn_trace.txt timeout 300 python bench.py --seconds 0.5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_wn_trace.log 2>&1
timeout 900 ncu --clock-control none --set full --import-source on -k regex:wavenet_warp -s 1 -c 1 -o gpurun_out/prof_wavenet_v3 -f \
    python bench.py --seconds 0.01 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_wavenet.log 2>&1
tail -2 gpurun_out/ncu_full_wavenet.log | cut -c1-200
