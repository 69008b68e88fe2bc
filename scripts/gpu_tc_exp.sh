#!/bin/bash
mkdir -p gpurun_out
for E in 0 1 2 3; do
MMK_TC_EXP=$E MMK_TC_TRACE_T=20000 MMK_TC_TRACE_FILE=gpurun_out/tc_trace_e$E.txt timeout 200 python bench.py --dtype bf16 --batch 128 --seconds 0.3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc_e$E.log 2>&1
echo "exp=$E $(grep -o '"p50_step_latency_us": [0-9.]*' gpurun_out/bench_tc_e$E.log)"
done
