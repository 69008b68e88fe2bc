#!/bin/bash
mkdir -p gpurun_out
MMK_TC_TRACE_T=3000 MMK_TC_TRACE_FILE=gpurun_out/wn7_trace.txt timeout 300 python bench.py --dtype bf16 --batch 64 --seconds 0.5 --steps 1 --warmup 2 --no-cpu-baseline --no-extras > gpurun_out/r2_tc_trace.log 2>&1
grep -o '"p50_step_latency_us": [0-9.]*' gpurun_out/r2_tc_trace.log
