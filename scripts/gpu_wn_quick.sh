#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_wavenet_gpu.py -m gpu -q -x > gpurun_out/pytest_wn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_wn.log
tail -4 gpurun_out/pytest_wn.log
timeout 300 python bench.py --seconds 0.5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_wn_quick.log 2>&1
tail -1 gpurun_out/bench_wn_quick.log | cut -c1-120; grep -o '"p50_step_latency_us": [0-9.]*' gpurun_out/bench_wn_quick.log
MMK_WN_TRACE_T=3000 MMK_WN_TRACE_FILE=gpurun_out/wn_trace.txt timeout 300 python bench.py --seconds 0.5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_wn_trace.log 2>&1
