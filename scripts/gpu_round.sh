#!/bin/bash
# GPU-box recipe (through gpurun, 1 GPU): parity tests, WaveNet bench + in-kernel timeline, ncu launch list + full capture.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
MMK_WN_TRACE_T=3000 MMK_WN_TRACE_FILE=gpurun_out/wn_trace.txt timeout 300 python bench.py --seconds 0.5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_wn_trace.log 2>&1
tail -1 gpurun_out/bench_wn_trace.log
