#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_wavenet_gpu.py -m gpu -q -x -k "6- or -6] or golden or w30 or stepwise or errors or default_geometry" > gpurun_out/pytest_wn6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_wn6.log
tail -4 gpurun_out/pytest_wn6.log
b() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --seconds 0.5 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_$name.log 2>&1
  echo "$name $(grep -o '"value": [0-9.]*' gpurun_out/r2_$name.log | head -1) $(grep -o '"p50_step_latency_us": [0-9.]*' gpurun_out/r2_$name.log) $(grep -o '"sm_used": [0-9]*' gpurun_out/r2_$name.log) $(tail -1 gpurun_out/r2_$name.log | cut -c1-200 | grep -v metric)"
}
b wn6 X=1
b wn6_trace MMK_WN_TRACE_T=3000 MMK_WN_TRACE_FILE=gpurun_out/wn6_trace.txt
env timeout 300 python bench.py --seconds 0.5 --steps 1 --warmup 3 --no-cpu-baseline --no-extras --batch 128 > gpurun_out/r2_wn6_b128.log 2>&1
echo "b128 $(grep -o '"value": [0-9.]*' gpurun_out/r2_wn6_b128.log | head -1) $(grep -o '"p50_step_latency_us": [0-9.]*' gpurun_out/r2_wn6_b128.log)"
env timeout 300 python bench.py --seconds 0.5 --steps 1 --warmup 3 --no-cpu-baseline --no-extras --batch 16 > gpurun_out/r2_wn6_b16.log 2>&1
echo "b16 $(grep -o '"value": [0-9.]*' gpurun_out/r2_wn6_b16.log | head -1) $(grep -o '"p50_step_latency_us": [0-9.]*' gpurun_out/r2_wn6_b16.log)"
[ "$1" = "ncu" ] && bash scripts/gpu_r2_ncu_wn6.sh
