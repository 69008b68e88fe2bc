#!/bin/bash
mkdir -p gpurun_out
for ST in 5 6 7 8 9; do for CS in 16 8; do
MMK_WN_STAGES=$ST MMK_WN_CLUSTER=$CS timeout 120 python bench.py --seconds 0.5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/wn_sweep_${ST}_${CS}.log 2>&1
echo "stages=$ST cluster=$CS $(grep -o '"value": [0-9.]*' gpurun_out/wn_sweep_${ST}_${CS}.log | head -1) $(grep -o '"p50_step_latency_us": [0-9.]*' gpurun_out/wn_sweep_${ST}_${CS}.log) $(grep -o '"sm_used": [0-9]*' gpurun_out/wn_sweep_${ST}_${CS}.log) $(tail -1 gpurun_out/wn_sweep_${ST}_${CS}.log | cut -c1-80 | grep -v metric)"
done; done
