#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_wavenet_tc_gpu.py -m gpu -q --maxfail=6 -s > gpurun_out/pytest_tc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tc.log
grep -v "^$" gpurun_out/pytest_tc.log | tail -40 | cut -c1-220
for B in 128 64; do
timeout 300 python bench.py --dtype bf16 --batch $B --seconds 0.5 --steps 1 --warmup 2 --no-cpu-baseline --no-extras > gpurun_out/r2_tc_b$B.log 2>&1
echo "tc b$B $(grep -o '"value": [0-9.]*' gpurun_out/r2_tc_b$B.log | head -1) $(grep -o '"p50_step_latency_us": [0-9.]*' gpurun_out/r2_tc_b$B.log) $(grep -o '"launch": {[^}]*}' gpurun_out/r2_tc_b$B.log) $(tail -1 gpurun_out/r2_tc_b$B.log | cut -c1-200 | grep -v metric)"
done
bash scripts/gpu_r2_tct.sh
