#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_samplernn_gpu.py tests/test_orchestration_gpu.py -m gpu -q --maxfail=5 > gpurun_out/pytest_sr.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_sr.log
tail -4 gpurun_out/pytest_sr.log
for B in 128 64 16; do
timeout 300 python bench.py --workload samplernn --batch $B --steps 1 --warmup 2 --no-cpu-baseline --no-extras > gpurun_out/r2_sr_b$B.log 2>&1
echo "sr b$B $(grep -o '"value": [0-9.]*' gpurun_out/r2_sr_b$B.log | head -1) $(grep -o '"p50_step_latency_us": [0-9.]*' gpurun_out/r2_sr_b$B.log) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_sr_b$B.log)"
done
