#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_samplernn_gpu.py -m gpu -q -s -k "tensor_core" > gpurun_out/pytest_srtc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_srtc.log
grep -v "^$" gpurun_out/pytest_srtc.log | tail -12 | cut -c1-250
for B in 128 16; do
timeout 300 python bench.py --workload samplernn --dtype bf16 --batch $B --steps 1 --warmup 2 --no-cpu-baseline --no-extras > gpurun_out/r2_srtc_b$B.log 2>&1
echo "sr bf16 b$B $(grep -o '"value": [0-9.]*' gpurun_out/r2_srtc_b$B.log | head -1) $(grep -o '"p50_step_latency_us": [0-9.]*' gpurun_out/r2_srtc_b$B.log) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_srtc_b$B.log) $(tail -1 gpurun_out/r2_srtc_b$B.log | cut -c1-300 | grep -v metric)"
done
