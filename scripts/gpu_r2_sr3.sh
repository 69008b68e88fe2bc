#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_samplernn_gpu.py -m gpu -q -x -k "lane_major or default_geometry or s3_full" > gpurun_out/pytest_sr3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_sr3.log
tail -15 gpurun_out/pytest_sr3.log
for B in 128 64 16; do
MMK_SR_DEBUG=1 timeout 300 python bench.py --workload samplernn --batch $B --seconds 1 --steps 1 --warmup 2 --no-cpu-baseline --no-extras > gpurun_out/r2_sr3_b$B.log 2>&1
echo "sr b$B $(grep -o '"value": [0-9.]*' gpurun_out/r2_sr3_b$B.log | head -1) $(grep -o '"p50_step_latency_us": [0-9.]*' gpurun_out/r2_sr3_b$B.log) $(grep 'sr2\]' gpurun_out/r2_sr3_b$B.log | tail -2)"
done
