#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_samplernn_gpu.py tests/test_orchestration_gpu.py -m gpu -q --maxfail=5 > gpurun_out/pytest_sr.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_sr.log
tail -4 gpurun_out/pytest_sr.log
for D in bf16 f32; do
MMK_SR_DEBUG=1 timeout 300 python bench.py --workload samplernn --dtype $D --batch 128 --steps 1 --warmup 2 --no-cpu-baseline --no-extras > gpurun_out/r2_sr_${D}_b128.log 2>&1
echo "sr $D b128 $(grep -o '"value": [0-9.]*' gpurun_out/r2_sr_${D}_b128.log | head -1) $(grep -o '"p50_step_latency_us": [0-9.]*' gpurun_out/r2_sr_${D}_b128.log) $(grep 'head Mcycles' gpurun_out/r2_sr_${D}_b128.log | tail -1)"
done
