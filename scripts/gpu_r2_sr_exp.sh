#!/bin/bash
# lane-major SampleRNN engine: parity tests, then section timers (SM cycles of CTA 0) and whole-run times
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_samplernn_gpu.py -m gpu -q -x -k "lane_major or default_geometry or s3_full" > gpurun_out/pytest_sr3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_sr3.log
tail -4 gpurun_out/pytest_sr3.log
run() { echo "$* : $(env "$@" timeout 200 python bench.py --workload samplernn --batch $B --seconds 1 --steps 1 --warmup 1 --no-cpu-baseline --no-extras 2>&1 | grep 'Mcycles\|"ms_per_step"' | sed 's/.*"ms_per_step": \([0-9.]*\).*/ms \1/' | tail -2 | tr '\n' ' ')"; }
for B in 128 64 16; do
echo "== B=$B"
run MMK_SR_DEBUG=1
done
