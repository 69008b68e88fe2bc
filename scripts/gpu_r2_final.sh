#!/bin/bash
# full GPU suite + smoke + the default bench lines (own arm and reference arm)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log | cut -c1-600
S=$(date +%s); timeout 900 python bench.py > gpurun_out/r02_bench_default.log 2>&1; echo "bench rc=$? wall=$(( $(date +%s) - S ))s"
tail -1 gpurun_out/r02_bench_default.log | cut -c1-3000
S=$(date +%s); timeout 900 python bench.py --impl reference > gpurun_out/r02_bench_reference.log 2>&1; echo "reference rc=$? wall=$(( $(date +%s) - S ))s"
tail -1 gpurun_out/r02_bench_reference.log | cut -c1-800
