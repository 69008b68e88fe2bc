#!/bin/bash
# the bench lines of the other workloads / modes, for profiles/r02_bench_*.json
mkdir -p gpurun_out
run() { # tag, args
  S=$(date +%s); timeout 900 python bench.py $2 > gpurun_out/r02_bench_$1.log 2>&1; echo "$1 rc=$? wall=$(( $(date +%s) - S ))s $(tail -1 gpurun_out/r02_bench_$1.log | grep -o '"value": [0-9.e+]*' | head -1) $(tail -1 gpurun_out/r02_bench_$1.log | grep -o '"cpu_baseline": {"value": [0-9.e+]*')"
}
run samplernn "--workload samplernn"
run samplernn_bf16 "--workload samplernn --dtype bf16 --no-extras"
run features "--workload features"
run wavenet_bf16_b128 "--dtype bf16 --batch 128 --steps 1 --warmup 3 --no-extras --no-cpu-baseline"
run wavenet_bf16_b64 "--dtype bf16 --batch 64 --steps 1 --warmup 3 --no-extras --no-cpu-baseline"
run reference_samplernn "--impl reference --workload samplernn"
run reference_features "--impl reference --workload features"
