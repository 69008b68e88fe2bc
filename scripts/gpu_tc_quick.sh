#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_wavenet_tc_gpu.py -m gpu -q -x 2>&1 | tail -3
for B in 128 ${TC_EXTRA_B}; do
timeout 300 python bench.py --dtype bf16 --batch $B --seconds 0.5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc_b$B.log 2>&1
tail -1 gpurun_out/bench_tc_b$B.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('B=$B', round(d['value']), 'samples/s', 'ms/step', round(d['ms_per_step'],1), 'p50us', d['p50_step_latency_us'], 'tensor frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']))
except Exception as e: print('failed', e)
"
done
if [ -n "$TC_NCU" ]; then
timeout 600 ncu --clock-control none --set full --import-source on -k regex:wavenet_tc -s 1 -c 1 -o gpurun_out/prof_wavenet_tc -f \
    python bench.py --dtype bf16 --batch 128 --seconds 0.02 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_wavenet_tc.log 2>&1
fi
