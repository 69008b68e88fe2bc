#!/bin/bash
# bf16 tensor-core WaveNet kernel: tests, batch sweep, timeline, ncu capture (1 GPU).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_wavenet_tc_gpu.py -m gpu -q -x 2>&1 | tail -2
for B in 64 128 1024 4096 18944; do
  timeout 300 python bench.py --dtype bf16 --batch $B --seconds 0.5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc_b$B.log 2>&1
  tail -1 gpurun_out/bench_tc_b$B.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('B=$B', round(d['value']), 'samples/s', 'ms/step', round(d['ms_per_step'],1), 'p50us', d['p50_step_latency_us'], 'tensor frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']), d['clocks'])
except Exception as e: print('B=$B failed', e)
"
done
MMK_TC_TRACE_T=20000 MMK_TC_TRACE_FILE=gpurun_out/tc_trace.txt timeout 200 python bench.py --dtype bf16 --batch 128 --seconds 0.3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc_trace.log 2>&1
timeout 600 ncu --clock-control none --set full --import-source on -k regex:wavenet_tc -s 1 -c 1 -o gpurun_out/prof_wavenet_tc -f \
    python bench.py --dtype bf16 --batch 18944 --seconds 0.02 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_wavenet_tc.log 2>&1
tail -2 gpurun_out/ncu_full_wavenet_tc.log | cut -c1-300
timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum -c 200 --csv --log-file gpurun_out/launches_wavenet_tc.csv \
    python bench.py --dtype bf16 --batch 18944 --seconds 0.05 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_wavenet_tc.log 2>&1
