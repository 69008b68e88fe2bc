#!/bin/bash
# GPU-box recipe (through gpurun, 1 GPU): parity tests, WaveNet bench + in-kernel timeline, ncu launch list + full capture.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
MMK_WN_TRACE_T=3000 MMK_WN_TRACE_FILE=gpurun_out/wn_trace.txt timeout 300 python bench.py --seconds 0.5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_wn_trace.log 2>&1
tail -1 gpurun_out/bench_wn_trace.log
NCU="ncu --clock-control none"
timeout 600 $NCU --set full --import-source on -k regex:wavenet_chain -s 1 -c 1 -o gpurun_out/prof_wavenet_v2 -f \
    python bench.py --seconds 0.02 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_wavenet.log 2>&1
tail -2 gpurun_out/ncu_full_wavenet.log
ls -la gpurun_out
