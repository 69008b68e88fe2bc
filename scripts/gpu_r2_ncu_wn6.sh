#!/bin/bash
# round 2: ncu full-set capture (with source) of the layer-pipelined WaveNet kernel on a short horizon
mkdir -p gpurun_out
timeout 900 ncu --clock-control none --set full --import-source on -k regex:wavenet6 -s 1 -c 1 -o gpurun_out/prof_wn6 -f \
    python bench.py --seconds 0.05 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_wn6.log 2>&1
tail -2 gpurun_out/ncu_full_wn6.log | cut -c1-300
ls -la gpurun_out/prof_wn6.ncu-rep
