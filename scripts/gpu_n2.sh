#!/bin/bash
# 2-GPU check of the three bench workloads (short horizons) — run with `gpurun --gpus 2`.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 2 --seconds 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/n2_wavenet.log 2>&1; tail -1 gpurun_out/n2_wavenet.log | cut -c1-400
timeout 300 $TR bench.py --gpus 2 --workload samplernn --seconds 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/n2_samplernn.log 2>&1; tail -1 gpurun_out/n2_samplernn.log | cut -c1-400
timeout 300 $TR bench.py --gpus 2 --workload features --steps 3 --warmup 3 > gpurun_out/n2_features.log 2>&1; tail -1 gpurun_out/n2_features.log | cut -c1-400
timeout 200 $TR bench.py --gpus 2 --impl reference --steps 1 --warmup 1 > gpurun_out/n2_reference.log 2>&1; tail -1 gpurun_out/n2_reference.log | cut -c1-200
timeout 300 python bench.py --batch 128 --seconds 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/n1_wavenet_b128.log 2>&1; tail -1 gpurun_out/n1_wavenet_b128.log | cut -c1-400
