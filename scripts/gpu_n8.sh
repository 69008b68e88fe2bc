#!/bin/bash
# 8-GPU check: default bench (cfg 2 shape per GPU) and cfg 4 (1024 prompts = 128 per GPU, bf16 tensor-core kernel).
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544"
timeout 200 $TR bench.py --gpus 8 --seconds 0.5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/n8_wavenet.log 2>&1; tail -1 gpurun_out/n8_wavenet.log | cut -c1-330
timeout 200 $TR bench.py --gpus 8 --dtype bf16 --batch 128 --seconds 0.5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/n8_wavenet_bf16_b128.log 2>&1; tail -1 gpurun_out/n8_wavenet_bf16_b128.log | cut -c1-330
