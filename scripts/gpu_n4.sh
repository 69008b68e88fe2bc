#!/bin/bash
# N-GPU check (N = $1, default 4) of the default bench and the cfg-4 style bf16 run — short horizons.
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR bench.py --gpus $N --seconds 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/n${N}_wavenet.log 2>&1; tail -1 gpurun_out/n${N}_wavenet.log | cut -c1-330
timeout 300 $TR bench.py --gpus $N --dtype bf16 --batch 128 --seconds 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/n${N}_wavenet_bf16_b128.log 2>&1; tail -1 gpurun_out/n${N}_wavenet_bf16_b128.log | cut -c1-330
timeout 300 $TR bench.py --gpus $N --workload samplernn --seconds 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/n${N}_samplernn.log 2>&1; tail -1 gpurun_out/n${N}_samplernn.log | cut -c1-330
timeout 300 $TR bench.py --gpus $N --workload features --steps 3 --warmup 3 > gpurun_out/n${N}_features.log 2>&1; tail -1 gpurun_out/n${N}_features.log | cut -c1-330
