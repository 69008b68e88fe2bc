#!/bin/bash
# 8 GPUs: one short bench line of the default workload with the side configs (weak scaling, one gather)
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --seconds 1 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_n${N}_wavenet.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/r2_n${N}_wavenet.log | cut -c1-400; tail -1 gpurun_out/r2_n${N}_wavenet.log | grep -o '"other_configs".*' | cut -c1-1500
