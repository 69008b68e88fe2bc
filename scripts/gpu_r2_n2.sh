#!/bin/bash
# 2 GPUs: the NCCL sharded-parity test and a short 2-rank bench line (fp32 default, with the side configs)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharding_nccl_gpu.py -m gpu -q > gpurun_out/pytest_nccl.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_nccl.log
tail -5 gpurun_out/pytest_nccl.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --seconds 1 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_n2_wavenet.log 2>&1
tail -1 gpurun_out/r2_n2_wavenet.log | cut -c1-1800
