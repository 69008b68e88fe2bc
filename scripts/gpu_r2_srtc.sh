#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_samplernn_gpu.py -m gpu -q -x -s -k "tensor_core" > gpurun_out/pytest_srtc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_srtc.log
grep -v "^$" gpurun_out/pytest_srtc.log | tail -30 | cut -c1-250
