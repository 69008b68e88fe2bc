#!/bin/bash
scripts/gpu_r2_ncu.sh wn7 wavenet7 --dtype bf16 --batch 128 --seconds 0.05
scripts/gpu_r2_ncu.sh srtc samplernn_cluster --workload samplernn --dtype bf16 --batch 128 --seconds 0.05
scripts/gpu_r2_ncu.sh sr samplernn_cluster --workload samplernn --batch 128 --seconds 0.05
