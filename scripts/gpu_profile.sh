#!/bin/bash
# GPU-box recipe: launch list + full ncu capture of the dominant kernels (run through gpurun, 1 GPU).
set -x
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches_wavenet.csv \
    python bench.py --seconds 0.1 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_wavenet.log 2>&1
$NCU --set full --import-source on -k regex:wavenet_pipe -s 1 -c 1 -o gpurun_out/prof_wavenet -f \
    python bench.py --seconds 0.02 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_wavenet.log 2>&1
python bench.py --workload features --steps 3 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_features.log
$NCU --metrics gpu__time_duration.sum -c 100 --csv --log-file gpurun_out/launches_features.csv \
    python bench.py --workload features --batch 360 --steps 2 --warmup 1 > gpurun_out/ncu_bench_features.log 2>&1
$NCU --set full --import-source on -k regex:"stft_mag_mel|mulaw_compress" -s 2 -c 2 -o gpurun_out/prof_features -f \
    python bench.py --workload features --batch 360 --steps 1 --warmup 1 > gpurun_out/ncu_full_features.log 2>&1
ls -la gpurun_out
