#!/bin/bash
# GPU-box recipe (through gpurun, 1 GPU): parity tests, the three bench workloads, ncu launch lists and full captures.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
timeout 600 python bench.py --steps 2 --warmup 3 > $O/bench_wavenet.log 2>&1; tail -1 $O/bench_wavenet.log | cut -c1-300
timeout 600 python bench.py --workload samplernn --steps 2 --warmup 3 > $O/bench_samplernn.log 2>&1; tail -1 $O/bench_samplernn.log | cut -c1-300
timeout 300 python bench.py --workload features --steps 3 --warmup 3 > $O/bench_features.log 2>&1; tail -1 $O/bench_features.log | cut -c1-300
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_wavenet.log 2>&1; tail -1 $O/bench_reference_wavenet.log | cut -c1-200
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_wavenet.csv \
    python bench.py --seconds 0.1 --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench_wavenet.log 2>&1
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_samplernn.csv \
    python bench.py --workload samplernn --seconds 0.1 --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench_samplernn.log 2>&1
timeout 600 $NCU --metrics gpu__time_duration.sum -c 100 --csv --log-file $O/launches_features.csv \
    python bench.py --workload features --batch 360 --steps 2 --warmup 1 > $O/ncu_bench_features.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:wavenet_ -s 1 -c 1 -o $O/prof_wavenet -f \
    python bench.py --seconds 0.02 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_full_wavenet.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:samplernn_ -s 1 -c 1 -o $O/prof_samplernn -f \
    python bench.py --workload samplernn --seconds 0.02 --prompt-seconds 0.05 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_full_samplernn.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:"stft_|mulaw_compress" -s 2 -c 2 -o $O/prof_features -f \
    python bench.py --workload features --batch 360 --steps 1 --warmup 1 > $O/ncu_full_features.log 2>&1
ls -la $O
