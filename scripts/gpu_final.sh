#!/bin/bash
# Round-end style check on one GPU: full parity suite, smoke(), the three default bench lines, reference arm, feature ncu.
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log | cut -c1-400
timeout 600 python bench.py --steps 2 --warmup 3 > $O/bench_wavenet.log 2>&1; tail -1 $O/bench_wavenet.log | cut -c1-200
timeout 600 python bench.py --workload samplernn --steps 2 --warmup 3 > $O/bench_samplernn.log 2>&1; tail -1 $O/bench_samplernn.log | cut -c1-200
timeout 300 python bench.py --workload features --steps 3 --warmup 3 > $O/bench_features.log 2>&1; tail -1 $O/bench_features.log | cut -c1-200
timeout 300 python bench.py --workload features --impl reference --steps 2 --warmup 1 > $O/bench_reference_features.log 2>&1; tail -1 $O/bench_reference_features.log | cut -c1-200
timeout 600 python bench.py --workload samplernn --impl reference --steps 1 --warmup 1 > $O/bench_reference_samplernn.log 2>&1; tail -1 $O/bench_reference_samplernn.log | cut -c1-200
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 100 --csv --log-file $O/launches_features.csv \
    python bench.py --workload features --batch 360 --steps 2 --warmup 1 > $O/ncu_bench_features.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:"stft2048|mulaw_compress_table" -c 2 -o $O/prof_features -f \
    python bench.py --workload features --batch 360 --steps 1 --warmup 1 > $O/ncu_full_features.log 2>&1
ls $O | head -50
