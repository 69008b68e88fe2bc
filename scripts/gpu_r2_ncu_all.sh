#!/bin/bash
scripts/gpu_r2_ncu.sh sr samplernn_cluster --workload samplernn --batch 128 --seconds 0.05
scripts/gpu_r2_ncu.sh wn7 wavenet7 --dtype bf16 --batch 128 --seconds 0.05
scripts/gpu_r2_ncu.sh wn6 wavenet6 --seconds 0.05
# launch lists of the default bench commands (short horizons; per-launch times are cold-cache and serialised)
for W in "wavenet:--seconds 0.05" "samplernn:--workload samplernn --seconds 0.05" "features:--workload features"; do
  N=${W%%:*}; A=${W#*:}
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_${N}_launches.csv python bench.py $A --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02_${N}_launches.log 2>&1
  echo "launch list $N rc=$? rows=$(wc -l < gpurun_out/r02_${N}_launches.csv)"
done
