#!/bin/bash
mkdir -p gpurun_out
run() { # name envs...
  local name=$1; shift
  env "$@" timeout 120 python bench.py --workload samplernn --seconds 0.5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/sr_sweep_$name.log 2>&1
  echo "$name $(grep -o '"value": [0-9.]*' gpurun_out/sr_sweep_$name.log | head -1) $(grep -o '"cluster_size": [0-9]*, "n_stages": [0-9]*' gpurun_out/sr_sweep_$name.log) $(tail -1 gpurun_out/sr_sweep_$name.log | grep -v metric | cut -c1-100)"
}
run default X=1
run cs2 MMK_SR_CLUSTER=2
run cs8 MMK_SR_CLUSTER=8
run kc32 MMK_SR_KC=32
run kc128 MMK_SR_KC=128
run st3 MMK_SR_NSTAGE=3
run st4kc32 MMK_SR_NSTAGE=4 MMK_SR_KC=32
run ctas144 MMK_SR_CTAS=144
run ctas96 MMK_SR_CTAS=96
