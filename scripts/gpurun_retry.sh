#!/bin/bash
# retries a gpurun call while the pod answers "no slot" (exit 3); usage: gpurun_retry.sh <timeout> <command string> [gpus]
T=$1; CMD=$2; G=${3:-1}
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$CMD" > /tmp/gpurun_last.txt 2>&1; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$CMD" > /tmp/gpurun_last.txt 2>&1; fi
  rc=$?
  if [ $rc -ne 3 ]; then tail -60 /tmp/gpurun_last.txt; exit $rc; fi
  sleep 45
done
echo "gave up after 40 tries"; exit 3
