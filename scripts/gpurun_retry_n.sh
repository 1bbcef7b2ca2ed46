#!/bin/bash
# dev tool: like gpurun_retry.sh for N GPUs; usage: gpurun_retry_n.sh <gpus> <timeout_s> <log> <command...>
N=$1; T=$2; LOG=$3; shift 3
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --gpus "$N" --timeout "$T" -- "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
