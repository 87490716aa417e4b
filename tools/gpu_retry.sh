#!/bin/bash
# usage: tools/gpu_retry.sh <timeout-seconds> <log-file> '<command>' [gpus]   — retries while the pod answers "busy" (exit 3)
T=$1; LOG=$2; CMD=$3; G=${4:-1}
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout "$T" -- "$CMD" > "$LOG" 2>&1; else /usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$CMD" > "$LOG" 2>&1; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
