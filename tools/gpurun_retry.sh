#!/bin/bash
# usage: [GPUS=N] tools/gpurun_retry.sh <timeout-seconds> <logfile> <command...>   (retries while the pod answers "busy")
T=$1; LOG=$2; shift 2
G=""
if [ -n "$GPUS" ]; then G="--gpus $GPUS"; fi
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $G --timeout "$T" -- "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc (attempt $i)" >> "$LOG"; exit $rc; fi
  sleep 90
done
echo "gave up" >> "$LOG"; exit 3
