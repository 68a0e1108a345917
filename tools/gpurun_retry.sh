#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <script> <log>   -- retries while the pod answers "transient" (nothing charged)
for i in $(seq 1 20); do
  gpurun --timeout "$1" -- "bash $2" > "$3" 2>&1
  if grep -q "status=transient" "$3" || grep -q "rc=3" "$3"; then sleep 60; else break; fi
done
tail -70 "$3"
