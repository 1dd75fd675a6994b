#!/bin/bash
# gpurun with retries while the pod has no free slot (exit 3 / "transient"); usage: gpurun_retry.sh <timeout> <command string>
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$1" -- "$2" 2>&1)
  echo "$out" | tail -${TAILN:-30}
  if echo "$out" | grep -q "status=transient\|nothing was charged"; then sleep 90; continue; fi
  break
done
