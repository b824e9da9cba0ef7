#!/bin/bash
# gpurun with retries while the pod has no free GPU slot (exit code 3 / "transient").  Usage: gpurun_retry.sh <timeout> '<command>'
T=$1; shift
for try in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$out"; exit $rc
done
echo "gpurun_retry: gave up"; exit 3
