#!/bin/bash
# usage: tools/gpu_retry.sh <tag> <timeout_s> [--gpus N] -- '<command>'
# Retries `gpurun` while it answers "no slot right now" (exit 3), every 3 minutes, for up to ~1 hour.
tag=$1; shift; tmo=$1; shift
extra=()
while [ "$1" != "--" ]; do extra+=("$1"); shift; done
shift
mkdir -p gpurun_out
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$tmo" "${extra[@]}" -- "$1" > "gpurun_out/${tag}_call.log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && [ $rc -ne 2 ]; then exit $rc; fi
  if [ $rc -eq 2 ] && ! grep -q "another call" "gpurun_out/${tag}_call.log"; then exit $rc; fi
  sleep 180
done
exit 3
