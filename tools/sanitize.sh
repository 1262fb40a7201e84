#!/bin/bash
# compute-sanitizer over every kernel of the engine (tools/sanitize_target.py drives them and checks the results against
# the oracle): memcheck, racecheck (shared-memory hazards: the bitonic sort / merge of the worker refresh, the per-warp
# tiles of policy_kernel), synccheck, initcheck.  Run on a GPU box:
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh > gpurun_out/sanitize.log 2>&1'
# The log is what profiles/r02_sanitize.log holds.
set -u
cd "$(dirname "$0")/.."
SAN=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck initcheck; do
  echo "=================================================================== $tool"
  n=2000; mode=full
  [ "$tool" = racecheck ] && { n=500; mode=light; }
  extra=""; [ "$tool" = initcheck ] && extra=""
  timeout ${SAN_TIMEOUT:-300} $SAN --tool $tool $extra --error-exitcode 3 --print-limit 20 python tools/sanitize_target.py $n $mode 2>&1 | grep -v "^$" | tail -40
  echo "$tool exit code: ${PIPESTATUS[0]}"
done
