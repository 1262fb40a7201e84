#!/bin/bash
# compute-sanitizer over the GPU parity tests (memcheck, then racecheck on the refresh / route tests).
# Run on a GPU box:  gpurun --timeout 1500 -- 'bash tools/sanitize.sh > gpurun_out/sanitize.log 2>&1'
set -u
cd "$(dirname "$0")/.."
SAN=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $SAN --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "not c3 and not c5" || echo "memcheck: FAILED ($?)"
timeout 500 $SAN --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "pool_sizes or ingest or ties" || echo "racecheck: FAILED ($?)"
