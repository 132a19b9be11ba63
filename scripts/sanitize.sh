#!/bin/bash
# dev: compute-sanitizer (memcheck, then racecheck) over small runs of every step path
set -o pipefail
for tool in memcheck racecheck; do
  echo "== $tool =="
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 3 python -m pytest tests/test_fast2.py -m gpu -x -q \
    -k "tiny or mixed or linked or job_modes or beside" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|Race|hazard" | head -12
done
