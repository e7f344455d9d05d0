#!/bin/bash
# Run each GPU test file in its own process (a CUDA trap in one file must not poison the others).
# Usage: tools/gpu_run_tests.sh [pytest-args...]; logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
rc=0
for f in tests/test_*gpu*.py; do
  n=$(basename "$f" .py)
  timeout 600 python -m pytest "$f" -m gpu -q -x --tb=short "$@" > "gpurun_out/$n.log" 2>&1
  r=$?
  echo "== $n exit $r"; tail -n 15 "gpurun_out/$n.log"
  [ $r -ne 0 ] && rc=$r
done
exit $rc
