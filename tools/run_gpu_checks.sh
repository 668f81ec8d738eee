#!/bin/bash
# Runs each GPU test file in its own process under a timeout (a hung kernel must not take the
# rest of the checks down) and collects logs under gpurun_out/.  Usage: tools/run_gpu_checks.sh [files...]
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
FILES=${@:-$(ls tests/test_gpu_*.py)}
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvidia_smi.csv 2>&1
for f in $FILES; do
  name=$(basename "$f" .py)
  echo "=== $f"
  timeout --signal=KILL ${DTG_TEST_TIMEOUT:-600} python -m pytest "$f" -m gpu -q --no-header -p no:cacheprovider \
    > "gpurun_out/${name}.log" 2>&1
  echo "exit=$?" >> "gpurun_out/${name}.log"
  tail -n 15 "gpurun_out/${name}.log"
done
