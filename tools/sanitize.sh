#!/bin/bash
# compute-sanitizer passes over the kernels at small shapes (one GPU).  memcheck catches out-of-bounds TMA boxes / TMEM
# epilogue stores, racecheck shared-memory hazards in the hand-rolled mbarrier pipelines, synccheck barrier misuse.
# tcgen05/TMA coverage of the tools varies by toolkit version; treat a clean run as necessary, not sufficient.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TESTS="tests/test_gpu_elementwise.py tests/test_gpu_gemm.py tests/test_gpu_attention.py"
for tool in memcheck racecheck synccheck; do
  echo "=== compute-sanitizer --tool $tool"
  timeout --signal=KILL 1200 compute-sanitizer --tool $tool --error-exitcode 7 --launch-timeout 0 \
      python -m pytest $TESTS -m gpu -q --no-header -p no:cacheprovider -x -k "not perf and not bandwidth" \
      > gpurun_out/sanitize_$tool.log 2>&1
  echo "exit=$?" >> gpurun_out/sanitize_$tool.log
  grep -E "ERROR SUMMARY|exit=|passed|failed" gpurun_out/sanitize_$tool.log | tail -n 4
done
