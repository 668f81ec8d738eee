#!/bin/bash
# round 2, session 7 (1 GPU): attention forward v2 after the warp-uniform rescale fix: numerics, timing.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention.py -q -x -m gpu > gpurun_out/r2s7_attn_pytest.log 2>&1
echo "attn pytest rc=$?"; tail -4 gpurun_out/r2s7_attn_pytest.log | cut -c1-300
timeout 300 python tools/prof_attn.py > gpurun_out/r2s7_attn_perf.log 2>&1; tail -8 gpurun_out/r2s7_attn_perf.log
exit 0
