#!/bin/bash
# 1-GPU session: L2-aware tile order: numerics, sustained GEMM with and without grouping, shapes test, bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "=== gemm tests"
timeout --signal=KILL 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_shapes.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_s17.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_s17.log; tail -n 8 gpurun_out/pytest_s17.log | cut -c1-300
echo "=== sustained gemm, grouped raster"
timeout --signal=KILL 400 python tools/gemm_sustained.py --seconds 1.5 --shapes square_8192 7b_down_fwd 7b_gateup_dgrad 7b_gateup_wgrad 7b_qkv_wgrad 7b_lmhead_wgrad --out gpurun_out/gemm_sustained_grouped.json > gpurun_out/gemm_sustained_grouped.log 2>&1; tail -n 8 gpurun_out/gemm_sustained_grouped.log | cut -c1-420
echo "=== bench own N=1"
DTG_PHASE_TIMING=1 timeout --signal=KILL 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_own_n1d.log 2>&1; tail -n 1 gpurun_out/bench_own_n1d.log | cut -c1-1700
echo "=== bench own N=1, grouping off"
DTG_GEMM_L2_BUDGET_MB=100000 DTG_PHASE_TIMING=1 timeout --signal=KILL 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_own_n1e.log 2>&1; tail -n 1 gpurun_out/bench_own_n1e.log | cut -c1-700
