#!/bin/bash
# 1-GPU session: sustained (power-capped) GEMM vs cuBLAS, large-model shapes, ncu of the fixed GEMM, bench sanity
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "=== sustained gemm"
timeout --signal=KILL 400 python tools/gemm_sustained.py --seconds 2.0 > gpurun_out/gemm_sustained.log 2>&1; tail -n 8 gpurun_out/gemm_sustained.log | cut -c1-400
echo "=== large-model shapes + chapter 01"
timeout --signal=KILL 900 python -m pytest tests/test_gpu_shapes.py tests/test_gpu_chapters.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_s16.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_s16.log; tail -n 12 gpurun_out/pytest_s16.log | cut -c1-300
echo "=== ncu gemm (qkv fwd shape, 2-CTA)"
timeout --signal=KILL 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 1 -f -o gpurun_out/prof_gemm_v3 python tools/prof_gemm.py 2 4096 12288 4096 > gpurun_out/ncu_gemm_v3.log 2>&1; tail -n 2 gpurun_out/ncu_gemm_v3.log
echo "=== bench own N=1"
DTG_PHASE_TIMING=1 timeout --signal=KILL 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_own_n1c.log 2>&1; tail -n 1 gpurun_out/bench_own_n1c.log | cut -c1-1700
