#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "=== attention numerics (3-stage bwd ring, ILP fwd)"
timeout --signal=KILL 300 python -m pytest tests/test_gpu_attention.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/attention.log 2>&1; echo "exit=$?" >> gpurun_out/attention.log; tail -n 4 gpurun_out/attention.log
echo "=== attention perf"
timeout --signal=KILL 200 python tools/prof_attn.py > gpurun_out/attn_perf.log 2>&1; tail -n 4 gpurun_out/attn_perf.log
echo "=== bench own N=1 with phase timing"
DTG_PHASE_TIMING=1 timeout --signal=KILL 600 python bench.py --steps 4 --warmup 3 > gpurun_out/bench_phases.log 2>&1; tail -n 1 gpurun_out/bench_phases.log
echo "=== ncu: attention bwd kernels"
timeout --signal=KILL 300 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -s 2 -c 2 -f -o gpurun_out/prof_attn_bwd python tools/prof_attn.py > gpurun_out/ncu_attn_bwd.log 2>&1; tail -n 2 gpurun_out/ncu_attn_bwd.log
echo "=== ncu: gemm v2 (fixed)"
timeout --signal=KILL 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 1 -f -o gpurun_out/prof_gemm_v2 python tools/prof_gemm.py 2 > gpurun_out/ncu_gemm_v2.log 2>&1; tail -n 2 gpurun_out/ncu_gemm_v2.log
echo "=== all 1-GPU gpu tests (driver-style)"
timeout --signal=KILL 900 python -m pytest tests/ -m gpu -x -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu_all.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_gpu_all.log; tail -n 6 gpurun_out/pytest_gpu_all.log
