#!/bin/bash
# Session 2 (1 GPU): re-check elementwise + gemm autograd, attention numerics/perf, ncu of both GEMM variants.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export DTG_TEST_TIMEOUT=300
tools/run_gpu_checks.sh tests/test_gpu_elementwise.py
timeout --signal=KILL 200 python -m pytest tests/test_gpu_gemm.py -m gpu -q --no-header -p no:cacheprovider -k "strided or autograd" > gpurun_out/gemm_misc.log 2>&1; tail -n 4 gpurun_out/gemm_misc.log
echo "=== attention numerics"
timeout --signal=KILL 300 python -m pytest tests/test_gpu_attention.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/attention.log 2>&1; echo "exit=$?" >> gpurun_out/attention.log; tail -n 25 gpurun_out/attention.log
echo "=== attention perf"
timeout --signal=KILL 200 python tools/prof_attn.py > gpurun_out/attn_perf.log 2>&1; tail -n 5 gpurun_out/attn_perf.log
echo "=== ncu gemm v1/v2"
for v in 1 2; do
  timeout --signal=KILL 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 1 -f -o gpurun_out/prof_gemm_v$v python tools/prof_gemm.py $v > gpurun_out/ncu_gemm_v$v.log 2>&1; tail -n 3 gpurun_out/ncu_gemm_v$v.log
done
echo "=== smoke with own attention"
timeout --signal=KILL 300 python __graft_entry__.py smoke > gpurun_out/smoke2.log 2>&1; tail -n 3 gpurun_out/smoke2.log
echo "=== bench own arm (own attention, optimizer-in-backward)"
timeout --signal=KILL 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_own2.log 2>&1; tail -n 3 gpurun_out/bench_own2.log
