#!/bin/bash
# First GPU session: numerics of the simple kernels + tcgen05 GEMM (both variants), GEMM throughput
# vs cuBLAS, and the two bench arms at N=1.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export DTG_TEST_TIMEOUT=300
tools/run_gpu_checks.sh tests/test_gpu_elementwise.py
echo "=== gemm variant 1"
timeout --signal=KILL 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q --no-header -p no:cacheprovider -k "not (True-2] or False-2])" > gpurun_out/gemm_v1.log 2>&1; echo "exit=$?" >> gpurun_out/gemm_v1.log; tail -n 12 gpurun_out/gemm_v1.log
echo "=== gemm variant 2"
timeout --signal=KILL 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q --no-header -p no:cacheprovider -k "True-2] or False-2]" > gpurun_out/gemm_v2.log 2>&1; echo "exit=$?" >> gpurun_out/gemm_v2.log; tail -n 12 gpurun_out/gemm_v2.log
echo "=== gemm bench"
timeout --signal=KILL 240 python tools/gemm_bench.py --variants 1 --out gpurun_out/gemm_bench_v1.json > gpurun_out/gemm_bench_v1.log 2>&1; tail -n 12 gpurun_out/gemm_bench_v1.log
timeout --signal=KILL 240 python tools/gemm_bench.py --variants 2 --out gpurun_out/gemm_bench_v2.json > gpurun_out/gemm_bench_v2.log 2>&1; tail -n 12 gpurun_out/gemm_bench_v2.log
echo "=== smoke (attention via SDPA for now)"
DTG_FORCE_REFERENCE=attention timeout --signal=KILL 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -n 5 gpurun_out/smoke.log
echo "=== bench own arm"
DTG_FORCE_REFERENCE=attention timeout --signal=KILL 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_own.log 2>&1; tail -n 5 gpurun_out/bench_own.log
echo "=== bench reference arm"
timeout --signal=KILL 1200 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.log 2>&1; tail -n 3 gpurun_out/bench_ref.log
