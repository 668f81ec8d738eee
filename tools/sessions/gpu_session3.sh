#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export DTG_TEST_TIMEOUT=300
timeout --signal=KILL 200 python -m pytest tests/test_gpu_elementwise.py -m gpu -q --no-header -p no:cacheprovider -k rope > gpurun_out/rope.log 2>&1; tail -n 3 gpurun_out/rope.log
echo "=== gemm variant 2 numerics"
timeout --signal=KILL 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q --no-header -p no:cacheprovider -k "True-2] or False-2]" > gpurun_out/gemm_v2.log 2>&1; tail -n 3 gpurun_out/gemm_v2.log
echo "=== gemm bench v2"
timeout --signal=KILL 240 python tools/gemm_bench.py --variants 1 2 --out gpurun_out/gemm_bench_v12.json > gpurun_out/gemm_bench_v12.log 2>&1; tail -n 12 gpurun_out/gemm_bench_v12.log
echo "=== bench own arm, gemm v2, own attention"
DTG_GEMM_VARIANT=2 timeout --signal=KILL 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_own3a.log 2>&1; tail -n 2 gpurun_out/bench_own3a.log
echo "=== bench own arm, gemm v2, SDPA attention (debug switch)"
DTG_GEMM_VARIANT=2 DTG_FORCE_REFERENCE=attention timeout --signal=KILL 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_own3b.log 2>&1; tail -n 2 gpurun_out/bench_own3b.log
