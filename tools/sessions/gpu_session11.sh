#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "=== all 1-GPU tests"
timeout --signal=KILL 900 python -m pytest tests/ -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu_all.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_gpu_all.log; tail -n 5 gpurun_out/pytest_gpu_all.log
echo "=== gemm bench"
timeout --signal=KILL 240 python tools/gemm_bench.py --variants 2 --out gpurun_out/gemm_bench_v2b.json > gpurun_out/gemm_bench_v2b.log 2>&1; tail -n 10 gpurun_out/gemm_bench_v2b.log | cut -c1-260
echo "=== bench own N=1"
DTG_PHASE_TIMING=1 timeout --signal=KILL 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_own_n1b.log 2>&1; tail -n 1 gpurun_out/bench_own_n1b.log | cut -c1-1700
echo "=== per-kernel device time of one training step"
timeout --signal=KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches2.csv python bench.py --steps 1 --warmup 3 > gpurun_out/launches_run2.log 2>&1; wc -l gpurun_out/launches2.csv
