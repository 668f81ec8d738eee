#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "=== attention numerics (2 threads/row rewrite)"
timeout --signal=KILL 300 python -m pytest tests/test_gpu_attention.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/attention.log 2>&1; echo "exit=$?" >> gpurun_out/attention.log; tail -n 12 gpurun_out/attention.log
echo "=== attention perf"
timeout --signal=KILL 200 python tools/prof_attn.py > gpurun_out/attn_perf.log 2>&1; tail -n 5 gpurun_out/attn_perf.log
echo "=== per-kernel device time of one training step (ncu, serialised)"
DTG_GEMM_VARIANT=2 timeout --signal=KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/launches_run.log 2>&1; tail -n 2 gpurun_out/launches_run.log; wc -l gpurun_out/launches.csv
echo "=== ncu attention fwd/bwd"
timeout --signal=KILL 300 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 3 -c 3 -f -o gpurun_out/prof_attn python tools/prof_attn.py > gpurun_out/ncu_attn.log 2>&1; tail -n 3 gpurun_out/ncu_attn.log
