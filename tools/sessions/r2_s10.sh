#!/bin/bash
# round 2, session 10 (1 GPU): attention v2 with the polynomial exp2 offload (numerics + timing); ncu --set full of
# every kernel of one 7B-shaped training step; the 1-GPU bench.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention.py -q -x -m gpu > gpurun_out/r2s10_attn_pytest.log 2>&1
echo "attn pytest rc=$?"; tail -3 gpurun_out/r2s10_attn_pytest.log | cut -c1-300
timeout 300 python tools/prof_attn.py > gpurun_out/r2s10_attn_perf.log 2>&1; tail -6 gpurun_out/r2s10_attn_perf.log
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_step python tools/prof_step.py > gpurun_out/ncu_step.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_step.log; ls -la gpurun_out/prof_step.ncu-rep
for v in 1 2; do
  DTG_ATTN_FWD=$v timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2s10_bench_attn$v.log 2>&1
  grep '^{"metric' gpurun_out/r2s10_bench_attn$v.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('attn v$v', round(d['ms_per_step'],1), round(d['value']), 'e2e', round(d['e2e']['value']), d['clocks'], 'loss', d['final_loss'])"
done
exit 0
