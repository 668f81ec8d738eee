#!/bin/bash
# round 2, session 12 (1 GPU): attention defaults flipped (forward v2, backward TS): tests, kernel timing, bench.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_shapes.py -q -m gpu > gpurun_out/r2s12_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2s12_pytest.log | cut -c1-300
timeout 300 python tools/prof_attn.py > gpurun_out/r2s12_attn_perf.log 2>&1; cat gpurun_out/r2s12_attn_perf.log | tail -9
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2s12_bench.log 2>&1
grep '^{"metric' gpurun_out/r2s12_bench.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('bench N=1:', round(d['ms_per_step'],1), 'ms', round(d['value']), 'tok/s e2e', round(d['e2e']['value']), d['clocks'], 'loss', round(d['final_loss'],4), 'launches', d['gpu_launches'])"
exit 0
