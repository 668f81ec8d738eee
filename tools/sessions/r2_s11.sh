#!/bin/bash
# round 2, session 11 (1 GPU): the driver's GPU tier (pytest -m gpu, smoke), attention variants (forward v1/v2 with the
# polynomial exp2 offload, backward with P/dS through smem / in TMEM) as kernels and inside the 1-GPU bench, and an
# ncu --set full capture of the elementwise / optimizer kernels of one 7B-shaped step.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/r2s11_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2s11_pytest.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python tools/prof_attn.py > gpurun_out/r2s11_attn_perf.log 2>&1; cat gpurun_out/r2s11_attn_perf.log | tail -9
for cfg in "1 ss" "2 ss" "1 ts" "2 ts"; do
  set -- $cfg
  DTG_ATTN_FWD=$1 DTG_ATTN_BWD=$2 timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2s11_bench_f$1_$2.log 2>&1
  grep '^{"metric' gpurun_out/r2s11_bench_f$1_$2.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('attn fwd v$1 bwd $2:', round(d['ms_per_step'],1), 'ms', round(d['value']), 'tok/s e2e', round(d['e2e']['value']), d['clocks'], 'loss', round(d['final_loss'],4))"
  grep -E "\[dtg\]|Error" gpurun_out/r2s11_bench_f$1_$2.log | head -3 | cut -c1-200
done
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:rmsnorm|rope|swiglu|cross_entropy|adamw|embedding" -c 16 -f -o gpurun_out/prof_step python tools/prof_step.py > gpurun_out/ncu_step.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_step.log | cut -c1-200; ls -la gpurun_out/prof_step.ncu-rep
exit 0
