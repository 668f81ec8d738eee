#!/bin/bash
# round 2, session 6 (2 GPUs): FSDP fused gather (after the version-counter fix) vs copy engine; attention forward v2
# (two query tiles per CTA, P in TMEM) numerics + timing against v1 / SDPA / flash-attn-2.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export DTG_BENCH_BUDGET="import=240,build=120,warmup=60,timed=60,e2e=60,teardown=40" DTG_BENCH_STALL_S=8
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests/test_gpu_attention.py -q -x -m gpu > gpurun_out/r2s6_attn_pytest.log 2>&1
echo "attn pytest rc=$?"; tail -4 gpurun_out/r2s6_attn_pytest.log | cut -c1-300
CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/prof_attn.py > gpurun_out/r2s6_attn_perf.log 2>&1; cat gpurun_out/r2s6_attn_perf.log | tail -8
timeout 600 python -m pytest "tests/test_gpu_tp.py::test_fsdp_gpu_matches_single_gpu" -q -s -m gpu > gpurun_out/r2s6_pytest.log 2>&1
echo "fsdp pytest rc=$?"; tail -4 gpurun_out/r2s6_pytest.log | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
run() { name=$1; shift
  echo "=== $name"
  "$@" > gpurun_out/$name.log 2>&1
  echo "rc=$?"
  grep -E '^\{"(metric|impl)' gpurun_out/$name.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)
    if 'unavailable' in d: print(d); continue
    print(round(d['ms_per_step'],1), round(d['value']), 'e2e', round(d['e2e']['value']), d.get('clocks'), d['config']['parallelism'], 'loss', d.get('final_loss'), 'launches', d.get('gpu_launches'))"
  grep -E "WATCHDOG|STALL|\[dtg\]|Error|timed out" gpurun_out/$name.log | cut -c1-300 | head -8
}
run r2s6_fsdp2_gemm timeout 300 $TR bench.py --gpus 2 --steps 8 --warmup 3 --parallelism fsdp
DTG_FSDP_GATHER=ce run r2s6_fsdp2_ce timeout 300 $TR bench.py --gpus 2 --steps 8 --warmup 3 --parallelism fsdp
run r2s6_ddp2 timeout 300 $TR bench.py --gpus 2 --steps 8 --warmup 3
exit 0
