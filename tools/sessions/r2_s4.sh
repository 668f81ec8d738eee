#!/bin/bash
# round 2, session 4 (2 GPUs): every multi-GPU test (engines, chapters, collectives, fused FSDP gather), then FSDP
# benches: unshard fused into the GEMMs vs copy-engine unshard.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export DTG_BENCH_BUDGET="import=240,build=120,warmup=60,timed=60,e2e=60,teardown=40" DTG_BENCH_STALL_S=8
timeout 1200 python -m pytest tests/test_gpu_comm.py tests/test_gpu_tp.py tests/test_gpu_loader.py tests/test_gpu_bgather.py tests/test_gpu_chapters.py -q -s -m gpu > gpurun_out/r2s4_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r2s4_pytest.log | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
run() { name=$1; shift
  echo "=== $name"
  "$@" > gpurun_out/$name.log 2>&1
  echo "rc=$?"
  grep '^{"metric' gpurun_out/$name.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(round(d['ms_per_step'],1), round(d['value']), 'e2e', round(d['e2e']['value']), d['clocks'], 'exposed', d.get('exposed_comm_ms'), d['config']['parallelism'], 'loss', d['final_loss'], 'launches', d['gpu_launches'], d.get('comm_trace'))"
  grep -E "WATCHDOG|STALL|\[dtg\]|Error|timed out" gpurun_out/$name.log | cut -c1-300 | head -8
}
run r2s4_fsdp2_gemm timeout 300 $TR bench.py --gpus 2 --steps 8 --warmup 3 --parallelism fsdp
DTG_FSDP_GATHER=ce run r2s4_fsdp2_ce timeout 300 $TR bench.py --gpus 2 --steps 8 --warmup 3 --parallelism fsdp
exit 0
