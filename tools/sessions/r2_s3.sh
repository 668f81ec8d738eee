#!/bin/bash
# round 2, session 3 (2 GPUs): new kernels against their references — B-gather GEMM, VMM arena + NVLS collectives,
# in-switch reduce-scatter, loader ring guard — then quick N=2 benches (ddp with/without multimem kernels, tp).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export DTG_BENCH_BUDGET="import=240,build=120,warmup=60,timed=60,e2e=60,teardown=40" DTG_BENCH_STALL_S=8 DTG_DEBUG_MARKERS=1
timeout 600 python -m pytest tests/test_gpu_bgather.py tests/test_gpu_comm.py tests/test_gpu_loader.py tests/test_gpu_tp.py -x -q -s -m gpu > gpurun_out/r2s3_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2s3_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
run() { name=$1; shift
  echo "=== $name"
  "$@" > gpurun_out/$name.log 2>&1
  echo "rc=$?"
  grep '^{"metric' gpurun_out/$name.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(round(d['ms_per_step'],1), round(d['value']), 'e2e', round(d['e2e']['value']), d['clocks'], 'exposed', d.get('exposed_comm_ms'), d['config']['parallelism'], 'loss', d['final_loss'], 'launches', d['gpu_launches'])"
  grep -E "WATCHDOG|STALL|\[dtg\]|Error|timed out" gpurun_out/$name.log | cut -c1-300 | head -8
}
run r2s3_ddp2 timeout 300 $TR bench.py --gpus 2 --steps 8 --warmup 3
DTG_NVLS_KERNELS=1 run r2s3_ddp2_nvls timeout 300 $TR bench.py --gpus 2 --steps 8 --warmup 3
run r2s3_tp2 timeout 300 $TR bench.py --gpus 2 --steps 8 --warmup 3 --parallelism tp --model meta-llama/Meta-Llama-3-8B --batch 1
DTG_TP_RS=push run r2s3_tp2_push timeout 300 $TR bench.py --gpus 2 --steps 8 --warmup 3 --parallelism tp --model meta-llama/Meta-Llama-3-8B --batch 1
exit 0
