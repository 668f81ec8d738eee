#!/bin/bash
# 2-GPU session: chapter scripts on GPUs, multi-GPU tests, FSDP unshard variants (copy engine vs SM kernel, prefetch depth)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "=== chapter + multi-GPU tests"
timeout --signal=KILL 1200 python -m pytest tests/test_gpu_chapters.py tests/test_gpu_tp.py tests/test_gpu_comm.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_s13.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_s13.log; tail -n 25 gpurun_out/pytest_s13.log | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
run() { # name, env..., -- args
  name=$1; shift
  echo "=== $name"
  env DTG_PHASE_TIMING=1 "$@" > gpurun_out/$name.log 2>&1
  grep '^{"metric' gpurun_out/$name.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(round(d['ms_per_step'],1), round(d['value']), d['clocks']['sm_mhz'], d.get('phases_ms'), d.get('comm_trace'))"
}
run fsdp_ce_d1 DTG_FSDP_AG=ce DTG_FSDP_PREFETCH=1 timeout --signal=KILL 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --parallelism fsdp
run fsdp_ce_d2 DTG_FSDP_AG=ce DTG_FSDP_PREFETCH=2 timeout --signal=KILL 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --parallelism fsdp
run fsdp_sm_d2 DTG_FSDP_AG=sm DTG_FSDP_PREFETCH=2 timeout --signal=KILL 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --parallelism fsdp
run ddp_trace DTG_COMM_TRACE=1 timeout --signal=KILL 600 $TR bench.py --gpus 2 --steps 5 --warmup 3
run ddp_cb48 DTG_COMM_BLOCKS=48 timeout --signal=KILL 600 $TR bench.py --gpus 2 --steps 5 --warmup 3
