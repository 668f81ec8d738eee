#!/bin/bash
# 8-GPU session: headline DDP+ZeRO-1 bench, FSDP, 2-D (dp4 x tp2), each with the communication trace
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
run() { name=$1; shift
  echo "=== $name"
  env DTG_PHASE_TIMING=1 DTG_COMM_TRACE=1 "$@" > gpurun_out/$name.log 2>&1
  grep '^{"metric' gpurun_out/$name.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(round(d['ms_per_step'],1), round(d['value']), 'e2e', round(d['e2e']['value']), d['clocks'], d.get('phases_ms'), d.get('comm_trace'), d['config']['parallelism'], 'loss', d['final_loss'])"
  grep -c "Error" gpurun_out/$name.log
}
run n8_ddp timeout --signal=KILL 600 $TR bench.py --gpus 8 --steps 5 --warmup 3
run n8_fsdp timeout --signal=KILL 600 $TR bench.py --gpus 8 --steps 5 --warmup 3 --parallelism fsdp
run n8_2d timeout --signal=KILL 600 $TR bench.py --gpus 8 --steps 5 --warmup 3 --parallelism 2d --tensor-parallel 2
exit 0
