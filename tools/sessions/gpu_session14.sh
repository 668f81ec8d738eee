#!/bin/bash
# 4-GPU session: first run at N=4 (DDP bench + trace, FSDP, 2-D dp2 x tp2 chapter run and 7B bench)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "=== 2-D chapter on 4 GPUs + comm tests"
timeout --signal=KILL 900 python -m pytest tests/test_gpu_chapters.py::test_2d_dp2_tp2_on_four_gpus tests/test_gpu_comm.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_s14.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_s14.log; tail -n 12 gpurun_out/pytest_s14.log | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511"
run() { name=$1; shift
  echo "=== $name"
  env DTG_PHASE_TIMING=1 "$@" > gpurun_out/$name.log 2>&1
  grep '^{"metric' gpurun_out/$name.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(round(d['ms_per_step'],1), round(d['value']), d['clocks']['sm_mhz'], d.get('phases_ms'), d.get('comm_trace'), d['config']['parallelism'])"
  grep -c Error gpurun_out/$name.log
}
run n4_ddp DTG_COMM_TRACE=1 timeout --signal=KILL 600 $TR bench.py --gpus 4 --steps 5 --warmup 3
run n4_fsdp DTG_COMM_TRACE=1 timeout --signal=KILL 600 $TR bench.py --gpus 4 --steps 5 --warmup 3 --parallelism fsdp
run n4_2d timeout --signal=KILL 600 $TR bench.py --gpus 4 --steps 5 --warmup 3 --parallelism 2d --tensor-parallel 2
