#!/bin/bash
# 2-GPU session: TP with optimizer-in-backward, cpu-offload chapter test, N=2 benches with fresh batches
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "=== chapter + TP tests"
timeout --signal=KILL 1200 python -m pytest tests/test_gpu_chapters.py tests/test_gpu_tp.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_s19.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_s19.log; tail -n 12 gpurun_out/pytest_s19.log | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
run() { name=$1; shift
  echo "=== $name"
  env DTG_PHASE_TIMING=1 "$@" > gpurun_out/$name.log 2>&1
  grep '^{"metric' gpurun_out/$name.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(round(d['ms_per_step'],1), round(d['value']), 'e2e', round(d['e2e']['value']), d['clocks']['sm_mhz'], d.get('phases_ms'), d['config']['parallelism'], 'loss', round(d['final_loss'],3))"
}
run n2_tp timeout --signal=KILL 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --parallelism tp
run n2_tp_noov DTG_TP_OVERLAP_OPT=0 timeout --signal=KILL 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --parallelism tp
run n2_ddp timeout --signal=KILL 600 $TR bench.py --gpus 2 --steps 5 --warmup 3
run n2_fsdp timeout --signal=KILL 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --parallelism fsdp
exit 0
