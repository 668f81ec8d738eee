#!/bin/bash
# round 2, session 8 (8 GPUs): the four BASELINE configs on 8xB200 — own arm (default environment, as the driver runs
# it) and the reference arm of chapters 04 / 06 — plus the NVLS on/off comparison and the 8-rank collective tests.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export DTG_BENCH_STALL_S=10 DTG_DEBUG_MARKERS=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
run() { name=$1; shift
  echo "=== $name"
  "$@" > gpurun_out/$name.log 2>&1
  echo "rc=$?"
  grep -E '^\{"(metric|impl)' gpurun_out/$name.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)
    if 'unavailable' in d: print(d); continue
    print(round(d['ms_per_step'],1), round(d['value']), 'e2e', round(d['e2e']['value']), d.get('clocks'), 'exposed', d.get('exposed_comm_ms'), d['config']['parallelism'], 'loss', d.get('final_loss'), 'launches', d.get('gpu_launches'), 'peakGB', d.get('peak_alloc_gb'), d.get('ref_breakdown_ms'))"
  grep -E "WATCHDOG|STALL|\[dtg\]|Error|timed out" gpurun_out/$name.log | cut -c1-300 | head -6
}
run r2s8_ddp8 timeout --signal=KILL 400 $TR bench.py --gpus 8 --steps 20 --warmup 5
DTG_NVLS_KERNELS=0 run r2s8_ddp8_nonvls timeout --signal=KILL 300 $TR bench.py --gpus 8 --steps 10 --warmup 3
run r2s8_fsdp8 timeout --signal=KILL 300 $TR bench.py --gpus 8 --steps 10 --warmup 3 --parallelism fsdp
DTG_FSDP_GATHER=ce run r2s8_fsdp8_ce timeout --signal=KILL 300 $TR bench.py --gpus 8 --steps 10 --warmup 3 --parallelism fsdp
run r2s8_tp8 timeout --signal=KILL 300 $TR bench.py --gpus 8 --steps 10 --warmup 3 --parallelism tp --model meta-llama/Meta-Llama-3-8B --batch 4
run r2s8_2d timeout --signal=KILL 400 $TR bench.py --gpus 8 --steps 5 --warmup 3 --parallelism 2d --tensor-parallel 4 --model meta-llama/Meta-Llama-3-70B
run r2s8_ref_fsdp timeout --signal=KILL 400 $TR bench.py --impl reference --gpus 8 --steps 8 --warmup 3 --parallelism fsdp
run r2s8_ref_tp timeout --signal=KILL 400 $TR bench.py --impl reference --gpus 8 --steps 8 --warmup 3 --parallelism tp --model meta-llama/Meta-Llama-3-8B --batch 4
timeout 300 python -m pytest tests/test_gpu_comm.py::test_symmetric_collectives tests/test_gpu_comm.py::test_nvls_collectives tests/test_gpu_bgather.py -q -s -m gpu > gpurun_out/r2s8_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|^\{" gpurun_out/r2s8_pytest.log | cut -c1-500
exit 0
