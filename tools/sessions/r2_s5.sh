#!/bin/bash
# round 2, session 5 (2 GPUs): FSDP fused gather v2 (remote pieces only, no per-GEMM barrier) vs copy engine; the
# reference arms of chapters 04 / 06 / 07 start up on this image (2-layer smoke).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export DTG_BENCH_BUDGET="import=240,build=120,warmup=60,timed=60,e2e=60,teardown=40" DTG_BENCH_STALL_S=8
timeout 600 python -m pytest tests/test_gpu_bgather.py "tests/test_gpu_tp.py::test_fsdp_gpu_matches_single_gpu" -q -s -m gpu > gpurun_out/r2s5_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2s5_pytest.log | cut -c1-300
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
    print(round(d['ms_per_step'],1), round(d['value']), 'e2e', round(d['e2e']['value']), d.get('clocks'), d['config']['parallelism'], 'loss', d.get('final_loss'), 'launches', d.get('gpu_launches'), d.get('ref_breakdown_ms'), d.get('environment_shims'))"
  grep -E "WATCHDOG|STALL|\[dtg\]|Error|timed out" gpurun_out/$name.log | cut -c1-300 | head -8
}
run r2s5_fsdp2_gemm timeout 300 $TR bench.py --gpus 2 --steps 8 --warmup 3 --parallelism fsdp
DTG_FSDP_GATHER=ce run r2s5_fsdp2_ce timeout 300 $TR bench.py --gpus 2 --steps 8 --warmup 3 --parallelism fsdp
run r2s5_ref_fsdp timeout 400 $TR bench.py --impl reference --gpus 2 --steps 3 --warmup 2 --parallelism fsdp --layers 2
run r2s5_ref_tp timeout 400 $TR bench.py --impl reference --gpus 2 --steps 3 --warmup 2 --parallelism tp --model meta-llama/Meta-Llama-3-8B --layers 2 --batch 1
run r2s5_ref_2d timeout 400 $TR bench.py --impl reference --gpus 2 --steps 3 --warmup 2 --parallelism 2d --tensor-parallel 2 --model meta-llama/Meta-Llama-3-70B --layers 2 --batch 1
exit 0
