#!/bin/bash
# 2-GPU session: chapter scripts on GPUs, multi-GPU tests, N=2 benches after the 128-bit fix
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "=== chapter + multi-GPU tests"
timeout --signal=KILL 1200 python -m pytest tests/test_gpu_chapters.py tests/test_gpu_tp.py tests/test_gpu_comm.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/pytest_s12.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_s12.log; tail -n 25 gpurun_out/pytest_s12.log | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
echo "=== bench ddp N=2"
DTG_PHASE_TIMING=1 timeout --signal=KILL 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_own_n2b.log 2>&1; tail -n 1 gpurun_out/bench_own_n2b.log | cut -c1-1700
echo "=== bench fsdp N=2"
DTG_PHASE_TIMING=1 timeout --signal=KILL 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --parallelism fsdp > gpurun_out/bench_fsdp_n2b.log 2>&1; tail -n 1 gpurun_out/bench_fsdp_n2b.log | cut -c1-1700
echo "=== bench tp N=2"
DTG_PHASE_TIMING=1 timeout --signal=KILL 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --parallelism tp > gpurun_out/bench_tp_n2b.log 2>&1; tail -n 1 gpurun_out/bench_tp_n2b.log | cut -c1-1700
