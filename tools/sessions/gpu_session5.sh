#!/bin/bash
# 2-GPU session: NVLink collectives vs NCCL, DDP+ZeRO-1 parity, bench at N=2 (own + reference), N=1 re-check
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
echo "=== comm tests (2 GPUs)"
timeout --signal=KILL 600 python -m pytest tests/test_gpu_comm.py -m gpu -q --no-header -p no:cacheprovider -s > gpurun_out/comm.log 2>&1; echo "exit=$?" >> gpurun_out/comm.log; tail -n 25 gpurun_out/comm.log
echo "=== bench own N=1"
timeout --signal=KILL 600 python bench.py --steps 4 --warmup 3 > gpurun_out/bench_own_n1.log 2>&1; tail -n 1 gpurun_out/bench_own_n1.log
echo "=== bench own N=2"
timeout --signal=KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/bench_own_n2.log 2>&1; tail -n 2 gpurun_out/bench_own_n2.log
echo "=== bench reference N=2"
timeout --signal=KILL 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_ref_n2.log 2>&1; tail -n 1 gpurun_out/bench_ref_n2.log
