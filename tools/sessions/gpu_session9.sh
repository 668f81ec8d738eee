#!/bin/bash
# 8-GPU session: headline bench (DDP+ZeRO-1) at N=8, FSDP at N=8, collectives bandwidth at N=8
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
echo "=== bench own N=8 (ddp + zero1)"
timeout --signal=KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_own_n8.log 2>&1; tail -n 2 gpurun_out/bench_own_n8.log | cut -c1-1600
echo "=== bench own N=8 (fsdp)"
timeout --signal=KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 8 --steps 5 --warmup 3 --parallelism fsdp > gpurun_out/bench_fsdp_n8.log 2>&1; tail -n 2 gpurun_out/bench_fsdp_n8.log | cut -c1-1600
echo "=== comm tests N=8"
timeout --signal=KILL 600 python -m pytest tests/test_gpu_comm.py -m gpu -q --no-header -p no:cacheprovider -s -k symmetric > gpurun_out/comm8.log 2>&1; tail -n 6 gpurun_out/comm8.log | cut -c1-900
