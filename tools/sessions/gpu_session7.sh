#!/bin/bash
# 2-GPU session: tensor-parallel fused GEMM modes + TP/FSDP engines vs single GPU, benches for fsdp / tp at N=2
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "=== TP / FSDP GPU tests"
timeout --signal=KILL 900 python -m pytest tests/test_gpu_tp.py -m gpu -q --no-header -p no:cacheprovider -s > gpurun_out/tp_tests.log 2>&1; echo "exit=$?" >> gpurun_out/tp_tests.log; tail -n 30 gpurun_out/tp_tests.log
echo "=== bench fsdp N=2"
timeout --signal=KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 4 --warmup 3 --parallelism fsdp > gpurun_out/bench_fsdp_n2.log 2>&1; tail -n 3 gpurun_out/bench_fsdp_n2.log | cut -c1-1500
echo "=== bench tp N=2"
timeout --signal=KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 4 --warmup 3 --parallelism tp > gpurun_out/bench_tp_n2.log 2>&1; tail -n 3 gpurun_out/bench_tp_n2.log | cut -c1-1500
