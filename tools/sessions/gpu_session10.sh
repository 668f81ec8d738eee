#!/bin/bash
# 2 GPUs: in-kernel all-gather GEMM (A_MODE 3) numerics, TP training parity, TP bench with phases
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout --signal=KILL 600 python -m pytest tests/test_gpu_tp.py -m gpu -q --no-header -p no:cacheprovider -s -k "gemm_modes or tensor_parallel" > gpurun_out/tp_tests2.log 2>&1; echo "exit=$?" >> gpurun_out/tp_tests2.log; tail -n 12 gpurun_out/tp_tests2.log | cut -c1-1200
DTG_PHASE_TIMING=1 timeout --signal=KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 4 --warmup 3 --parallelism tp > gpurun_out/bench_tp_n2b.log 2>&1; tail -n 2 gpurun_out/bench_tp_n2b.log | cut -c1-1700
