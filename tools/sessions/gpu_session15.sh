#!/bin/bash
# 4-GPU session: 2-D (dp2 x tp2) after the norm-scratch fix
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout --signal=KILL 600 python -m pytest tests/test_gpu_chapters.py::test_2d_dp2_tp2_on_four_gpus -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_s15.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_s15.log; tail -n 12 gpurun_out/pytest_s15.log | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511"
DTG_PHASE_TIMING=1 timeout --signal=KILL 600 $TR bench.py --gpus 4 --steps 5 --warmup 3 --parallelism 2d --tensor-parallel 2 > gpurun_out/n4_2d.log 2>&1
grep '^{"metric' gpurun_out/n4_2d.log | cut -c1-1500; grep -c Error gpurun_out/n4_2d.log
