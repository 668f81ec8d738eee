#!/bin/bash
# 2 GPUs: attention pipeline trace (GPU 0), host-side profile of a TP step, phase timing of TP / FSDP
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "=== attention bwd in-kernel trace"
timeout --signal=KILL 200 python tools/trace_attn.py > gpurun_out/attn_trace.log 2>&1; tail -n 34 gpurun_out/attn_trace.log
echo "=== TP N=2: cpu profile + phases"
DTG_PHASE_TIMING=1 DTG_CPU_PROFILE=gpurun_out/cpu_profile_tp.txt timeout --signal=KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 4 --warmup 3 --parallelism tp > gpurun_out/bench_tp_n2.log 2>&1; tail -n 2 gpurun_out/bench_tp_n2.log | cut -c1-1800; head -n 40 gpurun_out/cpu_profile_tp.txt
echo "=== single: cpu profile"
DTG_CPU_PROFILE=gpurun_out/cpu_profile_single.txt timeout --signal=KILL 600 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_single_prof.log 2>&1; head -n 12 gpurun_out/cpu_profile_single.txt
