#!/bin/bash
# round 2, session 9 (1 GPU): ncu --set full of the attention forward v2 kernel on the 7B shape
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd2_kernel -c 1 -f -o gpurun_out/prof_attn_fwd2 python tools/prof_attn.py > gpurun_out/ncu_attn_fwd2.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_attn_fwd2.log
ls -la gpurun_out/prof_attn_fwd2.ncu-rep
exit 0
