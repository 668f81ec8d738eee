"""One training step of a 7B-shaped (1 decoder layer) model on one GPU inside a cudaProfiler range, for

    ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_step \
        python tools/prof_step.py

Every hand-written kernel of the step (GEMM fwd/dgrad/wgrad, attention fwd/bwd, RMSNorm fwd/bwd, RoPE, SwiGLU fwd/bwd,
cross entropy, embedding fwd/bwd, fused bucket AdamW) is captured once with its shapes of the headline benchmark."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distributed_training_guide_b200.engine import TrainEngine  # noqa: E402

eng = TrainEngine.create("meta-llama/Llama-2-7b-hf", parallelism="single", batch_size=1, seq_length=4096, num_layers=1)
batches = [{k: v.cuda() for k, v in eng.synthetic_batch(seed=i, pinned=False).items()} for i in range(3)]
for b in batches[:2]:
    eng.step(b)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.step(batches[2])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step")
