"""Chapter 04 — fully sharded data parallel (ZeRO-3).

    torchrun --standalone --nproc-per-node gpu train_llm.py -d synthetic -m meta-llama/Llama-2-7b-hf -s 4096 [--cpu-offload]

The model is built on the meta device, every group (embedding, decoder layer, head) lives as 1/N flat
shards in NVLink-symmetric memory, is unsharded one layer ahead by a pull kernel and its gradients are
reduce-scattered + AdamW-updated by one fused kernel inside backward (parallel/fsdp.py).  Sharded
checkpoints use the torch.distributed.checkpoint directory layout of the reference chapter."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from distributed_training_guide_b200.parallel import strategies  # noqa: E402
from distributed_training_guide_b200.trainer import run_chapter  # noqa: E402

if __name__ == "__main__":
    run_chapter("04-fully-sharded-data-parallel", lambda args: strategies.FullyShardedDataParallel(args))
