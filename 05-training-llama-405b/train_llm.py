"""Chapter 05 — the large-model recipe (Llama-3.1-405B class): FSDP + optional CPU offload of the optimizer
(--cpu-offload), activation checkpointing (--checkpoint-activations) and explicit layer prefetch
(--prefetch-layers), launched on many nodes by launch.sh.

    bash launch.sh      # see README.md; single node: torchrun --standalone --nproc-per-node gpu train_llm.py ...

Same flags as the reference chapter (05-training-llama-405b/train_llm.py:455-472)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from distributed_training_guide_b200.parallel import strategies  # noqa: E402
from distributed_training_guide_b200.trainer import run_chapter  # noqa: E402

if __name__ == "__main__":
    run_chapter("05-training-llama-405b", lambda args: strategies.FullyShardedDataParallel(args))
