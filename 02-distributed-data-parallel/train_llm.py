"""Chapter 02 — data parallel training with ZeRO-1 optimizer sharding on N GPUs.

    torchrun --standalone --nproc-per-node gpu train_llm.py -d synthetic -m meta-llama/Llama-2-7b-hf -s 4096

One process per GPU.  Gradients of each bucket (embedding / decoder layer / head) are reduced, the
optimizer shard updated and the new parameters re-distributed by ONE NVLink kernel per bucket that
overlaps the rest of backward (parallel/ddp.py, csrc/comm.cu).  Flags, log records and checkpoint files
follow the reference chapter (02-distributed-data-parallel/train_llm.py)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from distributed_training_guide_b200.parallel import strategies  # noqa: E402
from distributed_training_guide_b200.trainer import run_chapter  # noqa: E402

if __name__ == "__main__":
    run_chapter("02-distributed-data-parallel", lambda args: strategies.DataParallelZero1(args))
