"""Chapter 07 — 2-D parallelism: FSDP over the data-parallel dimension x tensor parallel inside it.

    torchrun --standalone --nproc-per-node 8 train_llm.py -d synthetic -m meta-llama/Meta-Llama-3-70B -tp 4 -s 4096

-tp/--tensor-parallel sets the size of the (contiguous) tensor-parallel groups; world/tp data-parallel
replicas shard the TP-local parameters with the fused FSDP kernels."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from distributed_training_guide_b200.parallel import strategies  # noqa: E402
from distributed_training_guide_b200.trainer import run_chapter  # noqa: E402

if __name__ == "__main__":
    run_chapter("07-2d-parallel", lambda args: strategies.TwoDParallel(args))
