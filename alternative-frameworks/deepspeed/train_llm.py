"""DeepSpeed-style entry point: the run is described by a ZeRO JSON config (``--deepspeed_config
ds_config.json``) instead of flags — micro-batch size, AdamW hyper-parameters, WarmupCosineLR, bf16 and the
ZeRO stage all come from the file, exactly the keys the reference's ``ds_config.json`` uses.

    torchrun --standalone --nproc-per-node gpu train_llm.py -e ds-run -d synthetic -m meta-llama/Llama-2-7b-hf \\
        --deepspeed_config ds_config.json
    deepspeed --num_gpus 8 train_llm.py ...          # the deepspeed launcher's --local_rank is accepted

There is no DeepSpeed dependency: the ZeRO stage is mapped onto this repository's engines
(stage 0 -> fused-all-reduce DDP, 1/2 -> DDP + sharded optimizer, 3 -> FSDP with the fused
reduce-scatter+AdamW kernel; offload_optimizer.device == "cpu" -> --cpu-offload), and metrics go to wandb
(rank-0 run) like the reference's script.
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))

from distributed_training_guide_b200.parallel.zero_config import ZeroConfigured  # noqa: E402
from distributed_training_guide_b200.trainer import run_chapter  # noqa: E402

if __name__ == "__main__":
    run_chapter("deepspeed", lambda args: ZeroConfigured(args), require_experiment=True)
