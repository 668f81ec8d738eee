"""Command-line surface of the chapter scripts.

One builder covers the whole flag matrix of the reference (reference
parsers at ``01-single-gpu/train_llm.py:289-303``, ``04-...:370-385``, ``05-...:455-472``,
``07-2d-parallel/train_llm.py:388-403``, deepspeed ``train_llm.py:309-322``): every chapter
gets the common flags and opts into its extras by name.
"""
from __future__ import annotations

import argparse

COMMON = ("experiment-name", "dataset-name", "dataset-subset", "model-name", "save-dir", "seed",
          "num-epochs", "lr", "batch-size", "log-freq", "ckpt-freq", "seq-length")

CHAPTER_EXTRAS = {
    "01-single-gpu": (),
    "02-distributed-data-parallel": (),
    "04-fully-sharded-data-parallel": ("cpu-offload",),
    "05-training-llama-405b": ("cpu-offload", "checkpoint-activations", "prefetch-layers"),
    "06-tensor-parallel": (),
    "07-2d-parallel": ("tensor-parallel",),
    "deepspeed": ("local_rank", "zero_config"),
}


def get_parser(chapter: str = "01-single-gpu", require_experiment: bool = False) -> argparse.ArgumentParser:
    extras = CHAPTER_EXTRAS[chapter]
    p = argparse.ArgumentParser(description=f"{chapter}: causal-LM training on B200")
    p.add_argument("-e", "--experiment-name", default=None, required=require_experiment,
                   help="enables checkpointing/resume under <save-dir>/<experiment-name>")
    p.add_argument("-d", "--dataset-name", default=None, required=True,
                   help="'synthetic', a local .bin/.txt/.jsonl path, or a Hugging Face dataset id")
    p.add_argument("--dataset-subset", default=None)
    p.add_argument("-m", "--model-name", default=None, required=True,
                   help="embedded config id (e.g. meta-llama/Llama-2-7b-hf) or a directory with config.json")
    p.add_argument("--save-dir", default="../outputs")
    p.add_argument("--seed", default=0, type=int)
    p.add_argument("--num-epochs", default=100, type=int)
    p.add_argument("--lr", default=3e-5, type=float)
    p.add_argument("-b", "--batch-size", default=1, type=int)
    p.add_argument("--log-freq", default=10, type=int)
    p.add_argument("--ckpt-freq", default=500, type=int)
    p.add_argument("-s", "--seq-length", default=1024, type=int)
    # additions of this framework (absent from the reference; all default to its behaviour)
    p.add_argument("--max-steps", default=None, type=int, help="stop after this many optimizer steps")
    p.add_argument("--num-samples", default=None, type=int, help="size of the synthetic dataset")
    p.add_argument("--grad-accum-steps", default=1, type=int,
                   help="micro-batches per optimizer step (reference: related-topics/gradient-accumulation)")
    p.add_argument("--deterministic", action="store_true",
                   help="seeded loaders + rng.pt in checkpoints (reference: related-topics/determinism)")
    p.add_argument("--lr-scaling", choices=("none", "linear", "sqrt"), default="none",
                   help="scale --lr by the data-parallel size (reference: related-topics/effective-batch-size-and-lr)")
    p.add_argument("--wandb", choices=("off", "rank0", "local-rank0", "all"), default="off",
                   help="reference: related-topics/wandb-configurations")
    p.add_argument("--device", default=None, help="cuda (default when available) or cpu")
    p.add_argument("--num-workers", default=1, type=int, help="DataLoader worker processes (reference: 1)")
    p.add_argument("--pretrained", choices=("auto", "require", "never"), default=None,
                   help="load local Hugging Face safetensors for --model-name (chapter 05 defaults to auto: load "
                        "them when they exist on disk; other chapters default to never = random init)")
    if "cpu-offload" in extras:
        p.add_argument("--cpu-offload", default=False, action="store_true")
    if "checkpoint-activations" in extras:
        p.add_argument("--checkpoint-activations", default=False, action="store_true")
    if "prefetch-layers" in extras:
        p.add_argument("--prefetch-layers", default=False, action="store_true")
    if "tensor-parallel" in extras:
        p.add_argument("-tp", "--tensor-parallel", default=8, type=int)
    if "local_rank" in extras:
        p.add_argument("--local_rank", type=int, default=None)
    if "zero_config" in extras:
        p.add_argument("--deepspeed_config", default=None, help="DeepSpeed-style JSON (ds_config.json)")
        p.add_argument("--deepspeed", action="store_true", help="accepted for launcher compatibility")
    return p
