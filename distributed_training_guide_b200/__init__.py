"""distributed_training_guide_b200 — a B200-native (sm_100a) distributed causal-LM training
runtime with the capabilities of LambdaLabsML/distributed-training-guide.

Layout:  ``models/`` (Llama / GPT-2), ``ops/`` (autograd wrappers over the hand-written
kernels in ``csrc/``), ``parallel/`` (bootstrap, mesh, NVLink symmetric memory, DDP,
ZeRO-1, FSDP, TP/SP, 2-D, activation checkpointing, offload), ``utils/`` (CLI, logging,
timers, data, checkpointing), ``trainer.py`` (the loop every chapter script shares).
"""
__version__ = "0.1.0"
